// K6: quadratic forms  v_n = psi_n^T C psi_n  on the FP64 tensor pipe.
#include "pcq_mma.cuh"
#include <cstdlib>
#include <cstring>

namespace {

// K6: quadratic forms  v_n = psi_n^T C psi_n  (pushed-forward variance of lreg.predict, lreg/lreg.py:112-143,
// 164-198) on the FP64 tensor pipe.  With U = triu(C, 1) + diag(C) / 2,  v = 2 * rowsum((Psi U) o Psi): a GEMM
// whose B operand is upper triangular, so output tile (128 samples x column tile J) only contracts the basis
// indices i < 128 (J + 1) -- N K^2 flops instead of the 2 N K^2 of a factor-based evaluation through K2.
// Both operands use the SYRK's packed layout with the roles renamed (contraction index = basis term i):
//     Zq[i-block (8 terms)][sample (Nc)][8 terms]      Bq[i-block][column j (Kp)][8 terms]
// so the pipeline and the inner loop are those of pcq_syrk_kernel<128>; the epilogue multiplies the
// accumulators with Psi[sample][j] (read back from Zq), reduces over j and writes one partial per column
// tile; a second kernel adds the partials of a sample in tile order (deterministic).
struct QformParams {
    const double* zq;      // [Kp/8][Nc][8]
    const double* bq;      // [Kp/8][ldb][8]: U (quadratic form, ldb = Kp) or the coefficient vectors (ldb = Mp)
    int64_t Nc;            // samples in this chunk, padded to 128
    int Kp, nJ, ldb;       // padded number of terms, column tiles of the B operand, its row count
    int64_t units;         // (Nc / 128) * nJ
    double* part;          // quadratic form: [nJ][Nc]
    double* y; int64_t ldy; int64_t n; int m;     // linear forms: Y (n x m), leading dimension ldy
};

struct QfPackParams {
    const double* x; int64_t ldx; int64_t n; int64_t Nc; int64_t npairs;
    int K, Kp, sdim, pmax, nslots, width;
    const double* rec_a; const double* rec_b; const double* p1;
    const unsigned long long* desc4; const uint16_t* descw;
    double* zq;
};

// design-matrix chunk in term-block order: 16-sample groups, four lanes per (sample, block of 8 terms)
template <int W>
__global__ void __launch_bounds__(256) pcq_qf_pack_rows_kernel(QfPackParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TS = PACK_TS;
    double* T = reinterpret_cast<double*>(smem_raw);      // [nslots + 1][18]
    const int s = P.sdim, tid = threadIdx.x;
    const int KB = P.Kp / 8;
    for (int i = tid; i < (P.nslots + 1) * TS; i += 256) T[i] = 0.0;
    for (int64_t pair = blockIdx.x; pair < P.npairs; pair += gridDim.x) {
        const int64_t n0 = pair * 16;
        __syncthreads();
        if (n0 + 16 > P.n) {
            for (int i = tid; i < P.nslots * TS; i += 256) T[i] = 0.0;
            __syncthreads();
        }
        for (int idx = tid; idx < 16 * s; idx += 256) {
            const int sm = idx / s, c = idx - sm * s;
            const int64_t n = n0 + sm;
            if (n >= P.n) continue;
            if (c == 0) T[sm] = 1.0;
            pcq_fill_1d(P.x[n * P.ldx + c], c, s, P.pmax, P.rec_a, P.rec_b, P.p1, T + sm, TS);
        }
        __syncthreads();
        // four lanes per (sample, term block), a term pair each: one warp store instruction covers eight
        // samples x 64 B = 512 contiguous bytes
        for (int idx = tid; idx < 64 * KB; idx += 256) {
            const int q = idx & 3, sm = (idx >> 2) & 15, ib = idx >> 6;
            double v[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int term = ib * 8 + 2 * q + t;
                double val = 0.0;
                if (term < P.K) {
                    if (W > 0) {
                        const unsigned long long d = P.desc4[term];
                        val = T[(int)(d & 0xffffull) * TS + sm];
#pragma unroll
                        for (int w = 1; w < W; ++w) val = __dmul_rn(val, T[(int)((d >> (16 * w)) & 0xffffull) * TS + sm]);
                    } else {
                        const uint16_t* dw = P.descw + (size_t)term * P.width;
                        val = T[(int)dw[0] * TS + sm];
                        for (int w = 1; w < P.width; ++w) val = __dmul_rn(val, T[(int)dw[w] * TS + sm]);
                    }
                }
                v[t] = val;
            }
            __stcs(reinterpret_cast<double2*>(P.zq + ((size_t)ib * P.Nc + n0 + sm) * 8 + 2 * q), make_double2(v[0], v[1]));
        }
    }
}

// Bq[ib][j][t] = U[ib * 8 + t][j],  U = triu(C, 1) + diag(C) / 2  (zero outside K x K)
__global__ void __launch_bounds__(256) pcq_qf_pack_c_kernel(const double* __restrict__ c, int64_t ldc, int K, int Kp,
                                                            double* __restrict__ bq) {
    const int64_t total = (int64_t)Kp * Kp;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(e & 7);
        const int64_t r = e >> 3;
        const int j = (int)(r % Kp), ib = (int)(r / Kp);
        const int i = ib * 8 + t;
        double v = 0.0;
        if (i < K && j < K && i <= j) v = (i == j) ? 0.5 * c[(size_t)i * ldc + j] : c[(size_t)i * ldc + j];
        bq[e] = v;
    }
}

// Bq[ib][m][t] = cf[m][ib * 8 + t]  (zero outside M x K)
__global__ void __launch_bounds__(256) pcq_lf_pack_cf_kernel(const double* __restrict__ cf, int K, int M, int Kp, int Mp,
                                                             double* __restrict__ bq) {
    const int64_t total = (int64_t)Kp * Mp;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(e & 7);
        const int64_t r = e >> 3;
        const int m = (int)(r % Mp), ib = (int)(r / Mp);
        const int i = ib * 8 + t;
        bq[e] = (m < M && i < K) ? cf[(size_t)m * K + i] : 0.0;
    }
}


// Caller-supplied design matrix A (n x K row-major, leading dimension lda, any alignment) -> Zq[i-block][sample (Nc)][8
// terms], zero outside n x K.  Thread = (sample, term pair): a warp reads 512 contiguous bytes of a row and writes
// eight 64-byte runs -- whole sectors both ways.
__global__ void __launch_bounds__(256) pcq_qf_pack_matrix_kernel(const double* __restrict__ a, int64_t lda, int64_t n, int K,
                                                                 int Kp, int64_t Nc, double* __restrict__ zq) {
    const int pairs = Kp / 2;
    const int64_t total = Nc * pairs;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / pairs;
        const int j = (int)(e - row * pairs);
        const int t0 = 2 * j;
        double v0 = 0.0, v1 = 0.0;
        if (row < n) {
            if (t0 < K) v0 = __ldcs(a + row * lda + t0);
            if (t0 + 1 < K) v1 = __ldcs(a + row * lda + t0 + 1);
        }
        const int ib = t0 >> 3, q = (t0 & 7) >> 1;
        *reinterpret_cast<double2*>(zq + ((size_t)ib * Nc + row) * 8 + 2 * q) = make_double2(v0, v1);
    }
}

// Y[i][j] = sum_k A[i][k] cf[j][k] for a FEW coefficient vectors (m <= 8 per pass): one warp per row, the row read
// once with coalesced loads (HBM bound), fixed summation order (lane-strided partial sums, then a shuffle tree).
template <int MB>
__global__ void __launch_bounds__(256) pcq_matvec_rows_kernel(const double* __restrict__ a, int64_t lda, int64_t n, int K,
                                                              const double* __restrict__ cf, int64_t ldcf, int m0, int m,
                                                              double* __restrict__ y, int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n; row += nwarps) {
        double acc[MB];
#pragma unroll
        for (int j = 0; j < MB; ++j) acc[j] = 0.0;
        for (int k = lane; k < K; k += 32) {
            const double v = __ldcs(a + row * lda + k);
#pragma unroll
            for (int j = 0; j < MB; ++j)
                if (m0 + j < m) acc[j] = fma(v, __ldg(cf + (size_t)(m0 + j) * ldcf + k), acc[j]);
        }
#pragma unroll
        for (int j = 0; j < MB; ++j) {
            double t = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0 && m0 + j < m) y[row * ldy + m0 + j] = t;
        }
    }
}

// LF = false: quadratic forms (triangular U, row-wise product epilogue).
// LF = true:  linear forms Y = Psi C^T for many coefficient vectors (evalPC fast mode with M >= 64 outputs): same
//             pipeline, every unit contracts all Kp terms and the epilogue stores the 128 x 128 tile of Y.
template <bool LF>
__global__ void __launch_bounds__(SY_THREADS, 1) pcq_qform_kernel(QformParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ST = 128, CHUNK = ST * SB, FM = 4, FN = 8;
    double* ops = reinterpret_cast<double*>(smem_raw);                        // [NSTAGE][2][KS][CHUNK]
    uint64_t* full = reinterpret_cast<uint64_t*>(ops + (size_t)NSTAGE * 2 * KS * CHUNK);
    uint64_t* empty = full + NSTAGE;
    __shared__ double rowsum[2][ST];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int b = 0; b < NSTAGE; ++b) { mbar_init(full + b, 1); mbar_init(empty + b, CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t it = 0;
    // unit u -> (sample tile mt, column tile J); stages of a unit: the 32-term blocks below column 128 (J + 1)
    int64_t pu = blockIdx.x;      // producer cursor
    int pst = 0, pst1 = 0, pmt = 0, pJ = 0;
    uint32_t pit = 0;
    auto open_unit = [&]() {
        if (pu < P.units) {
            pmt = (int)(pu / P.nJ); pJ = (int)(pu - (int64_t)pmt * P.nJ);
            pst = 0; pst1 = LF ? P.Kp / (KS * SB) : (pJ + 1) * (ST / (KS * SB));
        }
    };
    if (warp == 0) open_unit();
    auto try_issue = [&](bool blocking) {
        if (pu >= P.units || pit >= it + NSTAGE) return;
        const int b = pit % NSTAGE;
        const uint32_t ph = (pit / NSTAGE) & 1u;
        if (blocking) mbar_wait(empty + b, ph ^ 1u);
        else if (!mbar_try(empty + b, ph ^ 1u)) return;
        if (lane == 0) mbar_expect_tx(full + b, 2u * KS * CHUNK * 8u);
        __syncwarp();
        if (lane < 2 * KS) {
            const int op = lane / KS, kb = lane % KS;
            const size_t ib = (size_t)pst * KS + kb;
            const double* src = op ? P.bq + (ib * P.ldb + (size_t)pJ * ST) * SB
                                   : P.zq + (ib * P.Nc + (size_t)pmt * ST) * SB;
            tma_bulk_g2s(ops + ((size_t)(b * 2 + op) * KS + kb) * CHUNK, src, CHUNK * 8u, full + b);
        }
        ++pit;
        if (++pst >= pst1) { pu += gridDim.x; open_unit(); }
    };
    if (warp == 0) {
        try_issue(true);
        try_issue(true);
    }

    const int g = lane >> 2, tig = lane & 3, wm = warp & 3, wn = warp >> 2;
    for (int64_t u = blockIdx.x; u < P.units; u += gridDim.x) {
        const int mt = (int)(u / P.nJ), J = (int)(u - (int64_t)mt * P.nJ);
        const int nst = LF ? P.Kp / (KS * SB) : (J + 1) * (ST / (KS * SB));
        double acc[FM][FN][2];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        for (int st = 0; st < nst; ++st, ++it) {
            const int b = it % NSTAGE;
            const uint32_t ph = (it / NSTAGE) & 1u;
            mbar_wait(full + b, ph);
            const double* oa = ops + (size_t)(b * 2) * KS * CHUNK;
            const double* ob = oa + (size_t)KS * CHUNK;
            const double* pa = oa + (wm * 32 + g) * SB + 2 * tig;
            const double* pb = ob + (wn * 64 + g) * SB + 2 * tig;
#pragma unroll
            for (int kb = 0; kb < KS; ++kb) {
                double2 af[FM], bf[FN];
#pragma unroll
                for (int i = 0; i < FM; ++i) af[i] = *reinterpret_cast<const double2*>(pa + kb * CHUNK + i * 8 * SB);
#pragma unroll
                for (int j = 0; j < FN; ++j) bf[j] = *reinterpret_cast<const double2*>(pb + kb * CHUNK + j * 8 * SB);
#pragma unroll
                for (int i = 0; i < FM; ++i)
#pragma unroll
                    for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
#pragma unroll
                for (int i = 0; i < FM; ++i)
#pragma unroll
                    for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i].y, bf[j].y);
                if (warp == 0) try_issue(kb == KS - 1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + b);
        }
        if (LF) {
#pragma unroll
            for (int i = 0; i < FM; ++i) {
                const int64_t row = (int64_t)mt * ST + wm * 32 + i * 8 + g;
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    const int col = J * ST + wn * 64 + j * 8 + 2 * tig;
                    if (row < P.n) {
                        if (col < P.m) P.y[row * P.ldy + col] = acc[i][j][0];
                        if (col + 1 < P.m) P.y[row * P.ldy + col + 1] = acc[i][j][1];
                    }
                }
            }
            continue;
        }
        // epilogue: rowsum_s = sum_j acc[s][j] * Psi[s][j] over this tile's 128 columns
        double rs[FM];
#pragma unroll
        for (int i = 0; i < FM; ++i) {
            const size_t srow = (size_t)mt * ST + wm * 32 + i * 8 + g;
            double sacc = 0.0;
#pragma unroll
            for (int j = 0; j < FN; ++j) {
                const int col = J * ST + wn * 64 + j * 8 + 2 * tig;        // even: the pair sits in one 8-term block
                const double2 psi = *reinterpret_cast<const double2*>(P.zq + ((size_t)(col >> 3) * P.Nc + srow) * 8 + (col & 7));
                sacc = fma(acc[i][j][0], psi.x, sacc);
                sacc = fma(acc[i][j][1], psi.y, sacc);
            }
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            rs[i] = sacc;
        }
        __syncthreads();       // previous unit's readers of rowsum are done
        if (tig == 0) {
#pragma unroll
            for (int i = 0; i < FM; ++i) rowsum[wn][wm * 32 + i * 8 + g] = rs[i];
        }
        __syncthreads();
        if (tid < ST) P.part[(size_t)J * P.Nc + (size_t)mt * ST + tid] = rowsum[0][tid] + rowsum[1][tid];
    }
}

// v[n] = 2 * sum_J part[J][n]  (tile order)
__global__ void __launch_bounds__(256) pcq_qf_reduce_kernel(const double* __restrict__ part, int64_t Nc, int nJ, int64_t n,
                                                            double* __restrict__ v, int64_t ldv) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int J = 0; J < nJ; ++J) acc += part[(size_t)J * Nc + i];
        v[i * ldv] = 2.0 * acc;
    }
}

}  // namespace

// v[i * ldv] = psi(x_i)^T C psi(x_i), i < n.  C: K x K symmetric, row-major (only its upper triangle is read).
int pcq_launch_qform(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* c, int64_t ldc, double* v,
                     int64_t ldv, double** bq, size_t* bq_bytes, double** ws, size_t* ws_bytes, double** zbuf,
                     size_t* zbuf_bytes, cudaStream_t st) {
    if (n == 0) return PCQ_OK;
    const int K = p->nterms;
    const int Kp = (K + 127) / 128 * 128;
    const int nJ = Kp / 128;
    const size_t smem = syrk_smem_bytes(128);
    const size_t smem_pack = (size_t)(p->nslots + 1) * PACK_TS * sizeof(double);
    if (smem > p->smem_optin || smem_pack > p->smem_optin) {
        pcq_set_error("quadform: slot table or pipeline does not fit in shared memory");
        return PCQ_ERR_UNSUPPORTED;
    }
    int rc = pcq_ensure_dev(bq, bq_bytes, (size_t)Kp * Kp * sizeof(double));
    if (rc != PCQ_OK) return rc;
    pcq_qf_pack_c_kernel<<<p->num_sms * 8, 256, 0, st>>>(c, ldc, K, Kp, *bq);
    PCQ_LAUNCH_CHECK("pcq_qf_pack_c_kernel");

    const char* env = getenv("PCQ_SYRK_CHUNK_MB");
    const int64_t target = (int64_t)(env ? atoi(env) : 2048) << 20;
    int64_t rows = (target / ((int64_t)Kp * 8)) / 128 * 128;
    if (rows < 128) rows = 128;
    const int64_t npad = (n + 127) / 128 * 128;
    if (rows > npad) rows = npad;
    if ((rc = pcq_ensure_dev(zbuf, zbuf_bytes, (size_t)rows * Kp * sizeof(double))) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(ws, ws_bytes, (size_t)nJ * rows * sizeof(double))) != PCQ_OK) return rc;
    PCQ_CUDA(cudaFuncSetAttribute(pcq_qform_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = (n - r0 < rows) ? n - r0 : rows;
        const int64_t Nc = (nr + 127) / 128 * 128;
        QfPackParams K1;
        memset(&K1, 0, sizeof(K1));
        K1.x = x + r0 * ldx; K1.ldx = ldx; K1.n = nr; K1.Nc = Nc; K1.npairs = Nc / 16;
        K1.K = K; K1.Kp = Kp; K1.sdim = p->sdim; K1.pmax = p->pmax; K1.nslots = p->nslots; K1.width = p->width;
        K1.rec_a = p->d_rec_a; K1.rec_b = p->d_rec_b; K1.p1 = p->d_p1; K1.desc4 = p->d_desc4; K1.descw = p->d_descw;
        K1.zq = *zbuf;
        const int64_t cap = (int64_t)p->num_sms * 8;
        const int pgrid = (int)(K1.npairs < cap ? K1.npairs : cap);
        auto launch = [&](auto kern) -> int {
            PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pack));
            kern<<<pgrid, 256, smem_pack, st>>>(K1);
            return PCQ_OK;
        };
        switch (p->width <= 4 ? p->width : 0) {
            case 1: rc = launch(pcq_qf_pack_rows_kernel<1>); break;
            case 2: rc = launch(pcq_qf_pack_rows_kernel<2>); break;
            case 3: rc = launch(pcq_qf_pack_rows_kernel<3>); break;
            case 4: rc = launch(pcq_qf_pack_rows_kernel<4>); break;
            default: rc = launch(pcq_qf_pack_rows_kernel<0>); break;
        }
        if (rc != PCQ_OK) return rc;
        PCQ_LAUNCH_CHECK("pcq_qf_pack_rows_kernel");

        QformParams Q;
        memset(&Q, 0, sizeof(Q));
        Q.zq = *zbuf; Q.bq = *bq; Q.Nc = Nc; Q.Kp = Kp; Q.nJ = nJ; Q.ldb = Kp; Q.units = (Nc / 128) * nJ; Q.part = *ws;
        const int grid = (int)(Q.units < p->num_sms ? Q.units : p->num_sms);
        pcq_qform_kernel<false><<<grid, SY_THREADS, smem, st>>>(Q);
        PCQ_LAUNCH_CHECK("pcq_qform_kernel");
        const int64_t want = (nr + 255) / 256;
        pcq_qf_reduce_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(*ws, Nc, nJ, nr, v + r0 * ldv, ldv);
        PCQ_LAUNCH_CHECK("pcq_qf_reduce_kernel");
    }
    return PCQ_OK;
}

// Y[i][j] = sum_k cf[j][k] Psi_k(x_i)  for m coefficient vectors at once (evalPC fast mode, many outputs): the
// design-matrix chunk is packed once and contracted with all vectors on the tensor pipe.
bool pcq_linforms_applicable(const pcq_plan* p, int m) {
    return m >= 64 && syrk_smem_bytes(128) <= p->smem_optin &&
           (size_t)(p->nslots + 1) * PACK_TS * sizeof(double) <= p->smem_optin && pcq_table_fits_smem(p);
}

int pcq_launch_linforms(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf, int m, double* y,
                        int64_t ldy, double** bq, size_t* bq_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st) {
    if (n == 0 || m == 0) return PCQ_OK;
    const int K = p->nterms;
    const int Kp = (K + 127) / 128 * 128, Mp = (m + 127) / 128 * 128;
    const int nJ = Mp / 128;
    const size_t smem = syrk_smem_bytes(128);
    const size_t smem_pack = (size_t)(p->nslots + 1) * PACK_TS * sizeof(double);
    int rc = pcq_ensure_dev(bq, bq_bytes, (size_t)Kp * Mp * sizeof(double));
    if (rc != PCQ_OK) return rc;
    pcq_lf_pack_cf_kernel<<<p->num_sms * 4, 256, 0, st>>>(cf, K, m, Kp, Mp, *bq);
    PCQ_LAUNCH_CHECK("pcq_lf_pack_cf_kernel");
    const char* env = getenv("PCQ_SYRK_CHUNK_MB");
    const int64_t target = (int64_t)(env ? atoi(env) : 2048) << 20;
    int64_t rows = (target / ((int64_t)Kp * 8)) / 128 * 128;
    if (rows < 128) rows = 128;
    const int64_t npad = (n + 127) / 128 * 128;
    if (rows > npad) rows = npad;
    if ((rc = pcq_ensure_dev(zbuf, zbuf_bytes, (size_t)rows * Kp * sizeof(double))) != PCQ_OK) return rc;
    PCQ_CUDA(cudaFuncSetAttribute(pcq_qform_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = (n - r0 < rows) ? n - r0 : rows;
        const int64_t Nc = (nr + 127) / 128 * 128;
        QfPackParams K1;
        memset(&K1, 0, sizeof(K1));
        K1.x = x + r0 * ldx; K1.ldx = ldx; K1.n = nr; K1.Nc = Nc; K1.npairs = Nc / 16;
        K1.K = K; K1.Kp = Kp; K1.sdim = p->sdim; K1.pmax = p->pmax; K1.nslots = p->nslots; K1.width = p->width;
        K1.rec_a = p->d_rec_a; K1.rec_b = p->d_rec_b; K1.p1 = p->d_p1; K1.desc4 = p->d_desc4; K1.descw = p->d_descw;
        K1.zq = *zbuf;
        const int64_t cap = (int64_t)p->num_sms * 8;
        const int pgrid = (int)(K1.npairs < cap ? K1.npairs : cap);
        auto launch = [&](auto kern) -> int {
            PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pack));
            kern<<<pgrid, 256, smem_pack, st>>>(K1);
            return PCQ_OK;
        };
        switch (p->width <= 4 ? p->width : 0) {
            case 1: rc = launch(pcq_qf_pack_rows_kernel<1>); break;
            case 2: rc = launch(pcq_qf_pack_rows_kernel<2>); break;
            case 3: rc = launch(pcq_qf_pack_rows_kernel<3>); break;
            case 4: rc = launch(pcq_qf_pack_rows_kernel<4>); break;
            default: rc = launch(pcq_qf_pack_rows_kernel<0>); break;
        }
        if (rc != PCQ_OK) return rc;
        PCQ_LAUNCH_CHECK("pcq_qf_pack_rows_kernel");
        QformParams Q;
        memset(&Q, 0, sizeof(Q));
        Q.zq = *zbuf; Q.bq = *bq; Q.Nc = Nc; Q.Kp = Kp; Q.nJ = nJ; Q.ldb = Mp; Q.units = (Nc / 128) * nJ;
        Q.y = y + r0 * ldy; Q.ldy = ldy; Q.n = nr; Q.m = m;
        const int grid = (int)(Q.units < p->num_sms ? Q.units : p->num_sms);
        pcq_qform_kernel<true><<<grid, SY_THREADS, smem, st>>>(Q);
        PCQ_LAUNCH_CHECK("pcq_qform_kernel");
    }
    return PCQ_OK;
}

// ---- the same two contractions for a design matrix handed over by the caller (lreg.predicta / compute_stdev /
// predicta_samples on Amat, lreg/lreg.py:112-198; apps/uqpc/uq_pc.py:313-315): no plan, A row-major in HBM.

// v[i * ldv] = a_i^T C a_i
int pcq_launch_qform_matrix(int num_sms, size_t smem_optin, const double* a, int64_t n, int64_t lda, int K, const double* c,
                            int64_t ldc, double* v, int64_t ldv, double** bq, size_t* bq_bytes, double** ws,
                            size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st) {
    if (n == 0) return PCQ_OK;
    const int Kp = (K + 127) / 128 * 128;
    const int nJ = Kp / 128;
    const size_t smem = syrk_smem_bytes(128);
    if (smem > smem_optin) { pcq_set_error("quadform_matrix: pipeline does not fit in shared memory"); return PCQ_ERR_UNSUPPORTED; }
    int rc = pcq_ensure_dev(bq, bq_bytes, (size_t)Kp * Kp * sizeof(double));
    if (rc != PCQ_OK) return rc;
    pcq_qf_pack_c_kernel<<<num_sms * 8, 256, 0, st>>>(c, ldc, K, Kp, *bq);
    PCQ_LAUNCH_CHECK("pcq_qf_pack_c_kernel");
    const char* env = getenv("PCQ_SYRK_CHUNK_MB");
    const int64_t target = (int64_t)(env ? atoi(env) : 2048) << 20;
    int64_t rows = (target / ((int64_t)Kp * 8)) / 128 * 128;
    if (rows < 128) rows = 128;
    const int64_t npad = (n + 127) / 128 * 128;
    if (rows > npad) rows = npad;
    if ((rc = pcq_ensure_dev(zbuf, zbuf_bytes, (size_t)rows * Kp * sizeof(double))) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(ws, ws_bytes, (size_t)nJ * rows * sizeof(double))) != PCQ_OK) return rc;
    PCQ_CUDA(cudaFuncSetAttribute(pcq_qform_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = (int64_t)num_sms * 8;
    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = (n - r0 < rows) ? n - r0 : rows;
        const int64_t Nc = (nr + 127) / 128 * 128;
        const int64_t wantp = (Nc * (Kp / 2) + 255) / 256;
        pcq_qf_pack_matrix_kernel<<<(int)(wantp < cap * 4 ? wantp : cap * 4), 256, 0, st>>>(a + r0 * lda, lda, nr, K, Kp, Nc, *zbuf);
        PCQ_LAUNCH_CHECK("pcq_qf_pack_matrix_kernel");
        QformParams Q;
        memset(&Q, 0, sizeof(Q));
        Q.zq = *zbuf; Q.bq = *bq; Q.Nc = Nc; Q.Kp = Kp; Q.nJ = nJ; Q.ldb = Kp; Q.units = (Nc / 128) * nJ; Q.part = *ws;
        const int grid = (int)(Q.units < num_sms ? Q.units : num_sms);
        pcq_qform_kernel<false><<<grid, SY_THREADS, smem, st>>>(Q);
        PCQ_LAUNCH_CHECK("pcq_qform_kernel");
        const int64_t want = (nr + 255) / 256;
        pcq_qf_reduce_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(*ws, Nc, nJ, nr, v + r0 * ldv, ldv);
        PCQ_LAUNCH_CHECK("pcq_qf_reduce_kernel");
    }
    return PCQ_OK;
}

// Y (n x m, ldy) = A (n x K, lda) CF^T, CF m x K with leading dimension ldcf
int pcq_launch_linforms_matrix(int num_sms, size_t smem_optin, const double* a, int64_t n, int64_t lda, int K,
                               const double* cf, int64_t ldcf, int m, double* y, int64_t ldy, double** bq, size_t* bq_bytes,
                               double** zbuf, size_t* zbuf_bytes, cudaStream_t st) {
    if (n == 0 || m == 0) return PCQ_OK;
    const int64_t cap = (int64_t)num_sms * 8;
    const size_t smem = syrk_smem_bytes(128);
    if (m < 48 || smem > smem_optin || ldcf != K) {
        // few vectors (the mean of predicta: m = 1): HBM-bound row dot products, eight vectors per pass over A
        const int64_t want = (n * 32 + 255) / 256;
        const int grid = (int)(want < cap ? want : cap);
        for (int m0 = 0; m0 < m; m0 += 8) {
            const int mb = m - m0;
            if (mb == 1) pcq_matvec_rows_kernel<1><<<grid, 256, 0, st>>>(a, lda, n, K, cf, ldcf, m0, m, y, ldy);
            else if (mb <= 4) pcq_matvec_rows_kernel<4><<<grid, 256, 0, st>>>(a, lda, n, K, cf, ldcf, m0, m, y, ldy);
            else pcq_matvec_rows_kernel<8><<<grid, 256, 0, st>>>(a, lda, n, K, cf, ldcf, m0, m, y, ldy);
            PCQ_LAUNCH_CHECK("pcq_matvec_rows_kernel");
        }
        return PCQ_OK;
    }
    const int Kp = (K + 127) / 128 * 128, Mp = (m + 127) / 128 * 128;
    const int nJ = Mp / 128;
    int rc = pcq_ensure_dev(bq, bq_bytes, (size_t)Kp * Mp * sizeof(double));
    if (rc != PCQ_OK) return rc;
    pcq_lf_pack_cf_kernel<<<num_sms * 4, 256, 0, st>>>(cf, K, m, Kp, Mp, *bq);
    PCQ_LAUNCH_CHECK("pcq_lf_pack_cf_kernel");
    const char* env = getenv("PCQ_SYRK_CHUNK_MB");
    const int64_t target = (int64_t)(env ? atoi(env) : 2048) << 20;
    int64_t rows = (target / ((int64_t)Kp * 8)) / 128 * 128;
    if (rows < 128) rows = 128;
    const int64_t npad = (n + 127) / 128 * 128;
    if (rows > npad) rows = npad;
    if ((rc = pcq_ensure_dev(zbuf, zbuf_bytes, (size_t)rows * Kp * sizeof(double))) != PCQ_OK) return rc;
    PCQ_CUDA(cudaFuncSetAttribute(pcq_qform_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = (n - r0 < rows) ? n - r0 : rows;
        const int64_t Nc = (nr + 127) / 128 * 128;
        const int64_t wantp = (Nc * (Kp / 2) + 255) / 256;
        pcq_qf_pack_matrix_kernel<<<(int)(wantp < cap * 4 ? wantp : cap * 4), 256, 0, st>>>(a + r0 * lda, lda, nr, K, Kp, Nc, *zbuf);
        PCQ_LAUNCH_CHECK("pcq_qf_pack_matrix_kernel");
        QformParams Q;
        memset(&Q, 0, sizeof(Q));
        Q.zq = *zbuf; Q.bq = *bq; Q.Nc = Nc; Q.Kp = Kp; Q.nJ = nJ; Q.ldb = Mp; Q.units = (Nc / 128) * nJ;
        Q.y = y + r0 * ldy; Q.ldy = ldy; Q.n = nr; Q.m = m;
        const int grid = (int)(Q.units < num_sms ? Q.units : num_sms);
        pcq_qform_kernel<true><<<grid, SY_THREADS, smem, st>>>(Q);
        PCQ_LAUNCH_CHECK("pcq_qform_kernel");
    }
    return PCQ_OK;
}
