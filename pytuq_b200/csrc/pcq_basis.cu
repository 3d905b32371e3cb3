// K1 (design matrix) and K2 (surrogate evaluation without A) for sm_100a.
//
// K1  pcq_bases_kernel   A[n,k] = prod_i phi_{mi[k,i]}(x[n,i])      PCRV.evalBases rv/pcrv.py:413-442
//     HBM-write bound: 8*N*(s+K) algorithmic bytes.  A CTA owns blocks of R consecutive
//     samples; the 1-D recurrences run once per (sample, dim) into a shared-memory slot
//     table, then each warp sweeps one row of A with lanes along k (the packed term
//     descriptors are staged in shared memory) and writes 16-byte streaming stores.
// K2  pcq_evalpc_kernel  y[n,m] = sum_k cf[m,k] prod_i phi(...)      PCRV.evalPC   rv/pcrv.py:469-498
//     FP64 bound.  Thread per sample, slot table [slot][lane] in shared memory (conflict
//     free), descriptors and coefficients are warp-uniform loads.
#include "pcq_internal.cuh"
#include <cstring>
#include <cmath>

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_ROWS = 16;   // samples per CTA iteration (2 per warp)

struct BasesParams {
    const double* x;
    double* a;
    int64_t n, ldx, lda;
    int sdim, nterms, pmax, nslots, width;
    const double* rec_a;
    const double* rec_b;
    const double* p1;
    const unsigned long long* desc4;
    const uint16_t* descw;
};

template <int W, bool DESC_SMEM>
__global__ void __launch_bounds__(K1_THREADS) pcq_bases_kernel(BasesParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [desc4: K x 8B (if DESC_SMEM)] [tab: K1_ROWS x nslots doubles]
    unsigned long long* sdesc = reinterpret_cast<unsigned long long*>(smem_raw);
    double* tab = reinterpret_cast<double*>(smem_raw + (DESC_SMEM ? (size_t)P.nterms * 8 : 0));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int NWARPS = K1_THREADS / 32;
    const int K = P.nterms;
    const int s = P.sdim;

    if (DESC_SMEM) {
        for (int k = tid; k < K; k += K1_THREADS) sdesc[k] = P.desc4[k];
    }
    const unsigned long long* desc = DESC_SMEM ? sdesc : P.desc4;

    const int64_t nblocks = (P.n + K1_ROWS - 1) / K1_ROWS;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t n0 = blk * K1_ROWS;
        const int rows = (int)min((int64_t)K1_ROWS, P.n - n0);
        __syncthreads();   // previous iteration's readers are done with tab (and desc is staged)
        // phase 1: 1-D recurrences, one (sample, dim) pair per thread
        for (int idx = tid; idx < rows * s; idx += K1_THREADS) {
            const int r = idx / s;
            const int i = idx - r * s;
            double* t = tab + (size_t)r * P.nslots;
            if (i == 0) t[0] = 1.0;
            const double xv = P.x[(n0 + r) * P.ldx + i];
            pcq_fill_1d(xv, i, s, P.pmax, P.rec_a, P.rec_b, P.p1, t, 1);
        }
        __syncthreads();
        // phase 2: one warp per row, lanes along k
        for (int r = warp; r < rows; r += NWARPS) {
            const double* t = tab + (size_t)r * P.nslots;
            double* out = P.a + (n0 + r) * P.lda;
            if (W > 0) {
                // 16-byte stores need an even element offset from a 16-byte aligned base
                const int head = (int)((reinterpret_cast<uintptr_t>(out) >> 3) & 1);
                if (head && lane == 0) __stcs(out, pcq_term4<W>(t, 1, desc[0]));
                const int npairs = (K - head) >> 1;
                for (int q = lane; q < npairs; q += 32) {
                    const int k = head + 2 * q;
                    double2 v;
                    v.x = pcq_term4<W>(t, 1, desc[k]);
                    v.y = pcq_term4<W>(t, 1, desc[k + 1]);
                    __stcs(reinterpret_cast<double2*>(out + k), v);
                }
                if (((K - head) & 1) && lane == 31) {
                    __stcs(out + (K - 1), pcq_term4<W>(t, 1, desc[K - 1]));
                }
            } else {
                for (int k = lane; k < K; k += 32) {
                    __stcs(out + k, pcq_termw(t, 1, P.descw + (size_t)k * P.width, P.width));
                }
            }
        }
    }
}

template <int W>
int launch_bases_w(pcq_plan* p, const BasesParams& P, cudaStream_t st) {
    const size_t tab_bytes = (size_t)K1_ROWS * p->nslots * sizeof(double);
    const size_t desc_bytes = (size_t)p->nterms * 8;
    const int64_t nblocks = (P.n + K1_ROWS - 1) / K1_ROWS;
    const bool desc_smem = (W > 0) && (desc_bytes + tab_bytes <= 96 * 1024);
    const size_t smem = tab_bytes + (desc_smem ? desc_bytes : 0);
    if (smem > p->smem_optin) {
        pcq_set_error("evalBases: slot table of %zu bytes exceeds shared memory", smem);
        return PCQ_ERR_UNSUPPORTED;
    }
    auto kern = desc_smem ? pcq_bases_kernel<W, true> : pcq_bases_kernel<W, false>;
    PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PCQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K1_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t cap = (int64_t)p->num_sms * occ;
    const int grid = (int)(nblocks < cap ? nblocks : cap);
    kern<<<grid, K1_THREADS, smem, st>>>(P);
    PCQ_LAUNCH_CHECK("pcq_bases_kernel");
    return PCQ_OK;
}

// ---------------------------------------------------------------------------- K2 / K5
struct EvalParams {
    const double* x;
    const double* cf;   // M x K
    double* y;          // may be null in germ mode
    int64_t n, ldx, ldy;
    int sdim, nterms, pmax, nslots, width, m;
    const double* rec_a;
    const double* rec_b;
    const double* p1;
    const unsigned long long* desc4;
    const uint16_t* descw;
    // germ mode (K5)
    unsigned long long seed;
    int64_t first_index;
    const int* germ_kind;   // device, s entries: 0 uniform(-1,1), 1 normal(0,1)
    double* part;           // [grid*warps][m][4] partial (sum, sumsq, min, max)
};

// Philox4x32-10 (Salmon et al., SC'11), counter-based: the draw for (sample, dim pair) is a
// pure function of (seed, global sample index, pair index), so shards of one logical
// stream can be generated independently on different GPUs.
__device__ __forceinline__ void philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2,
                                              unsigned int c3, unsigned int k0, unsigned int k1,
                                              unsigned int (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned int n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two germ values for dimensions (2*pair, 2*pair+1) of global sample `gidx`
__device__ __forceinline__ void pcq_germ_pair(unsigned long long seed, unsigned long long gidx,
                                              unsigned int pair, int kind0, int kind1,
                                              double& z0, double& z1) {
    unsigned int r[4];
    philox4x32_10((unsigned int)gidx, (unsigned int)(gidx >> 32), pair, 0x50435121u,
                  (unsigned int)seed, (unsigned int)(seed >> 32), r);
    const unsigned long long a = ((unsigned long long)r[1] << 32) | r[0];
    const unsigned long long b = ((unsigned long long)r[3] << 32) | r[2];
    const double u0 = (double)(a >> 11) * 0x1.0p-53;   // [0,1)
    const double u1 = (double)(b >> 11) * 0x1.0p-53;
    // Box-Muller on (1-u0) in (0,1] and u1
    const double rad = sqrt(-2.0 * log(1.0 - u0));
    double sn, cs;
    sincospi(2.0 * u1, &sn, &cs);
    z0 = kind0 ? rad * cs : 2.0 * u0 - 1.0;
    z1 = kind1 ? rad * sn : 2.0 * u1 - 1.0;
}

// W = 0: generic width from descw.  MB outputs per pass.  FAST: shared basis product + FMA.
// GERM: x comes from the Philox stream and per-output power sums are reduced (K5).
template <int W, int MB, bool FAST, int THREADS, bool GERM>
__global__ void __launch_bounds__(THREADS) pcq_evalpc_kernel(EvalParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);   // [nslots][THREADS]
    const int tid = threadIdx.x;
    const int s = P.sdim;
    const int K = P.nterms;
    double* mytab = tab + tid;
    constexpr int WW = W > 0 ? W : 1;
    double* mypart = nullptr;
    if (GERM) mypart = P.part + ((size_t)blockIdx.x * (THREADS / 32) + (tid >> 5)) * P.m * 4;
    bool first_iter = true;

    const int64_t nblocks = (P.n + THREADS - 1) / THREADS;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t n = blk * THREADS + tid;
        const bool live = n < P.n;
        // each thread only ever touches its own column of the table: no barrier needed
        mytab[0] = 1.0;
        if (GERM) {
            for (int i = 0; i < s; i += 2) {
                double z0, z1;
                const int k0 = P.germ_kind[i];
                const int k1 = (i + 1 < s) ? P.germ_kind[i + 1] : 0;
                pcq_germ_pair(P.seed, (unsigned long long)(P.first_index + n), (unsigned int)(i >> 1),
                              k0, k1, z0, z1);
                pcq_fill_1d(z0, i, s, P.pmax, P.rec_a, P.rec_b, P.p1, mytab, THREADS);
                if (i + 1 < s)
                    pcq_fill_1d(z1, i + 1, s, P.pmax, P.rec_a, P.rec_b, P.p1, mytab, THREADS);
            }
        } else {
            for (int i = 0; i < s; ++i) {
                const double xv = live ? P.x[n * P.ldx + i] : 0.0;
                pcq_fill_1d(xv, i, s, P.pmax, P.rec_a, P.rec_b, P.p1, mytab, THREADS);
            }
        }
        for (int m0 = 0; m0 < P.m; m0 += MB) {
            double acc[MB];
#pragma unroll
            for (int j = 0; j < MB; ++j) acc[j] = 0.0;
            const double* cfp[MB];
#pragma unroll
            for (int j = 0; j < MB; ++j) cfp[j] = P.cf + (size_t)min(m0 + j, P.m - 1) * K;
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                if constexpr (W > 0) {
                    const unsigned long long d = __ldg(P.desc4 + k);
                    double f[WW];
                    f[0] = mytab[(size_t)(d & 0xffffull) * THREADS];
                    if constexpr (W >= 2) f[1] = mytab[(size_t)((d >> 16) & 0xffffull) * THREADS];
                    if constexpr (W >= 3) f[2] = mytab[(size_t)((d >> 32) & 0xffffull) * THREADS];
                    if constexpr (W >= 4) f[3] = mytab[(size_t)((d >> 48) & 0xffffull) * THREADS];
                    if constexpr (FAST) {
                        double psi = f[0];
#pragma unroll
                        for (int q = 1; q < W; ++q) psi = __dmul_rn(psi, f[q]);
#pragma unroll
                        for (int j = 0; j < MB; ++j) acc[j] = __fma_rn(__ldg(cfp[j] + k), psi, acc[j]);
                    } else {
                        // rv/pcrv.py:492-495: prd = cf[k]; prd *= phi_i (i ascending); val += prd
#pragma unroll
                        for (int j = 0; j < MB; ++j) {
                            double prd = __ldg(cfp[j] + k);
#pragma unroll
                            for (int q = 0; q < W; ++q) prd = __dmul_rn(prd, f[q]);
                            acc[j] = __dadd_rn(acc[j], prd);
                        }
                    }
                } else {
                    const uint16_t* dw = P.descw + (size_t)k * P.width;
                    if constexpr (FAST) {
                        const double psi = pcq_termw(mytab, THREADS, dw, P.width);
#pragma unroll
                        for (int j = 0; j < MB; ++j) acc[j] = __fma_rn(__ldg(cfp[j] + k), psi, acc[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < MB; ++j) {
                            double prd = __ldg(cfp[j] + k);
                            for (int q = 0; q < P.width; ++q)
                                prd = __dmul_rn(prd, mytab[(size_t)__ldg(dw + q) * THREADS]);
                            acc[j] = __dadd_rn(acc[j], prd);
                        }
                    }
                }
            }
            if (live && P.y != nullptr) {
#pragma unroll
                for (int j = 0; j < MB; ++j)
                    if (m0 + j < P.m) P.y[n * P.ldy + m0 + j] = acc[j];
            }
            if constexpr (GERM) {
                // warp-shuffle reduction of (sum, sum of squares, min, max); lane 0 keeps the
                // warp's running partial in its private workspace row (no atomics)
#pragma unroll
                for (int j = 0; j < MB; ++j) {
                    double v1 = live ? acc[j] : 0.0;
                    double v2 = live ? acc[j] * acc[j] : 0.0;
                    double mn = live ? acc[j] : INFINITY;
                    double mx = live ? acc[j] : -INFINITY;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    }
                    if ((tid & 31) == 0 && m0 + j < P.m) {
                        double* q = mypart + (size_t)(m0 + j) * 4;
                        if (first_iter) { q[0] = v1; q[1] = v2; q[2] = mn; q[3] = mx; }
                        else { q[0] += v1; q[1] += v2; q[2] = fmin(q[2], mn); q[3] = fmax(q[3], mx); }
                    }
                }
            }
        }
        first_iter = false;
    }
    if constexpr (GERM) {
        // CTAs (or warps) that never ran an iteration still own a workspace row
        if (first_iter && (tid & 31) == 0) {
            for (int j = 0; j < P.m; ++j) {
                double* q = mypart + (size_t)j * 4;
                q[0] = 0.0; q[1] = 0.0; q[2] = INFINITY; q[3] = -INFINITY;
            }
        }
    }
}

// fixed-order reduction of the per-warp partials -> sums[m][4]
__global__ void pcq_prop_reduce_kernel(const double* part, int nrows, int m, double* sums) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double s1 = 0.0, s2 = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int r = 0; r < nrows; ++r) {
        const double* q = part + ((size_t)r * m + j) * 4;
        s1 += q[0]; s2 += q[1]; mn = fmin(mn, q[2]); mx = fmax(mx, q[3]);
    }
    sums[j * 4 + 0] = s1; sums[j * 4 + 1] = s2; sums[j * 4 + 2] = mn; sums[j * 4 + 3] = mx;
}

__global__ void pcq_germ_kernel(unsigned long long seed, int64_t first, int64_t n, int s,
                                const int* kind, double* x, int64_t ldx) {
    const int npair = (s + 1) / 2;
    const int64_t total = n * npair;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / npair;
        const int pr = (int)(idx - r * npair);
        const int i = 2 * pr;
        double z0, z1;
        pcq_germ_pair(seed, (unsigned long long)(first + r), (unsigned int)pr, kind[i],
                      (i + 1 < s) ? kind[i + 1] : 0, z0, z1);
        x[r * ldx + i] = z0;
        if (i + 1 < s) x[r * ldx + i + 1] = z1;
    }
}

template <int W, int MB, bool FAST, int THREADS, bool GERM>
int launch_evalpc_t(pcq_plan* p, EvalParams P, cudaStream_t st, double* sums) {
    const size_t smem = (size_t)p->nslots * THREADS * sizeof(double);
    auto kern = pcq_evalpc_kernel<W, MB, FAST, THREADS, GERM>;
    PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PCQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t nblocks = (P.n + THREADS - 1) / THREADS;
    const int64_t cap = (int64_t)p->num_sms * occ;
    const int grid = (int)(nblocks < cap ? nblocks : cap);
    if (GERM) {
        const int nrows = grid * (THREADS / 32);
        int rc = pcq_ensure_dev(&p->d_ws, &p->ws_bytes, (size_t)nrows * P.m * 4 * sizeof(double));
        if (rc != PCQ_OK) return rc;
        P.part = p->d_ws;
        kern<<<grid, THREADS, smem, st>>>(P);
        PCQ_LAUNCH_CHECK("pcq_evalpc_kernel<germ>");
        pcq_prop_reduce_kernel<<<(P.m + 127) / 128, 128, 0, st>>>(P.part, nrows, P.m, sums);
        PCQ_LAUNCH_CHECK("pcq_prop_reduce_kernel");
        return PCQ_OK;
    }
    kern<<<grid, THREADS, smem, st>>>(P);
    PCQ_LAUNCH_CHECK("pcq_evalpc_kernel");
    return PCQ_OK;
}

template <int W, int MB, bool FAST, bool GERM>
int launch_evalpc_threads(pcq_plan* p, const EvalParams& P, cudaStream_t st, double* sums) {
    // the [slot][lane] table costs nslots*8 bytes per thread; pick the widest CTA that
    // still leaves room for two CTAs per SM
    const size_t per_thread = (size_t)p->nslots * sizeof(double);
    const size_t budget = p->smem_optin;
    if (per_thread * 128 * 2 <= budget) return launch_evalpc_t<W, MB, FAST, 128, GERM>(p, P, st, sums);
    if (per_thread * 64 <= budget) return launch_evalpc_t<W, MB, FAST, 64, GERM>(p, P, st, sums);
    if (per_thread * 32 <= budget) return launch_evalpc_t<W, MB, FAST, 32, GERM>(p, P, st, sums);
    pcq_set_error("evalPC: %d table slots per sample do not fit shared memory", p->nslots);
    return PCQ_ERR_UNSUPPORTED;
}

template <int W, bool GERM>
int launch_evalpc_w(pcq_plan* p, const EvalParams& P, int mode, cudaStream_t st, double* sums) {
    const bool fast = mode != 0;
    if (P.m >= 4) {
        return fast ? launch_evalpc_threads<W, 4, true, GERM>(p, P, st, sums)
                    : launch_evalpc_threads<W, 4, false, GERM>(p, P, st, sums);
    }
    return fast ? launch_evalpc_threads<W, 1, true, GERM>(p, P, st, sums)
                : launch_evalpc_threads<W, 1, false, GERM>(p, P, st, sums);
}

template <bool GERM>
int launch_evalpc_any(pcq_plan* p, const EvalParams& P, int mode, cudaStream_t st, double* sums) {
    switch (p->width <= PCQ_MAXW_FAST ? p->width : 0) {
        case 1: return launch_evalpc_w<1, GERM>(p, P, mode, st, sums);
        case 2: return launch_evalpc_w<2, GERM>(p, P, mode, st, sums);
        case 3: return launch_evalpc_w<3, GERM>(p, P, mode, st, sums);
        case 4: return launch_evalpc_w<4, GERM>(p, P, mode, st, sums);
        default: return launch_evalpc_w<0, GERM>(p, P, mode, st, sums);
    }
}

void fill_eval_params(pcq_plan* p, EvalParams& P) {
    memset(&P, 0, sizeof(P));
    P.sdim = p->sdim; P.nterms = p->nterms; P.pmax = p->pmax; P.nslots = p->nslots;
    P.width = p->width;
    P.rec_a = p->d_rec_a; P.rec_b = p->d_rec_b; P.p1 = p->d_p1;
    P.desc4 = p->d_desc4; P.descw = p->d_descw;
}

}  // namespace

int pcq_launch_bases(pcq_plan* p, const double* x, int64_t n, int64_t ldx, double* a,
                     int64_t lda, cudaStream_t st) {
    if (n == 0) return PCQ_OK;
    BasesParams P;
    P.x = x; P.a = a; P.n = n; P.ldx = ldx; P.lda = lda;
    P.sdim = p->sdim; P.nterms = p->nterms; P.pmax = p->pmax; P.nslots = p->nslots;
    P.width = p->width;
    P.rec_a = p->d_rec_a; P.rec_b = p->d_rec_b; P.p1 = p->d_p1;
    P.desc4 = p->d_desc4; P.descw = p->d_descw;
    switch (p->width <= PCQ_MAXW_FAST ? p->width : 0) {
        case 1: return launch_bases_w<1>(p, P, st);
        case 2: return launch_bases_w<2>(p, P, st);
        case 3: return launch_bases_w<3>(p, P, st);
        case 4: return launch_bases_w<4>(p, P, st);
        default: return launch_bases_w<0>(p, P, st);
    }
}

int pcq_launch_evalpc(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf,
                      int m, double* y, int64_t ldy, int mode, cudaStream_t st) {
    if (n == 0 || m == 0) return PCQ_OK;
    EvalParams P;
    fill_eval_params(p, P);
    P.x = x; P.cf = cf; P.y = y; P.n = n; P.ldx = ldx; P.ldy = ldy; P.m = m;
    return launch_evalpc_any<false>(p, P, mode, st, nullptr);
}

static int upload_kinds(pcq_plan* p, const int32_t* germ_kind, const int** dev, cudaStream_t st) {
    PCQ_CUDA(cudaMemcpyAsync(p->d_kinds, germ_kind, (size_t)p->sdim * sizeof(int),
                             cudaMemcpyHostToDevice, st));
    *dev = p->d_kinds;
    return PCQ_OK;
}

int pcq_launch_propagate(pcq_plan* p, uint64_t seed, int64_t first, int64_t n,
                         const int32_t* germ_kind, const double* cf, int m, double* sums,
                         double* y, int64_t ldy, int mode, cudaStream_t st) {
    if (m == 0) return PCQ_OK;
    EvalParams P;
    fill_eval_params(p, P);
    const int* dkind = nullptr;
    int rc = upload_kinds(p, germ_kind, &dkind, st);
    if (rc != PCQ_OK) return rc;
    P.cf = cf; P.y = y; P.n = n; P.ldy = ldy; P.m = m;
    P.seed = seed; P.first_index = first; P.germ_kind = dkind;
    return launch_evalpc_any<true>(p, P, mode, st, sums);
}

int pcq_launch_germ(pcq_plan* p, uint64_t seed, int64_t first, int64_t n,
                    const int32_t* germ_kind, double* x, int64_t ldx, cudaStream_t st) {
    if (n == 0) return PCQ_OK;
    const int* dkind = nullptr;
    int rc = upload_kinds(p, germ_kind, &dkind, st);
    if (rc != PCQ_OK) return rc;
    const int64_t total = n * ((p->sdim + 1) / 2);
    const int64_t want = (total + 255) / 256;
    const int64_t cap = (int64_t)p->num_sms * 8;
    pcq_germ_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(seed, first, n, p->sdim, dkind, x, ldx);
    PCQ_LAUNCH_CHECK("pcq_germ_kernel");
    return PCQ_OK;
}
