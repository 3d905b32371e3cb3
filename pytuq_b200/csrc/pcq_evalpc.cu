// K2 (surrogate evaluation without A) and K5 (propagation with on-device germ) for sm_100a.
//
//   y[n,m] = sum_k cf[m,k] prod_i phi_{mi[k,i]}(x[n,i])          PCRV.evalPC  rv/pcrv.py:469-498
//
// FP64 bound: F_min = 4 s (p-1) + 3 K M flop per sample.  Thread per sample (SPT samples per
// thread), per-sample slot table [slot][lane] in shared memory (conflict free).  The term
// descriptors and coefficients are staged chunk-wise in shared memory as pre-multiplied byte
// offsets, so one broadcast LDS.128 + LDS.64 per term serves SPT*32 samples.
//   mode 0 (exact): prd = cf[k]; prd *= phi (dims ascending); val += prd   -- bit-identical
//   mode 1 (fast) : psi = prod phi; acc[m] = fma(cf[m][k], psi, acc[m])
#include "pcq_internal.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cmath>

namespace {

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
    const uint8_t* flags;
    int sumsq;              // 1: y[n] = sum_m (output m)^2 instead of the M outputs (quadratic forms)
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
        double vsq1 = 0.0;
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
            if (P.sumsq) {
#pragma unroll
                for (int j = 0; j < MB; ++j)
                    if (m0 + j < P.m) vsq1 = __fma_rn(acc[j], acc[j], vsq1);
            } else if (live && P.y != nullptr) {
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
        if (P.sumsq && live && P.y != nullptr) P.y[n * P.ldy] = vsq1;
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


// ------------------------------------------------------------------------------------------------
// K2 v2: staged descriptors, SPT samples per thread.
constexpr int K2_TC = 512;     // terms per staged chunk

template <int W, int MB, bool FAST, int THREADS, int SPT, bool GERM>
__global__ void __launch_bounds__(THREADS) pcq_evalpc2_kernel(EvalParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int COLS = THREADS * SPT;                 // samples per CTA iteration
    // layout: tab[nslots][COLS] | sdesc[K2_TC] (uint4: W byte offsets) | scf[MB][K2_TC]
    double* tab = reinterpret_cast<double*>(smem_raw);
    uint4* sdesc = reinterpret_cast<uint4*>(tab + (size_t)P.nslots * COLS);
    double* scf = reinterpret_cast<double*>(sdesc + K2_TC);
    const int tid = threadIdx.x;
    const int s = P.sdim;
    const int K = P.nterms;
    double* mytab = tab + tid;
    const char* mybase = reinterpret_cast<const char*>(mytab);
    double* mypart = nullptr;
    if (GERM) mypart = P.part + ((size_t)blockIdx.x * (THREADS / 32) + (tid >> 5)) * P.m * 4;
    bool first_iter = true;

    const int64_t nblocks = (P.n + COLS - 1) / COLS;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        bool live[SPT];
        int64_t nidx[SPT];
        // each thread only ever touches its own SPT columns of the table
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
            nidx[q] = blk * COLS + q * THREADS + tid;
            live[q] = nidx[q] < P.n;
            double* t = mytab + q * THREADS;
            t[0] = 1.0;
            if (GERM) {
                for (int i = 0; i < s; i += 2) {
                    double z0, z1;
                    const int k0 = P.germ_kind[i];
                    const int k1 = (i + 1 < s) ? P.germ_kind[i + 1] : 0;
                    pcq_germ_pair(P.seed, (unsigned long long)(P.first_index + nidx[q]),
                                  (unsigned int)(i >> 1), k0, k1, z0, z1);
                    pcq_fill_1d(z0, i, s, P.pmax, P.rec_a, P.rec_b, P.p1, t, COLS);
                    if (i + 1 < s) pcq_fill_1d(z1, i + 1, s, P.pmax, P.rec_a, P.rec_b, P.p1, t, COLS);
                }
            } else {
                for (int i = 0; i < s; ++i) {
                    const double xv = live[q] ? P.x[nidx[q] * P.ldx + i] : 0.0;
                    pcq_fill_1d(xv, i, s, P.pmax, P.rec_a, P.rec_b, P.p1, t, COLS);
                }
            }
        }
        double vsq[SPT];
#pragma unroll
        for (int q = 0; q < SPT; ++q) vsq[q] = 0.0;
        for (int m0 = 0; m0 < P.m; m0 += MB) {
            double acc[MB][SPT];
#pragma unroll
            for (int j = 0; j < MB; ++j)
#pragma unroll
                for (int q = 0; q < SPT; ++q) acc[j][q] = 0.0;
            double pf[W > 1 ? W - 1 : 1][SPT];
            double ppsi[SPT];
#pragma unroll
            for (int q = 0; q < SPT; ++q) { ppsi[q] = 1.0; pf[0][q] = 1.0; }
            for (int kc = 0; kc < K; kc += K2_TC) {
                const int nk = min(K2_TC, K - kc);
                __syncthreads();       // previous chunk fully consumed
                for (int k = tid; k < nk; k += THREADS) {
                    const unsigned long long d = P.desc4[kc + k];
                    // x,y,z: byte offsets of the W-1 prefix slots; w: offset of the last slot with
                    // the "prefix changed" flag in bit 0 (offsets are multiples of 8)
                    unsigned o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int w = 0; w + 1 < W; ++w) o[w] = (unsigned)((d >> (16 * w)) & 0xffffull) * (COLS * 8);
                    o[3] = ((unsigned)((d >> (16 * (W - 1))) & 0xffffull) * (COLS * 8)) |
                           ((kc + k == 0) ? 1u : (unsigned)(P.flags[kc + k] & 1));
                    sdesc[k] = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
                    for (int j = 0; j < MB; ++j)
                        scf[j * K2_TC + k] = P.cf[(size_t)min(m0 + j, P.m - 1) * K + kc + k];
                }
                __syncthreads();
#pragma unroll 4
                for (int k = 0; k < nk; ++k) {
                    const uint4 o = sdesc[k];
                    if (W > 1 && (o.w & 1u)) {       // warp-uniform: new prefix
#pragma unroll
                        for (int q = 0; q < SPT; ++q) {
                            pf[0][q] = *reinterpret_cast<const double*>(mybase + o.x + q * THREADS * 8);
                            if constexpr (W >= 3) pf[1][q] = *reinterpret_cast<const double*>(mybase + o.y + q * THREADS * 8);
                            if constexpr (W >= 4) pf[2][q] = *reinterpret_cast<const double*>(mybase + o.z + q * THREADS * 8);
                            if constexpr (FAST) {
                                double pp = pf[0][q];
#pragma unroll
                                for (int w = 1; w + 1 < W; ++w) pp = __dmul_rn(pp, pf[w][q]);
                                ppsi[q] = pp;
                            }
                        }
                    }
                    double last[SPT];
#pragma unroll
                    for (int q = 0; q < SPT; ++q)
                        last[q] = *reinterpret_cast<const double*>(mybase + (o.w & ~1u) + q * THREADS * 8);
                    if constexpr (FAST) {
#pragma unroll
                        for (int q = 0; q < SPT; ++q) {
                            const double psi = (W > 1) ? __dmul_rn(ppsi[q], last[q]) : last[q];
#pragma unroll
                            for (int j = 0; j < MB; ++j) acc[j][q] = __fma_rn(scf[j * K2_TC + k], psi, acc[j][q]);
                        }
                    } else {
                        // rv/pcrv.py:492-495: prd = cf[k]; prd *= phi_i (i ascending); val += prd
#pragma unroll
                        for (int j = 0; j < MB; ++j) {
                            const double c = scf[j * K2_TC + k];
#pragma unroll
                            for (int q = 0; q < SPT; ++q) {
                                double prd = c;
#pragma unroll
                                for (int w = 0; w + 1 < W; ++w) prd = __dmul_rn(prd, pf[w][q]);
                                prd = __dmul_rn(prd, last[q]);
                                acc[j][q] = __dadd_rn(acc[j][q], prd);
                            }
                        }
                    }
                }
            }
            if (P.sumsq) {
#pragma unroll
                for (int q = 0; q < SPT; ++q)
#pragma unroll
                    for (int j = 0; j < MB; ++j)
                        if (m0 + j < P.m) vsq[q] = __fma_rn(acc[j][q], acc[j][q], vsq[q]);
            } else if (P.y != nullptr) {
#pragma unroll
                for (int q = 0; q < SPT; ++q)
                    if (live[q]) {
#pragma unroll
                        for (int j = 0; j < MB; ++j)
                            if (m0 + j < P.m) P.y[nidx[q] * P.ldy + m0 + j] = acc[j][q];
                    }
            }
            if constexpr (GERM) {
#pragma unroll
                for (int j = 0; j < MB; ++j) {
                    double v1 = 0.0, v2 = 0.0, mn = INFINITY, mx = -INFINITY;
#pragma unroll
                    for (int q = 0; q < SPT; ++q) {
                        if (live[q]) {
                            v1 += acc[j][q]; v2 += acc[j][q] * acc[j][q];
                            mn = fmin(mn, acc[j][q]); mx = fmax(mx, acc[j][q]);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    }
                    if ((tid & 31) == 0 && m0 + j < P.m) {
                        double* qp = mypart + (size_t)(m0 + j) * 4;
                        if (first_iter) { qp[0] = v1; qp[1] = v2; qp[2] = mn; qp[3] = mx; }
                        else { qp[0] += v1; qp[1] += v2; qp[2] = fmin(qp[2], mn); qp[3] = fmax(qp[3], mx); }
                    }
                }
            }
        }
        if (P.sumsq && P.y != nullptr) {
#pragma unroll
            for (int q = 0; q < SPT; ++q)
                if (live[q]) P.y[nidx[q] * P.ldy] = vsq[q];
        }
        first_iter = false;
    }
    if constexpr (GERM) {
        if (first_iter && (tid & 31) == 0) {
            for (int j = 0; j < P.m; ++j) {
                double* qp = mypart + (size_t)j * 4;
                qp[0] = 0.0; qp[1] = 0.0; qp[2] = INFINITY; qp[3] = -INFINITY;
            }
        }
    }
}

template <int W, int MB, bool FAST, int THREADS, int SPT, bool GERM>
int launch_evalpc2_t(pcq_plan* p, EvalParams P, cudaStream_t st, double* sums, size_t smem) {
    auto kern = pcq_evalpc2_kernel<W, MB, FAST, THREADS, SPT, GERM>;
    PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PCQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t cols = (int64_t)THREADS * SPT;
    const int64_t nblocks = (P.n + cols - 1) / cols;
    const int64_t cap = (int64_t)p->num_sms * occ;
    const int grid = (int)(nblocks < cap ? (nblocks > 0 ? nblocks : 1) : cap);
    if (GERM) {
        const int nrows = grid * (THREADS / 32);
        int rc = pcq_ensure_dev(&p->d_ws, &p->ws_bytes, (size_t)nrows * P.m * 4 * sizeof(double));
        if (rc != PCQ_OK) return rc;
        P.part = p->d_ws;
        kern<<<grid, THREADS, smem, st>>>(P);
        PCQ_LAUNCH_CHECK("pcq_evalpc2_kernel<germ>");
        pcq_prop_reduce_kernel<<<(P.m + 127) / 128, 128, 0, st>>>(P.part, nrows, P.m, sums);
        PCQ_LAUNCH_CHECK("pcq_prop_reduce_kernel");
        return PCQ_OK;
    }
    kern<<<grid, THREADS, smem, st>>>(P);
    PCQ_LAUNCH_CHECK("pcq_evalpc2_kernel");
    return PCQ_OK;
}

inline size_t evalpc2_smem(const pcq_plan* p, int threads, int spt, int mb) {
    return (size_t)p->nslots * threads * spt * 8 + (size_t)K2_TC * 16 + (size_t)mb * K2_TC * 8;
}

// picks (THREADS, SPT): prefer 2 samples per thread when that still leaves >= 8 warps per SM
template <int W, int MB, bool FAST, bool GERM>
int launch_evalpc2_cfg(pcq_plan* p, const EvalParams& P, cudaStream_t st, double* sums, bool* handled) {
    *handled = true;
    const size_t budget = p->smem_optin;
    const size_t per_sm = 228 * 1024 - 1024;
    int cfg = -1;    // 0:(128,2) 1:(64,2) 2:(128,1) 3:(64,1)
    const char* env = getenv("PCQ_K2_CFG");
    if (env) cfg = atoi(env);
    if (cfg < 0) {
        const int th[4] = {128, 64, 128, 64}, sp[4] = {2, 2, 1, 1};
        int best = -1; double best_score = -1.0;
        for (int c = 0; c < 4; ++c) {
            const size_t sm = evalpc2_smem(p, th[c], sp[c], MB);
            if (sm > budget) continue;
            int ctas = (int)(per_sm / (sm + 1024));
            ctas = ctas > 16 ? 16 : ctas;
            const int warps = ctas * th[c] / 32;
            if (warps < 1) continue;
            // throughput model: issue cost per sample-term falls with SPT, latency hiding grows with warps
            const double score = (sp[c] == 2 ? 1.25 : 1.0) * (warps >= 8 ? 1.0 : warps / 8.0);
            if (score > best_score) { best_score = score; best = c; }
        }
        cfg = best;
    }
    switch (cfg) {
        case 0: return launch_evalpc2_t<W, MB, FAST, 128, 2, GERM>(p, P, st, sums, evalpc2_smem(p, 128, 2, MB));
        case 1: return launch_evalpc2_t<W, MB, FAST, 64, 2, GERM>(p, P, st, sums, evalpc2_smem(p, 64, 2, MB));
        case 2: return launch_evalpc2_t<W, MB, FAST, 128, 1, GERM>(p, P, st, sums, evalpc2_smem(p, 128, 1, MB));
        case 3: return launch_evalpc2_t<W, MB, FAST, 64, 1, GERM>(p, P, st, sums, evalpc2_smem(p, 64, 1, MB));
        default: *handled = false; return PCQ_OK;
    }
}

template <int W, bool GERM>
int launch_evalpc2_w(pcq_plan* p, const EvalParams& P, int mode, cudaStream_t st, double* sums, bool* handled) {
    const bool fast = (mode & 1) != 0;
    if (fast && !GERM && P.m >= 16) {
        // many outputs (quadratic forms, KL modes): 16 accumulators per sample amortise the basis
        // product and the descriptor traffic over 16 fused multiply-adds
        return launch_evalpc2_cfg<W, 16, true, false>(p, P, st, sums, handled);
    }
    if (P.m >= 4) {
        return fast ? launch_evalpc2_cfg<W, 4, true, GERM>(p, P, st, sums, handled)
                    : launch_evalpc2_cfg<W, 4, false, GERM>(p, P, st, sums, handled);
    }
    return fast ? launch_evalpc2_cfg<W, 1, true, GERM>(p, P, st, sums, handled)
                : launch_evalpc2_cfg<W, 1, false, GERM>(p, P, st, sums, handled);
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
    const bool fast = (mode & 1) != 0;
    if (P.m >= 4) {
        return fast ? launch_evalpc_threads<W, 4, true, GERM>(p, P, st, sums)
                    : launch_evalpc_threads<W, 4, false, GERM>(p, P, st, sums);
    }
    return fast ? launch_evalpc_threads<W, 1, true, GERM>(p, P, st, sums)
                : launch_evalpc_threads<W, 1, false, GERM>(p, P, st, sums);
}

template <bool GERM>
int launch_evalpc_any(pcq_plan* p, const EvalParams& P, int mode, cudaStream_t st, double* sums) {
    bool handled = false;
    int rc = PCQ_OK;
    switch (p->width <= PCQ_MAXW_FAST ? p->width : 0) {
        case 1: rc = launch_evalpc2_w<1, GERM>(p, P, mode, st, sums, &handled); break;
        case 2: rc = launch_evalpc2_w<2, GERM>(p, P, mode, st, sums, &handled); break;
        case 3: rc = launch_evalpc2_w<3, GERM>(p, P, mode, st, sums, &handled); break;
        case 4: rc = launch_evalpc2_w<4, GERM>(p, P, mode, st, sums, &handled); break;
        default: break;
    }
    if (handled || rc != PCQ_OK) return rc;
    // generic width (or a table too large for the staged kernel): v1 kernel reading descw
    return launch_evalpc_w<0, GERM>(p, P, mode, st, sums);
}

// per-output power sums of a materialised y chunk (specialised propagation path): per-warp partial rows
__global__ void __launch_bounds__(256) pcq_moments_kernel(const double* __restrict__ y, int64_t n, int64_t ldy,
                                                          int m, double* part) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int j = 0; j < m; ++j) {
        double v1 = 0.0, v2 = 0.0, mn = INFINITY, mx = -INFINITY;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
            const double v = y[i * ldy + j];
            v1 += v; v2 = fma(v, v, v2); mn = fmin(mn, v); mx = fmax(mx, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) {
            double* q = part + ((size_t)warp * m + j) * 4;
            q[0] = v1; q[1] = v2; q[2] = mn; q[3] = mx;
        }
    }
}

// sums <- sums (+) tmp  (first != 0: sums <- tmp)
__global__ void pcq_moments_combine_kernel(double* sums, const double* tmp, int m, int first) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    if (first) {
        for (int q = 0; q < 4; ++q) sums[j * 4 + q] = tmp[j * 4 + q];
    } else {
        sums[j * 4 + 0] += tmp[j * 4 + 0];
        sums[j * 4 + 1] += tmp[j * 4 + 1];
        sums[j * 4 + 2] = fmin(sums[j * 4 + 2], tmp[j * 4 + 2]);
        sums[j * 4 + 3] = fmax(sums[j * 4 + 3], tmp[j * 4 + 3]);
    }
}

// residual sums of one output: per-warp partials of (sum r^2, sum y r), r = y - yhat
__global__ void __launch_bounds__(256) pcq_resid_kernel(const double* __restrict__ y, int64_t ldy,
                                                        const double* __restrict__ yhat, int64_t n, double* part) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double rr = 0.0, yr = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const double yv = y[i * ldy];
        const double r = yv - yhat[i];
        rr = fma(r, r, rr);
        yr = fma(yv, r, yr);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rr += __shfl_xor_sync(0xffffffffu, rr, o);
        yr += __shfl_xor_sync(0xffffffffu, yr, o);
    }
    if (lane == 0) { part[2 * warp] = rr; part[2 * warp + 1] = yr; }
}

// fixed-order sum of the per-warp partials; sums (+)= result
__global__ void __launch_bounds__(64) pcq_resid_reduce_kernel(const double* part, int nwarps, double* sums, int accumulate) {
    const int q = threadIdx.x;
    if (q >= 2) return;
    double acc = 0.0;
    for (int w = 0; w < nwarps; ++w) acc += part[2 * w + q];
    sums[q] = (accumulate ? sums[q] : 0.0) + acc;
}

// Large germ dimension: evalPC from slot tables in global memory ([slot][nc], pcq_launch_gtab).  One thread per
// sample, terms in k order, coefficient first, factors in descriptor order -- the reference's rounding sequence,
// used for both modes (the fast mode may differ from it by rounding only, it need not).
__global__ void __launch_bounds__(128) pcq_evalpc_g_kernel(EvalParams P, const double* __restrict__ tab, int64_t nc) {
    const int K = P.nterms, W = P.width;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * blockDim.x) {
        double vsq = 0.0;
        for (int m = 0; m < P.m; ++m) {
            const double* cf = P.cf + (size_t)m * K;
            double acc = 0.0;
            for (int k = 0; k < K; ++k) {
                const uint16_t* dw = P.descw + (size_t)k * W;
                double prd = __ldg(cf + k);
                for (int q = 0; q < W; ++q) prd = __dmul_rn(prd, tab[(size_t)__ldg(dw + q) * nc + i]);
                acc = __dadd_rn(acc, prd);
            }
            if (P.sumsq) vsq = fma(acc, acc, vsq);
            else P.y[i * P.ldy + m] = acc;
        }
        if (P.sumsq) P.y[i * P.ldy] = vsq;
    }
}

void fill_eval_params(pcq_plan* p, EvalParams& P) {
    memset(&P, 0, sizeof(P));
    P.sdim = p->sdim; P.nterms = p->nterms; P.pmax = p->pmax; P.nslots = p->nslots;
    P.width = p->width;
    P.rec_a = p->d_rec_a; P.rec_b = p->d_rec_b; P.p1 = p->d_p1;
    P.desc4 = p->d_desc4; P.descw = p->d_descw; P.flags = p->d_flags;
}

}  // namespace

int pcq_launch_evalpc(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf,
                      int m, double* y, int64_t ldy, int mode, cudaStream_t st) {
    if (n == 0 || m == 0) return PCQ_OK;
    EvalParams P;
    fill_eval_params(p, P);
    P.x = x; P.cf = cf; P.y = y; P.n = n; P.ldx = ldx; P.ldy = ldy; P.m = m;
    P.sumsq = (mode & 2) ? 1 : 0;
    const int fast = mode & 1;
    if (!pcq_table_fits_smem(p)) {
        const int64_t nc = PCQ_GTAB_ROWS;
        for (int64_t r0 = 0; r0 < n; r0 += nc) {
            EvalParams Q = P;
            Q.n = std::min<int64_t>(nc, n - r0);
            Q.x = x + r0 * ldx;
            Q.y = y + r0 * ldy;
            double* tab = nullptr;
            int rc = pcq_launch_gtab(p, Q.x, Q.n, ldx, nc, &tab, st);
            if (rc != PCQ_OK) return rc;
            pcq_evalpc_g_kernel<<<(int)((Q.n + 127) / 128), 128, 0, st>>>(Q, tab, nc);
            PCQ_LAUNCH_CHECK("pcq_evalpc_g_kernel");
        }
        return PCQ_OK;
    }
    if (!P.sumsq && p->spec[fast].load(std::memory_order_acquire) && m <= 8) {
        // multi-index-specialised kernels (pcq_plan_specialize): one output per launch sequence
        for (int j = 0; j < m; ++j) {
            int rc = pcq_spec_launch(p, fast, x, n, ldx, cf + (size_t)j * p->nterms, y + j, ldy, st);
            if (rc != PCQ_OK) return rc;
        }
        return PCQ_OK;
    }
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
    const int fast = mode & 1;
    const bool bigdim = !pcq_table_fits_smem(p);
    if ((p->spec[fast].load(std::memory_order_acquire) && m <= 8) || bigdim) {
        // materialised-germ path: germ chunk -> HBM (160 B per sample, far below the evaluation cost) ->
        // specialised evalPC (or, for a germ too wide for shared-memory tables, the global-table evalPC) ->
        // power sums; chunks of 4M samples (64k for wide germs), sums combined on the device
        const int64_t chunk = bigdim ? ((int64_t)1 << 16) : ((int64_t)1 << 22);
        const int s = p->sdim;
        if ((rc = pcq_ensure_dev(&p->d_stage[0], &p->stage_bytes[0], (size_t)chunk * s * 8)) != PCQ_OK) return rc;
        if (!y && (rc = pcq_ensure_dev(&p->d_out[0], &p->out_bytes[0], (size_t)chunk * m * 8)) != PCQ_OK) return rc;
        const int mgrid = p->num_sms * 4;
        const int nrows = mgrid * 8;
        if ((rc = pcq_ensure_dev(&p->d_ws, &p->ws_bytes, ((size_t)nrows * m * 4 + (size_t)m * 4) * 8)) != PCQ_OK) return rc;
        double* part = p->d_ws;
        double* tmp = p->d_ws + (size_t)nrows * m * 4;
        bool firstc = true;
        for (int64_t r0 = 0; r0 < n || firstc; r0 += chunk) {
            const int64_t nr = (n - r0 < chunk) ? (n - r0 > 0 ? n - r0 : 0) : chunk;
            double* yc = y ? y + r0 * ldy : p->d_out[0];
            const int64_t ldc = y ? ldy : m;
            if (nr > 0) {
                if ((rc = pcq_launch_germ(p, seed, first + r0, nr, germ_kind, p->d_stage[0], s, st)) != PCQ_OK) return rc;
                if (bigdim) {
                    if ((rc = pcq_launch_evalpc(p, p->d_stage[0], nr, s, cf, m, yc, ldc, mode & 1, st)) != PCQ_OK) return rc;
                } else {
                    for (int j = 0; j < m; ++j)
                        if ((rc = pcq_spec_launch(p, fast, p->d_stage[0], nr, s, cf + (size_t)j * p->nterms, yc + j, ldc, st)) != PCQ_OK)
                            return rc;
                }
            }
            pcq_moments_kernel<<<mgrid, 256, 0, st>>>(yc, nr, ldc, m, part);
            PCQ_LAUNCH_CHECK("pcq_moments_kernel");
            pcq_prop_reduce_kernel<<<(m + 127) / 128, 128, 0, st>>>(part, nrows, m, tmp);
            PCQ_LAUNCH_CHECK("pcq_prop_reduce_kernel");
            pcq_moments_combine_kernel<<<(m + 127) / 128, 128, 0, st>>>(sums, tmp, m, firstc ? 1 : 0);
            PCQ_LAUNCH_CHECK("pcq_moments_combine_kernel");
            firstc = false;
        }
        return PCQ_OK;
    }
    return launch_evalpc_any<true>(p, P, mode, st, sums);
}

int pcq_launch_residual(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy,
                        const double* cf, double* sums, int accumulate, cudaStream_t st) {
    const int64_t chunk = 1 << 22;
    const int grid = p->num_sms * 4;
    const int nwarps = grid * 8;
    int rc;
    if ((rc = pcq_ensure_dev(&p->d_out[0], &p->out_bytes[0], (size_t)std::min<int64_t>(chunk, std::max<int64_t>(n, 1)) * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&p->d_ws, &p->ws_bytes, (size_t)nwarps * 2 * 8)) != PCQ_OK) return rc;
    if (n == 0) {
        pcq_resid_reduce_kernel<<<1, 64, 0, st>>>(p->d_ws, 0, sums, accumulate);
        PCQ_LAUNCH_CHECK("pcq_resid_reduce_kernel");
        return PCQ_OK;
    }
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t nr = std::min(chunk, n - r0);
        if ((rc = pcq_launch_evalpc(p, x + r0 * ldx, nr, ldx, cf, 1, p->d_out[0], 1, 1, st)) != PCQ_OK) return rc;
        pcq_resid_kernel<<<grid, 256, 0, st>>>(y + r0 * ldy, ldy, p->d_out[0], nr, p->d_ws);
        PCQ_LAUNCH_CHECK("pcq_resid_kernel");
        pcq_resid_reduce_kernel<<<1, 64, 0, st>>>(p->d_ws, nwarps, sums, (accumulate || r0 > 0) ? 1 : 0);
        PCQ_LAUNCH_CHECK("pcq_resid_reduce_kernel");
    }
    return PCQ_OK;
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
