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
#include <algorithm>
#include <cstring>
#include <cstring>
#include <cstdlib>
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


// ------------------------------------------------------------------------------------------------
// K1 v2: the term descriptors live in registers.  A CTA owns 16 consecutive rows; the slot table
// is laid out [slot][row] (stride 17 doubles) so that with the row loop unrolled every gather is
// register + immediate; each warp walks 64-term chunks (two terms per lane), loads the two
// descriptors once and produces the chunk for all 16 rows: shared-memory traffic per value is the
// W gathers only.
constexpr int K1_TS = K1_ROWS + 1;

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

template <int W>
__global__ void __launch_bounds__(K1_THREADS) pcq_bases2_kernel(BasesParams P, int base_odd) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);     // [nslots][K1_TS]
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int NWARPS = K1_THREADS / 32;
    const int K = P.nterms;
    const int s = P.sdim;
    const int nchunks = (K + 63) / 64;
    const bool lda_odd = (P.lda & 1) != 0;

    const int64_t nblocks = (P.n + K1_ROWS - 1) / K1_ROWS;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t n0 = blk * K1_ROWS;
        const int rows = (int)min((int64_t)K1_ROWS, P.n - n0);
        __syncthreads();   // previous iteration's readers are done with tab
        for (int idx = tid; idx < K1_ROWS * s; idx += K1_THREADS) {
            const int r = idx / s;
            const int i = idx - r * s;
            if (i == 0) tab[r] = 1.0;
            const double xv = (r < rows) ? P.x[(n0 + r) * P.ldx + i] : 0.0;
            pcq_fill_1d(xv, i, s, P.pmax, P.rec_a, P.rec_b, P.p1, tab + r, K1_TS);
        }
        __syncthreads();
        // 32-bit shared-space addresses + ld.shared: keeps the generic->shared window arithmetic (S2R /
        // LEA per access, ncu r01-g) out of the inner loop
        const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab);
        for (int c = warp; c < nchunks; c += NWARPS) {
            const int k = c * 64 + 2 * lane;
            const bool v0 = k < K, v1 = k + 1 < K;
            const unsigned long long d0 = v0 ? __ldg(P.desc4 + k) : 0ull;
            const unsigned long long d1 = v1 ? __ldg(P.desc4 + k + 1) : 0ull;
            uint32_t o0[W], o1[W];
#pragma unroll
            for (int q = 0; q < W; ++q) {
                o0[q] = tab_s + (uint32_t)((d0 >> (16 * q)) & 0xffffull) * (K1_TS * 8);
                o1[q] = tab_s + (uint32_t)((d1 >> (16 * q)) & 0xffffull) * (K1_TS * 8);
            }
            double* out = P.a + n0 * P.lda + k;
#pragma unroll
            for (int r = 0; r < K1_ROWS; ++r) {
                if (r < rows) {
                    double a = lds_f64(o0[0] + r * 8), b = lds_f64(o1[0] + r * 8);
#pragma unroll
                    for (int q = 1; q < W; ++q) {
                        a = __dmul_rn(a, lds_f64(o0[q] + r * 8));
                        b = __dmul_rn(b, lds_f64(o1[q] + r * 8));
                    }
                    double* dst = out + (size_t)r * P.lda;
                    // n0 is a multiple of 16 and k is even: element k of this row is 16-byte aligned unless
                    // exactly one of (base pointer, odd leading dimension on an odd row) shifts it by 8 bytes
                    const bool shifted = (base_odd != 0) != (lda_odd && (r & 1));
                    if (!shifted) {
                        if (v1) __stcs(reinterpret_cast<double2*>(dst), make_double2(a, b));
                        else if (v0) __stcs(dst, a);
                    } else {
                        // aligned pairs are (k+1, k+2): own second value + the next lane's first; the chunk's
                        // first value and the last lane's second value go out as 8-byte stores
                        const double an = __shfl_down_sync(0xffffffffu, a, 1);
                        if (lane == 0 && v0) __stcs(dst, a);
                        if (lane < 31 && k + 2 < K) __stcs(reinterpret_cast<double2*>(dst + 1), make_double2(b, an));
                        else if (v1) __stcs(dst + 1, b);
                    }
                }
            }
        }
    }
}

template <int W>
int launch_bases2_w(pcq_plan* p, const BasesParams& P, cudaStream_t st, bool* handled) {
    const size_t smem = (size_t)p->nslots * K1_TS * sizeof(double);
    *handled = false;
    if (smem > p->smem_optin || getenv("PCQ_K1_V1")) return PCQ_OK;
    *handled = true;
    auto kern = pcq_bases2_kernel<W>;
    PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PCQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K1_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t nblocks = (P.n + K1_ROWS - 1) / K1_ROWS;
    const int64_t cap = (int64_t)p->num_sms * occ;
    const int grid = (int)(nblocks < cap ? nblocks : cap);
    const int base_odd = (int)((reinterpret_cast<uintptr_t>(P.a) >> 3) & 1);
    kern<<<grid, K1_THREADS, smem, st>>>(P, base_odd);
    PCQ_LAUNCH_CHECK("pcq_bases2_kernel");
    return PCQ_OK;
}


// ------------------------------------------------------------------------------------------------
// Large germ dimension (slot table of a few samples does not fit shared memory, e.g. d = 1000):
// the per-sample tables go to a global scratch laid out [slot][sample of the chunk] -- warps that walk
// samples read it coalesced, as they do the shared [slot][lane] tables -- and generic kernels gather
// from it through L1/L2.  Same recurrence, same product order: results are bit-identical to the
// shared-memory kernels.  This is the capacity route, not a tuned one.
__global__ void __launch_bounds__(256) pcq_gtab_kernel(BasesParams P, double* tab, int64_t nc) {
    const int s = P.sdim;
    const int64_t total = P.n * s;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx % P.n;            // sample fastest: coalesced table writes
        const int c = (int)(idx / P.n);
        if (c == 0) tab[i] = 1.0;
        pcq_fill_1d(P.x[i * P.ldx + c], c, s, P.pmax, P.rec_a, P.rec_b, P.p1, tab + i, (int)nc);
    }
}

// A[i][k] = prod_q tab[slot_q(k)][i]: a 32 x 32 (sample x term) tile per iteration, transposed through
// shared memory so that table reads run along samples and the stores along terms
__global__ void __launch_bounds__(256) pcq_bases_g_kernel(BasesParams P, const double* __restrict__ tab, int64_t nc) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t nbi = (P.n + 31) / 32;
    const int nbk = (P.nterms + 31) / 32;
    for (int64_t blk = blockIdx.x; blk < nbi * nbk; blk += gridDim.x) {
        const int64_t i0 = (blk / nbk) * 32;
        const int k0 = (int)(blk % nbk) * 32;
        __syncthreads();
        for (int kk = ty; kk < 32; kk += 8) {
            const int k = k0 + kk;
            const int64_t i = i0 + tx;
            double v = 0.0;
            if (k < P.nterms && i < P.n) {
                const uint16_t* dw = P.descw + (size_t)k * P.width;
                v = tab[(size_t)dw[0] * nc + i];
                for (int q = 1; q < P.width; ++q) v = __dmul_rn(v, tab[(size_t)dw[q] * nc + i]);
            }
            tile[kk][tx] = v;
        }
        __syncthreads();
        for (int ii = ty; ii < 32; ii += 8) {
            const int64_t i = i0 + ii;
            const int k = k0 + tx;
            if (i < P.n && k < P.nterms) P.a[i * P.lda + k] = tile[tx][ii];
        }
    }
}

}  // namespace

bool pcq_table_fits_smem(const pcq_plan* p) {
    // the tightest consumer is evalPC's [slot][lane] table with 32 threads per CTA
    return (size_t)p->nslots * sizeof(double) * 32 <= p->smem_optin &&
           (size_t)K1_ROWS * p->nslots * sizeof(double) <= p->smem_optin;
}

// slot tables of n samples in global scratch, [slot][nc]; returns the table through *tab
int pcq_launch_gtab(pcq_plan* p, const double* x, int64_t n, int64_t ldx, int64_t nc, double** tab, cudaStream_t st) {
    const int b = (st == p->streams[1] && st != nullptr) ? 1 : 0;     // table owned by the launching plan stream
    int rc = pcq_ensure_dev(&p->d_gtab[b], &p->gtab_bytes[b], (size_t)p->nslots * nc * sizeof(double));
    if (rc != PCQ_OK) return rc;
    BasesParams P;
    memset(&P, 0, sizeof(P));
    P.x = x; P.n = n; P.ldx = ldx; P.sdim = p->sdim; P.pmax = p->pmax; P.nslots = p->nslots;
    P.rec_a = p->d_rec_a; P.rec_b = p->d_rec_b; P.p1 = p->d_p1;
    const int64_t want = (n * p->sdim + 255) / 256, cap = (int64_t)p->num_sms * 8;
    pcq_gtab_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(P, p->d_gtab[b], nc);
    PCQ_LAUNCH_CHECK("pcq_gtab_kernel");
    *tab = p->d_gtab[b];
    return PCQ_OK;
}

static int launch_bases_bigdim(pcq_plan* p, const BasesParams& P0, cudaStream_t st) {
    const int64_t nc = PCQ_GTAB_ROWS;
    for (int64_t r0 = 0; r0 < P0.n; r0 += nc) {
        BasesParams P = P0;
        P.n = std::min<int64_t>(nc, P0.n - r0);
        P.x = P0.x + r0 * P0.ldx;
        P.a = P0.a + r0 * P0.lda;
        double* tab = nullptr;
        int rc = pcq_launch_gtab(p, P.x, P.n, P.ldx, nc, &tab, st);
        if (rc != PCQ_OK) return rc;
        const int64_t want = ((P.n + 31) / 32) * ((P.nterms + 31) / 32), cap = (int64_t)p->num_sms * 8;
        pcq_bases_g_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(P, tab, nc);
        PCQ_LAUNCH_CHECK("pcq_bases_g_kernel");
    }
    return PCQ_OK;
}

int pcq_launch_bases(pcq_plan* p, const double* x, int64_t n, int64_t ldx, double* a,
                     int64_t lda, cudaStream_t st) {
    if (n == 0) return PCQ_OK;
    BasesParams P;
    P.x = x; P.a = a; P.n = n; P.ldx = ldx; P.lda = lda;
    P.sdim = p->sdim; P.nterms = p->nterms; P.pmax = p->pmax; P.nslots = p->nslots;
    P.width = p->width;
    P.rec_a = p->d_rec_a; P.rec_b = p->d_rec_b; P.p1 = p->d_p1;
    P.desc4 = p->d_desc4; P.descw = p->d_descw;
    if (!pcq_table_fits_smem(p)) return launch_bases_bigdim(p, P, st);
    bool handled = false;
    int rc = PCQ_OK;
    switch (p->width <= PCQ_MAXW_FAST ? p->width : 0) {
        case 1: rc = launch_bases2_w<1>(p, P, st, &handled); break;
        case 2: rc = launch_bases2_w<2>(p, P, st, &handled); break;
        case 3: rc = launch_bases2_w<3>(p, P, st, &handled); break;
        case 4: rc = launch_bases2_w<4>(p, P, st, &handled); break;
        default: break;
    }
    if (handled || rc != PCQ_OK) return rc;
    switch (p->width <= PCQ_MAXW_FAST ? p->width : 0) {
        case 1: return launch_bases_w<1>(p, P, st);
        case 2: return launch_bases_w<2>(p, P, st);
        case 3: return launch_bases_w<3>(p, P, st);
        case 4: return launch_bases_w<4>(p, P, st);
        default: return launch_bases_w<0>(p, P, st);
    }
}
