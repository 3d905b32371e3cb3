// K3 / K4, fused-generator form: S = sum_n z_n z_n^T on the FP64 tensor pipe (DMMA) with the design
// matrix generated on chip.  Since r01-j the DEFAULT path is pack + SYRK (pcq_syrk.cu, ~15 % faster);
// this file is what runs with PCQ_GRAM_SYRK=0 -- it needs no scratch buffer for the packed chunk and
// never writes a design-matrix element to HBM -- and for n == 0.
//
//   z_n = [ Psi(x_n) (K) | y_n (M) | 1 ],  L = K + M + 1
//   S[:K,:K] = A^T A (anl.fita lreg/anl.py:78, bcs_fit lreg/bcs.py:84,95,100)
//   S[:K,K:K+M] = A^T Y (lreg/anl.py:83, lreg/bcs.py:83),  S[K:,K:] = Y^T Y, sums, N
//
// Work decomposition (stream-K over the upper triangle):
//   the L x L output is cut into 128 x 128 tiles, only tiles (I <= J) are computed; the
//   sample axis is cut into blocks of GRAM_SB samples; the ntri * nblk (tile, block)
//   units are laid out tile-major and divided evenly and contiguously among the CTAs
//   (one CTA per SM).  A CTA therefore works on at most two partial tiles plus complete
//   tiles in between.  Partial tiles go to a workspace and are summed in CTA order by a
//   fix-up kernel (deterministic, no atomics); a mirror kernel fills the lower triangle.
// Per CTA (256 threads, 8 warps as 4 x 2, warp tile 32 x 64, 64 accumulators per thread): operand
//   tiles [16 samples][128 columns] are GENERATED in shared memory as products of per-sample
//   slot-table entries and contracted with mma.sync.m8n8k4.f64 (DMMA); generation of stage it+1 is
//   interleaved into the DMMA stream of stage it.
// Two kernels share this skeleton (dispatch in pcq_launch_gram):
//   pcq_gram_tma_kernel   slot tables precomputed per L2-sized sample chunk by pcq_tables_kernel and
//                         fetched per stage with cp.async.bulk (TMA) + mbarrier (<= 4 non-zero orders)
//   pcq_gram_kernel       generic: any width, and a materialised A; plain generate / barrier / DMMA
//                         phases
#include "pcq_internal.cuh"
#include "pcq_moment.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int GT = 128;          // tile edge (terms)
constexpr int KT = 16;           // samples per pipeline stage
constexpr int LDT = GT + 4;      // padded operand row: conflict-free DMMA fragment loads
constexpr int G_THREADS = 256;
constexpr int GRAM_SB = 64;      // samples per scheduling unit (multiple of KT)
constexpr int XPF = 4;           // x values prefetched per thread per stage (pcq_gram_kernel)

struct GramParams {
    const double* x; int64_t ldx;
    const double* a; int64_t lda;
    const double* y; int64_t ldy;
    int kb;            // number of basis columns (K, or kcols of a materialised A)
    int m;
    int L;             // kb + m + 1
    int ntile;         // ceil(L / GT)
    int64_t n;
    int64_t nblk;      // ceil(n / GRAM_SB)
    int64_t total;     // ntri * nblk
    int sdim, pmax, nslots, width;
    const double* rec_a; const double* rec_b; const double* p1;
    const unsigned long long* desc4; const uint16_t* descw;
    double* s_out; int accumulate;
    double* ws;        // [grid][2][GT*GT]
    int mirror;        // fill the lower triangle after this launch
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void tile_decode(int64_t t, int ntile, int& I, int& J) {
    int i = 0;
    int64_t rem = t;
    while (rem >= (int64_t)(ntile - i)) { rem -= (ntile - i); ++i; }
    I = i;
    J = i + (int)rem;
}


// complete tiles go straight to S, partial ones to this CTA's workspace slot
__device__ __forceinline__ void flush_tile(const GramParams& P, double (&acc)[4][8][2], int I, int J,
                                           bool complete, int slot, int wm, int wn, int g, int tig) {
    if (complete) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = I * GT + wm * 32 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = J * GT + wn * 64 + j * 8 + 2 * tig;
                if (row < P.L) {
                    double* dst = P.s_out + (size_t)row * P.L + c;
                    if (c < P.L) dst[0] = (P.accumulate ? dst[0] : 0.0) + acc[i][j][0];
                    if (c + 1 < P.L) dst[1] = (P.accumulate ? dst[1] : 0.0) + acc[i][j][1];
                }
            }
        }
    } else {
        double* w = P.ws + ((size_t)blockIdx.x * 2 + slot) * GT * GT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = wm * 32 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = wn * 64 + j * 8 + 2 * tig;
                *reinterpret_cast<double2*>(w + (size_t)row * GT + c) =
                    make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
    }
}

template <int W, bool FUSED>
__global__ void __launch_bounds__(G_THREADS, 1) pcq_gram_kernel(GramParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: op[2 stages][2 operands][KT][LDT] | tab[2 stages][KT][nslots]
    double* op = reinterpret_cast<double*>(smem_raw);
    double* tab = op + 2 * 2 * KT * LDT;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2;      // groupID
    const int tig = lane & 3;     // thread in group
    const int wm = warp & 3;      // 4 warps along M (32 rows each)
    const int wn = warp >> 2;     // 2 warps along N (64 cols each)
    const int half = tid >> 7;    // which operand tile this thread generates
    const int col = tid & 127;
    const int s = P.sdim;

    const int64_t G = gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * P.total / G;
    const int64_t hi = ((int64_t)blockIdx.x + 1) * P.total / G;
    const int64_t first_tile = (hi > lo) ? lo / P.nblk : -1;

    int64_t u = lo;
    while (u < hi) {
        const int64_t t = u / P.nblk;
        const int64_t b0 = u - t * P.nblk;
        const int64_t seg_end = min(hi, (t + 1) * P.nblk);
        const int64_t nb = seg_end - u;
        const int64_t sbeg = b0 * GRAM_SB;
        const int64_t send = min(P.n, (b0 + nb) * GRAM_SB);
        int I, J;
        tile_decode(t, P.ntile, I, J);
        const bool diag = (I == J);

        // this thread's generated column
        const int gc = ((half == 0 || diag) ? I : J) * GT + col;
        int kind;   // 0 basis, 1 y, 2 ones, 3 zero padding
        if (gc < P.kb) kind = 0;
        else if (gc < P.kb + P.m) kind = 1;
        else if (gc == P.kb + P.m) kind = 2;
        else kind = 3;
        unsigned long long mydesc = 0;
        if (FUSED && W > 0 && kind == 0) mydesc = P.desc4[gc];
        const int rstart = diag ? half : 0;
        const int rstep = diag ? 2 : 1;
        double* opw_base = op + (size_t)((diag ? 0 : half) * KT) * LDT + col;

        double acc[4][8][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        const int nst = (int)((send - sbeg + KT - 1) / KT);
        const int npairs = KT * s;

        auto load_x = [&](int st, double (&xr)[XPF]) {
            if (!FUSED) return;
            const int64_t n0 = sbeg + (int64_t)st * KT;
#pragma unroll
            for (int j = 0; j < XPF; ++j) {
                const int idx = tid + j * G_THREADS;
                xr[j] = 0.0;
                if (idx < npairs && st < nst) {
                    const int r = idx / s;
                    const int i = idx - r * s;
                    if (n0 + r < send) xr[j] = P.x[(n0 + r) * P.ldx + i];
                }
            }
        };
        auto build_tables = [&](int st, const double (&xr)[XPF]) {
            if (!FUSED) return;
            double* T = tab + (size_t)(st & 1) * KT * P.nslots;
            const int64_t n0 = sbeg + (int64_t)st * KT;
#pragma unroll
            for (int j = 0; j < XPF; ++j) {
                const int idx = tid + j * G_THREADS;
                if (idx < npairs) {
                    const int r = idx / s;
                    const int i = idx - r * s;
                    double* tr = T + (size_t)r * P.nslots;
                    if (i == 0) tr[0] = 1.0;
                    pcq_fill_1d(xr[j], i, s, P.pmax, P.rec_a, P.rec_b, P.p1, tr, 1);
                }
            }
            for (int idx = tid + XPF * G_THREADS; idx < npairs; idx += G_THREADS) {
                const int r = idx / s;
                const int i = idx - r * s;
                double* tr = T + (size_t)r * P.nslots;
                if (i == 0) tr[0] = 1.0;
                const double xv = (n0 + r < send) ? P.x[(n0 + r) * P.ldx + i] : 0.0;
                pcq_fill_1d(xv, i, s, P.pmax, P.rec_a, P.rec_b, P.p1, tr, 1);
            }
        };
        auto generate = [&](int st) {
            const double* T = tab + (size_t)(st & 1) * KT * P.nslots;
            double* o = opw_base + (size_t)(st & 1) * 2 * KT * LDT;
            const int64_t n0 = sbeg + (int64_t)st * KT;
#pragma unroll 4
            for (int r = rstart; r < KT; r += rstep) {
                const bool live = (n0 + r) < send;
                double v;
                if (kind == 0) {
                    if (FUSED) {
                        if (W > 0) v = pcq_term4<W>(T + (size_t)r * P.nslots, 1, mydesc);
                        else v = pcq_termw(T + (size_t)r * P.nslots, 1,
                                           P.descw + (size_t)gc * P.width, P.width);
                    } else {
                        v = live ? P.a[(n0 + r) * P.lda + gc] : 0.0;
                    }
                } else if (kind == 1) {
                    v = live ? P.y[(n0 + r) * P.ldy + (gc - P.kb)] : 0.0;
                } else if (kind == 2) {
                    v = 1.0;
                } else {
                    v = 0.0;
                }
                o[(size_t)r * LDT] = live ? v : 0.0;
            }
        };
        auto mma_stage = [&](int st) {
            const double* oa = op + (size_t)(st & 1) * 2 * KT * LDT;
            const double* ob = oa + (diag ? 0 : KT * LDT);
            const double* pa = oa + (size_t)tig * LDT + wm * 32 + g;
            const double* pb = ob + (size_t)tig * LDT + wn * 64 + g;
#pragma unroll
            for (int kk = 0; kk < KT; kk += 4) {
                double af[4], bf[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = pa[(size_t)kk * LDT + i * 8];
#pragma unroll
                for (int j = 0; j < 8; ++j) bf[j] = pb[(size_t)kk * LDT + j * 8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        };

        __syncthreads();   // previous segment's MMA readers are done with op / tab
        double xr[XPF];
        load_x(0, xr);
        build_tables(0, xr);
        load_x(1, xr);
        __syncthreads();
        generate(0);
        for (int it = 0; it < nst; ++it) {
            if (it + 1 < nst) {
                build_tables(it + 1, xr);
                load_x(it + 2, xr);
            }
            __syncthreads();
            if (it + 1 < nst) generate(it + 1);
            mma_stage(it);
        }

        flush_tile(P, acc, I, J, nb == P.nblk, (t == first_tile) ? 0 : 1, wm, wn, g, tig);
        u = seg_end;
    }
}

// sums the partial tiles of every tile that was split across CTAs, in CTA order
__global__ void __launch_bounds__(256) pcq_gram_fixup_kernel(GramParams P, int grid_main) {
    const int64_t t = blockIdx.x;
    const int64_t G = grid_main;
    const int64_t u0 = t * P.nblk;
    const int64_t u1 = (t + 1) * P.nblk - 1;
    const int64_t c_first = ((u0 + 1) * G - 1) / P.total;
    const int64_t c_last = ((u1 + 1) * G - 1) / P.total;
    if (c_first == c_last) return;   // complete tile, already written by its owner
    int I, J;
    tile_decode(t, P.ntile, I, J);
    for (int e = threadIdx.x + blockIdx.y * blockDim.x; e < GT * GT; e += blockDim.x * gridDim.y) {
        const int r = e / GT;
        const int c = e - r * GT;
        const int row = I * GT + r;
        const int cc = J * GT + c;
        if (row >= P.L || cc >= P.L) continue;
        double sum = 0.0;
        for (int64_t cta = c_first; cta <= c_last; ++cta) {
            const int64_t lo = cta * P.total / G;
            const int slot = (lo / P.nblk == t) ? 0 : 1;
            sum += P.ws[((size_t)cta * 2 + slot) * GT * GT + e];
        }
        double* dst = P.s_out + (size_t)row * P.L + cc;
        *dst = (P.accumulate ? *dst : 0.0) + sum;
    }
}

// lower triangle <- upper triangle
__global__ void __launch_bounds__(256) pcq_gram_mirror_kernel(double* s, int L) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;   // source blocks: upper triangle (incl. diagonal blocks)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int row = bi * 32 + r, c = bj * 32 + tx;
        tile[r][tx] = (row < L && c < L) ? s[(size_t)row * L + c] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int row = bj * 32 + r, c = bi * 32 + tx;   // transposed destination
        if (row < L && c < L && row > c) s[(size_t)row * L + c] = tile[tx][r];
    }
}

template <int W, bool FUSED>
int launch_gram_t(const GramParams& P0, int num_sms, size_t smem_optin, double** ws,
                  size_t* ws_bytes, cudaStream_t st) {
    GramParams P = P0;
    const size_t smem = (size_t)(2 * 2 * KT * LDT + (FUSED ? 2 * KT * P.nslots : 0)) * sizeof(double);
    if (smem > smem_optin) {
        pcq_set_error("suffstats: %zu bytes of shared memory needed, %zu available", smem, smem_optin);
        return PCQ_ERR_UNSUPPORTED;
    }
    auto kern = pcq_gram_kernel<W, FUSED>;
    PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid64 = P.total < num_sms ? P.total : num_sms;
    const int grid = (int)grid64;
    const size_t need = (size_t)grid * 2 * GT * GT * sizeof(double);
    int rc = pcq_ensure_dev(ws, ws_bytes, need);
    if (rc != PCQ_OK) return rc;
    P.ws = *ws;
    kern<<<grid, G_THREADS, smem, st>>>(P);
    PCQ_LAUNCH_CHECK("pcq_gram_kernel");
    const int ntri = P.ntile * (P.ntile + 1) / 2;
    pcq_gram_fixup_kernel<<<dim3(ntri, 4), 256, 0, st>>>(P, grid);
    PCQ_LAUNCH_CHECK("pcq_gram_fixup_kernel");
    const int nb32 = (P.L + 31) / 32;
    if (P.mirror) {
        pcq_gram_mirror_kernel<<<dim3(nb32, nb32), 256, 0, st>>>(P.s_out, P.L);
        PCQ_LAUNCH_CHECK("pcq_gram_mirror_kernel");
    }
    return PCQ_OK;
}


// ------------------------------------------------------------------------------------------------
// K3 v6: slot tables precomputed per L2-sized chunk and fed to the Gram kernel by TMA.
//   The recurrences depend on the sample only, yet every one of the ntri tiles rebuilt them, and on
//   sm_100a those FP64 vector instructions share the pipe with DMMA (profiles/r01_c: ~10 % of the pipe
//   plus bubbles).  Here a small kernel writes, once per chunk, one blob per 16 samples that is already
//   in the generator's shared-memory layout [slot][17] (slots: 1, phi..., y..., 0); the Gram kernel then
//   fetches each stage's blob with ONE cp.async.bulk (TMA, mbarrier complete_tx) two stages ahead:
//   no recurrences, no x/y prefetch registers and no global-load address arithmetic in the DMMA warps.
//   The chunk's blobs (518 B per sample at C2) stay L2-resident while the tiles sweep them.
constexpr int TB_KT = 16;
constexpr int TB_TS = TB_KT + 1;

__host__ __device__ inline int tb_slots_even(int nslots, int m) { return (nslots + m + 1 + 1) & ~1; }

// one CTA = 8 blobs (128 samples)
__global__ void __launch_bounds__(256) pcq_tables_kernel(GramParams P, double* blobs, int64_t nblocks16) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);          // [8][NTe][17]
    const int NTe = tb_slots_even(P.nslots, P.m);
    const int blob = NTe * TB_TS;
    const int s = P.sdim;
    const int tid = threadIdx.x;
    for (int64_t b0 = (int64_t)blockIdx.x * 8; b0 < nblocks16; b0 += (int64_t)gridDim.x * 8) {
        const int nb = (int)min((int64_t)8, nblocks16 - b0);
        __syncthreads();
        for (int i = tid; i < nb * blob; i += 256) stage[i] = 0.0;
        __syncthreads();
        const int64_t n0 = b0 * TB_KT;
        // recurrences: one (sample, dim) pair per thread
        for (int idx = tid; idx < nb * TB_KT * s; idx += 256) {
            const int sm = idx / s, c = idx - sm * s;
            const int64_t n = n0 + sm;
            if (n >= P.n) continue;
            double* T = stage + (size_t)(sm / TB_KT) * blob + (sm % TB_KT);
            if (c == 0) T[0] = 1.0;
            const double v = P.x[n * P.ldx + c];
            pcq_fill_1d(v, c, s, P.pmax, P.rec_a, P.rec_b, P.p1, T, TB_TS);
        }
        for (int idx = tid; idx < nb * TB_KT * P.m; idx += 256) {
            const int sm = idx / P.m, c = idx - sm * P.m;
            const int64_t n = n0 + sm;
            if (n >= P.n) continue;
            stage[(size_t)(sm / TB_KT) * blob + (size_t)(P.nslots + c) * TB_TS + (sm % TB_KT)] = P.y[n * P.ldy + c];
        }
        __syncthreads();
        double* dst = blobs + (size_t)b0 * blob;
        for (int i = tid; i < nb * blob; i += 256) dst[i] = stage[i];
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct TmaSmem {
    double* op;        // [2][2][16][LDT]
    double* tab;       // [3][NTe][17]
    uint64_t* bars;    // [3]
};

template <int W, bool DIAG>
__device__ __forceinline__ void fused_tma_segment(const GramParams& P, const TmaSmem& sm, const double* blobs,
                                                  int blob, int I, int J, int64_t blk0, int nst,
                                                  double (&acc)[4][8][2], int& gbase, uint32_t& phases) {
    constexpr int KT = TB_KT;
    constexpr int TS = TB_TS;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int half = tid >> 7, col = tid & 127;
    const uint32_t blob_bytes = (uint32_t)blob * 8u;

    const int gc = ((half == 0 || DIAG) ? I : J) * GT + col;
    int off[W];
#pragma unroll
    for (int q = 0; q < W; ++q) off[q] = 0;
    if (gc < P.kb) {
        const unsigned long long d = P.desc4[gc];
#pragma unroll
        for (int q = 0; q < W; ++q) off[q] = (int)((d >> (16 * q)) & 0xffffull) * TS;
    } else if (gc < P.kb + P.m) {
        off[0] = (P.nslots + gc - P.kb) * TS;
    } else if (gc > P.kb + P.m) {
        off[0] = (P.nslots + P.m) * TS;
    }
    const int rbase = DIAG ? half : 0;

    auto issue = [&](int st) {          // one elected thread: TMA bulk copy of stage st's blob
        if (tid == 0 && st < nst) {
            const int b = (gbase + st) % 3;
            mbar_expect_tx(sm.bars + b, blob_bytes);
            tma_bulk_g2s(sm.tab + (size_t)b * blob, blobs + (size_t)(blk0 + st) * blob, blob_bytes, sm.bars + b);
        }
    };
    auto wait_stage = [&](int st) {
        if (st < nst) {
            const int b = (gbase + st) % 3;
            mbar_wait(sm.bars + b, (phases >> b) & 1u);
            phases ^= (1u << b);
        }
    };
    auto gen_value = [&](const double* Tn, int r) {
        double v = Tn[off[0] + r];
#pragma unroll
        for (int q = 1; q < W; ++q) v = __dmul_rn(v, Tn[off[q] + r]);
        return v;
    };

    __syncthreads();   // previous segment's readers are done with op / tab
    issue(0);
    issue(1);
    wait_stage(0);
    {
        const double* T0 = sm.tab + (size_t)((gbase + 0) % 3) * blob + rbase;
        double* o0 = sm.op + (size_t)((DIAG ? 0 : half) * KT) * LDT + col + rbase * LDT;
#pragma unroll
        for (int r = 0; r < KT; r += (DIAG ? 2 : 1)) o0[(size_t)r * LDT] = gen_value(T0, r);
    }
    __syncthreads();

    for (int it = 0; it < nst; ++it) {
        const int cur = it & 1, nxt = cur ^ 1;
        issue(it + 2);             // its buffer was last read by the generator two iterations ago
        wait_stage(it + 1);
        const double* Tn = sm.tab + (size_t)((gbase + it + 1) % 3) * blob + rbase;
        double* on = sm.op + (size_t)nxt * 2 * KT * LDT + (size_t)((DIAG ? 0 : half) * KT) * LDT + col +
                     rbase * LDT;
        const double* oa = sm.op + (size_t)cur * 2 * KT * LDT;
        const double* ob = oa + (DIAG ? 0 : KT * LDT);
        const double* pa = oa + (size_t)tig * LDT + wm * 32 + g;
        const double* pb = ob + (size_t)tig * LDT + wn * 64 + g;
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = pa[(size_t)kk * LDT + i * 8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = pb[(size_t)kk * LDT + j * 8];
            double gv[4];
#pragma unroll
            for (int q = 0; q < (DIAG ? 2 : 4); ++q) {
                const int r = DIAG ? (kk + 2 * q) : (kk + q);
                gv[q] = gen_value(Tn, r);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
#pragma unroll
            for (int q = 0; q < (DIAG ? 2 : 4); ++q) {
                const int r = DIAG ? (kk + 2 * q) : (kk + q);
                on[(size_t)r * LDT] = gv[q];
            }
        }
        __syncthreads();
    }
    gbase = (gbase + nst) % 3;
}

template <int W>
__global__ void __launch_bounds__(G_THREADS, 1) pcq_gram_tma_kernel(GramParams P, const double* blobs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NTe = tb_slots_even(P.nslots, P.m);
    const int blob = NTe * TB_TS;
    TmaSmem sm;
    sm.op = reinterpret_cast<double*>(smem_raw);
    sm.tab = sm.op + 2 * 2 * TB_KT * LDT;
    sm.bars = reinterpret_cast<uint64_t*>(sm.tab + (size_t)3 * blob);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < 3; ++b) mbar_init(sm.bars + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3, wm = warp & 3, wn = warp >> 2;
    const int64_t G = gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * P.total / G;
    const int64_t hi = ((int64_t)blockIdx.x + 1) * P.total / G;
    const int64_t first_tile = (hi > lo) ? lo / P.nblk : -1;
    int gbase = 0;
    uint32_t phases = 0;

    int64_t u = lo;
    while (u < hi) {
        const int64_t t = u / P.nblk;
        const int64_t b0 = u - t * P.nblk;
        const int64_t seg_end = min(hi, (t + 1) * P.nblk);
        const int64_t nb = seg_end - u;
        const int64_t sbeg = b0 * GRAM_SB;
        const int64_t send = min(P.n, (b0 + nb) * GRAM_SB);
        const int nst = (int)((send - sbeg + TB_KT - 1) / TB_KT);
        int I, J;
        tile_decode(t, P.ntile, I, J);
        double acc[4][8][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        if (I == J) fused_tma_segment<W, true>(P, sm, blobs, blob, I, J, sbeg / TB_KT, nst, acc, gbase, phases);
        else fused_tma_segment<W, false>(P, sm, blobs, blob, I, J, sbeg / TB_KT, nst, acc, gbase, phases);
        flush_tile(P, acc, I, J, nb == P.nblk, (t == first_tile) ? 0 : 1, wm, wn, g, tig);
        u = seg_end;
    }
}

template <int W>
int launch_gram_tma(const GramParams& P0, int num_sms, size_t smem_optin, double** ws,
                    size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st, bool* handled) {
    GramParams P = P0;
    const int NTe = tb_slots_even(P.nslots, P.m);
    const size_t blob = (size_t)NTe * TB_TS;
    const size_t smem = ((size_t)2 * 2 * TB_KT * LDT + 3 * blob) * sizeof(double) + 3 * sizeof(uint64_t) + 128;
    const size_t smem_tab = 8 * blob * sizeof(double);
    *handled = false;
    if (smem > smem_optin || smem_tab > smem_optin) return PCQ_OK;
    *handled = true;
    const int64_t nblocks16 = (P.n + TB_KT - 1) / TB_KT;
    int rc = pcq_ensure_dev(zbuf, zbuf_bytes, (size_t)nblocks16 * blob * sizeof(double));
    if (rc != PCQ_OK) return rc;
    PCQ_CUDA(cudaFuncSetAttribute(pcq_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tab));
    const int64_t want = (nblocks16 + 7) / 8;
    const int tgrid = (int)(want < (int64_t)num_sms * 2 ? want : (int64_t)num_sms * 2);
    pcq_tables_kernel<<<tgrid, 256, smem_tab, st>>>(P, *zbuf, nblocks16);
    PCQ_LAUNCH_CHECK("pcq_tables_kernel");
    auto kern = pcq_gram_tma_kernel<W>;
    PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(P.total < num_sms ? P.total : num_sms);
    rc = pcq_ensure_dev(ws, ws_bytes, (size_t)grid * 2 * GT * GT * sizeof(double));
    if (rc != PCQ_OK) return rc;
    P.ws = *ws;
    kern<<<grid, G_THREADS, smem, st>>>(P, *zbuf);
    PCQ_LAUNCH_CHECK("pcq_gram_tma_kernel");
    const int ntri = P.ntile * (P.ntile + 1) / 2;
    pcq_gram_fixup_kernel<<<dim3(ntri, 4), 256, 0, st>>>(P, grid);
    PCQ_LAUNCH_CHECK("pcq_gram_fixup_kernel");
    const int nb32 = (P.L + 31) / 32;
    if (P.mirror) {
        pcq_gram_mirror_kernel<<<dim3(nb32, nb32), 256, 0, st>>>(P.s_out, P.L);
        PCQ_LAUNCH_CHECK("pcq_gram_mirror_kernel");
    }
    return PCQ_OK;
}

__global__ void pcq_fill_kernel(double* p, size_t n, double v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}

}  // namespace

static int launch_gram_once(pcq_plan* p, int num_sms, size_t smem_optin, const double* x, int64_t n,
                            int64_t ldx, const double* a, int64_t lda, int kcols, const double* y,
                            int64_t ldy, int m, double* s, int accumulate, int mirror, double** ws,
                            size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st) {
    GramParams P;
    memset(&P, 0, sizeof(P));
    P.mirror = mirror;
    const bool fused = (a == nullptr);
    P.x = x; P.ldx = ldx; P.a = a; P.lda = lda; P.y = y; P.ldy = ldy;
    P.kb = fused ? p->nterms : kcols;
    P.m = m;
    P.L = P.kb + m + 1;
    P.ntile = (P.L + GT - 1) / GT;
    P.n = n;
    P.s_out = s; P.accumulate = accumulate;
    if (n == 0) {
        if (!accumulate) {
            pcq_fill_kernel<<<256, 256, 0, st>>>(s, (size_t)P.L * P.L, 0.0);
            PCQ_LAUNCH_CHECK("pcq_fill_kernel");
        }
        return PCQ_OK;
    }
    P.nblk = (n + GRAM_SB - 1) / GRAM_SB;
    P.total = (int64_t)(P.ntile * (P.ntile + 1) / 2) * P.nblk;
    if (fused) {
        P.sdim = p->sdim; P.pmax = p->pmax; P.nslots = p->nslots; P.width = p->width;
        P.rec_a = p->d_rec_a; P.rec_b = p->d_rec_b; P.p1 = p->d_p1;
        P.desc4 = p->d_desc4; P.descw = p->d_descw;
        bool handled = false;
        int rc = PCQ_OK;
        const int v1_only = getenv("PCQ_GRAM_V1") != nullptr;
        if (!v1_only) {
            switch (p->width <= PCQ_MAXW_FAST ? p->width : 0) {
                case 1: rc = launch_gram_tma<1>(P, num_sms, smem_optin, ws, ws_bytes, zbuf, zbuf_bytes, st, &handled); break;
                case 2: rc = launch_gram_tma<2>(P, num_sms, smem_optin, ws, ws_bytes, zbuf, zbuf_bytes, st, &handled); break;
                case 3: rc = launch_gram_tma<3>(P, num_sms, smem_optin, ws, ws_bytes, zbuf, zbuf_bytes, st, &handled); break;
                case 4: rc = launch_gram_tma<4>(P, num_sms, smem_optin, ws, ws_bytes, zbuf, zbuf_bytes, st, &handled); break;
                default: break;
            }
            if (handled || rc != PCQ_OK) return rc;
        }
        switch (p->width <= PCQ_MAXW_FAST ? p->width : 0) {
            case 1: return launch_gram_t<1, true>(P, num_sms, smem_optin, ws, ws_bytes, st);
            case 2: return launch_gram_t<2, true>(P, num_sms, smem_optin, ws, ws_bytes, st);
            case 3: return launch_gram_t<3, true>(P, num_sms, smem_optin, ws, ws_bytes, st);
            case 4: return launch_gram_t<4, true>(P, num_sms, smem_optin, ws, ws_bytes, st);
            default: return launch_gram_t<0, true>(P, num_sms, smem_optin, ws, ws_bytes, st);
        }
    }
    return launch_gram_t<0, false>(P, num_sms, smem_optin, ws, ws_bytes, st);
}

// Every sample block is read by all ntri output tiles.  With one launch over all samples the tiles
// reach a given block at different times and x is re-read from HBM ~ntri times (ncu, r01: 18.8 GB for
// 210 MB of input).  Processing the samples in L2-sized chunks (default 48 MB of x/y per launch,
// accumulating into S) keeps each chunk L2-resident while the tiles sweep it.
int pcq_launch_gram(pcq_plan* p, int num_sms, size_t smem_optin, const double* x, int64_t n,
                    int64_t ldx, const double* a, int64_t lda, int kcols, const double* y,
                    int64_t ldy, int m, double* s, int accumulate, double** ws,
                    size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st) {
    if (n > 0 && a == nullptr && !pcq_table_fits_smem(p)) {
        // germ too wide for shared-memory slot tables: design-matrix chunks through the global-table K1
        // (pcq_launch_bases takes that route by itself), then the materialised-matrix path on each chunk
        const int64_t rows = std::max<int64_t>(32, std::min<int64_t>(n, ((int64_t)1 << 29) / ((int64_t)p->nterms * 8)));
        int rc = pcq_ensure_dev(&p->d_achunk, &p->achunk_bytes, (size_t)rows * p->nterms * sizeof(double));
        if (rc != PCQ_OK) return rc;
        for (int64_t r0 = 0; r0 < n; r0 += rows) {
            const int64_t nr = std::min(rows, n - r0);
            if ((rc = pcq_launch_bases(p, x + r0 * ldx, nr, ldx, p->d_achunk, p->nterms, st)) != PCQ_OK) return rc;
            bool handled = false;
            rc = pcq_launch_syrk(p, num_sms, smem_optin, nullptr, nr, 0, p->d_achunk, p->nterms, p->nterms,
                                 y ? y + r0 * ldy : nullptr, ldy, m, s, accumulate || r0 > 0, ws, ws_bytes, zbuf,
                                 zbuf_bytes, st, &handled);
            if (rc != PCQ_OK) return rc;
            if (!handled) { pcq_set_error("suffstats: SYRK pipeline unavailable for the wide-germ route"); return PCQ_ERR_UNSUPPORTED; }
        }
        return PCQ_OK;
    }
    if (n > 0 && a == nullptr && p->mom) {
        // moment form (pcq_moment.cu): N |Gamma| instead of N K^2 / 2 multiply-adds where the planner expects a gain
        bool handled = false;
        int rc = pcq_launch_moments(p, x, n, ldx, y, ldy, m, s, accumulate, ws, ws_bytes, zbuf, zbuf_bytes, st, &handled);
        if (handled || rc != PCQ_OK) return rc;
        // more outputs than the moment form carries (pc_fit of a field, C5): G, A^T 1 and N in moment form, the output
        // columns (A^T Y, Y^T Y, Y^T 1) by the SYRK over the tile columns that hold them
        const char* he = getenv("PCQ_GRAM_HYBRID");
        if (m > MOM_MAXOUT && y && !(he && atoi(he) == 0)) {
            rc = pcq_launch_moments(p, x, n, ldx, nullptr, 0, 0, s, accumulate, ws, ws_bytes, zbuf, zbuf_bytes, st, &handled, m);
            if (rc != PCQ_OK) return rc;
            if (handled) {
                bool h2 = false;
                rc = pcq_launch_syrk(p, num_sms, smem_optin, x, n, ldx, nullptr, 0, 0, y, ldy, m, s, accumulate, ws, ws_bytes,
                                     zbuf, zbuf_bytes, st, &h2, true);
                if (rc != PCQ_OK) return rc;
                if (!h2) { pcq_set_error("suffstats: SYRK pipeline unavailable for the output columns"); return PCQ_ERR_UNSUPPORTED; }
                return PCQ_OK;
            }
        }
    }
    {
        const char* e = getenv("PCQ_GRAM_SYRK");
        if (n > 0 && !(e && atoi(e) == 0)) {
            bool handled = false;
            const int rc = pcq_launch_syrk(p, num_sms, smem_optin, x, n, ldx, a, lda, kcols, y, ldy, m, s, accumulate,
                                           ws, ws_bytes, zbuf, zbuf_bytes, st, &handled);
            if (handled || rc != PCQ_OK) return rc;
        }
    }
    if (n == 0)
        return launch_gram_once(p, num_sms, smem_optin, x, 0, ldx, a, lda, kcols, y, ldy, m, s, accumulate, 1,
                                ws, ws_bytes, zbuf, zbuf_bytes, st);
    const bool fused = (a == nullptr);
    const bool use_tma = fused && p->width <= PCQ_MAXW_FAST && getenv("PCQ_GRAM_V1") == nullptr;
    const int64_t row_bytes = use_tma ? (int64_t)tb_slots_even(p->nslots, m) * TB_TS * 8 / TB_KT
                                      : 8 * (int64_t)((fused ? p->sdim : kcols) + m);
    const char* env = getenv("PCQ_GRAM_CHUNK_MB");
    const int64_t target = (int64_t)(env ? atoi(env) : 48) << 20;
    int64_t rows = target > 0 ? target / row_bytes : n;
    rows = (rows / GRAM_SB) * GRAM_SB;
    if (rows < 8192) rows = 8192;
    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = (n - r0 < rows) ? n - r0 : rows;
        const int last = (r0 + nr >= n);
        int rc = launch_gram_once(p, num_sms, smem_optin, fused ? x + r0 * ldx : nullptr, nr, ldx,
                                  fused ? nullptr : a + r0 * lda, lda, kcols, y ? y + r0 * ldy : nullptr, ldy, m,
                                  s, accumulate || r0 > 0, last, ws, ws_bytes, zbuf, zbuf_bytes, st);
        if (rc != PCQ_OK) return rc;
    }
    return PCQ_OK;
}
