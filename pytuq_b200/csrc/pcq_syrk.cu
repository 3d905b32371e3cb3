// K3 / K4, "pack + SYRK" form: S = Z^T Z on the FP64 tensor pipe with NOTHING but DMMA and 128-bit
// shared-memory loads in the inner loop.
//
//   z_n = [ Psi(x_n) (K) | y_n (M) | 1 | 0-padding ],  Lp = 128 * ceil((K + M + 1) / 128) columns
//
// The fused generator (pcq_gram.cu) rebuilds every design-matrix element once per output tile that
// uses it (~L/128 times); its products and table gathers share the FP64 pipe and the issue slots with
// the DMMAs and cap it at ~80 % of the executed-DMMA peak.  Here the two concerns are separated:
//
//   pack   (pcq_pack_bases_kernel / pcq_pack_matrix_kernel, HBM-bound, ~3 % of the time)
//          writes a chunk of Z ONCE, in the order the tensor-pipe fragments want it:
//              Zp[sample block b (8 samples)][column c (Lp)][8 samples]         (doubles)
//          so that a (128 columns x 8 samples) operand chunk is one contiguous 8 KB run, and a lane's
//          two k-step values of a column are adjacent (one LDS.128, conflict free).  The k index of
//          mma.m8n8k4 is a summation index, so lane tig may feed samples (2 tig, 2 tig + 1) of a
//          block in its two k-steps as long as both operands agree.
//   syrk   (pcq_syrk_kernel)  one persistent CTA per SM: 8 warps (4 x 2, warp tile 32 x 64, 64
//          accumulators per thread); warp 0 also keeps a 3-stage ring of (2 operands x 4 blocks x
//          8 KB) filled with cp.async.bulk (TMA) on "full" mbarriers and every warp releases a stage
//          on an "empty" mbarrier, so there is no CTA-wide barrier anywhere in the main loop.
//          Per chunk it runs twice: over the off-diagonal tiles (full 128 x 128 units) and over the
//          diagonal tiles, whose upper-triangle 32 x 32 sub-blocks form one list cut into units of
//          eight (one per warp, two per scheduler: half the cost of a full tile, 1.25 units per tile).
//          Work units are (sample part, tile or sub-block group), part-major, dealt round-robin: the
//          number of parts S is chosen per pass so that units * S fills the SMs in whole waves
//          (K = 1771: 91 tiles x 13 parts = 7.99 waves; 18 groups x 8 parts = 0.97), and CTAs of a
//          wave sweep the same sample window at the same time, so the operand chunks are shared
//          through L2 instead of re-read from HBM.
//   fixup  sums the S partial tiles of every output tile in part order (deterministic, no atomics);
//          the mirror kernel fills the lower triangle.
#include "pcq_mma.cuh"
#include <cstdlib>
#include <cstring>

namespace {

struct SyrkParams {
    const double* zp;
    int Lp, L, ntile, ntri, nsplit;   // ntri: tiles of THIS launch (all, off-diagonal only, or diagonal only)
    int kind;                         // 0 all upper-triangle tiles, 1 off-diagonal tiles, 2 diagonal tiles,
                                      // 3 + t0: the tile columns J >= t0 only (I <= J), full tiles ("rect" mode)
    int cmin;                         // rect mode: only entries in columns >= cmin are written, and of the last column
                                      // (the ones) only the rows in [cmin, L - 1): the rest of S belongs to the moment form
    int64_t nstages;       // stages (of STAGE_SAMPLES samples) in this chunk
    double* s_out;
    int accumulate;
    double* ws;            // [ntri][nsplit][ST*ST], used when nsplit > 1
};

struct PackParams {
    const double* x; int64_t ldx;
    const double* a; int64_t lda;
    const double* y; int64_t ldy;
    int kb, m, Lp;
    int64_t n;             // live samples in this chunk
    int64_t npairs;        // 16-sample groups to write (covers the stage padding)
    int sdim, pmax, nslots, width;
    const double* rec_a; const double* rec_b; const double* p1;
    const unsigned long long* desc4; const uint16_t* descw;
    double* zp;
};

__device__ __forceinline__ void tile_decode(int t, int ntile, int& I, int& J) {
    int i = 0, rem = t;
    while (rem >= ntile - i) { rem -= (ntile - i); ++i; }
    I = i;
    J = i + rem;
}

// ------------------------------------------------------------------------------------------ pack
// K3: basis values generated from x (bit-identical to K1: same recurrence, same product order)
template <int W>
__global__ void __launch_bounds__(256) pcq_pack_bases_kernel(PackParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [nslots + m + 1][18]: 1, phi..., y..., 0.  Row stride 18: sample pairs are 16-byte aligned
    // (LDS.128) and eight consecutive slots fall into distinct bank groups.
    constexpr int TS = PACK_TS;
    double* T = reinterpret_cast<double*>(smem_raw);
    const int s = P.sdim, tid = threadIdx.x;
    const int nrow = P.nslots + P.m + 1;
    const int zero_slot = P.nslots + P.m;
    for (int i = tid; i < nrow * TS; i += 256) T[i] = 0.0;
    for (int64_t pair = blockIdx.x; pair < P.npairs; pair += gridDim.x) {
        const int64_t n0 = pair * 16;
        __syncthreads();
        if (n0 + 16 > P.n) {     // ragged tail: dead samples must contribute exact zeros
            for (int i = tid; i < zero_slot * TS; i += 256) T[i] = 0.0;
            __syncthreads();
        }
        for (int idx = tid; idx < 16 * s; idx += 256) {
            const int sm = idx / s, c = idx - sm * s;
            const int64_t n = n0 + sm;
            if (n >= P.n) continue;
            if (c == 0) T[sm] = 1.0;
            pcq_fill_1d(P.x[n * P.ldx + c], c, s, P.pmax, P.rec_a, P.rec_b, P.p1, T + sm, TS);
        }
        for (int idx = tid; idx < 16 * P.m; idx += 256) {
            const int sm = idx / P.m, c = idx - sm * P.m;
            const int64_t n = n0 + sm;
            if (n < P.n) T[(P.nslots + c) * TS + sm] = P.y[n * P.ldy + c];
        }
        __syncthreads();
        // four lanes per column (lane & 3 = sample pair): a warp store instruction covers 8 columns x 64 B
        // = 512 contiguous bytes of the packed chunk
        double* out = P.zp + (size_t)pair * 2 * P.Lp * SB;
        const int q2 = 2 * (tid & 3);
        for (int col = tid >> 2; col < P.Lp; col += 64) {
            double2 v[2];
            if (col < P.kb) {
                if (W > 0) {
                    const unsigned long long d = P.desc4[col];
                    int off[W > 0 ? W : 1];
#pragma unroll
                    for (int q = 0; q < W; ++q) off[q] = (int)((d >> (16 * q)) & 0xffffull) * TS + q2;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double2 t = *reinterpret_cast<const double2*>(T + off[0] + 8 * h);
#pragma unroll
                        for (int q = 1; q < W; ++q) {
                            const double2 f = *reinterpret_cast<const double2*>(T + off[q] + 8 * h);
                            t.x = __dmul_rn(t.x, f.x);
                            t.y = __dmul_rn(t.y, f.y);
                        }
                        v[h] = t;
                    }
                } else {
                    const uint16_t* dw = P.descw + (size_t)col * P.width;
#pragma unroll
                    for (int h = 0; h < 2; ++h) v[h] = *reinterpret_cast<const double2*>(T + (int)dw[0] * TS + q2 + 8 * h);
                    for (int q = 1; q < P.width; ++q) {
                        const int o = (int)dw[q] * TS + q2;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const double2 f = *reinterpret_cast<const double2*>(T + o + 8 * h);
                            v[h].x = __dmul_rn(v[h].x, f.x);
                            v[h].y = __dmul_rn(v[h].y, f.y);
                        }
                    }
                }
            } else {
                const int slot = col < P.kb + P.m ? P.nslots + (col - P.kb) : (col == P.kb + P.m ? 0 : zero_slot);
#pragma unroll
                for (int h = 0; h < 2; ++h) v[h] = *reinterpret_cast<const double2*>(T + slot * TS + q2 + 8 * h);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
                __stcs(reinterpret_cast<double2*>(out + ((size_t)h * P.Lp + col) * SB + q2), v[h]);
        }
    }
}

// K4: the caller's design matrix (row-major, any lda) transposed into block order through shared memory
__global__ void __launch_bounds__(256) pcq_pack_matrix_kernel(PackParams P) {
    __shared__ double tile[16][257];
    const int tid = threadIdx.x;
    const int ncg = (P.Lp + 255) / 256;
    const int64_t total = P.npairs * ncg;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t pair = w / ncg;
        const int c0 = (int)(w - pair * ncg) * 256;
        const int64_t n0 = pair * 16;
        const int col = c0 + tid;
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < 16; ++r) {
            const int64_t n = n0 + r;
            double v = 0.0;
            if (n < P.n) {
                if (col < P.kb) v = P.a[n * P.lda + col];
                else if (col < P.kb + P.m) v = P.y[n * P.ldy + (col - P.kb)];
                else if (col == P.kb + P.m) v = 1.0;
            }
            tile[r][tid] = v;
        }
        __syncthreads();
        if (col < P.Lp) {
            double* out = P.zp + (size_t)pair * 2 * P.Lp * SB;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double2* dst = reinterpret_cast<double2*>(out + ((size_t)h * P.Lp + col) * SB);
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    dst[r] = make_double2(tile[h * 8 + 2 * r][tid], tile[h * 8 + 2 * r + 1][tid]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ syrk
// rect mode (cmin > 0): is entry (row, c) of S this launch's to write?
__device__ __forceinline__ bool syrk_keeps(const SyrkParams& P, int row, int c) {
    if (P.cmin <= 0) return true;
    return c >= P.cmin && (c < P.L - 1 || (row >= P.cmin && row < P.L - 1));
}
__device__ __forceinline__ void tile_of(int kind, int t, int ntile, int& I, int& J) {
    if (kind == 2) { I = J = t; return; }
    if (kind == 0) { tile_decode(t, ntile, I, J); return; }
    if (kind >= 3) {                          // column J holds J + 1 tiles, columns from t0 on
        int j = kind - 3, rem = t;
        while (rem > j) { rem -= j + 1; ++j; }
        I = rem;
        J = j;
        return;
    }
    int i = 0, rem = t;                       // off-diagonal: row i holds ntile - 1 - i tiles
    while (rem >= ntile - 1 - i) { rem -= ntile - 1 - i; ++i; }
    I = i;
    J = i + 1 + rem;
}

// Diagonal tiles only need their upper triangle: 10 of the 16 32 x 32 sub-blocks.  The sub-blocks of ALL diagonal
// tiles are laid out in one list (tile-major) and cut into units of eight, one sub-block per warp, so that every
// scheduler (two warps) carries exactly two of them: a diagonal unit costs half a full tile and 1.25 units cover a
// tile.  Eight consecutive list entries touch at most two tiles, which arrive in the two operand slots of a stage.
__device__ __forceinline__ void diag_subblock(int h, int& hr, int& hc) {
    // h: 0..9 -> (0,0) (0,1) (0,2) (0,3) (1,1) (1,2) (1,3) (2,2) (2,3) (3,3)
    hr = h < 4 ? 0 : (h < 7 ? 1 : (h < 9 ? 2 : 3));
    hc = h < 4 ? h : (h < 7 ? h - 3 : (h < 9 ? h - 5 : 3));
}
__device__ __forceinline__ void diag_unit_tiles(int t, int ntile, int& I, int& J) {
    I = (t * 8) / 10;
    J = (t * 8 + 7) / 10;
    if (J > ntile - 1) J = ntile - 1;
}

// producer cursor: the (unit, stage) sequence this CTA walks, one step ahead of the consumers
struct SyCursor {
    int u; int I, J; int64_t st, st1; uint32_t pit;
};
__device__ __forceinline__ void cursor_open(const SyrkParams& P, SyCursor& c, int units) {
    while (c.u < units) {
        const int part = c.u / P.ntri, t = c.u - part * P.ntri;
        if (P.kind == 2) diag_unit_tiles(t, P.ntile, c.I, c.J);
        else tile_of(P.kind, t, P.ntile, c.I, c.J);
        c.st = (int64_t)part * P.nstages / P.nsplit;
        c.st1 = (int64_t)(part + 1) * P.nstages / P.nsplit;
        if (c.st < c.st1) return;
        c.u += gridDim.x;
    }
}

template <int ST, bool DIAGH>      // DIAGH: diagonal tiles computed as balanced 32 x 32 sub-blocks (P.kind == 2, ST == 128)
__global__ void __launch_bounds__(SY_THREADS, 1) pcq_syrk_kernel(SyrkParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int CHUNK = ST * SB;        // doubles per operand chunk (8 KB at ST = 128)
    constexpr int FM = ST / 32;           // 8-row fragments per warp along M (4 warps)
    constexpr int FN = ST / 16;           // 8-column fragments per warp along N (2 warps)
    double* ops = reinterpret_cast<double*>(smem_raw);                        // [NSTAGE][2][KS][CHUNK]
    uint64_t* full = reinterpret_cast<uint64_t*>(ops + (size_t)NSTAGE * 2 * KS * CHUNK);
    uint64_t* empty = full + NSTAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int b = 0; b < NSTAGE; ++b) { mbar_init(full + b, 1); mbar_init(empty + b, CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int units = P.ntri * P.nsplit;
    uint32_t it = 0;      // stages consumed so far by this CTA (same sequence in every warp)

    // Warp 0 doubles as the producer (a ninth warp would cost a whole 4-warp register allocation
    // granule: 168 registers/thread instead of 255).  It refills the ring buffer that stage it-1
    // used as soon as all eight warps have released it, polling between its own DMMA blocks.
    SyCursor cur;
    cur.u = blockIdx.x; cur.pit = 0; cur.I = cur.J = 0; cur.st = cur.st1 = 0;
    if (warp == 0) cursor_open(P, cur, units);
    auto try_issue = [&](bool blocking) {
        if (cur.u >= units || cur.pit >= it + NSTAGE) return;
        const int b = cur.pit % NSTAGE;
        const uint32_t ph = (cur.pit / NSTAGE) & 1u;
        if (blocking) mbar_wait(empty + b, ph ^ 1u);
        else if (!mbar_try(empty + b, ph ^ 1u)) return;
        const bool diag = (cur.I == cur.J);
        if (lane == 0) mbar_expect_tx(full + b, (diag ? 1u : 2u) * KS * CHUNK * 8u);
        __syncwarp();
        if (lane < 2 * KS && (lane < KS || !diag)) {
            const int op = lane / KS, kb = lane % KS;
            const double* src = P.zp + ((size_t)(cur.st * KS + kb) * P.Lp + (size_t)(op ? cur.J : cur.I) * ST) * SB;
            tma_bulk_g2s(ops + ((size_t)(b * 2 + op) * KS + kb) * CHUNK, src, CHUNK * 8u, full + b);
        }
        ++cur.pit;
        if (++cur.st >= cur.st1) { cur.u += gridDim.x; cursor_open(P, cur, units); }
    };
    if (warp == 0) {
        try_issue(true);
        try_issue(true);
    }

    // -------------------- consumer warps
    const int g = lane >> 2, tig = lane & 3, wm = warp & 3, wn = warp >> 2;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int part = u / P.ntri, t = u - part * P.ntri;
        int I, J;
        // DIAGH: this warp's sub-block (hr, hc) of diagonal tile mytile, operand slot `which`; idle past the list end
        int mytile = 0, hr = 0, hc = 0, which = 0;
        bool active = true;
        if (DIAGH) {
            diag_unit_tiles(t, P.ntile, I, J);
            const int e = t * 8 + warp;
            active = e < 10 * P.ntile;
            mytile = active ? e / 10 : I;
            diag_subblock(active ? e - mytile * 10 : 0, hr, hc);
            which = (mytile != I) ? 1 : 0;
        } else {
            tile_of(P.kind, t, P.ntile, I, J);
        }
        const bool diag = (I == J);
        const int64_t st0 = (int64_t)part * P.nstages / P.nsplit;
        const int64_t st1 = (int64_t)(part + 1) * P.nstages / P.nsplit;
        double acc[FM][FN][2];            // DIAGH uses acc[i][j], i, j < 4 (one 32 x 32 sub-block)
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        for (int64_t st = st0; st < st1; ++st, ++it) {
            const int b = it % NSTAGE;
            const uint32_t ph = (it / NSTAGE) & 1u;
            mbar_wait(full + b, ph);
            const double* oa = ops + (size_t)(b * 2) * KS * CHUNK;
            if (DIAGH) {
                const double* q = oa + (which ? (size_t)KS * CHUNK : 0) + g * SB + 2 * tig;
#pragma unroll
                for (int kb = 0; kb < KS; ++kb) {
                    if (active) {
                        double2 af[4], bf[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) af[i] = *reinterpret_cast<const double2*>(q + kb * CHUNK + (hr * 32 + i * 8) * SB);
#pragma unroll
                        for (int j = 0; j < 4; ++j) bf[j] = *reinterpret_cast<const double2*>(q + kb * CHUNK + (hc * 32 + j * 8) * SB);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i].y, bf[j].y);
                    }
                    if (warp == 0) try_issue(kb == KS - 1);
                }
            } else {
                const double* ob = diag ? oa : oa + (size_t)KS * CHUNK;
                const double* pa = oa + (wm * (FM * 8) + g) * SB + 2 * tig;
                const double* pb = ob + (wn * (FN * 8) + g) * SB + 2 * tig;
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
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + b);
        }

        // tile-local (row, column) of accumulator (i, j): the regular 4 x 2 warp grid, or this warp's sub-block
        auto rc_of = [&](int i, int j, int& row, int& c) {
            if (DIAGH) {
                row = hr * 32 + i * 8 + g;
                c = hc * 32 + j * 8 + 2 * tig;
            } else {
                row = wm * (FM * 8) + i * 8 + g;
                c = wn * (FN * 8) + j * 8 + 2 * tig;
            }
        };
        const int ti = DIAGH ? mytile : I, tj = DIAGH ? mytile : J;     // output tile this warp writes
        const int tslot = DIAGH ? mytile : t;                            // its workspace slot
        double* w = P.nsplit == 1 ? nullptr : P.ws + ((size_t)tslot * P.nsplit + part) * ST * ST;
        if (!DIAGH || active) {
#pragma unroll
            for (int i = 0; i < FM; ++i) {
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    if (DIAGH && j >= 4) continue;
                    int rl, cl;
                    rc_of(i, j, rl, cl);
                    if (w) {
                        *reinterpret_cast<double2*>(w + (size_t)rl * ST + cl) = make_double2(acc[i][j][0], acc[i][j][1]);
                    } else {
                        const int row = ti * ST + rl, c = tj * ST + cl;
                        if (row < P.L) {
                            double* dst = P.s_out + (size_t)row * P.L + c;
                            if (c < P.L && syrk_keeps(P, row, c)) dst[0] = (P.accumulate ? dst[0] : 0.0) + acc[i][j][0];
                            if (c + 1 < P.L && syrk_keeps(P, row, c + 1)) dst[1] = (P.accumulate ? dst[1] : 0.0) + acc[i][j][1];
                        }
                    }
                }
            }
        }
    }
}

// sums the nsplit partial tiles of every output tile, in part order; grid (ntri, 8)
template <int ST>
__global__ void __launch_bounds__(256) pcq_syrk_fixup_kernel(SyrkParams P) {
    const int t = blockIdx.x;
    int I, J;
    tile_of(P.kind, t, P.ntile, I, J);
    const double* w = P.ws + (size_t)t * P.nsplit * ST * ST;
    constexpr int PER = ST * ST / 8;      // elements per blockIdx.y slice
    for (int e = blockIdx.y * PER + threadIdx.x; e < (blockIdx.y + 1) * PER; e += 256) {
        const int r = e / ST, c = e - r * ST;
        const int row = I * ST + r, cc = J * ST + c;
        if (row >= P.L || cc >= P.L || row > cc || !syrk_keeps(P, row, cc)) continue;      // the mirror kernel fills the lower triangle
        double sum = 0.0;
        for (int q = 0; q < P.nsplit; ++q) sum += w[(size_t)q * ST * ST + e];
        double* dst = P.s_out + (size_t)row * P.L + cc;
        *dst = (P.accumulate ? *dst : 0.0) + sum;
    }
}

// lower triangle <- upper triangle
__global__ void __launch_bounds__(256) pcq_syrk_mirror_kernel(double* s, int L) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int row = bi * 32 + r, c = bj * 32 + tx;
        tile[r][tx] = (row < L && c < L) ? s[(size_t)row * L + c] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int row = bj * 32 + r, c = bi * 32 + tx;
        if (row < L && c < L && row > c) s[(size_t)row * L + c] = tile[tx][r];
    }
}

// number of sample parts: fill the SMs in whole waves (time ~ ceil(ntri * S / G) / S), ties -> fewer parts
int choose_split(int ntri, int grid, int64_t nstages) {
    int best = 1;
    double best_cost = 1e300;
    int smax = (int)(nstages < 32 ? nstages : 32);
    const int mem_cap = 8192 / (ntri > 0 ? ntri : 1);       // partial tiles of one pass stay below ~1 GB of workspace
    if (smax > mem_cap) smax = mem_cap < 1 ? 1 : mem_cap;
    for (int S = 1; S <= smax; ++S) {
        const int64_t waves = ((int64_t)ntri * S + grid - 1) / grid;
        // every part is ceil(nstages / S) stages at worst; a small per-unit term for the tile flush
        const double cost = (double)waves * ((double)((nstages + S - 1) / S) + 0.5);
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = S; }
    }
    return best;
}

}  // namespace


// S (+)= Z^T Z over n samples, generated from the plan (a == nullptr) or taken from a materialised A.
// zbuf: lazily grown device scratch for the packed chunk; ws: workspace for the partial tiles.
int pcq_launch_syrk(pcq_plan* p, int num_sms, size_t smem_optin, const double* x, int64_t n, int64_t ldx,
                    const double* a, int64_t lda, int kcols, const double* y, int64_t ldy, int m, double* s,
                    int accumulate, double** ws, size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes,
                    cudaStream_t st, bool* handled, bool rect) {
    *handled = false;
    const bool fused = (a == nullptr);
    const int kb = fused ? p->nterms : kcols;
    const int L = kb + m + 1;
    // tile edge: 96 only where it executes clearly fewer flops -- measured per executed flop it is ~7 % slower
    // than 128 (36 instead of 64 independent DMMAs between fragment loads); PCQ_SYRK_TILE forces it
    int ST = 128;
    {
        // cost in flops: 128-wide tiles with the diagonal tiles at 5/8 (sub-block pass), 96-wide tiles in full
        const double n128 = (L + 127) / 128, n96 = (L + 95) / 96;
        const double cost128 = (n128 * (n128 - 1) / 2 + 0.625 * n128) * 128.0 * 128.0;
        const double cost96 = n96 * (n96 + 1) / 2 * 96.0 * 96.0;
        if (cost96 * 1.07 < cost128) ST = 96;
        if (const char* e = getenv("PCQ_SYRK_TILE")) { const int v = atoi(e); if (v == 96 || v == 128) ST = v; }
    }
    const int ntile = (L + ST - 1) / ST;
    const int Lp = ntile * ST;
    const int ntri = ntile * (ntile + 1) / 2;
    const size_t smem = syrk_smem_bytes(ST);
    size_t smem_pack = 0;
    if (fused) smem_pack = (size_t)(p->nslots + m + 1) * PACK_TS * sizeof(double);
    if (smem > smem_optin || smem_pack > smem_optin || n <= 0) return PCQ_OK;
    *handled = true;

    const char* env = getenv("PCQ_SYRK_CHUNK_MB");
    const int64_t target = (int64_t)(env ? atoi(env) : 2048) << 20;
    int64_t rows = target / ((int64_t)Lp * 8);
    rows = (rows / STAGE_SAMPLES) * STAGE_SAMPLES;
    if (rows < STAGE_SAMPLES) rows = STAGE_SAMPLES;
    const int64_t npad = (n + STAGE_SAMPLES - 1) / STAGE_SAMPLES * STAGE_SAMPLES;
    if (rows > npad) rows = npad;
    {   // equal chunks instead of full ones plus a small remainder (the remainder launch quantises badly)
        const int64_t nchunks = (n + rows - 1) / rows;
        const int64_t even = ((n + nchunks - 1) / nchunks + STAGE_SAMPLES - 1) / STAGE_SAMPLES * STAGE_SAMPLES;
        if (even < rows) rows = even;
    }
    int rc = pcq_ensure_dev(zbuf, zbuf_bytes, (size_t)rows * Lp * sizeof(double));
    if (rc != PCQ_OK) return rc;
    PCQ_CUDA(cudaFuncSetAttribute(pcq_syrk_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)syrk_smem_bytes(128)));
    PCQ_CUDA(cudaFuncSetAttribute(pcq_syrk_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)syrk_smem_bytes(128)));
    PCQ_CUDA(cudaFuncSetAttribute(pcq_syrk_kernel<96, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)syrk_smem_bytes(96)));

    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = (n - r0 < rows) ? n - r0 : rows;
        const int64_t nstages = (nr + STAGE_SAMPLES - 1) / STAGE_SAMPLES;
        PackParams K;
        memset(&K, 0, sizeof(K));
        K.kb = kb; K.m = m; K.Lp = Lp; K.n = nr; K.npairs = nstages * (STAGE_SAMPLES / 16); K.zp = *zbuf;
        K.y = y ? y + r0 * ldy : nullptr; K.ldy = ldy;
        if (fused) {
            K.x = x + r0 * ldx; K.ldx = ldx;
            K.sdim = p->sdim; K.pmax = p->pmax; K.nslots = p->nslots; K.width = p->width;
            K.rec_a = p->d_rec_a; K.rec_b = p->d_rec_b; K.p1 = p->d_p1; K.desc4 = p->d_desc4; K.descw = p->d_descw;
            const int64_t cap = (int64_t)num_sms * 8;
            const int grid = (int)(K.npairs < cap ? K.npairs : cap);
            auto launch = [&](auto kern) -> int {
                PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pack));
                kern<<<grid, 256, smem_pack, st>>>(K);
                return PCQ_OK;
            };
            switch (p->width <= 4 ? p->width : 0) {
                case 1: rc = launch(pcq_pack_bases_kernel<1>); break;
                case 2: rc = launch(pcq_pack_bases_kernel<2>); break;
                case 3: rc = launch(pcq_pack_bases_kernel<3>); break;
                case 4: rc = launch(pcq_pack_bases_kernel<4>); break;
                default: rc = launch(pcq_pack_bases_kernel<0>); break;
            }
            if (rc != PCQ_OK) return rc;
            PCQ_LAUNCH_CHECK("pcq_pack_bases_kernel");
        } else {
            K.a = a + r0 * lda; K.lda = lda;
            const int64_t want = K.npairs * ((Lp + 255) / 256);
            const int64_t cap = (int64_t)num_sms * 16;
            pcq_pack_matrix_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(K);
            PCQ_LAUNCH_CHECK("pcq_pack_matrix_kernel");
        }

        // One launch over all upper-triangle tiles, or -- 128-wide tiles, more than one tile -- two: the
        // off-diagonal tiles, then the diagonal tiles as balanced sub-blocks (3/4 of the cost each).  Each
        // launch gets its own number of sample parts, so both fill the SMs in whole waves
        // (K = 1771: 91 tiles x 13 parts = 7.99 waves, 14 diagonal tiles x 21 parts = 1.99 waves).
        const char* denv = getenv("PCQ_SYRK_DIAG");
        const bool split_diag = ST == 128 && !(denv && atoi(denv) == 0) && !rect;
        struct Pass { int kind, ntiles; bool diagh; int wstiles; };     // ntiles: scheduling units per part
        Pass passes[2];
        int npass = 0;
        if (split_diag) {
            if (ntri - ntile > 0) passes[npass++] = {1, ntri - ntile, false, ntri - ntile};
            passes[npass++] = {2, (10 * ntile + 7) / 8, true, ntile};
        } else if (rect) {
            // the y columns against everything (A^T Y, Y^T Y, Y^T 1): tile columns from the one holding column kb on
            const int t0 = kb / ST;
            const int nt = ntri - t0 * (t0 + 1) / 2;
            passes[npass++] = {3 + t0, nt, false, nt};
        } else {
            passes[npass++] = {0, ntri, false, ntri};
        }
        const char* senv = getenv("PCQ_SYRK_SPLIT");
        int nsplit[2] = {1, 1};
        size_t ws_need = 0;
        for (int q = 0; q < npass; ++q) {
            int S = senv ? atoi(senv) : choose_split(passes[q].ntiles, num_sms, nstages);
            if (S < 1) S = 1;
            if (S > nstages) S = (int)nstages;
            nsplit[q] = S;
            if (S > 1) ws_need += (size_t)passes[q].wstiles * S * ST * ST * sizeof(double);
        }
        if (ws_need) {
            rc = pcq_ensure_dev(ws, ws_bytes, ws_need);
            if (rc != PCQ_OK) return rc;
        }
        size_t ws_off = 0;
        for (int q = 0; q < npass; ++q) {
            SyrkParams P;
            memset(&P, 0, sizeof(P));
            P.zp = *zbuf; P.Lp = Lp; P.L = L; P.ntile = ntile; P.ntri = passes[q].ntiles; P.kind = passes[q].kind;
            P.nstages = nstages; P.nsplit = nsplit[q];
            P.s_out = s; P.accumulate = (accumulate || r0 > 0) ? 1 : 0;
            P.cmin = rect ? kb : 0;
            if (P.nsplit > 1) {
                P.ws = *ws + ws_off;
                ws_off += (size_t)passes[q].wstiles * P.nsplit * ST * ST;
            }
            const int units = P.ntri * P.nsplit;
            const int sgrid = units < num_sms ? units : num_sms;
            if (passes[q].diagh) pcq_syrk_kernel<128, true><<<sgrid, SY_THREADS, smem, st>>>(P);
            else if (ST == 128) pcq_syrk_kernel<128, false><<<sgrid, SY_THREADS, smem, st>>>(P);
            else pcq_syrk_kernel<96, false><<<sgrid, SY_THREADS, smem, st>>>(P);
            PCQ_LAUNCH_CHECK("pcq_syrk_kernel");
            if (P.nsplit > 1) {
                if (ST == 128) pcq_syrk_fixup_kernel<128><<<dim3(passes[q].wstiles, 8), 256, 0, st>>>(P);
                else pcq_syrk_fixup_kernel<96><<<dim3(passes[q].wstiles, 8), 256, 0, st>>>(P);
                PCQ_LAUNCH_CHECK("pcq_syrk_fixup_kernel");
            }
        }
    }
    const int nb32 = (L + 31) / 32;
    pcq_syrk_mirror_kernel<<<dim3(nb32, nb32), 256, 0, st>>>(s, L);
    PCQ_LAUNCH_CHECK("pcq_syrk_mirror_kernel");
    return PCQ_OK;
}
