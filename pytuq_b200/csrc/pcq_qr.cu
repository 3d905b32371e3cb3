// QR statistic of a tall row-major matrix on the device: R of  Z = Q R  by blocked Householder
// reflections, hand-written -- the sample-axis reduction of the least-squares route that must follow the reference's
// SVD solve (scipy.linalg.lstsq(cond=1e-13), lreg/lreg.py:335) on ill-conditioned / square / rank-deficient systems,
// where the Gram form (K3) squares the condition number.
//
//   for every panel of NB = 32 columns:
//     qr_panel_kernel   (cooperative, one thread per row, the row's 32 panel values in registers): column by column --
//                       partial dot products of the pivot column with the panel, ONE grid-wide barrier, every CTA
//                       forms the reflector (beta, tau, w) from the partials in the same fixed order, updates its rows;
//                       the Gram of the reflectors needed for T rides on the same barriers.  Builds T (compact WY).
//     qr_vtc_kernel     W_part = V^T C over row ranges (DMMA m8n8k4, operands staged in shared memory)
//     qr_w2_kernel      W2 = T^T (sum of the parts, fixed order)
//     qr_update_kernel  C -= V W2   (DMMA, 64 x 128 tiles, k = 32), in place
//   qr_copy_r_kernel    upper triangle -> R.
// All reductions have a fixed order: the factor is reproducible run to run.  Chunks and ranks are combined by
// factorising stacked triangles with the same routine (pytuq_b200/lreg/tsqr.py: balanced tree / butterfly).
#include "pcq_mma.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>
#include <mutex>

namespace cg = cooperative_groups;

namespace {

constexpr int QR_NB = 32;          // panel width
constexpr int QR_ROWS = 256;       // rows (= threads) per CTA of the panel kernel
constexpr int VS = QR_NB + 4;      // shared row stride of a V tile   (== 4 mod 16: conflict-free fragment loads)
constexpr int CS = 128 + 4;        // shared row stride of a 128-wide tile

struct QrWs {
    std::mutex mu;
    double* part = nullptr; size_t part_bytes = 0;     // [2][grid][NB] partial sums + [2][NB] pivot rows
    double* tmat = nullptr; size_t tmat_bytes = 0;     // T (NB x NB) + tau (NB)
    double* wpart = nullptr; size_t wpart_bytes = 0;   // [RS][NB][ncp]
    double* w2 = nullptr; size_t w2_bytes = 0;         // [NB][ncp]
};
QrWs g_qr_ws[64];

struct PanelParams {
    double* z; int64_t ldz; int64_t m; int j0; int jb;
    double* part;      // [2][gridDim.x][QR_NB]
    double* piv;       // [2][QR_NB]
    double* tmat;      // [QR_NB][QR_NB] then tau[QR_NB]
};

// sum over the CTA of QR_NB per-thread values (static register indexing); out[c] is written by thread c
__device__ __forceinline__ void block_sums(const double (&vals)[QR_NB], double (*sh)[QR_NB], double* out, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int c = 0; c < QR_NB; ++c) {
        double t = vals[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sh[warp][c] = t;
    }
    __syncthreads();
    if (tid < QR_NB) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < QR_ROWS / 32; ++w) t += sh[w][tid];
        out[tid] = t;
    }
}

__global__ void __launch_bounds__(QR_ROWS) qr_panel_kernel(PanelParams P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[QR_ROWS / 32][QR_NB];
    __shared__ double s_red[8][QR_NB];
    __shared__ double s_sum[QR_NB];         // [c >= jj]: dots of the pivot column jj;  [c < jj-1]: Gram column of reflector jj-1
    __shared__ double s_piv[QR_NB];
    __shared__ double s_w[QR_NB];
    __shared__ double s_sc[3];              // scale, beta, tau
    __shared__ double s_tau[QR_NB];
    __shared__ double s_T[QR_NB][QR_NB + 1];   // CTA 0: the compact-WY factor under construction
    static_assert(QR_ROWS == 8 * QR_NB, "the partial-sum reduction maps eight threads to each of the NB values");
    const int tid = threadIdx.x;
    const int jb = P.jb;
    const int64_t row = (int64_t)P.j0 + (int64_t)blockIdx.x * QR_ROWS + tid;
    const bool live = row < P.m;
    const int rloc = (int)(row - P.j0);     // row index inside the panel block (0 = first diagonal row)
    double a[QR_NB];
#pragma unroll
    for (int c = 0; c < QR_NB; ++c) a[c] = (live && c < jb) ? P.z[row * P.ldz + P.j0 + c] : 0.0;
    if (blockIdx.x == 0)
        for (int e = tid; e < QR_NB * (QR_NB + 1); e += QR_ROWS) (&s_T[0][0])[e] = 0.0;

    for (int jj = 0; jj <= jb; ++jj) {
        const int par = jj & 1;
        // (the row lives in registers: a[jj] with a run-time jj is a select over the unrolled index, never an address)
        double ajj = 0.0, ajm1 = 0.0;
#pragma unroll
        for (int c = 0; c < QR_NB; ++c) {
            if (c == jj) ajj = a[c];
            if (c == jj - 1) ajm1 = a[c];
        }
        // per-thread contributions, one value per panel column c:
        //   c >= jj    dot of column jj below the diagonal with column c              (reflector jj)
        //   c <  jj-1  product of the finished reflector jj-1 with reflector c        (Gram column for T)
        double vals[QR_NB];
#pragma unroll
        for (int c = 0; c < QR_NB; ++c) {
            double v = 0.0;
            if (c >= jj) {
                if (jj < jb && c < jb && live && rloc > jj) v = ajj * a[c];
            } else if (c < jj - 1 && live && rloc >= jj - 1) {
                const double vj = (rloc == jj - 1) ? 1.0 : ajm1;                      // reflector jj-1: unit diagonal
                const double vc = (rloc == c) ? 1.0 : (rloc > c ? a[c] : 0.0);
                v = vj * vc;
            }
            vals[c] = v;
        }
        double* mine = P.part + ((size_t)par * gridDim.x + blockIdx.x) * QR_NB;
        block_sums(vals, sh, mine, tid);
        // the pivot row's current values (owned by one thread) for everybody
        if (jj < jb && live && rloc == jj) {
#pragma unroll
            for (int c = 0; c < QR_NB; ++c) P.piv[par * QR_NB + c] = a[c];
        }
        __threadfence();
        grid.sync();
        // every CTA reduces the partials in a fixed order (identical result everywhere): eight threads per value, each
        // over every eighth block with four loads in flight, then the eight partial sums in order
        {
            const int v = tid & (QR_NB - 1), q = tid >> 5;
            const double* src = P.part + (size_t)par * gridDim.x * QR_NB + v;
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            unsigned b = q;
            for (; b + 24 < gridDim.x; b += 32) {
                t0 += src[(size_t)b * QR_NB];         t1 += src[(size_t)(b + 8) * QR_NB];
                t2 += src[(size_t)(b + 16) * QR_NB];  t3 += src[(size_t)(b + 24) * QR_NB];
            }
            for (; b < gridDim.x; b += 8) t0 += src[(size_t)b * QR_NB];
            s_red[q][v] = (t0 + t1) + (t2 + t3);
            if (tid < QR_NB) s_piv[tid] = (jj < jb) ? P.piv[par * QR_NB + tid] : 0.0;
        }
        __syncthreads();
        if (tid < QR_NB)
            s_sum[tid] = (((s_red[0][tid] + s_red[1][tid]) + (s_red[2][tid] + s_red[3][tid])) +
                          ((s_red[4][tid] + s_red[5][tid]) + (s_red[6][tid] + s_red[7][tid])));
        __syncthreads();
        // T column jj-1 (CTA 0, thread r = row r): T[0:cc, cc] = -tau_cc * T[0:cc, 0:cc] * (V[:,0:cc]^T v_cc),  T[cc][cc] = tau_cc
        if (blockIdx.x == 0 && jj >= 1 && tid < QR_NB) {
            const int cc = jj - 1, r = tid;
            const double tj = s_tau[cc];
            if (r < cc) {
                double acc = 0.0;
                for (int k = r; k < cc; ++k) acc += s_T[r][k] * s_sum[k];
                s_T[r][cc] = -tj * acc;
            } else if (r == cc) {
                s_T[r][cc] = tj;
            }
        }
        if (jj == jb) break;
        // reflector jj from (alpha, sigma): H = I - tau v v^T, v = [1; x / (alpha - beta)]
        if (tid == 0) {
            const double alpha = s_piv[jj];
            const double sigma = s_sum[jj];
            double beta = alpha, t = 0.0, scale = 0.0;
            if (sigma > 0.0) {
                beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
                t = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            s_sc[0] = scale; s_sc[1] = beta; s_sc[2] = t;
            s_tau[jj] = t;
        }
        __syncthreads();
        if (tid < QR_NB) s_w[tid] = (tid > jj && tid < jb) ? s_sc[2] * (s_piv[tid] + s_sc[0] * s_sum[tid]) : 0.0;
        __syncthreads();
        if (live && rloc >= jj) {
            const bool pivot = (rloc == jj);
            const double v = pivot ? 1.0 : ajj * s_sc[0];
            const double diag = pivot ? s_sc[1] : v;
#pragma unroll
            for (int c = 0; c < QR_NB; ++c) {
                if (c == jj) a[c] = diag;
                else if (c > jj) a[c] -= v * s_w[c];
            }
        }
        __syncthreads();
    }
    if (live) {
#pragma unroll
        for (int c = 0; c < QR_NB; ++c) if (c < jb) P.z[row * P.ldz + P.j0 + c] = a[c];
    }
    if (blockIdx.x == 0) {       // the finished T (upper triangular, zero elsewhere) for the trailing update
        __syncthreads();
        for (int e = tid; e < QR_NB * QR_NB; e += QR_ROWS) P.tmat[e] = s_T[e / QR_NB][e % QR_NB];
    }
}

// element (r, c) of V for panel rows r = 0.. (r relative to j0): unit lower trapezoid stored below the diagonal of z
__device__ __forceinline__ double v_elem(const double* z, int64_t ldz, int j0, int jb, int64_t m, int64_t r, int c) {
    const int64_t row = (int64_t)j0 + r;
    if (row >= m || c >= jb || r < c) return 0.0;
    if (r == c) return 1.0;
    return z[row * ldz + j0 + c];
}

struct TrailParams {
    double* z; int64_t ldz; int64_t m; int j0; int jb; int c0; int nc; int ncp;
    int rs; int64_t rows_per_split;
    double* wpart; double* w2; const double* tmat;
};

// W_part[split][k][n] = sum over the split's rows of V[r][k] * C[r][n]
__global__ void __launch_bounds__(256) qr_vtc_kernel(TrailParams P) {
    __shared__ __align__(16) double Vt[32 * VS];
    __shared__ __align__(16) double Ct[32 * CS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * 128;
    const int split = blockIdx.y;
    const int64_t rbeg = (int64_t)split * P.rows_per_split;
    const int64_t rend = min(rbeg + P.rows_per_split, P.m - P.j0);
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int g = lane >> 2, t4 = lane & 3;
    for (int64_t r0 = rbeg; r0 < rend; r0 += 32) {
        for (int e = tid; e < 32 * 32; e += 256) {
            const int rr = e >> 5, cc = e & 31;
            Vt[rr * VS + cc] = (r0 + rr < rend) ? v_elem(P.z, P.ldz, P.j0, P.jb, P.m, r0 + rr, cc) : 0.0;
        }
        for (int e = tid; e < 32 * 128; e += 256) {
            const int rr = e >> 7, cc = e & 127;
            const int64_t row = (int64_t)P.j0 + r0 + rr;
            Ct[rr * CS + cc] = (r0 + rr < rend && n0 + cc < P.nc) ? P.z[row * P.ldz + P.c0 + n0 + cc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 4) {
            double bfr[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) bfr[j] = Ct[(k0 + t4) * CS + warp * 16 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double afr = Vt[(k0 + t4) * VS + i * 8 + g];       // A = V^T: A[mi][k] = V[k][mi]
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(acc[i][j][0], acc[i][j][1], afr, bfr[j]);
            }
        }
        __syncthreads();
    }
    double* out = P.wpart + (size_t)split * QR_NB * P.ncp;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int k = i * 8 + g, n = n0 + warp * 16 + j * 8 + 2 * t4;
            out[(size_t)k * P.ncp + n] = acc[i][j][0];
            out[(size_t)k * P.ncp + n + 1] = acc[i][j][1];
        }
}

// W2[i][n] = sum_k T[k][i] * (sum over splits, in order, of W_part[split][k][n]); a CTA takes 8 columns, thread = (k, n)
__global__ void __launch_bounds__(256) qr_w2_kernel(TrailParams P) {
    __shared__ double Ts[QR_NB * QR_NB];
    __shared__ double Ws[QR_NB][8];
    for (int e = threadIdx.x; e < QR_NB * QR_NB; e += 256) Ts[e] = P.tmat[e];
    const int nl = threadIdx.x & 7, k = threadIdx.x >> 3;
    const int n = blockIdx.x * 8 + nl;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    const double* src = P.wpart + (size_t)k * P.ncp + n;
    const size_t stride = (size_t)QR_NB * P.ncp;
    int sp = 0;
    for (; sp + 3 < P.rs; sp += 4) {
        t0 += src[(size_t)sp * stride]; t1 += src[(size_t)(sp + 1) * stride];
        t2 += src[(size_t)(sp + 2) * stride]; t3 += src[(size_t)(sp + 3) * stride];
    }
    for (; sp < P.rs; ++sp) t0 += src[(size_t)sp * stride];
    Ws[k][nl] = (t0 + t1) + (t2 + t3);
    __syncthreads();
    const int i = k;
    double acc = 0.0;
#pragma unroll
    for (int kk = 0; kk < QR_NB; ++kk) if (kk <= i) acc += Ts[kk * QR_NB + i] * Ws[kk][nl];       // T upper triangular
    P.w2[(size_t)i * P.ncp + n] = (i < P.jb) ? acc : 0.0;
}

// C[r][n] -= sum_k V[r][k] W2[k][n]   on a 64 x 128 tile (8 warps as 2 x 4, warp tile 32 x 32: 32 accumulators, so that
// several CTAs are resident per SM and one CTA's loads overlap another's read-modify-write of C)
constexpr int UPD_ROWS = 64;
__global__ void __launch_bounds__(256) qr_update_kernel(TrailParams P) {
    extern __shared__ __align__(16) double smem_qr[];
    double* Vt = smem_qr;                       // [UPD_ROWS][VS]
    double* Wt = smem_qr + UPD_ROWS * VS;       // [32][CS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * 128;
    const int64_t r0 = (int64_t)blockIdx.y * UPD_ROWS;
    for (int e = tid; e < UPD_ROWS * 32; e += 256) {
        const int rr = e >> 5, cc = e & 31;
        Vt[rr * VS + cc] = v_elem(P.z, P.ldz, P.j0, P.jb, P.m, r0 + rr, cc);
    }
    for (int e = tid; e < 32 * 128; e += 256) {
        const int kk = e >> 7, cc = e & 127;
        Wt[kk * CS + cc] = P.w2[(size_t)kk * P.ncp + n0 + cc];
    }
    __syncthreads();
    const int wm = warp >> 2, wn = warp & 3;          // warp tile: rows wm*32.., cols wn*32..
    const int g = lane >> 2, t4 = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 4) {
        double afr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) afr[i] = Vt[(wm * 32 + i * 8 + g) * VS + k0 + t4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double bfr = Wt[(k0 + t4) * CS + wn * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(acc[i][j][0], acc[i][j][1], afr[i], bfr);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = (int64_t)P.j0 + r0 + wm * 32 + i * 8 + g;
        if (row >= P.m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + wn * 32 + j * 8 + 2 * t4;
            double* dst = P.z + row * P.ldz + P.c0 + n;
            if (n < P.nc) dst[0] -= acc[i][j][0];
            if (n + 1 < P.nc) dst[1] -= acc[i][j][1];
        }
    }
}

__global__ void __launch_bounds__(256) qr_copy_r_kernel(const double* __restrict__ z, int64_t ldz, int64_t m, int L,
                                                        double* __restrict__ r, int64_t ldr) {
    const int64_t total = (int64_t)L * L;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / L), k = (int)(e - (int64_t)i * L);
        r[(size_t)i * ldr + k] = (k >= i && i < m) ? z[(size_t)i * ldz + k] : 0.0;
    }
}

}  // namespace

extern "C" {

// largest number of rows one pcq_qr_r_dev call accepts on this device (the panel kernel is cooperative: all its
// CTAs must be resident at once)
int64_t pcq_qr_max_rows(int device) {
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) { pcq_set_error("qr: cudaSetDevice(%d) failed", device); return PCQ_ERR_NODEV; }
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qr_panel_kernel, QR_ROWS, 0);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) { pcq_set_error("qr: occupancy query failed: %s", cudaGetErrorString(e)); return PCQ_ERR_CUDA; }
    return (int64_t)per_sm * sms * QR_ROWS;
}

// R factor (L x L upper triangular, row-major, leading dimension ldr) of the m x L row-major matrix z (leading
// dimension ldz), which is overwritten with the Householder data.  m may be smaller than L (rows >= m of R are zero).
int pcq_qr_r_dev(int device, double* z, int64_t m, int64_t ldz, int L, double* r, int64_t ldr, void* stream) {
    if (device < 0 || device >= 64 || !z || !r || m < 1 || L < 1 || ldz < L || ldr < L) {
        pcq_set_error("qr_r: bad argument");
        return PCQ_ERR_ARG;
    }
    const int64_t cap = pcq_qr_max_rows(device);
    if (cap < 0) return (int)cap;
    if (m > cap) { pcq_set_error("qr_r: %lld rows exceed the %lld one call can take (chunk the rows)", (long long)m, (long long)cap); return PCQ_ERR_ARG; }
    int prev = -1;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
    PCQ_CUDA(cudaSetDevice(device));
    int sms = 0;
    PCQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t st = (cudaStream_t)stream;
    QrWs& W = g_qr_ws[device];
    std::lock_guard<std::mutex> lock(W.mu);
    const int maxgrid = (int)((m + QR_ROWS - 1) / QR_ROWS);
    const int ncp_max = (L + 127) / 128 * 128;
    const int RS = 64;
    int rc;
    if ((rc = pcq_ensure_dev(&W.part, &W.part_bytes, ((size_t)2 * maxgrid * QR_NB + 2 * QR_NB) * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&W.tmat, &W.tmat_bytes, (size_t)(QR_NB * QR_NB + QR_NB) * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&W.wpart, &W.wpart_bytes, (size_t)RS * QR_NB * ncp_max * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&W.w2, &W.w2_bytes, (size_t)QR_NB * ncp_max * 8)) != PCQ_OK) return rc;
    PCQ_CUDA(cudaMemsetAsync(W.part, 0, ((size_t)2 * maxgrid * QR_NB + 2 * QR_NB) * 8, st));
    PCQ_CUDA(cudaMemsetAsync(W.tmat, 0, (size_t)(QR_NB * QR_NB + QR_NB) * 8, st));
    const size_t upd_smem = (size_t)(UPD_ROWS * VS + 32 * CS) * sizeof(double);
    PCQ_CUDA(cudaFuncSetAttribute(qr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
    const int ncols = (int)std::min<int64_t>(L, m);         // columns that have a row to pivot on
    for (int j0 = 0; j0 < ncols; j0 += QR_NB) {
        const int jb = std::min(QR_NB, L - j0);
        PanelParams P;
        P.z = z; P.ldz = ldz; P.m = m; P.j0 = j0; P.jb = jb;
        P.part = W.part; P.piv = W.part + (size_t)2 * maxgrid * QR_NB; P.tmat = W.tmat;
        const int grid = (int)((m - j0 + QR_ROWS - 1) / QR_ROWS);
        void* args[] = {(void*)&P};
        PCQ_CUDA(cudaLaunchCooperativeKernel((const void*)qr_panel_kernel, dim3(grid), dim3(QR_ROWS), args, 0, st));
        g_pcq_launches.fetch_add(1, std::memory_order_relaxed);
        const int c0 = j0 + jb, nc = L - c0;
        if (nc <= 0) continue;
        TrailParams T;
        T.z = z; T.ldz = ldz; T.m = m; T.j0 = j0; T.jb = jb; T.c0 = c0; T.nc = nc; T.ncp = (nc + 127) / 128 * 128;
        const int64_t rows = m - j0;
        int rs = (int)std::min<int64_t>(RS, (rows + 127) / 128);
        if (rs < 1) rs = 1;
        T.rows_per_split = ((rows + rs - 1) / rs + 31) / 32 * 32;
        rs = (int)((rows + T.rows_per_split - 1) / T.rows_per_split);
        T.rs = rs; T.wpart = W.wpart; T.w2 = W.w2; T.tmat = W.tmat;
        const int ctiles = T.ncp / 128;
        qr_vtc_kernel<<<dim3(ctiles, rs), 256, 0, st>>>(T);
        PCQ_LAUNCH_CHECK("qr_vtc_kernel");
        qr_w2_kernel<<<T.ncp / 8, 256, 0, st>>>(T);
        PCQ_LAUNCH_CHECK("qr_w2_kernel");
        qr_update_kernel<<<dim3(ctiles, (unsigned)((rows + UPD_ROWS - 1) / UPD_ROWS)), 256, upd_smem, st>>>(T);
        PCQ_LAUNCH_CHECK("qr_update_kernel");
    }
    qr_copy_r_kernel<<<sms * 4, 256, 0, st>>>(z, ldz, m, L, r, ldr);
    PCQ_LAUNCH_CHECK("qr_copy_r_kernel");
    return PCQ_OK;
}

}  // extern "C"
