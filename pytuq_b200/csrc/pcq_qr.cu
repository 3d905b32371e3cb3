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
//     qr_update_kernel  C -= V W2   (DMMA, 128 x 128 tiles, k = 32), in place
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
    double* part = nullptr; size_t part_bytes = 0;     // [2][grid][2*NB] partial sums + [2][NB] pivot rows
    double* tmat = nullptr; size_t tmat_bytes = 0;     // T (NB x NB) + tau (NB)
    double* wpart = nullptr; size_t wpart_bytes = 0;   // [RS][NB][ncp]
    double* w2 = nullptr; size_t w2_bytes = 0;         // [NB][ncp]
};
QrWs g_qr_ws[64];

struct PanelParams {
    double* z; int64_t ldz; int64_t m; int j0; int jb;
    double* part;      // [2][gridDim.x][2*QR_NB]
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
    __shared__ double s_sum[2 * QR_NB];     // [0..NB): dots of the pivot column, [NB..2NB): Gram column of the reflectors
    __shared__ double s_red[4][2 * QR_NB];
    __shared__ double s_w[QR_NB];
    __shared__ double s_sc[2];              // scale, beta
    static_assert(QR_ROWS == 4 * 2 * QR_NB, "the partial-sum reduction maps four threads to each of the 2*NB values");
    const int tid = threadIdx.x;
    const int jb = P.jb;
    const int64_t row = (int64_t)P.j0 + (int64_t)blockIdx.x * QR_ROWS + tid;
    const bool live = row < P.m;
    const int rloc = (int)(row - P.j0);     // row index inside the panel block (0 = first diagonal row)
    double a[QR_NB];
#pragma unroll
    for (int c = 0; c < QR_NB; ++c) a[c] = (live && c < jb) ? P.z[row * P.ldz + P.j0 + c] : 0.0;
    double* tau = P.tmat + QR_NB * QR_NB;

    for (int jj = 0; jj <= jb; ++jj) {
        // per-thread contributions: (a) dots of column jj below the diagonal with columns jj..jb-1 (reflector jj),
        // (b) Gram column of the finished reflector jj-1 with reflectors 0..jj-2 (for T)
        const int par = jj & 1;
        // (the row lives in registers: a[jj] with a run-time jj is a select over the unrolled index, never an address)
        double ajj = 0.0, ajm1 = 0.0;
#pragma unroll
        for (int c = 0; c < QR_NB; ++c) {
            if (c == jj) ajj = a[c];
            if (c == jj - 1) ajm1 = a[c];
        }
        double pv[QR_NB], qv[QR_NB];
#pragma unroll
        for (int c = 0; c < QR_NB; ++c) {
            double p = 0.0;
            if (jj < jb && c >= jj && c < jb && live && rloc > jj) p = ajj * a[c];
            pv[c] = p;
            double q = 0.0;
            if (jj >= 1 && c < jj - 1 && live && rloc >= jj - 1) {
                const double vj = (rloc == jj - 1) ? 1.0 : ajm1;                      // reflector jj-1: unit diagonal
                const double vc = (rloc == c) ? 1.0 : (rloc > c ? a[c] : 0.0);
                q = vj * vc;
            }
            qv[c] = q;
        }
        // block sums -> this CTA's slot of the partial buffer (two halves)
        double* mine = P.part + ((size_t)par * gridDim.x + blockIdx.x) * (2 * QR_NB);
        block_sums(pv, sh, mine, tid);
        __syncthreads();
        block_sums(qv, sh, mine + QR_NB, tid);
        // the pivot row's current values (owned by one thread) for everybody
        if (jj < jb && live && rloc == jj) {
#pragma unroll
            for (int c = 0; c < QR_NB; ++c) P.piv[par * QR_NB + c] = a[c];
        }
        __threadfence();
        grid.sync();
        // every CTA reduces the partials in a fixed order (identical result everywhere): four threads per value, each
        // over every fourth block, then the four partial sums in order
        {
            const int v = tid & (2 * QR_NB - 1), q = tid >> 6;       // QR_ROWS = 4 * 2 * QR_NB
            double t = 0.0;
            for (unsigned b = q; b < gridDim.x; b += 4) t += P.part[((size_t)par * gridDim.x + b) * (2 * QR_NB) + v];
            s_red[q][v] = t;
        }
        __syncthreads();
        if (tid < 2 * QR_NB) s_sum[tid] = ((s_red[0][tid] + s_red[1][tid]) + s_red[2][tid]) + s_red[3][tid];
        __syncthreads();
        // T column jj-1 (CTA 0): T[0:jj-1, jj-1] = -tau * T[0:jj-1, 0:jj-1] * (V[:,0:jj-1]^T v_{jj-1}),  T[jj-1][jj-1] = tau
        if (blockIdx.x == 0 && jj >= 1 && tid == 0) {
            const int cc = jj - 1;
            const double tj = tau[cc];
            for (int r = 0; r < cc; ++r) {
                double acc = 0.0;
                for (int k = r; k < cc; ++k) acc += P.tmat[r * QR_NB + k] * s_sum[QR_NB + k];
                P.tmat[r * QR_NB + cc] = -tj * acc;
            }
            P.tmat[cc * QR_NB + cc] = tj;
            for (int r = cc + 1; r < QR_NB; ++r) P.tmat[r * QR_NB + cc] = 0.0;
        }
        if (jj == jb) break;
        // reflector jj from (alpha, sigma): H = I - tau v v^T, v = [1; x / (alpha - beta)]
        if (tid == 0) {
            const double alpha = P.piv[par * QR_NB + jj];
            const double sigma = s_sum[jj];
            double beta = alpha, t = 0.0, scale = 0.0;
            if (sigma > 0.0) {
                beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
                t = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            s_sc[0] = scale; s_sc[1] = beta;
            if (blockIdx.x == 0) tau[jj] = t;
            for (int c = 0; c < QR_NB; ++c)
                s_w[c] = (c > jj && c < jb) ? t * (P.piv[par * QR_NB + c] + scale * s_sum[c]) : 0.0;
        }
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

// W2[i][n] = sum_k T[k][i] * (sum over splits, in order, of W_part[split][k][n])
__global__ void __launch_bounds__(128) qr_w2_kernel(TrailParams P) {
    __shared__ double Ts[QR_NB * QR_NB];
    for (int e = threadIdx.x; e < QR_NB * QR_NB; e += 128) Ts[e] = P.tmat[e];
    __syncthreads();
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= P.ncp) return;
    double w[QR_NB];
#pragma unroll
    for (int k = 0; k < QR_NB; ++k) {
        double t = 0.0;
        for (int s = 0; s < P.rs; ++s) t += P.wpart[((size_t)s * QR_NB + k) * P.ncp + n];
        w[k] = t;
    }
#pragma unroll
    for (int i = 0; i < QR_NB; ++i) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < QR_NB; ++k) if (k <= i) t += Ts[k * QR_NB + i] * w[k];      // T upper triangular
        P.w2[(size_t)i * P.ncp + n] = (i < P.jb) ? t : 0.0;
    }
}

// C[r][n] -= sum_k V[r][k] W2[k][n]   on a 128 x 128 tile
__global__ void __launch_bounds__(256) qr_update_kernel(TrailParams P) {
    extern __shared__ __align__(16) double smem_qr[];
    double* Vt = smem_qr;                  // [128][VS]
    double* Wt = smem_qr + 128 * VS;       // [32][CS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * 128;
    const int64_t r0 = (int64_t)blockIdx.y * 128;
    for (int e = tid; e < 128 * 32; e += 256) {
        const int rr = e >> 5, cc = e & 31;
        Vt[rr * VS + cc] = v_elem(P.z, P.ldz, P.j0, P.jb, P.m, r0 + rr, cc);
    }
    for (int e = tid; e < 32 * 128; e += 256) {
        const int kk = e >> 7, cc = e & 127;
        Wt[kk * CS + cc] = P.w2[(size_t)kk * P.ncp + n0 + cc];
    }
    __syncthreads();
    const int wm = warp >> 1, wn = warp & 1;          // warp tile: rows wm*32.., cols wn*64..
    const int g = lane >> 2, t4 = lane & 3;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 4) {
        double afr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) afr[i] = Vt[(wm * 32 + i * 8 + g) * VS + k0 + t4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double bfr = Wt[(k0 + t4) * CS + wn * 64 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(acc[i][j][0], acc[i][j][1], afr[i], bfr);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = (int64_t)P.j0 + r0 + wm * 32 + i * 8 + g;
        if (row >= P.m) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + wn * 64 + j * 8 + 2 * t4;
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
    const int RS = 16;
    int rc;
    if ((rc = pcq_ensure_dev(&W.part, &W.part_bytes, ((size_t)2 * maxgrid * 2 * QR_NB + 2 * QR_NB) * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&W.tmat, &W.tmat_bytes, (size_t)(QR_NB * QR_NB + QR_NB) * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&W.wpart, &W.wpart_bytes, (size_t)RS * QR_NB * ncp_max * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&W.w2, &W.w2_bytes, (size_t)QR_NB * ncp_max * 8)) != PCQ_OK) return rc;
    PCQ_CUDA(cudaMemsetAsync(W.part, 0, ((size_t)2 * maxgrid * 2 * QR_NB + 2 * QR_NB) * 8, st));
    PCQ_CUDA(cudaMemsetAsync(W.tmat, 0, (size_t)(QR_NB * QR_NB + QR_NB) * 8, st));
    const size_t upd_smem = (size_t)(128 * VS + 32 * CS) * sizeof(double);
    PCQ_CUDA(cudaFuncSetAttribute(qr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
    const int ncols = (int)std::min<int64_t>(L, m);         // columns that have a row to pivot on
    for (int j0 = 0; j0 < ncols; j0 += QR_NB) {
        const int jb = std::min(QR_NB, L - j0);
        PanelParams P;
        P.z = z; P.ldz = ldz; P.m = m; P.j0 = j0; P.jb = jb;
        P.part = W.part; P.piv = W.part + (size_t)2 * maxgrid * 2 * QR_NB; P.tmat = W.tmat;
        const int grid = (int)((m - j0 + QR_ROWS - 1) / QR_ROWS);
        void* args[] = {(void*)&P};
        PCQ_CUDA(cudaLaunchCooperativeKernel((const void*)qr_panel_kernel, dim3(grid), dim3(QR_ROWS), args, 0, st));
        g_pcq_launches.fetch_add(1, std::memory_order_relaxed);
        const int c0 = j0 + jb, nc = L - c0;
        if (nc <= 0) continue;
        TrailParams T;
        T.z = z; T.ldz = ldz; T.m = m; T.j0 = j0; T.jb = jb; T.c0 = c0; T.nc = nc; T.ncp = (nc + 127) / 128 * 128;
        const int64_t rows = m - j0;
        int rs = (int)std::min<int64_t>(RS, (rows + 255) / 256);
        if (rs < 1) rs = 1;
        T.rows_per_split = ((rows + rs - 1) / rs + 31) / 32 * 32;
        rs = (int)((rows + T.rows_per_split - 1) / T.rows_per_split);
        T.rs = rs; T.wpart = W.wpart; T.w2 = W.w2; T.tmat = W.tmat;
        const int ctiles = T.ncp / 128;
        qr_vtc_kernel<<<dim3(ctiles, rs), 256, 0, st>>>(T);
        PCQ_LAUNCH_CHECK("qr_vtc_kernel");
        qr_w2_kernel<<<(T.ncp + 127) / 128, 128, 0, st>>>(T);
        PCQ_LAUNCH_CHECK("qr_w2_kernel");
        qr_update_kernel<<<dim3(ctiles, (unsigned)((rows + 127) / 128)), 256, upd_smem, st>>>(T);
        PCQ_LAUNCH_CHECK("qr_update_kernel");
    }
    qr_copy_r_kernel<<<sms * 4, 256, 0, st>>>(z, ldz, m, L, r, ldr);
    PCQ_LAUNCH_CHECK("qr_copy_r_kernel");
    return PCQ_OK;
}

}  // extern "C"
