// One-sided Jacobi SVD of the K x K triangular factor of the QR route: the small problem
// min || R_A c - Q^T y || is solved through the singular values of R_A with the reference's cut-off
// (scipy.linalg.lstsq(cond=1e-13), lreg/lreg.py:335), so the factorisation has to deliver small singular values to
// high relative accuracy -- which is what Hestenes' method does.
//
// Storage: G = A^T row-major (row j of G = column j of A, contiguous) and V^T likewise, so a plane rotation of two
// columns touches two contiguous rows of each.  One kernel launch per round of a round-robin tournament (n / 2 disjoint
// pairs, one CTA per pair: three dot products by block reduction, the rotation that makes the two columns orthogonal,
// both row pairs updated); n - 1 rounds make a sweep that visits every pair once.  The largest |cos angle| met in a
// sweep is kept on the device; the host stops when it falls below the tolerance.  At the end the rows of G are the
// columns of U Sigma: sigma_j = || row j ||.  Both matrices stay L2-resident for the sizes of the path (K ~ 2000: 64 MB).
#include "pcq_internal.cuh"

#include <algorithm>
#include <cmath>

namespace {

__device__ __forceinline__ double block_sum3(double& a, double& b, double& c, double (*sh)[3]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { sh[warp][0] = a; sh[warp][1] = b; sh[warp][2] = c; }
    __syncthreads();
    a = b = c = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[w][0]; b += sh[w][1]; c += sh[w][2]; }
    return a;
}

// round r of the tournament on np players (np even, player np - 1 fixed): pair i of np / 2
__device__ __forceinline__ void pair_of(int r, int i, int np, int& p, int& q) {
    const int m = np - 1;
    if (i == 0) { p = m; q = r; }
    else { p = (r + i) % m; q = (r - i + m) % m; }
    if (p > q) { const int t = p; p = q; q = t; }
}

__global__ void __launch_bounds__(256) pcq_jacobi_round_kernel(double* __restrict__ g, double* __restrict__ vt, int n, int np,
                                                               int round, double tol, unsigned long long* __restrict__ offmax) {
    __shared__ double sh[8][3];
    __shared__ double s_cs[2];
    int p, q;
    pair_of(round, blockIdx.x, np, p, q);
    if (q >= n) return;                                   // the padding player of an odd n
    double* gp = g + (size_t)p * n;
    double* gq = g + (size_t)q * n;
    double alpha = 0.0, beta = 0.0, gamma = 0.0;
    for (int k = threadIdx.x; k < n; k += 256) {
        const double x = gp[k], y = gq[k];
        alpha = fma(x, x, alpha); beta = fma(y, y, beta); gamma = fma(x, y, gamma);
    }
    block_sum3(alpha, beta, gamma, sh);
    if (threadIdx.x == 0) {
        double c = 1.0, s = 0.0;
        const double denom = sqrt(alpha) * sqrt(beta);
        if (denom > 0.0 && fabs(gamma) > tol * denom) {
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            c = 1.0 / sqrt(1.0 + t * t);
            s = c * t;
            atomicMax(offmax, (unsigned long long)__double_as_longlong(fabs(gamma) / denom));     // positive doubles order like integers
        }
        s_cs[0] = c; s_cs[1] = s;
    }
    __syncthreads();
    const double c = s_cs[0], s = s_cs[1];
    if (s == 0.0) return;
    double* vp = vt + (size_t)p * n;
    double* vq = vt + (size_t)q * n;
    for (int k = threadIdx.x; k < n; k += 256) {
        const double x = gp[k], y = gq[k];
        gp[k] = c * x - s * y;
        gq[k] = s * x + c * y;
        const double u = vp[k], w = vq[k];
        vp[k] = c * u - s * w;
        vq[k] = s * u + c * w;
    }
}

__global__ void __launch_bounds__(256) pcq_jacobi_norms_kernel(const double* __restrict__ g, int n, double* __restrict__ sigma) {
    __shared__ double sh[8][3];
    const double* row = g + (size_t)blockIdx.x * n;
    double a = 0.0, b = 0.0, c = 0.0;
    for (int k = threadIdx.x; k < n; k += 256) a = fma(row[k], row[k], a);
    block_sum3(a, b, c, sh);
    if (threadIdx.x == 0) sigma[blockIdx.x] = sqrt(a);
}

__global__ void __launch_bounds__(256) pcq_jacobi_eye_kernel(double* __restrict__ vt, int n) {
    const int64_t total = (int64_t)n * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        vt[e] = (e / n == e % n) ? 1.0 : 0.0;
}

}  // namespace

extern "C" {

// In: g = A^T (n x n, row-major, contiguous).  Out: g = (U Sigma)^T (row j = sigma_j u_j), vt = V^T, sigma (n, unsorted),
// *sweeps = sweeps used.  A = U Sigma V^T.  Synchronises the stream once per sweep (convergence test on the host).
int pcq_svd_jacobi_dev(int device, double* g_dev, int n, double* vt_dev, double* sigma_dev, int max_sweeps, double tol,
                       int* sweeps, void* stream) {
    if (device < 0 || device >= 64 || !g_dev || !vt_dev || !sigma_dev || n < 1 || max_sweeps < 1 || !(tol > 0.0)) {
        pcq_set_error("svd_jacobi: bad argument");
        return PCQ_ERR_ARG;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
    PCQ_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0;
    PCQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    unsigned long long* d_off = nullptr;
    PCQ_CUDA(cudaMalloc((void**)&d_off, sizeof(unsigned long long)));
    struct Free { void* p; ~Free() { cudaFree(p); } } freer{d_off};
    pcq_jacobi_eye_kernel<<<sms * 4, 256, 0, st>>>(vt_dev, n);
    PCQ_LAUNCH_CHECK("pcq_jacobi_eye_kernel");
    const int np = n + (n & 1);
    int used = 0;
    if (n > 1) {
        for (int sw = 0; sw < max_sweeps; ++sw) {
            PCQ_CUDA(cudaMemsetAsync(d_off, 0, sizeof(unsigned long long), st));
            for (int r = 0; r < np - 1; ++r) {
                pcq_jacobi_round_kernel<<<np / 2, 256, 0, st>>>(g_dev, vt_dev, n, np, r, tol, d_off);
                PCQ_LAUNCH_CHECK("pcq_jacobi_round_kernel");
            }
            unsigned long long bits = 0;
            PCQ_CUDA(cudaMemcpyAsync(&bits, d_off, sizeof(bits), cudaMemcpyDeviceToHost, st));
            PCQ_CUDA(cudaStreamSynchronize(st));
            used = sw + 1;
            if (bits == 0ull) break;                      // no pair needed a rotation: converged
        }
    }
    pcq_jacobi_norms_kernel<<<n, 256, 0, st>>>(g_dev, n, sigma_dev);
    PCQ_LAUNCH_CHECK("pcq_jacobi_norms_kernel");
    if (sweeps) *sweeps = used;
    return PCQ_OK;
}

}  // extern "C"
