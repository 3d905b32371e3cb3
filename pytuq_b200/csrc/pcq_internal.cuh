// Internal declarations shared by the libpcq translation units (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <thread>
#include <string>
#include <vector>

#include "pcq.h"

struct pcq_spec;   // NVRTC-specialised evalPC kernels of one plan (pcq_spec.cu)

#define PCQ_MAXW_FAST 4   // terms with <= 4 non-zero orders use the packed 8-byte descriptor

// Slot numbering of the per-sample table of 1-D polynomial values:
//   slot 0                      = 1.0   (order 0 of every dimension, and descriptor padding)
//   slot 1 + (ord-1)*s + dim    = phi^{(dim)}_{ord}(x_dim),  1 <= ord <= pmax
// nslots = 1 + pmax*s.  Dimension is the fast index so that neighbouring terms of a
// total-order set (which differ in one dimension) gather from neighbouring banks.
struct pcq_plan {
    int device = 0;
    int num_sms = 0;
    int sdim = 0;        // s
    int nterms = 0;      // K
    int pmax = 0;        // highest 1-D order tabulated
    int nslots = 0;
    int width = 0;       // W = max number of non-zero orders in a term (>=1)
    int downward_closed = 0;
    size_t smem_optin = 0;

    // device-resident tables
    double* d_rec_a = nullptr;    // (pmax+1) x s
    double* d_rec_b = nullptr;    // (pmax+1) x s
    double* d_p1 = nullptr;       // s
    int* d_kinds = nullptr;       // s, germ kinds of the last propagate call
    // term descriptors: W slot indices per term: padding (slot 0) first, then ascending dimension
    unsigned long long* d_desc4 = nullptr;  // K packed 4 x uint16 (valid when width <= 4)
    uint16_t* d_descw = nullptr;            // K x width (always valid)
    uint8_t* d_flags = nullptr;             // K, bit 0: first W-1 slots differ from the previous term's

    // host copies kept for the (opt-in) multi-index-specialised kernels
    std::vector<int32_t> h_mi;
    std::vector<double> h_rec_a, h_rec_b, h_p1;
    // [0] exact, [1] fast; published with release stores (a background build may install them while
    // another host thread is launching kernels through this plan)
    std::atomic<pcq_spec*> spec[2]{};
    std::thread spec_thread;                  // background build started by pcq_plan_specialize_async
    std::atomic<int> spec_building{0};        // modes being built in the background (bit 0 exact, bit 1 fast)
    // Gram workspace (lazily grown)
    double* d_ws = nullptr;
    size_t ws_bytes = 0;
    // large germ dimension: slot tables of a sample chunk in global memory, and a design-matrix chunk
    // (one table per plan stream: the host-pointer entry points alternate their chunks between streams[0] and
    //  streams[1], and the table of chunk c must survive while chunk c+1 is being tabulated on the other stream)
    double* d_gtab[2] = {nullptr, nullptr};
    size_t gtab_bytes[2] = {0, 0};
    double* d_achunk = nullptr;
    size_t achunk_bytes = 0;
    // packed upper-triangular operand of the quadratic-form kernel (lazily grown)
    double* d_q = nullptr;
    size_t q_bytes = 0;
    // staging for the host-pointer entry points (lazily grown)
    double* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes[2] = {0, 0};
    double* d_out[2] = {nullptr, nullptr};
    size_t out_bytes[2] = {0, 0};
    double* h_pin[2] = {nullptr, nullptr};   // pinned D2H staging of the host-pointer K1
    size_t pin_bytes[2] = {0, 0};
    double* d_misc = nullptr;      // coefficients / S matrix for host-pointer calls
    size_t misc_bytes = 0;
    void* mom = nullptr;           // moment form of K3 (pcq_moment.cu): extended recurrences + per-output-count plans
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t events[2] = {nullptr, nullptr};
};

// error plumbing --------------------------------------------------------------
void pcq_set_error(const char* fmt, ...);
extern std::atomic<long long> g_pcq_launches;
extern std::atomic<int> g_pcq_host_thread_div;      // > 1 while a pcq_group call runs one host pipeline per device

#define PCQ_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) {                                                  \
            pcq_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                    \
            return PCQ_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

#define PCQ_LAUNCH_CHECK(name)                                                     \
    do {                                                                           \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) {                                                   \
            pcq_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return PCQ_ERR_CUDA;                                                   \
        }                                                                          \
        g_pcq_launches.fetch_add(1, std::memory_order_relaxed);                    \
    } while (0)

// grows *buf to at least `bytes` (device memory); returns pcq_status
int pcq_ensure_dev(double** buf, size_t* have, size_t bytes);

// kernel launchers (defined in the kernel translation units) -----------------------
int pcq_launch_bases(pcq_plan* p, const double* x, int64_t n, int64_t ldx, double* a,
                     int64_t lda, cudaStream_t st);
// large germ dimension (slot tables in global memory, pcq_basis.cu)
constexpr int64_t PCQ_GTAB_ROWS = 4096;
bool pcq_table_fits_smem(const pcq_plan* p);
int pcq_launch_gtab(pcq_plan* p, const double* x, int64_t n, int64_t ldx, int64_t nc, double** tab, cudaStream_t st);
int pcq_launch_evalpc(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf,
                      int m, double* y, int64_t ldy, int mode, cudaStream_t st);
// generator == plan (fused, p != nullptr) or materialised A (a != nullptr)
int pcq_launch_gram(pcq_plan* p, int num_sms, size_t smem_optin, const double* x, int64_t n,
                    int64_t ldx, const double* a, int64_t lda, int kcols, const double* y,
                    int64_t ldy, int m, double* s, int accumulate, double** ws,
                    size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st);
// pack + SYRK form of the same (pcq_syrk.cu); *handled = false when it does not apply
int pcq_launch_syrk(pcq_plan* p, int num_sms, size_t smem_optin, const double* x, int64_t n, int64_t ldx,
                    const double* a, int64_t lda, int kcols, const double* y, int64_t ldy, int m, double* s,
                    int accumulate, double** ws, size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes,
                    cudaStream_t st, bool* handled, bool rect = false);
// rect: only the output columns of S (A^T Y, Y^T Y, Y^T 1) -- the part the moment form leaves when it serves G alone
// moment form of the same statistics (pcq_moment.cu); *handled = false when it does not apply or would be slower
int pcq_launch_moments(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy, int m,
                       double* s, int accumulate, double** ws, size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes,
                       cudaStream_t st, bool* handled, int layout_m = -1);
// layout_m >= 0 with m == 0: G, A^T 1 and N only, written into an S laid out for layout_m outputs (L = K + layout_m + 1)
void pcq_mom_free(pcq_plan* p);
int pcq_launch_propagate(pcq_plan* p, uint64_t seed, int64_t first, int64_t n,
                         const int32_t* germ_kind, const double* cf, int m, double* sums,
                         double* y, int64_t ldy, int mode, cudaStream_t st);
int pcq_launch_qform(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* c, int64_t ldc, double* v,
                     int64_t ldv, double** bq, size_t* bq_bytes, double** ws, size_t* ws_bytes, double** zbuf,
                     size_t* zbuf_bytes, cudaStream_t st);
bool pcq_linforms_applicable(const pcq_plan* p, int m);
// the same contractions on a caller-supplied design matrix (no plan)
int pcq_launch_qform_matrix(int num_sms, size_t smem_optin, const double* a, int64_t n, int64_t lda, int K, const double* c,
                            int64_t ldc, double* v, int64_t ldv, double** bq, size_t* bq_bytes, double** ws,
                            size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st);
int pcq_launch_linforms_matrix(int num_sms, size_t smem_optin, const double* a, int64_t n, int64_t lda, int K,
                               const double* cf, int64_t ldcf, int m, double* y, int64_t ldy, double** bq, size_t* bq_bytes,
                               double** zbuf, size_t* zbuf_bytes, cudaStream_t st);
int pcq_launch_linforms(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf, int m, double* y,
                        int64_t ldy, double** bq, size_t* bq_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st);
int pcq_launch_residual(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy,
                        const double* cf, double* sums, int accumulate, cudaStream_t st);
int pcq_launch_germ(pcq_plan* p, uint64_t seed, int64_t first, int64_t n,
                    const int32_t* germ_kind, double* x, int64_t ldx, cudaStream_t st);
// multi-index-specialised evalPC (pcq_spec.cu)
int pcq_spec_build(pcq_plan* p, int fast);
int pcq_spec_launch(pcq_plan* p, int fast, const double* x, int64_t n, int64_t ldx, const double* cf,
                    double* y, int64_t ldy, cudaStream_t st);
void pcq_spec_free(pcq_spec* sp);

// device helpers -----------------------------------------------------------------------
#ifdef __CUDACC__
// One 1-D family evaluated at one point, written into a per-sample slot table whose
// consecutive slots are `stride` doubles apart (tab points at slot 0 of this sample).
// Rounding sequence of PC1d.__call__ (rv/pcrv.py:794-802):
//   p_{n+1} = ((a_n*x)*p_n) + (b_n*p_{n-1})
// i.e. four separately rounded operations, never contracted to FMA.  Slot 0 (=1.0) is
// written by the caller once per sample.
__device__ __forceinline__ void pcq_fill_1d(double xv, int dim, int sdim, int pmax,
                                            const double* __restrict__ rec_a,
                                            const double* __restrict__ rec_b,
                                            const double* __restrict__ p1, double* tab,
                                            int stride) {
    if (pmax >= 1) {
        double pm = 1.0;
        double pc = __dmul_rn(xv, p1[dim]);
        tab[(size_t)(1 + dim) * stride] = pc;
        for (int nn = 1; nn < pmax; ++nn) {
            const double a = rec_a[nn * sdim + dim];
            const double b = rec_b[nn * sdim + dim];
            const double pn = __dadd_rn(__dmul_rn(__dmul_rn(a, xv), pc), __dmul_rn(b, pm));
            tab[(size_t)(1 + nn * sdim + dim) * stride] = pn;
            pm = pc;
            pc = pn;
        }
    }
}

// product of the W tabulated factors of one term, left to right (rv/pcrv.py:436-438);
// the padding slots hold exactly 1.0 so the value equals the reference's full product.
template <int W>
__device__ __forceinline__ double pcq_term4(const double* __restrict__ tab, int stride,
                                            unsigned long long d) {
    double v = tab[(size_t)(d & 0xffffull) * stride];
    if (W >= 2) v = __dmul_rn(v, tab[(size_t)((d >> 16) & 0xffffull) * stride]);
    if (W >= 3) v = __dmul_rn(v, tab[(size_t)((d >> 32) & 0xffffull) * stride]);
    if (W >= 4) v = __dmul_rn(v, tab[(size_t)((d >> 48) & 0xffffull) * stride]);
    return v;
}
__device__ __forceinline__ double pcq_termw(const double* __restrict__ tab, int stride,
                                            const uint16_t* __restrict__ dw, int width) {
    double v = tab[(size_t)dw[0] * stride];
    for (int j = 1; j < width; ++j) v = __dmul_rn(v, tab[(size_t)dw[j] * stride]);
    return v;
}
#endif
