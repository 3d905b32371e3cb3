// C-ABI of libpcq (see include/pcq.h): plans, host-pointer entry points, error plumbing.
#include "pcq_internal.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <set>
#include <thread>

std::atomic<long long> g_pcq_launches{0};
std::atomic<int> g_pcq_host_thread_div{1};      // see pcq_host_threads()

namespace {
thread_local char g_err[512] = "";

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int device_props(int device, int* num_sms, size_t* smem_optin) {
    int v = 0;
    PCQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    *num_sms = v;
    PCQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    *smem_optin = (size_t)v;
    return PCQ_OK;
}

template <typename T>
int dev_upload(T** dst, const T* src, size_t count) {
    PCQ_CUDA(cudaMalloc((void**)dst, std::max<size_t>(count, 1) * sizeof(T)));
    if (count) PCQ_CUDA(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return PCQ_OK;
}

// rows of x processed per chunk by the host-pointer entry points
int64_t chunk_rows(int64_t bytes_per_row, int64_t n) {
    const int64_t target = (int64_t)1 << 28;   // 256 MiB of output/input per chunk
    int64_t r = target / std::max<int64_t>(bytes_per_row, 1);
    r = std::max<int64_t>(r, 1024);
    r &= ~(int64_t)15;
    return std::min(r, std::max<int64_t>(n, 1));
}


// host <-> device copy of `rows` rows of `width` doubles; flat when both sides are contiguous
// (2-D copies from pageable memory are staged row by row and are several times slower)
int copy_rows(double* dst, int64_t ld_dst, const double* src, int64_t ld_src, int64_t width, int64_t rows,
              cudaMemcpyKind kind, cudaStream_t st) {
    if (rows == 0 || width == 0) return PCQ_OK;
    if (ld_dst == width && ld_src == width) {
        PCQ_CUDA(cudaMemcpyAsync(dst, src, (size_t)rows * width * 8, kind, st));
    } else {
        PCQ_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld_dst * 8, src, (size_t)ld_src * 8, (size_t)width * 8,
                                   (size_t)rows, kind, st));
    }
    return PCQ_OK;
}

// host threads for the pinned <-> pageable copies of the host-pointer entry points (PCQ_HOST_THREADS, default 16, at
// most the hardware concurrency; measured on the 16-core GPU box: evalBases host-to-host 14.6 / 21.0 / 24.3 GB/s and
// evalPC of 1e7 samples 0.092 / 0.075 / 0.072 s with 4 / 8 / 16 threads).  Divided among the members of a pcq_group
// call, which run one such pipeline per device at the same time.
int pcq_host_threads() {
    static int n = 0;
    if (n == 0) {
        const char* e = getenv("PCQ_HOST_THREADS");
        int v = e ? atoi(e) : 16;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && v > hw) v = hw;
        n = v < 1 ? 1 : (v > 64 ? 64 : v);
    }
    const int div = std::max(1, g_pcq_host_thread_div.load(std::memory_order_relaxed));
    return std::max(2, n / div);
}

int ensure_streams(pcq_plan* p) {
    for (int i = 0; i < 2; ++i) {
        if (!p->streams[i]) PCQ_CUDA(cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking));
        if (!p->events[i]) PCQ_CUDA(cudaEventCreateWithFlags(&p->events[i], cudaEventDisableTiming));
    }
    return PCQ_OK;
}
}  // namespace

void pcq_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int pcq_ensure_dev(double** buf, size_t* have, size_t bytes) {
    if (*have >= bytes && *buf) return PCQ_OK;
    if (*buf) {
        PCQ_CUDA(cudaFree(*buf));
        *buf = nullptr;
        *have = 0;
    }
    cudaError_t e = cudaMalloc((void**)buf, bytes);
    if (e != cudaSuccess) {
        pcq_set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        *buf = nullptr;
        return PCQ_ERR_ALLOC;
    }
    *have = bytes;
    return PCQ_OK;
}

extern "C" {

int pcq_version(void) { return PCQ_VERSION; }
const char* pcq_last_error(void) { return g_err; }
int64_t pcq_launch_count(void) { return g_pcq_launches.load(); }

int pcq_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        pcq_set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        return PCQ_ERR_NODEV;
    }
    return n;
}

int pcq_plan_create(pcq_plan** out, int device, int sdim, int nterms, const int32_t* mi, int pmax,
                    const double* rec_a, const double* rec_b, const double* p1_scale) {
    if (!out) { pcq_set_error("plan_create: out is null"); return PCQ_ERR_ARG; }
    *out = nullptr;
    if (sdim <= 0 || nterms <= 0 || pmax < 0 || !mi || !rec_a || !rec_b || !p1_scale) {
        pcq_set_error("plan_create: bad argument (sdim=%d nterms=%d pmax=%d)", sdim, nterms, pmax);
        return PCQ_ERR_ARG;
    }
    const int64_t nslots64 = 1 + (int64_t)pmax * sdim;
    if (nslots64 > 65535) {
        pcq_set_error("plan_create: 1 + pmax*sdim = %lld table slots exceed 65535", (long long)nslots64);
        return PCQ_ERR_UNSUPPORTED;
    }
    int ndev = pcq_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) {
        pcq_set_error("plan_create: device %d not available (%d visible)", device, ndev);
        return PCQ_ERR_NODEV;
    }
    // host-side compression of the multi-index: per term, the slots of its non-zero orders in
    // ascending dimension order (rv/pcrv.py:436-438 multiplies in that order; x1.0 factors drop out)
    int width = 1;
    for (int k = 0; k < nterms; ++k) {
        int nnz = 0;
        for (int i = 0; i < sdim; ++i) {
            const int o = mi[(size_t)k * sdim + i];
            if (o < 0 || o > pmax) {
                pcq_set_error("plan_create: mi[%d,%d]=%d outside [0,%d]", k, i, o, pmax);
                return PCQ_ERR_ARG;
            }
            nnz += (o > 0);
        }
        width = std::max(width, nnz);
    }
    // Canonical descriptor: W slots, padding (slot 0 = 1.0) FIRST, then the non-zero orders in
    // ascending dimension.  1.0 * a * b rounds exactly like a * b, so the value is unchanged, and
    // neighbouring terms of graded sets then share their first W-1 slots ("prefix"): flag bit 0 of
    // term k says that its prefix differs from term k-1's (K2 reloads the prefix only then).
    std::vector<uint16_t> descw((size_t)nterms * width, 0);
    std::vector<unsigned long long> desc4(nterms, 0ull);
    std::vector<uint8_t> flags(nterms, 1);
    for (int k = 0; k < nterms; ++k) {
        int nnz = 0;
        for (int i = 0; i < sdim; ++i) nnz += (mi[(size_t)k * sdim + i] > 0);
        int j = width - nnz;
        for (int i = 0; i < sdim; ++i) {
            const int o = mi[(size_t)k * sdim + i];
            if (o > 0) descw[(size_t)k * width + j++] = (uint16_t)(1 + (o - 1) * sdim + i);
        }
        if (width <= PCQ_MAXW_FAST) {
            unsigned long long d = 0;
            for (int q = 0; q < width; ++q) d |= (unsigned long long)descw[(size_t)k * width + q] << (16 * q);
            desc4[k] = d;
        }
        if (k > 0) {
            bool same = true;
            for (int q = 0; q + 1 < width; ++q)
                same = same && descw[(size_t)k * width + q] == descw[(size_t)(k - 1) * width + q];
            flags[k] = same ? 0 : 1;
        }
    }
    // downward-closed check (informational; lets callers pick parent-product algorithms)
    int closed = 1;
    {
        std::set<std::vector<int32_t>> rows;
        for (int k = 0; k < nterms; ++k)
            rows.insert(std::vector<int32_t>(mi + (size_t)k * sdim, mi + (size_t)(k + 1) * sdim));
        for (int k = 0; k < nterms && closed; ++k) {
            std::vector<int32_t> r(mi + (size_t)k * sdim, mi + (size_t)(k + 1) * sdim);
            for (int i = 0; i < sdim && closed; ++i) {
                if (r[i] > 0) {
                    r[i] -= 1;
                    if (!rows.count(r)) closed = 0;
                    r[i] += 1;
                }
            }
        }
    }

    DeviceGuard guard(device);
    if (!guard.ok) { pcq_set_error("plan_create: cudaSetDevice(%d) failed", device); return PCQ_ERR_CUDA; }
    pcq_plan* p = new (std::nothrow) pcq_plan();
    if (!p) { pcq_set_error("plan_create: out of host memory"); return PCQ_ERR_ALLOC; }
    p->device = device;
    p->sdim = sdim; p->nterms = nterms; p->pmax = pmax; p->nslots = (int)nslots64;
    p->width = width; p->downward_closed = closed;
    p->h_mi.assign(mi, mi + (size_t)nterms * sdim);
    p->h_rec_a.assign(rec_a, rec_a + (size_t)(pmax + 1) * sdim);
    p->h_rec_b.assign(rec_b, rec_b + (size_t)(pmax + 1) * sdim);
    p->h_p1.assign(p1_scale, p1_scale + sdim);
    int rc = device_props(device, &p->num_sms, &p->smem_optin);
    const size_t nrec = (size_t)(pmax + 1) * sdim;
    if (rc == PCQ_OK) rc = dev_upload(&p->d_rec_a, rec_a, nrec);
    if (rc == PCQ_OK) rc = dev_upload(&p->d_rec_b, rec_b, nrec);
    if (rc == PCQ_OK) rc = dev_upload(&p->d_p1, p1_scale, (size_t)sdim);
    if (rc == PCQ_OK) rc = dev_upload(&p->d_descw, descw.data(), descw.size());
    if (rc == PCQ_OK) rc = dev_upload(&p->d_desc4, desc4.data(), desc4.size());
    if (rc == PCQ_OK) rc = dev_upload(&p->d_flags, flags.data(), flags.size());
    if (rc == PCQ_OK) {
        std::vector<int> zeros(sdim, 0);
        rc = dev_upload(&p->d_kinds, zeros.data(), zeros.size());
    }
    if (rc != PCQ_OK) { pcq_plan_destroy(p); return rc; }
    *out = p;
    return PCQ_OK;
}

int pcq_plan_destroy(pcq_plan* p) {
    if (!p) return PCQ_OK;
    DeviceGuard guard(p->device);
    if (p->spec_thread.joinable()) p->spec_thread.join();
    pcq_spec_free(p->spec[0].load());
    pcq_spec_free(p->spec[1].load());
    pcq_mom_free(p);
    cudaFree(p->d_rec_a); cudaFree(p->d_rec_b); cudaFree(p->d_p1); cudaFree(p->d_kinds);
    cudaFree(p->d_desc4); cudaFree(p->d_descw); cudaFree(p->d_flags); cudaFree(p->d_ws); cudaFree(p->d_misc); cudaFree(p->d_q); cudaFree(p->d_gtab[0]); cudaFree(p->d_gtab[1]); cudaFree(p->d_achunk);
    for (int i = 0; i < 2; ++i) {
        cudaFree(p->d_stage[i]); cudaFree(p->d_out[i]);
        if (p->h_pin[i]) cudaFreeHost(p->h_pin[i]);
        if (p->streams[i]) cudaStreamDestroy(p->streams[i]);
        if (p->events[i]) cudaEventDestroy(p->events[i]);
    }
    delete p;
    return PCQ_OK;
}

int64_t pcq_plan_info(const pcq_plan* p, int what) {
    if (!p) return PCQ_ERR_ARG;
    switch (what) {
        case 0: return p->sdim;
        case 1: return p->nterms;
        case 2: return p->pmax;
        case 3: return p->width;
        case 4: return p->device;
        case 5: return p->downward_closed;
        case 6: return p->num_sms;
        case 7: return (p->spec[0].load(std::memory_order_acquire) ? 1 : 0) | (p->spec[1].load(std::memory_order_acquire) ? 2 : 0);
        case 8: return p->spec_building.load(std::memory_order_acquire);
        default: return PCQ_ERR_ARG;
    }
}

int pcq_plan_specialize_async(pcq_plan* p, int modes) {
    if (!p || (modes & ~3)) { pcq_set_error("specialize_async: bad argument"); return PCQ_ERR_ARG; }
    if (p->spec[0].load(std::memory_order_acquire)) modes &= ~1;
    if (p->spec[1].load(std::memory_order_acquire)) modes &= ~2;
    if (modes == 0 || p->spec_building.load(std::memory_order_acquire)) return PCQ_OK;   // built or under way
    if (p->spec_thread.joinable()) p->spec_thread.join();
    p->spec_building.store(modes, std::memory_order_release);
    p->spec_thread = std::thread([p, modes] {
        cudaSetDevice(p->device);
        if (modes & 1) pcq_spec_build(p, 0);      // failure (no NVRTC): the generic kernels simply stay in use
        if (modes & 2) pcq_spec_build(p, 1);
        p->spec_building.store(0, std::memory_order_release);
    });
    return PCQ_OK;
}

int pcq_plan_specialize(pcq_plan* p, int modes) {
    if (!p || (modes & ~3)) { pcq_set_error("specialize: bad argument"); return PCQ_ERR_ARG; }
    if (p->spec_thread.joinable()) p->spec_thread.join();      // a background build finishes first
    DeviceGuard guard(p->device);
    int rc = PCQ_OK;
    if ((modes & 1) && (rc = pcq_spec_build(p, 0)) != PCQ_OK) return rc;
    if ((modes & 2) && (rc = pcq_spec_build(p, 1)) != PCQ_OK) return rc;
    return PCQ_OK;
}

// ------------------------------------------------------------------ device-pointer entry points
int pcq_eval_bases_dev(pcq_plan* p, const double* x, int64_t n, int64_t ldx, double* a, int64_t lda,
                       void* stream) {
    if (!p || n < 0 || (n > 0 && (!x || !a)) || ldx < p->sdim || lda < p->nterms) {
        pcq_set_error("eval_bases: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    return pcq_launch_bases(p, x, n, ldx, a, lda, (cudaStream_t)stream);
}

// fast mode with many outputs goes to the tensor pipe (packed design-matrix chunk x coefficient matrix),
// everything else to the K2 kernels
static int evalpc_any(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf, int m, double* y,
                      int64_t ldy, int mode, cudaStream_t st);

int pcq_eval_pc_dev(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf, int m,
                    double* y, int64_t ldy, int mode, void* stream) {
    if (!p || n < 0 || m < 0 || (n > 0 && m > 0 && (!x || !cf || !y)) || ldx < p->sdim ||
        ldy < ((mode & 2) ? 1 : m)) {
        pcq_set_error("eval_pc: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    return evalpc_any(p, x, n, ldx, cf, m, y, ldy, mode, (cudaStream_t)stream);
}

namespace {
// The packed chunk of the pack + SYRK path can be 2 GB: it is ONE lazily grown buffer per device, shared by
// all plans and by K4, not a per-plan allocation.  A lease serialises the enqueueing host threads and orders
// the streams that use the buffer one after the other with an event.
struct DevScratch {
    std::mutex mu;
    double* zp = nullptr; size_t zp_bytes = 0;
    cudaEvent_t last = nullptr; cudaStream_t last_stream = nullptr; bool used = false;
};
DevScratch g_scratch[64];
struct ScratchLease {
    DevScratch& d; cudaStream_t st;
    ScratchLease(int device, cudaStream_t stream) : d(g_scratch[device & 63]), st(stream) {
        d.mu.lock();
        if (d.used && d.last_stream != st) cudaStreamWaitEvent(st, d.last, 0);
    }
    ~ScratchLease() {
        if (!d.last) cudaEventCreateWithFlags(&d.last, cudaEventDisableTiming);
        if (d.last) { cudaEventRecord(d.last, st); d.last_stream = st; d.used = true; }
        d.mu.unlock();
    }
};
}  // namespace

static int evalpc_any(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf, int m, double* y,
                      int64_t ldy, int mode, cudaStream_t st) {
    if (mode == 1 && n >= 1024 && pcq_linforms_applicable(p, m)) {
        ScratchLease lease(p->device, st);
        return pcq_launch_linforms(p, x, n, ldx, cf, m, y, ldy, &p->d_q, &p->q_bytes, &lease.d.zp, &lease.d.zp_bytes, st);
    }
    return pcq_launch_evalpc(p, x, n, ldx, cf, m, y, ldy, mode, st);
}

int pcq_quadform_dev(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* c, int64_t ldc,
                     double* v, int64_t ldv, void* stream) {
    if (!p || n < 0 || !c || (n > 0 && (!x || !v)) || ldx < p->sdim || ldc < p->nterms || ldv < 1) {
        pcq_set_error("quadform: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    ScratchLease lease(p->device, (cudaStream_t)stream);
    return pcq_launch_qform(p, x, n, ldx, c, ldc, v, ldv, &p->d_q, &p->q_bytes, &p->d_ws, &p->ws_bytes,
                            &lease.d.zp, &lease.d.zp_bytes, (cudaStream_t)stream);
}

int pcq_residual_ss_dev(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy,
                        const double* cf, double* sums, int accumulate, void* stream) {
    if (!p || n < 0 || !cf || !sums || (n > 0 && (!x || !y)) || ldx < p->sdim || ldy < 1) {
        pcq_set_error("residual_ss: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    return pcq_launch_residual(p, x, n, ldx, y, ldy, cf, sums, accumulate, (cudaStream_t)stream);
}

int64_t pcq_suffstats_dim(const pcq_plan* p, int m) {
    if (!p || m < 0) return PCQ_ERR_ARG;
    return (int64_t)p->nterms + m + 1;
}

int pcq_suffstats_dev(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* y,
                      int64_t ldy, int m, double* s, int accumulate, void* stream) {
    if (!p || n < 0 || m < 0 || !s || (n > 0 && !x) || (n > 0 && m > 0 && !y) || ldx < p->sdim ||
        (m > 0 && ldy < m)) {
        pcq_set_error("suffstats: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    ScratchLease lease(p->device, (cudaStream_t)stream);
    return pcq_launch_gram(p, p->num_sms, p->smem_optin, x, n, ldx, nullptr, 0, 0, y, ldy, m, s,
                           accumulate, &p->d_ws, &p->ws_bytes, &lease.d.zp, &lease.d.zp_bytes, (cudaStream_t)stream);
}

namespace {
// K4 has no plan: keep one workspace per device
struct GramCtx { double* ws = nullptr; size_t ws_bytes = 0; double* stage[2] = {nullptr, nullptr};
                 size_t stage_bytes[2] = {0, 0}; double* s = nullptr; size_t s_bytes = 0;
                 double* ys[2] = {nullptr, nullptr}; size_t ys_bytes[2] = {0, 0};
                 double* bq = nullptr; size_t bq_bytes = 0;        // packed B operand of the matrix forms
                 double* out = nullptr; size_t out_bytes = 0; };   // results of the host-pointer matrix forms
GramCtx g_gram_ctx[64];
}

int pcq_gram_dev(int device, const double* a, int64_t n, int64_t lda, int kcols, const double* y,
                 int64_t ldy, int m, double* s, int accumulate, void* stream) {
    if (device < 0 || device >= 64 || n < 0 || kcols <= 0 || m < 0 || !s || (n > 0 && !a) ||
        (n > 0 && m > 0 && !y) || lda < kcols || (m > 0 && ldy < m)) {
        pcq_set_error("gram: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(device);
    if (!guard.ok) { pcq_set_error("gram: cudaSetDevice(%d) failed", device); return PCQ_ERR_NODEV; }
    int num_sms = 0; size_t smem_optin = 0;
    int rc = device_props(device, &num_sms, &smem_optin);
    if (rc != PCQ_OK) return rc;
    GramCtx& c = g_gram_ctx[device];
    ScratchLease lease(device, (cudaStream_t)stream);     // also guards the per-device K4 workspace
    return pcq_launch_gram(nullptr, num_sms, smem_optin, nullptr, n, 0, a, lda, kcols, y, ldy, m, s,
                           accumulate, &c.ws, &c.ws_bytes, &lease.d.zp, &lease.d.zp_bytes, (cudaStream_t)stream);
}

int pcq_matrix_forms_dev(int device, const double* a, int64_t n, int64_t lda, int kcols, const double* cf, int64_t ldcf,
                         int m, double* y, int64_t ldy, const double* c, int64_t ldc, double* v, int64_t ldv,
                         void* stream) {
    if (device < 0 || device >= 64 || n < 0 || kcols <= 0 || m < 0 || lda < kcols || (n > 0 && !a) ||
        (m > 0 && (!cf || !y || ldcf < kcols || ldy < m)) || (c && (!v || ldc < kcols || ldv < 1))) {
        pcq_set_error("matrix_forms: bad argument");
        return PCQ_ERR_ARG;
    }
    if (n == 0) return PCQ_OK;
    DeviceGuard guard(device);
    if (!guard.ok) { pcq_set_error("matrix_forms: cudaSetDevice(%d) failed", device); return PCQ_ERR_NODEV; }
    int num_sms = 0; size_t smem_optin = 0;
    int rc = device_props(device, &num_sms, &smem_optin);
    if (rc != PCQ_OK) return rc;
    GramCtx& g = g_gram_ctx[device];
    ScratchLease lease(device, (cudaStream_t)stream);     // also guards the per-device workspace
    if (m > 0 && (rc = pcq_launch_linforms_matrix(num_sms, smem_optin, a, n, lda, kcols, cf, ldcf, m, y, ldy, &g.bq,
                                                  &g.bq_bytes, &lease.d.zp, &lease.d.zp_bytes, (cudaStream_t)stream)) != PCQ_OK)
        return rc;
    if (c && (rc = pcq_launch_qform_matrix(num_sms, smem_optin, a, n, lda, kcols, c, ldc, v, ldv, &g.bq, &g.bq_bytes, &g.ws,
                                           &g.ws_bytes, &lease.d.zp, &lease.d.zp_bytes, (cudaStream_t)stream)) != PCQ_OK)
        return rc;
    return PCQ_OK;
}

int pcq_matrix_forms(int device, const double* a, int64_t n, int64_t lda, int kcols, const double* cf, int64_t ldcf, int m,
                     double* y, int64_t ldy, const double* c, int64_t ldc, double* v, int64_t ldv) {
    if (device < 0 || device >= 64 || n < 0 || kcols <= 0 || m < 0 || lda < kcols || (n > 0 && !a) ||
        (m > 0 && (!cf || !y || ldcf < kcols || ldy < m)) || (c && (!v || ldc < kcols || ldv < 1))) {
        pcq_set_error("matrix_forms: bad argument");
        return PCQ_ERR_ARG;
    }
    if (n == 0) return PCQ_OK;
    DeviceGuard guard(device);
    if (!guard.ok) { pcq_set_error("matrix_forms: cudaSetDevice(%d) failed", device); return PCQ_ERR_NODEV; }
    GramCtx& g = g_gram_ctx[device];
    const int K = kcols;
    int rc;
    // operands: CF (m x K) and C (K x K) into the per-device S scratch; A chunk by chunk through the K4 staging buffer
    const size_t opnd = ((size_t)m * K + (c ? (size_t)K * K : 0)) * 8;
    if ((rc = pcq_ensure_dev(&g.s, &g.s_bytes, opnd ? opnd : 8)) != PCQ_OK) return rc;
    double* d_cf = g.s;
    double* d_c = g.s + (size_t)m * K;
    if (m > 0) PCQ_CUDA(cudaMemcpy2D(d_cf, (size_t)K * 8, cf, (size_t)ldcf * 8, (size_t)K * 8, (size_t)m, cudaMemcpyHostToDevice));
    if (c) PCQ_CUDA(cudaMemcpy2D(d_c, (size_t)K * 8, c, (size_t)ldc * 8, (size_t)K * 8, (size_t)K, cudaMemcpyHostToDevice));
    const int64_t rows = chunk_rows((int64_t)(K + m + 1) * 8, n);
    if ((rc = pcq_ensure_dev(&g.stage[0], &g.stage_bytes[0], (size_t)rows * K * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&g.out, &g.out_bytes, (size_t)rows * (m + 1) * 8)) != PCQ_OK) return rc;
    double* d_y = g.out;
    double* d_v = g.out + (size_t)rows * m;
    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = std::min(rows, n - r0);
        PCQ_CUDA(cudaMemcpy2D(g.stage[0], (size_t)K * 8, a + r0 * lda, (size_t)lda * 8, (size_t)K * 8, (size_t)nr,
                              cudaMemcpyHostToDevice));
        if ((rc = pcq_matrix_forms_dev(device, g.stage[0], nr, K, K, m ? d_cf : nullptr, K, m, d_y, std::max(m, 1),
                                       c ? d_c : nullptr, K, d_v, 1, nullptr)) != PCQ_OK)
            return rc;
        if (m > 0) PCQ_CUDA(cudaMemcpy2D(y + r0 * ldy, (size_t)ldy * 8, d_y, (size_t)m * 8, (size_t)m * 8, (size_t)nr, cudaMemcpyDeviceToHost));
        if (c) PCQ_CUDA(cudaMemcpy2D(v + r0 * ldv, (size_t)ldv * 8, d_v, 8, 8, (size_t)nr, cudaMemcpyDeviceToHost));
    }
    return PCQ_OK;
}

int pcq_propagate_dev(pcq_plan* p, uint64_t seed, int64_t first_index, int64_t n,
                      const int32_t* germ_kind, const double* cf, int m, double* sums, double* y,
                      int64_t ldy, int mode, void* stream) {
    if (!p || n < 0 || m <= 0 || !germ_kind || !cf || !sums || (y && ldy < m)) {
        pcq_set_error("propagate: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    return pcq_launch_propagate(p, seed, first_index, n, germ_kind, cf, m, sums, y, ldy, mode,
                                (cudaStream_t)stream);
}

int pcq_sample_germ_dev(pcq_plan* p, uint64_t seed, int64_t first_index, int64_t n,
                        const int32_t* germ_kind, double* x, int64_t ldx, void* stream) {
    if (!p || n < 0 || !germ_kind || (n > 0 && !x) || ldx < p->sdim) {
        pcq_set_error("sample_germ: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    return pcq_launch_germ(p, seed, first_index, n, germ_kind, x, ldx, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ host-pointer entry points
// Double-buffered chunk pipeline: chunk c uses stream c&1; H2D(x) -> kernel -> D2H(result).
// Pageable numpy memory is accepted as is (the driver stages it); the two streams overlap
// the copies of one chunk with the kernel of the other.
int pcq_eval_bases(pcq_plan* p, const double* x, int64_t n, int64_t ldx, double* a, int64_t lda) {
    if (!p || n < 0 || (n > 0 && (!x || !a)) || ldx < p->sdim || lda < p->nterms) {
        pcq_set_error("eval_bases: bad argument");
        return PCQ_ERR_ARG;
    }
    if (n == 0) return PCQ_OK;
    DeviceGuard guard(p->device);
    int rc = ensure_streams(p);
    if (rc != PCQ_OK) return rc;
    const int K = p->nterms, s = p->sdim;
    // The result is PCIe bound (8*K bytes per row back to the host).  D2H goes to two pinned staging
    // buffers at full link speed; the host copies chunk c-1 into the caller's (pageable) array with a few
    // threads while the GPU computes and ships chunk c.
    const int64_t rows = std::min<int64_t>(std::max<int64_t>(((int64_t)64 << 20) / ((int64_t)K * 8), 256), n);
    for (int b = 0; b < 2; ++b) {
        if (p->pin_bytes[b] < (size_t)rows * K * 8) {
            if (p->h_pin[b]) cudaFreeHost(p->h_pin[b]);
            p->h_pin[b] = nullptr; p->pin_bytes[b] = 0;
            if (cudaHostAlloc((void**)&p->h_pin[b], (size_t)rows * K * 8, cudaHostAllocDefault) != cudaSuccess) {
                pcq_set_error("eval_bases: pinned staging allocation of %zu bytes failed", (size_t)rows * K * 8);
                return PCQ_ERR_ALLOC;
            }
            p->pin_bytes[b] = (size_t)rows * K * 8;
        }
    }
    auto drain = [&](int b, int64_t r0, int64_t nr) -> int {      // pinned buffer b -> caller rows [r0, r0+nr)
        PCQ_CUDA(cudaEventSynchronize(p->events[b]));
        const int nthreads = nr * K * 8 >= (8 << 20) ? pcq_host_threads() : 1;
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) {
            const int64_t lo = nr * t / nthreads, hi = nr * (t + 1) / nthreads;
            auto work = [=]() {
                if (lda == K) {
                    memcpy(a + (r0 + lo) * lda, p->h_pin[b] + lo * K, (size_t)(hi - lo) * K * 8);
                } else {
                    for (int64_t r = lo; r < hi; ++r) memcpy(a + (r0 + r) * lda, p->h_pin[b] + r * K, (size_t)K * 8);
                }
            };
            if (nthreads == 1) work(); else pool.emplace_back(work);
        }
        for (auto& th : pool) th.join();
        return PCQ_OK;
    };
    // an error in the middle of the pipeline must not leave copies / kernels in flight on buffers the next call may
    // reallocate
    auto quiesce = [&](int status) { cudaStreamSynchronize(p->streams[0]); cudaStreamSynchronize(p->streams[1]); return status; };
    int c = 0;
    int64_t prev_r0 = 0, prev_nr = 0;
    for (int64_t r0 = 0; r0 < n; r0 += rows, ++c) {
        const int b = c & 1;
        const int64_t nr = std::min(rows, n - r0);
        cudaStream_t st = p->streams[b];
        if ((rc = pcq_ensure_dev(&p->d_stage[b], &p->stage_bytes[b], (size_t)rows * s * 8)) != PCQ_OK) return quiesce(rc);
        if ((rc = pcq_ensure_dev(&p->d_out[b], &p->out_bytes[b], (size_t)rows * K * 8)) != PCQ_OK) return quiesce(rc);
        if ((rc = copy_rows(p->d_stage[b], s, x + r0 * ldx, ldx, s, nr, cudaMemcpyHostToDevice, st)) != PCQ_OK) return quiesce(rc);
        if ((rc = pcq_launch_bases(p, p->d_stage[b], nr, s, p->d_out[b], K, st)) != PCQ_OK) return quiesce(rc);
        PCQ_CUDA(cudaMemcpyAsync(p->h_pin[b], p->d_out[b], (size_t)nr * K * 8, cudaMemcpyDeviceToHost, st));
        PCQ_CUDA(cudaEventRecord(p->events[b], st));
        if (c >= 1 && (rc = drain(b ^ 1, prev_r0, prev_nr)) != PCQ_OK) return quiesce(rc);
        prev_r0 = r0; prev_nr = nr;
    }
    return drain((c - 1) & 1, prev_r0, prev_nr);
}

int pcq_eval_pc(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf, int m,
                double* y, int64_t ldy, int mode) {
    if (!p || n < 0 || m < 0 || (n > 0 && m > 0 && (!x || !cf || !y)) || ldx < p->sdim || ldy < ((mode & 2) ? 1 : m)) {
        pcq_set_error("eval_pc: bad argument");
        return PCQ_ERR_ARG;
    }
    if (n == 0 || m == 0) return PCQ_OK;
    DeviceGuard guard(p->device);
    int rc = ensure_streams(p);
    if (rc != PCQ_OK) return rc;
    const int K = p->nterms, s = p->sdim;
    if ((rc = pcq_ensure_dev(&p->d_misc, &p->misc_bytes, (size_t)m * K * 8)) != PCQ_OK) return rc;
    PCQ_CUDA(cudaMemcpyAsync(p->d_misc, cf, (size_t)m * K * 8, cudaMemcpyHostToDevice, p->streams[0]));
    PCQ_CUDA(cudaEventRecord(p->events[0], p->streams[0]));
    PCQ_CUDA(cudaStreamWaitEvent(p->streams[1], p->events[0], 0));
    const int mo = (mode & 2) ? 1 : m;          // output columns per sample
    // The call is bound by getting x to the device (8 s bytes per sample against ~1 ns of kernel time).  A copy
    // straight from the caller's pageable array is staged by the driver on ONE thread (~9 GB/s); here a few host
    // threads gather each chunk into one of two pinned buffers (which also packs a strided x) and the DMA engine
    // sends it while the next chunk is being gathered: ~2.5x the rate.  Results return through pinned memory too.
    const int64_t rows = std::min<int64_t>(std::max<int64_t>(((int64_t)32 << 20) / ((int64_t)(s + mo) * 8), 1024), n);
    const size_t in_bytes = (size_t)rows * s * 8, out_bytes = (size_t)rows * mo * 8;
    for (int b = 0; b < 2; ++b) {
        if (p->pin_bytes[b] < in_bytes + out_bytes) {
            if (p->h_pin[b]) cudaFreeHost(p->h_pin[b]);
            p->h_pin[b] = nullptr; p->pin_bytes[b] = 0;
            if (cudaHostAlloc((void**)&p->h_pin[b], in_bytes + out_bytes, cudaHostAllocDefault) != cudaSuccess) {
                pcq_set_error("eval_pc: pinned staging allocation of %zu bytes failed", in_bytes + out_bytes);
                return PCQ_ERR_ALLOC;
            }
            p->pin_bytes[b] = in_bytes + out_bytes;
        }
        if ((rc = pcq_ensure_dev(&p->d_stage[b], &p->stage_bytes[b], in_bytes)) != PCQ_OK) return rc;
        if ((rc = pcq_ensure_dev(&p->d_out[b], &p->out_bytes[b], out_bytes)) != PCQ_OK) return rc;
    }
    auto quiesce = [&](int status) { cudaStreamSynchronize(p->streams[0]); cudaStreamSynchronize(p->streams[1]); return status; };
    auto parallel_rows = [&](int64_t nr, size_t bytes, auto&& body) {     // body(lo, hi) over row ranges, a few threads
        const int nthreads = bytes >= ((size_t)4 << 20) ? pcq_host_threads() : 1;
        if (nthreads == 1) { body((int64_t)0, nr); return; }
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back([&, t] { body(nr * t / nthreads, nr * (t + 1) / nthreads); });
        for (auto& th : pool) th.join();
    };
    auto drain = [&](int b, int64_t r0, int64_t nr) -> int {               // results of buffer b -> caller rows [r0, r0+nr)
        PCQ_CUDA(cudaEventSynchronize(p->events[b]));
        const double* src = p->h_pin[b] + (size_t)rows * s;
        parallel_rows(nr, (size_t)nr * mo * 8, [&](int64_t lo, int64_t hi) {
            if (ldy == mo) memcpy(y + (r0 + lo) * ldy, src + lo * mo, (size_t)(hi - lo) * mo * 8);
            else for (int64_t r = lo; r < hi; ++r) memcpy(y + (r0 + r) * ldy, src + r * mo, (size_t)mo * 8);
        });
        return PCQ_OK;
    };
    int c = 0;
    int64_t prev_r0 = 0, prev_nr = 0;
    for (int64_t r0 = 0; r0 < n; r0 += rows, ++c) {
        const int b = c & 1;
        const int64_t nr = std::min(rows, n - r0);
        cudaStream_t st = p->streams[b];
        // buffer b was last used by chunk c-2, whose results were drained (after its event) in iteration c-1
        double* hx = p->h_pin[b];
        parallel_rows(nr, (size_t)nr * s * 8, [&](int64_t lo, int64_t hi) {
            if (ldx == s) memcpy(hx + lo * s, x + (r0 + lo) * ldx, (size_t)(hi - lo) * s * 8);
            else for (int64_t r = lo; r < hi; ++r) memcpy(hx + r * s, x + (r0 + r) * ldx, (size_t)s * 8);
        });
        if (cudaMemcpyAsync(p->d_stage[b], hx, (size_t)nr * s * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            pcq_set_error("eval_pc: host-to-device copy failed");
            return quiesce(PCQ_ERR_CUDA);
        }
        if ((rc = evalpc_any(p, p->d_stage[b], nr, s, p->d_misc, m, p->d_out[b], mo, mode, st)) != PCQ_OK) return quiesce(rc);
        if (cudaMemcpyAsync(p->h_pin[b] + (size_t)rows * s, p->d_out[b], (size_t)nr * mo * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaEventRecord(p->events[b], st) != cudaSuccess) {
            pcq_set_error("eval_pc: device-to-host copy failed");
            return quiesce(PCQ_ERR_CUDA);
        }
        if (c >= 1 && (rc = drain(b ^ 1, prev_r0, prev_nr)) != PCQ_OK) return quiesce(rc);
        prev_r0 = r0; prev_nr = nr;
    }
    if ((rc = drain((c - 1) & 1, prev_r0, prev_nr)) != PCQ_OK) return quiesce(rc);
    PCQ_CUDA(cudaStreamSynchronize(p->streams[0]));
    PCQ_CUDA(cudaStreamSynchronize(p->streams[1]));
    return PCQ_OK;
}

int pcq_quadform(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* c, int64_t ldc,
                 double* v, int64_t ldv) {
    if (!p || n < 0 || !c || (n > 0 && (!x || !v)) || ldx < p->sdim || ldc < p->nterms || ldv < 1) {
        pcq_set_error("quadform: bad argument");
        return PCQ_ERR_ARG;
    }
    if (n == 0) return PCQ_OK;
    DeviceGuard guard(p->device);
    int rc = ensure_streams(p);
    if (rc != PCQ_OK) return rc;
    const int K = p->nterms, s = p->sdim;
    cudaStream_t st = p->streams[0];
    // C goes to the plan's small scratch, x / v through the staging buffers, chunk by chunk
    if ((rc = pcq_ensure_dev(&p->d_misc, &p->misc_bytes, (size_t)K * K * 8)) != PCQ_OK) return rc;
    if ((rc = copy_rows(p->d_misc, K, c, ldc, K, K, cudaMemcpyHostToDevice, st)) != PCQ_OK) return rc;
    const int64_t rows = chunk_rows((int64_t)(s + 1) * 8, n);
    if ((rc = pcq_ensure_dev(&p->d_stage[0], &p->stage_bytes[0], (size_t)rows * s * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&p->d_out[0], &p->out_bytes[0], (size_t)rows * 8)) != PCQ_OK) return rc;
    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = std::min(rows, n - r0);
        if ((rc = copy_rows(p->d_stage[0], s, x + r0 * ldx, ldx, s, nr, cudaMemcpyHostToDevice, st)) != PCQ_OK) return rc;
        if ((rc = pcq_quadform_dev(p, p->d_stage[0], nr, s, p->d_misc, K, p->d_out[0], 1, st)) != PCQ_OK) return rc;
        if ((rc = copy_rows(v + r0 * ldv, ldv, p->d_out[0], 1, 1, nr, cudaMemcpyDeviceToHost, st)) != PCQ_OK) return rc;
    }
    PCQ_CUDA(cudaStreamSynchronize(st));
    return PCQ_OK;
}

int pcq_residual_ss(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy,
                    const double* cf, double* sums) {
    if (!p || n < 0 || !cf || !sums || (n > 0 && (!x || !y)) || ldx < p->sdim || ldy < 1) {
        pcq_set_error("residual_ss: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    int rc = ensure_streams(p);
    if (rc != PCQ_OK) return rc;
    const int K = p->nterms, s = p->sdim;
    cudaStream_t st = p->streams[0];
    if ((rc = pcq_ensure_dev(&p->d_misc, &p->misc_bytes, (size_t)(K + 2) * 8)) != PCQ_OK) return rc;
    double* d_sums = p->d_misc + K;
    PCQ_CUDA(cudaMemcpyAsync(p->d_misc, cf, (size_t)K * 8, cudaMemcpyHostToDevice, st));
    const int64_t rows = chunk_rows((int64_t)(s + 1) * 8, n);
    if ((rc = pcq_ensure_dev(&p->d_stage[0], &p->stage_bytes[0], (size_t)rows * s * 8)) != PCQ_OK) return rc;
    if ((rc = pcq_ensure_dev(&p->d_stage[1], &p->stage_bytes[1], (size_t)rows * 8)) != PCQ_OK) return rc;
    if (n == 0 && (rc = pcq_launch_residual(p, nullptr, 0, s, nullptr, 1, p->d_misc, d_sums, 0, st)) != PCQ_OK) return rc;
    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = std::min(rows, n - r0);
        if ((rc = copy_rows(p->d_stage[0], s, x + r0 * ldx, ldx, s, nr, cudaMemcpyHostToDevice, st)) != PCQ_OK) return rc;
        if ((rc = copy_rows(p->d_stage[1], 1, y + r0 * ldy, ldy, 1, nr, cudaMemcpyHostToDevice, st)) != PCQ_OK) return rc;
        if ((rc = pcq_launch_residual(p, p->d_stage[0], nr, s, p->d_stage[1], 1, p->d_misc, d_sums, r0 > 0, st)) != PCQ_OK)
            return rc;
    }
    PCQ_CUDA(cudaMemcpyAsync(sums, d_sums, 2 * 8, cudaMemcpyDeviceToHost, st));
    PCQ_CUDA(cudaStreamSynchronize(st));
    return PCQ_OK;
}

int pcq_propagate(pcq_plan* p, uint64_t seed, int64_t first_index, int64_t n, const int32_t* germ_kind,
                  const double* cf, int m, double* sums, int mode) {
    if (!p || n < 0 || m <= 0 || !germ_kind || !cf || !sums) {
        pcq_set_error("propagate: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    int rc = ensure_streams(p);
    if (rc != PCQ_OK) return rc;
    const int K = p->nterms;
    // coefficients and the 4 m sums share the plan's small scratch buffer
    if ((rc = pcq_ensure_dev(&p->d_misc, &p->misc_bytes, ((size_t)m * K + (size_t)m * 4) * 8)) != PCQ_OK) return rc;
    cudaStream_t st = p->streams[0];
    double* d_sums = p->d_misc + (size_t)m * K;
    PCQ_CUDA(cudaMemcpyAsync(p->d_misc, cf, (size_t)m * K * 8, cudaMemcpyHostToDevice, st));
    if ((rc = pcq_launch_propagate(p, seed, first_index, n, germ_kind, p->d_misc, m, d_sums, nullptr, m, mode, st)) != PCQ_OK)
        return rc;
    PCQ_CUDA(cudaMemcpyAsync(sums, d_sums, (size_t)m * 4 * 8, cudaMemcpyDeviceToHost, st));
    PCQ_CUDA(cudaStreamSynchronize(st));
    return PCQ_OK;
}

int pcq_suffstats(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy,
                  int m, double* s_host) {
    if (!p || n < 0 || m < 0 || !s_host || (n > 0 && !x) || (n > 0 && m > 0 && !y) ||
        ldx < p->sdim || (m > 0 && ldy < m)) {
        pcq_set_error("suffstats: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(p->device);
    int rc = ensure_streams(p);
    if (rc != PCQ_OK) return rc;
    const int s = p->sdim;
    const int64_t L = (int64_t)p->nterms + m + 1;
    if ((rc = pcq_ensure_dev(&p->d_misc, &p->misc_bytes, (size_t)L * L * 8)) != PCQ_OK) return rc;
    // all H2D copies go on stream 1, all Gram launches on stream 0 (they accumulate into the same S); the two
    // are ordered through events: copied[b] (stream 1 -> 0) and p->events[b] (buffer b consumed, stream 0 -> 1)
    cudaStream_t sk = p->streams[0];
    struct EventPair {
        cudaEvent_t e[2] = {nullptr, nullptr};
        ~EventPair() { for (auto ev : e) if (ev) cudaEventDestroy(ev); }
    } copied;
    for (auto& ev : copied.e) PCQ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ScratchLease lease(p->device, sk);
    if (n == 0) {
        if ((rc = pcq_launch_gram(p, p->num_sms, p->smem_optin, nullptr, 0, s, nullptr, 0, 0, nullptr,
                                  std::max(m, 1), m, p->d_misc, 0, &p->d_ws, &p->ws_bytes, &lease.d.zp, &lease.d.zp_bytes, sk)) != PCQ_OK)
            return rc;
    }
    const int64_t rows = chunk_rows((int64_t)(s + m) * 8, n);
    int c = 0;
    for (int64_t r0 = 0; r0 < n; r0 += rows, ++c) {
        const int b = c & 1;
        const int64_t nr = std::min(rows, n - r0);
        cudaStream_t sc = p->streams[1];
        if ((rc = pcq_ensure_dev(&p->d_stage[b], &p->stage_bytes[b], (size_t)rows * s * 8)) != PCQ_OK) return rc;
        if (m > 0 && (rc = pcq_ensure_dev(&p->d_out[b], &p->out_bytes[b], (size_t)rows * m * 8)) != PCQ_OK) return rc;
        // buffer b was last read by the Gram launch of chunk c-2
        if (c >= 2) PCQ_CUDA(cudaStreamWaitEvent(sc, p->events[b], 0));
        if ((rc = copy_rows(p->d_stage[b], s, x + r0 * ldx, ldx, s, nr, cudaMemcpyHostToDevice, sc)) != PCQ_OK) return rc;
        if (m > 0 && (rc = copy_rows(p->d_out[b], m, y + r0 * ldy, ldy, m, nr, cudaMemcpyHostToDevice, sc)) != PCQ_OK) return rc;
        PCQ_CUDA(cudaEventRecord(copied.e[b], sc));
        PCQ_CUDA(cudaStreamWaitEvent(sk, copied.e[b], 0));
        if ((rc = pcq_launch_gram(p, p->num_sms, p->smem_optin, p->d_stage[b], nr, s, nullptr, 0, 0,
                                  p->d_out[b], std::max(m, 1), m, p->d_misc, c > 0, &p->d_ws,
                                  &p->ws_bytes, &lease.d.zp, &lease.d.zp_bytes, sk)) != PCQ_OK)
            return rc;
        PCQ_CUDA(cudaEventRecord(p->events[b], sk));
    }
    PCQ_CUDA(cudaMemcpyAsync(s_host, p->d_misc, (size_t)L * L * 8, cudaMemcpyDeviceToHost, sk));
    PCQ_CUDA(cudaStreamSynchronize(sk));
    PCQ_CUDA(cudaStreamSynchronize(p->streams[1]));
    return PCQ_OK;
}

int pcq_host_copy_rows(double* dst, int64_t ld_dst, const double* src, int64_t ld_src, int64_t width, int64_t rows) {
    if (rows < 0 || width < 0 || (rows > 0 && width > 0 && (!dst || !src)) || ld_dst < width || ld_src < width) {
        pcq_set_error("host_copy_rows: bad argument");
        return PCQ_ERR_ARG;
    }
    if (rows == 0 || width == 0) return PCQ_OK;
    const int nthreads = (int64_t)rows * width * 8 >= (8 << 20) ? (int)std::min<int64_t>(pcq_host_threads(), rows) : 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t lo = rows * t / nthreads, hi = rows * (t + 1) / nthreads;
        auto work = [=]() {
            if (ld_dst == width && ld_src == width) {
                memcpy(dst + lo * width, src + lo * width, (size_t)(hi - lo) * width * 8);
            } else {
                for (int64_t r = lo; r < hi; ++r) memcpy(dst + r * ld_dst, src + r * ld_src, (size_t)width * 8);
            }
        };
        if (nthreads == 1) work(); else pool.emplace_back(work);
    }
    for (auto& th : pool) th.join();
    return PCQ_OK;
}

int pcq_gram(int device, const double* a, int64_t n, int64_t lda, int kcols, const double* y,
             int64_t ldy, int m, double* s_host) {
    if (device < 0 || device >= 64 || n < 0 || kcols <= 0 || m < 0 || !s_host || (n > 0 && !a) ||
        (n > 0 && m > 0 && !y) || lda < kcols || (m > 0 && ldy < m)) {
        pcq_set_error("gram: bad argument");
        return PCQ_ERR_ARG;
    }
    DeviceGuard guard(device);
    if (!guard.ok) { pcq_set_error("gram: cudaSetDevice(%d) failed", device); return PCQ_ERR_NODEV; }
    GramCtx& c = g_gram_ctx[device];
    const int64_t L = (int64_t)kcols + m + 1;
    int rc;
    if ((rc = pcq_ensure_dev(&c.s, &c.s_bytes, (size_t)L * L * 8)) != PCQ_OK) return rc;
    const int64_t rows = chunk_rows((int64_t)(kcols + m) * 8, n);
    if (n == 0) {
        if ((rc = pcq_gram_dev(device, nullptr, 0, lda, kcols, nullptr, ldy, m, c.s, 0, nullptr)) != PCQ_OK) return rc;
    }
    int ci = 0;
    for (int64_t r0 = 0; r0 < n; r0 += rows, ++ci) {
        const int64_t nr = std::min(rows, n - r0);
        if ((rc = pcq_ensure_dev(&c.stage[0], &c.stage_bytes[0], (size_t)rows * kcols * 8)) != PCQ_OK) return rc;
        if (m > 0 && (rc = pcq_ensure_dev(&c.ys[0], &c.ys_bytes[0], (size_t)rows * m * 8)) != PCQ_OK) return rc;
        PCQ_CUDA(cudaMemcpy2D(c.stage[0], (size_t)kcols * 8, a + r0 * lda, (size_t)lda * 8,
                              (size_t)kcols * 8, (size_t)nr, cudaMemcpyHostToDevice));
        if (m > 0)
            PCQ_CUDA(cudaMemcpy2D(c.ys[0], (size_t)m * 8, y + r0 * ldy, (size_t)ldy * 8, (size_t)m * 8,
                                  (size_t)nr, cudaMemcpyHostToDevice));
        if ((rc = pcq_gram_dev(device, c.stage[0], nr, kcols, kcols, c.ys[0], std::max(m, 1), m, c.s,
                               ci > 0, nullptr)) != PCQ_OK)
            return rc;
    }
    PCQ_CUDA(cudaMemcpy(s_host, c.s, (size_t)L * L * 8, cudaMemcpyDeviceToHost));
    return PCQ_OK;
}

}  // extern "C"
