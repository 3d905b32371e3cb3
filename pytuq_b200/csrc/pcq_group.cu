// pcq_group: ONE process driving several GPUs behind the unchanged host-pointer API (`ndev, devs` of the
// plan; 8(e): the path shards over samples).
//
// A group owns one pcq_plan per device.  Every entry point cuts the caller's rows into contiguous blocks, one per
// device, and runs them concurrently on one host thread per device:
//   * evalBases / evalPC / quadratic forms / residual sums / propagation need no exchange at all (rows are
//     independent; propagation blocks are disjoint index ranges of one logical Philox stream);
//   * the statistics pass reduces each block to a partial S on its own device (pinned double-buffered H2D of the block
//     overlapping K3), then the partials are summed ON DEVICE 0 in device order -- peer copies over NVLink + an add
//     kernel, deterministic -- which is the single exchange step of the path (the all-reduce of the multi-process
//     form, pytuq_b200/parallel.py).
// Nothing here touches NCCL: inside one process the peer copy is the cheaper primitive (25 MB .. 261 MB once per fit).
#include "pcq_internal.cuh"

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Member {
    int device = 0;
    pcq_plan* plan = nullptr;
    // statistics pass: pinned staging pair, device stage pair, partial S, streams and events
    double* h_pin[2] = {nullptr, nullptr};
    size_t pin_bytes[2] = {0, 0};
    double* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes[2] = {0, 0};
    double* d_S = nullptr;
    size_t S_bytes = 0;
    double* d_tmp = nullptr;      // device 0 only: landing buffer of the peer copies
    size_t tmp_bytes = 0;
    cudaStream_t copy = nullptr, comp = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    // result of the last call
    int rc = PCQ_OK;
    std::string err;
};

}  // namespace

struct pcq_group {
    std::vector<Member> m;
    int active = 0;        // members the next calls are spread over (the first `active` of m)
    int sdim = 0, nterms = 0;
};

namespace {

constexpr size_t GROUP_STAGE_BYTES = (size_t)128 << 20;

__global__ void __launch_bounds__(256) pcq_group_add_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] += src[i];
}

void block_of(int64_t n, int i, int parts, int64_t* lo, int64_t* hi) {
    const int64_t base = n / parts, extra = n % parts;
    *lo = i * base + std::min<int64_t>(i, extra);
    *hi = *lo + base + (i < extra ? 1 : 0);
}

int ensure_pinned(double** buf, size_t* have, size_t bytes) {
    if (*buf && *have >= bytes) return PCQ_OK;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr; *have = 0;
    if (cudaHostAlloc((void**)buf, bytes, cudaHostAllocPortable) != cudaSuccess) {
        pcq_set_error("group: pinned staging allocation of %zu bytes failed", bytes);
        return PCQ_ERR_ALLOC;
    }
    *have = bytes;
    return PCQ_OK;
}

int member_streams(Member& M) {
    if (!M.copy) PCQ_CUDA(cudaStreamCreateWithFlags(&M.copy, cudaStreamNonBlocking));
    if (!M.comp) PCQ_CUDA(cudaStreamCreateWithFlags(&M.comp, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        if (!M.copied[b]) PCQ_CUDA(cudaEventCreateWithFlags(&M.copied[b], cudaEventDisableTiming));
        if (!M.consumed[b]) PCQ_CUDA(cudaEventCreateWithFlags(&M.consumed[b], cudaEventDisableTiming));
    }
    return PCQ_OK;
}

// partial statistics of rows [lo, hi) on M's device, left in M.d_S
int member_suffstats(Member& M, const double* x, int64_t ldx, const double* y, int64_t ldy, int m, int64_t lo, int64_t hi) {
    if (cudaSetDevice(M.device) != cudaSuccess) { pcq_set_error("group: cudaSetDevice(%d) failed", M.device); return PCQ_ERR_NODEV; }
    int rc = member_streams(M);
    if (rc != PCQ_OK) return rc;
    const int s = M.plan->sdim;
    const int64_t L = (int64_t)M.plan->nterms + m + 1;
    if ((rc = pcq_ensure_dev(&M.d_S, &M.S_bytes, (size_t)L * L * 8)) != PCQ_OK) return rc;
    const int64_t n = hi - lo;
    if (n == 0) {
        rc = pcq_suffstats_dev(M.plan, nullptr, 0, s, nullptr, std::max(m, 1), m, M.d_S, 0, M.comp);
        if (rc == PCQ_OK) PCQ_CUDA(cudaStreamSynchronize(M.comp));
        return rc;
    }
    const int64_t ncol = s + m;
    const int64_t rows = std::min<int64_t>(std::max<int64_t>((int64_t)(GROUP_STAGE_BYTES / (ncol * 8)), 4096), n);
    for (int b = 0; b < 2; ++b) {
        if ((rc = ensure_pinned(&M.h_pin[b], &M.pin_bytes[b], (size_t)rows * ncol * 8)) != PCQ_OK) return rc;
        if ((rc = pcq_ensure_dev(&M.d_stage[b], &M.stage_bytes[b], (size_t)rows * ncol * 8)) != PCQ_OK) return rc;
    }
    // chunk schedule: the first host copy + H2D cannot overlap anything, so the pipeline ramps up (1/8, 1/4, 1/2 of a
    // stage) before it settles on full stages
    std::vector<std::pair<int64_t, int64_t>> chunks;
    int64_t r0 = lo;
    for (int frac : {8, 4, 2}) {
        if (hi - r0 > rows) { const int64_t nr = std::max<int64_t>(rows / frac, 4096); chunks.push_back({r0, nr}); r0 += nr; }
    }
    while (r0 < hi) { const int64_t nr = std::min(rows, hi - r0); chunks.push_back({r0, nr}); r0 += nr; }
    int c = 0;
    for (auto& ch : chunks) {
        const int b = c & 1;
        const int64_t q0 = ch.first, nr = ch.second;
        if (c >= 2) PCQ_CUDA(cudaEventSynchronize(M.consumed[b]));       // staging pair b free again (host and device side)
        double* hx = M.h_pin[b];
        double* hy = hx + nr * s;
        if (ldx == s) memcpy(hx, x + q0 * ldx, (size_t)nr * s * 8);
        else for (int64_t r = 0; r < nr; ++r) memcpy(hx + r * s, x + (q0 + r) * ldx, (size_t)s * 8);
        if (m > 0) {
            if (ldy == m) memcpy(hy, y + q0 * ldy, (size_t)nr * m * 8);
            else for (int64_t r = 0; r < nr; ++r) memcpy(hy + r * m, y + (q0 + r) * ldy, (size_t)m * 8);
        }
        PCQ_CUDA(cudaMemcpyAsync(M.d_stage[b], hx, (size_t)nr * ncol * 8, cudaMemcpyHostToDevice, M.copy));
        PCQ_CUDA(cudaEventRecord(M.copied[b], M.copy));
        PCQ_CUDA(cudaStreamWaitEvent(M.comp, M.copied[b], 0));
        rc = pcq_suffstats_dev(M.plan, M.d_stage[b], nr, s, m > 0 ? M.d_stage[b] + nr * s : nullptr, std::max(m, 1), m, M.d_S,
                               c > 0 ? 1 : 0, M.comp);
        if (rc != PCQ_OK) { cudaStreamSynchronize(M.copy); cudaStreamSynchronize(M.comp); return rc; }
        PCQ_CUDA(cudaEventRecord(M.consumed[b], M.comp));
        PCQ_CUDA(cudaStreamWaitEvent(M.copy, M.consumed[b], 0));         // the next H2D into stage b waits for the kernel
        ++c;
    }
    PCQ_CUDA(cudaStreamSynchronize(M.comp));
    PCQ_CUDA(cudaStreamSynchronize(M.copy));
    return PCQ_OK;
}

// runs fn(member index) on one host thread per member; returns the first failure (its message becomes the caller's)
template <typename F>
int run_members(pcq_group* g, F fn) {
    const int nd = g->active;
    struct Share { Share(int n) { g_pcq_host_thread_div.store(n); } ~Share() { g_pcq_host_thread_div.store(1); } } share(nd);
    std::vector<std::thread> pool;
    for (int i = 0; i < nd; ++i) {
        pool.emplace_back([g, i, &fn] {
            Member& M = g->m[i];
            cudaSetDevice(M.device);
            M.rc = fn(i);
            M.err = (M.rc != PCQ_OK) ? std::string(pcq_last_error()) : std::string();
        });
    }
    for (auto& t : pool) t.join();
    for (int i = 0; i < nd; ++i)
        if (g->m[i].rc != PCQ_OK) {
            pcq_set_error("group member %d (device %d): %s", i, g->m[i].device, g->m[i].err.c_str());
            return g->m[i].rc;
        }
    return PCQ_OK;
}

}  // namespace

extern "C" {

int pcq_group_create(pcq_group** out, int ndev, const int* devices, int sdim, int nterms, const int32_t* mi, int pmax,
                     const double* rec_a, const double* rec_b, const double* p1_scale) {
    if (!out) { pcq_set_error("group_create: out is null"); return PCQ_ERR_ARG; }
    *out = nullptr;
    if (ndev <= 0 || ndev > 64 || !devices) { pcq_set_error("group_create: bad device list"); return PCQ_ERR_ARG; }
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) { pcq_set_error("group_create: device %d listed twice", devices[i]); return PCQ_ERR_ARG; }
    pcq_group* g = new pcq_group();
    g->sdim = sdim; g->nterms = nterms;
    g->m.resize(ndev);
    g->active = ndev;
    for (int i = 0; i < ndev; ++i) {
        g->m[i].device = devices[i];
        int rc = pcq_plan_create(&g->m[i].plan, devices[i], sdim, nterms, mi, pmax, rec_a, rec_b, p1_scale);
        if (rc != PCQ_OK) { pcq_group_destroy(g); return rc; }
    }
    // peer access from the reducing device to every other one (an already enabled pair is not an error)
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(devices[0]) == cudaSuccess) {
        for (int i = 1; i < ndev; ++i) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[0], devices[i]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[i], 0);
                if (e != cudaSuccess) cudaGetLastError();
            }
        }
    }
    if (prev >= 0) cudaSetDevice(prev);
    *out = g;
    return PCQ_OK;
}

int pcq_group_destroy(pcq_group* g) {
    if (!g) return PCQ_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    for (Member& M : g->m) {
        if (cudaSetDevice(M.device) != cudaSuccess) continue;
        for (int b = 0; b < 2; ++b) {
            if (M.h_pin[b]) cudaFreeHost(M.h_pin[b]);
            cudaFree(M.d_stage[b]);
            if (M.copied[b]) cudaEventDestroy(M.copied[b]);
            if (M.consumed[b]) cudaEventDestroy(M.consumed[b]);
        }
        cudaFree(M.d_S);
        cudaFree(M.d_tmp);
        if (M.copy) cudaStreamDestroy(M.copy);
        if (M.comp) cudaStreamDestroy(M.comp);
        if (M.plan) pcq_plan_destroy(M.plan);
    }
    if (prev >= 0) cudaSetDevice(prev);
    delete g;
    return PCQ_OK;
}

int pcq_group_size(const pcq_group* g) { return g ? (int)g->m.size() : PCQ_ERR_ARG; }

int pcq_group_enable_moments(pcq_group* g, int nord, const double* rec_a_ext, const double* rec_b_ext) {
    if (!g) { pcq_set_error("group_enable_moments: group is null"); return PCQ_ERR_ARG; }
    for (auto& mem : g->m) {
        const int rc = pcq_plan_enable_moments(mem.plan, nord, rec_a_ext, rec_b_ext);
        if (rc != PCQ_OK) return rc;
    }
    return PCQ_OK;
}

int pcq_group_set_active(pcq_group* g, int n) {
    if (!g || n < 1 || n > (int)g->m.size()) { pcq_set_error("group_set_active: 1 <= n <= size"); return PCQ_ERR_ARG; }
    g->active = n;
    return PCQ_OK;
}

pcq_plan* pcq_group_plan(pcq_group* g, int i) { return (g && i >= 0 && i < (int)g->m.size()) ? g->m[i].plan : nullptr; }

int pcq_group_device(const pcq_group* g, int i) { return (g && i >= 0 && i < (int)g->m.size()) ? g->m[i].device : PCQ_ERR_ARG; }

int pcq_group_specialize(pcq_group* g, int modes) {
    if (!g) { pcq_set_error("group_specialize: null group"); return PCQ_ERR_ARG; }
    // the first member compiles (and fills the cubin cache on disk), the others then load from it in parallel
    int rc = pcq_plan_specialize(g->m[0].plan, modes);
    if (rc != PCQ_OK || g->m.size() == 1) return rc;
    const int saved = g->active;
    g->active = (int)g->m.size();
    rc = run_members(g, [g, modes](int i) { return i == 0 ? PCQ_OK : pcq_plan_specialize(g->m[i].plan, modes); });
    g->active = saved;
    return rc;
}

int pcq_group_suffstats(pcq_group* g, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy, int m,
                        double* s_dev0, double* s_host) {
    if (!g || n < 0 || m < 0 || (!s_dev0 && !s_host) || (n > 0 && !x) || (n > 0 && m > 0 && !y) || ldx < g->sdim ||
        (m > 0 && ldy < m)) {
        pcq_set_error("group_suffstats: bad argument");
        return PCQ_ERR_ARG;
    }
    const int nd = g->active;
    int rc = run_members(g, [&](int i) {
        int64_t lo, hi;
        block_of(n, i, nd, &lo, &hi);
        return member_suffstats(g->m[i], x, ldx, y, ldy, m, lo, hi);
    });
    if (rc != PCQ_OK) return rc;
    // the exchange step: partial S of device i -> device 0 (peer copy), summed there in device order
    Member& R = g->m[0];
    int prev = -1;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
    PCQ_CUDA(cudaSetDevice(R.device));
    const int64_t L = (int64_t)g->nterms + m + 1;
    const size_t bytes = (size_t)L * L * 8;
    if (nd > 1) {
        if ((rc = pcq_ensure_dev(&R.d_tmp, &R.tmp_bytes, bytes)) != PCQ_OK) return rc;
        for (int i = 1; i < nd; ++i) {
            PCQ_CUDA(cudaMemcpyPeerAsync(R.d_tmp, R.device, g->m[i].d_S, g->m[i].device, bytes, R.comp));
            pcq_group_add_kernel<<<R.plan->num_sms * 4, 256, 0, R.comp>>>(R.d_S, R.d_tmp, L * L);
            PCQ_LAUNCH_CHECK("pcq_group_add_kernel");
        }
    }
    if (s_dev0) PCQ_CUDA(cudaMemcpyAsync(s_dev0, R.d_S, bytes, cudaMemcpyDeviceToDevice, R.comp));
    if (s_host) PCQ_CUDA(cudaMemcpyAsync(s_host, R.d_S, bytes, cudaMemcpyDeviceToHost, R.comp));
    PCQ_CUDA(cudaStreamSynchronize(R.comp));
    return PCQ_OK;
}

int pcq_group_eval_bases(pcq_group* g, const double* x, int64_t n, int64_t ldx, double* a, int64_t lda) {
    if (!g || n < 0 || (n > 0 && (!x || !a)) || ldx < g->sdim || lda < g->nterms) {
        pcq_set_error("group_eval_bases: bad argument");
        return PCQ_ERR_ARG;
    }
    const int nd = g->active;
    return run_members(g, [&](int i) {
        int64_t lo, hi;
        block_of(n, i, nd, &lo, &hi);
        return hi > lo ? pcq_eval_bases(g->m[i].plan, x + lo * ldx, hi - lo, ldx, a + lo * lda, lda) : PCQ_OK;
    });
}

int pcq_group_eval_pc(pcq_group* g, const double* x, int64_t n, int64_t ldx, const double* cf, int m, double* y, int64_t ldy,
                      int mode) {
    if (!g || n < 0 || m < 0 || (n > 0 && m > 0 && (!x || !cf || !y)) || ldx < g->sdim || ldy < ((mode & 2) ? 1 : m)) {
        pcq_set_error("group_eval_pc: bad argument");
        return PCQ_ERR_ARG;
    }
    const int nd = g->active;
    return run_members(g, [&](int i) {
        int64_t lo, hi;
        block_of(n, i, nd, &lo, &hi);
        return hi > lo ? pcq_eval_pc(g->m[i].plan, x + lo * ldx, hi - lo, ldx, cf, m, y + lo * ldy, ldy, mode) : PCQ_OK;
    });
}

int pcq_group_quadform(pcq_group* g, const double* x, int64_t n, int64_t ldx, const double* c, int64_t ldc, double* v,
                       int64_t ldv) {
    if (!g || n < 0 || !c || (n > 0 && (!x || !v)) || ldx < g->sdim || ldc < g->nterms || ldv < 1) {
        pcq_set_error("group_quadform: bad argument");
        return PCQ_ERR_ARG;
    }
    const int nd = g->active;
    return run_members(g, [&](int i) {
        int64_t lo, hi;
        block_of(n, i, nd, &lo, &hi);
        return hi > lo ? pcq_quadform(g->m[i].plan, x + lo * ldx, hi - lo, ldx, c, ldc, v + lo * ldv, ldv) : PCQ_OK;
    });
}

int pcq_group_residual_ss(pcq_group* g, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy, const double* cf,
                          double* sums) {
    if (!g || n < 0 || !cf || !sums || (n > 0 && (!x || !y)) || ldx < g->sdim || ldy < 1) {
        pcq_set_error("group_residual_ss: bad argument");
        return PCQ_ERR_ARG;
    }
    const int nd = g->active;
    std::vector<double> part((size_t)nd * 2, 0.0);
    int rc = run_members(g, [&](int i) {
        int64_t lo, hi;
        block_of(n, i, nd, &lo, &hi);
        return hi > lo ? pcq_residual_ss(g->m[i].plan, x + lo * ldx, hi - lo, ldx, y + lo * ldy, ldy, cf, &part[(size_t)i * 2]) : PCQ_OK;
    });
    if (rc != PCQ_OK) return rc;
    sums[0] = sums[1] = 0.0;
    for (int i = 0; i < nd; ++i) { sums[0] += part[(size_t)i * 2]; sums[1] += part[(size_t)i * 2 + 1]; }
    return PCQ_OK;
}

int pcq_group_propagate(pcq_group* g, uint64_t seed, int64_t first_index, int64_t n, const int32_t* germ_kind, const double* cf,
                        int m, double* sums, int mode) {
    if (!g || n < 0 || m <= 0 || !germ_kind || !cf || !sums) {
        pcq_set_error("group_propagate: bad argument");
        return PCQ_ERR_ARG;
    }
    const int nd = g->active;
    std::vector<double> part((size_t)nd * m * 4, 0.0);
    std::vector<int64_t> cnt(nd, 0);
    int rc = run_members(g, [&](int i) {
        int64_t lo, hi;
        block_of(n, i, nd, &lo, &hi);
        cnt[i] = hi - lo;
        return hi > lo ? pcq_propagate(g->m[i].plan, seed, first_index + lo, hi - lo, germ_kind, cf, m, &part[(size_t)i * m * 4], mode)
                       : PCQ_OK;
    });
    if (rc != PCQ_OK) return rc;
    bool first = true;
    for (int i = 0; i < nd; ++i) {
        if (cnt[i] == 0) continue;
        for (int j = 0; j < m; ++j) {
            const double* p = &part[((size_t)i * m + j) * 4];
            double* o = sums + (size_t)j * 4;
            if (first) { o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3]; }
            else { o[0] += p[0]; o[1] += p[1]; o[2] = std::min(o[2], p[2]); o[3] = std::max(o[3], p[3]); }
        }
        first = false;
    }
    if (first)      // n == 0: the single-device call defines the empty result
        return pcq_propagate(g->m[0].plan, seed, first_index, 0, germ_kind, cf, m, sums, mode);
    return PCQ_OK;
}

}  // extern "C"
