// K2s: evalPC kernels specialised per multi-index (opt-in, pcq_plan_specialize).
//
// The generic K2 is bound by shared-memory traffic and latency: every factor of every term is a
// per-sample gather from the [slot][lane] table.  For a FIXED multi-index the term list can be
// unrolled into straight-line code whose 1-D table lives in registers (named scalars, static
// indexing): 3.2-3.8x faster and still bit-identical in exact mode (profiles/r01_d_k2_notes.md).
// The price is compile time (~quadratic in the block size), so the list is cut into chunks of
// SPEC_CHUNK terms, one kernel each, accumulating through y in term order (exactness needs only
// the k order of the sum, which is kept).  Source is generated here from the plan, compiled with
// NVRTC (dlopen'ed: libpcq has no link-time dependency on it) to an sm_100a cubin, loaded with
// cudaLibraryLoadData.  Constants are emitted as hexadecimal float literals, i.e. exactly.
#include "pcq_internal.cuh"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int SPEC_CHUNK = 512;
constexpr int SPEC_THREADS = 128;

struct SpecChunk { cudaKernel_t kern = nullptr; void* cfc = nullptr; int k0 = 0, k1 = 0; };

}  // namespace

struct pcq_spec {
    std::vector<cudaLibrary_t> libs;
    std::vector<SpecChunk> chunks;
};

namespace {

// ---- NVRTC through dlopen ------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    int (*Version)(int*, int*) = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc N;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* nm : names) {
            N.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (N.h) break;
        }
        if (!N.h) return;
#define PCQ_SYM(field, sym) *(void**)(&N.field) = dlsym(N.h, sym)
        PCQ_SYM(CreateProgram, "nvrtcCreateProgram");
        PCQ_SYM(CompileProgram, "nvrtcCompileProgram");
        PCQ_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
        PCQ_SYM(GetCUBIN, "nvrtcGetCUBIN");
        PCQ_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        PCQ_SYM(GetProgramLog, "nvrtcGetProgramLog");
        PCQ_SYM(DestroyProgram, "nvrtcDestroyProgram");
        PCQ_SYM(Version, "nvrtcVersion");
#undef PCQ_SYM
        N.ok = N.CreateProgram && N.CompileProgram && N.GetCUBINSize && N.GetCUBIN && N.GetProgramLogSize &&
               N.GetProgramLog && N.DestroyProgram;
    });
    return N;
}

std::string hexd(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%a", v);
    return buf;
}

// source of the kernel for terms [k0, k1)
std::string gen_source(const pcq_plan* p, int k0, int k1, bool fast, const std::string& name) {
    const int s = p->sdim;
    std::string o;
    o.reserve((size_t)(k1 - k0) * 96 + 4096);
    // the chunk's coefficients live in the constant bank (copied there in-stream before every launch):
    // FP64 instructions take them as direct operands, no load instruction per term
    o += "__constant__ double " + name + "_cf[" + std::to_string(SPEC_CHUNK) + "];\n";
    o += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(SPEC_THREADS) + ") " + name +
         "(const double* __restrict__ x, long long n, long long ldx, double* __restrict__ y, long long ldy) {\n";
    o += "  const long long nb = (n + " + std::to_string(SPEC_THREADS - 1) + ") / " + std::to_string(SPEC_THREADS) + ";\n";
    o += "  for (long long blk = blockIdx.x; blk < nb; blk += gridDim.x) {\n";
    o += "    const long long i = blk * " + std::to_string(SPEC_THREADS) + " + threadIdx.x;\n";
    o += "    const bool live = i < n;\n";
    o += "    const double* xr = x + (live ? i : 0) * ldx;\n";
    // highest order of every dimension used by this chunk
    std::vector<int> top(s, 0);
    for (int k = k0; k < k1; ++k)
        for (int i = 0; i < s; ++i) top[i] = std::max(top[i], (int)p->h_mi[(size_t)k * s + i]);
    for (int i = 0; i < s; ++i) {
        if (top[i] == 0) continue;
        const std::string xi = "x" + std::to_string(i), t = "t" + std::to_string(i) + "_";
        o += "    const double " + xi + " = xr[" + std::to_string(i) + "];\n";
        // PC1d.__call__ (rv/pcrv.py:794-802): p1 = x * scale ; p_{n+1} = ((a_n x) p_n) + (b_n p_{n-1})
        o += "    const double " + t + "1 = __dmul_rn(" + xi + ", " + hexd(p->h_p1[i]) + ");\n";
        for (int nn = 1; nn < top[i]; ++nn) {
            const std::string pm = (nn == 1) ? std::string("1.0") : t + std::to_string(nn - 1);
            o += "    const double " + t + std::to_string(nn + 1) + " = __dadd_rn(__dmul_rn(__dmul_rn(" +
                 hexd(p->h_rec_a[(size_t)nn * s + i]) + ", " + xi + "), " + t + std::to_string(nn) + "), __dmul_rn(" +
                 hexd(p->h_rec_b[(size_t)nn * s + i]) + ", " + pm + "));\n";
        }
    }
    o += std::string("    double acc = ") + (k0 == 0 ? "0.0" : "(live ? y[i * ldy] : 0.0)") + ";\n";
    // A CTA barrier every few terms: it bounds how far ptxas hoists the (table x coefficient) products
    // ahead of the serial accumulation chain -- without it the first chunk of an exact kernel needs
    // >255 registers and spills 2.5 KB per thread -- and keeps the warps of a CTA on the same cache lines
    // of this long straight-line code.
    int sync_every = 64;
    if (const char* e = getenv("PCQ_SPEC_SYNC")) sync_every = atoi(e);
    // (independent partial sums in fast mode were measured SLOWER: 0.50 -> 0.52 (2) -> 0.76 ms (4), the
    // extra scheduling freedom costs registers; one serial accumulator it is)
    for (int k = k0; k < k1; ++k) {
        if (sync_every > 0 && k > k0 && (k - k0) % sync_every == 0) o += "    __syncthreads();\n";
        const std::string c = name + "_cf[" + std::to_string(k - k0) + "]";
        if (!fast) {
            // rv/pcrv.py:492-495: prd = cf[k]; prd *= phi_i (i ascending); val += prd  (x1.0 factors dropped: exact)
            std::string e = c;
            for (int i = 0; i < s; ++i) {
                const int od = p->h_mi[(size_t)k * s + i];
                if (od > 0) e = "__dmul_rn(" + e + ", t" + std::to_string(i) + "_" + std::to_string(od) + ")";
            }
            o += "    acc = __dadd_rn(acc, " + e + ");\n";
        } else {
            std::string e;
            for (int i = 0; i < s; ++i) {
                const int od = p->h_mi[(size_t)k * s + i];
                if (od > 0) {
                    const std::string f = "t" + std::to_string(i) + "_" + std::to_string(od);
                    e = e.empty() ? f : "__dmul_rn(" + e + ", " + f + ")";
                }
            }
            o += e.empty() ? "    acc = __dadd_rn(acc, " + c + ");\n" : "    acc = __fma_rn(" + c + ", " + e + ", acc);\n";
        }
    }
    o += "    if (live) y[i * ldy] = acc;\n  }\n}\n";
    return o;
}

// ---- cubin cache on disk -------------------------------------------------------------------
// The generated source is a pure function of (multi-index chunk, recurrence constants, mode), so its cubin can
// be reused across processes: $PCQ_SPEC_CACHE_DIR (default $HOME/.cache/pcq_k2s; empty string disables) holds
// one file per chunk, named by a 128-bit FNV-1a hash of the source and the NVRTC version.  Writes go through a
// temporary file + rename, so concurrent ranks never read a partial cubin.
std::string cache_dir() {
    const char* e = getenv("PCQ_SPEC_CACHE_DIR");
    if (e) return std::string(e);
    const char* home = getenv("HOME");
    return home ? std::string(home) + "/.cache/pcq_k2s" : std::string();
}

std::string cache_key(const std::string& src) {
    unsigned long long h1 = 1469598103934665603ull, h2 = 1099511628211ull * 7919ull;
    int major = 0, minor = 0;
    if (nvrtc().Version) nvrtc().Version(&major, &minor);
    const std::string salted = src + "|sm_100a|nvrtc" + std::to_string(major) + "." + std::to_string(minor);
    for (unsigned char ch : salted) {
        h1 = (h1 ^ ch) * 1099511628211ull;
        h2 = (h2 ^ (ch + 0x9eu)) * 0x100000001b3ull + 0x632be59bd9b4e019ull;
    }
    char buf[40];
    snprintf(buf, sizeof(buf), "%016llx%016llx", h1, h2);
    return buf;
}

bool cache_load(const std::string& key, std::vector<char>& cubin) {
    const std::string dir = cache_dir();
    if (dir.empty()) return false;
    FILE* f = fopen((dir + "/" + key + ".cubin").c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    bool ok = sz > 0;
    if (ok) {
        cubin.resize((size_t)sz);
        ok = fread(cubin.data(), 1, (size_t)sz, f) == (size_t)sz;
    }
    fclose(f);
    return ok;
}

void cache_store(const std::string& key, const std::vector<char>& cubin) {
    const std::string dir = cache_dir();
    if (dir.empty()) return;
    std::string partial;
    for (size_t i = 1; i <= dir.size(); ++i)          // mkdir -p
        if (i == dir.size() || dir[i] == '/') { partial = dir.substr(0, i); mkdir(partial.c_str(), 0755); }
    const std::string tmp = dir + "/" + key + "." + std::to_string((long)getpid()) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    fclose(f);
    if (ok) rename(tmp.c_str(), (dir + "/" + key + ".cubin").c_str());
    else remove(tmp.c_str());
}

int compile_chunk(const std::string& src, const std::string& name, std::vector<char>& cubin, std::string& err) {
    Nvrtc& N = nvrtc();
    const std::string key = cache_key(src);
    if (cache_load(key, cubin)) return PCQ_OK;
    nvrtcProgram prog = nullptr;
    if (N.CreateProgram(&prog, src.c_str(), (name + ".cu").c_str(), 0, nullptr, nullptr) != 0) {
        err = "nvrtcCreateProgram failed";
        return PCQ_ERR_CUDA;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17"};
    const int rc = N.CompileProgram(prog, 2, opts);
    if (rc != 0) {
        size_t ls = 0;
        N.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) N.GetProgramLog(prog, &log[0]);
        err = "nvrtcCompileProgram failed: " + log.substr(0, 300);
        N.DestroyProgram(&prog);
        return PCQ_ERR_CUDA;
    }
    size_t cs = 0;
    N.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    N.GetCUBIN(prog, cubin.data());
    N.DestroyProgram(&prog);
    cache_store(key, cubin);
    return PCQ_OK;
}

}  // namespace

void pcq_spec_free(pcq_spec* sp) {
    if (!sp) return;
    for (cudaLibrary_t l : sp->libs) cudaLibraryUnload(l);
    delete sp;
}

int pcq_spec_build(pcq_plan* p, int fast) {
    if (p->spec[fast].load(std::memory_order_acquire)) return PCQ_OK;
    if (!nvrtc().ok) {
        pcq_set_error("specialize: libnvrtc.so.12 not found (the generic kernels stay in use)");
        return PCQ_ERR_UNSUPPORTED;
    }
    const int K = p->nterms;
    const int nchunk = (K + SPEC_CHUNK - 1) / SPEC_CHUNK;
    std::vector<std::vector<char>> cubins(nchunk);
    std::vector<std::string> errs(nchunk);
    std::vector<int> rcs(nchunk, PCQ_OK);
    std::vector<std::string> names(nchunk);
    for (int j = 0; j < nchunk; ++j) names[j] = std::string(fast ? "pcq_k2s_fast_" : "pcq_k2s_exact_") + std::to_string(j);
    // NVRTC programs are independent: compile the chunks on a few host threads
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const int nthreads = (int)std::min<unsigned>(hw, (unsigned)nchunk);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t] {
            for (int j = t; j < nchunk; j += nthreads) {
                const int k0 = j * SPEC_CHUNK, k1 = std::min(K, k0 + SPEC_CHUNK);
                rcs[j] = compile_chunk(gen_source(p, k0, k1, fast != 0, names[j]), names[j], cubins[j], errs[j]);
            }
        });
    }
    for (auto& th : pool) th.join();
    for (int j = 0; j < nchunk; ++j) {
        if (rcs[j] != PCQ_OK) {
            pcq_set_error("specialize: chunk %d: %s", j, errs[j].c_str());
            return rcs[j];
        }
    }
    pcq_spec* sp = new pcq_spec();
    for (int j = 0; j < nchunk; ++j) {
        cudaLibrary_t lib = nullptr;
        cudaError_t e = cudaLibraryLoadData(&lib, cubins[j].data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
        SpecChunk c;
        if (e == cudaSuccess) {
            sp->libs.push_back(lib);
            e = cudaLibraryGetKernel(&c.kern, lib, names[j].c_str());
        }
        if (e == cudaSuccess) {
            size_t bytes = 0;
            e = cudaLibraryGetGlobal(&c.cfc, &bytes, lib, (names[j] + "_cf").c_str());
        }
        if (e != cudaSuccess) {
            pcq_set_error("specialize: loading chunk %d failed: %s", j, cudaGetErrorString(e));
            pcq_spec_free(sp);
            return PCQ_ERR_CUDA;
        }
        c.k0 = j * SPEC_CHUNK;
        c.k1 = std::min(K, c.k0 + SPEC_CHUNK);
        sp->chunks.push_back(c);
    }
    p->spec[fast].store(sp, std::memory_order_release);
    return PCQ_OK;
}

// y[n*ldy] = sum_k cf[k] Psi_k(x_n) for ONE output; chunks accumulate through y in term order
int pcq_spec_launch(pcq_plan* p, int fast, const double* x, int64_t n, int64_t ldx, const double* cf, double* y,
                    int64_t ldy, cudaStream_t st) {
    pcq_spec* sp = p->spec[fast].load(std::memory_order_acquire);
    if (!sp) return PCQ_ERR_ARG;
    if (n == 0) return PCQ_OK;
    const int64_t nblocks = (n + SPEC_THREADS - 1) / SPEC_THREADS;
    const int64_t cap = (int64_t)p->num_sms * 8;
    const int grid = (int)(nblocks < cap ? nblocks : cap);
    long long nn = n, lx = ldx, ly = ldy;
    for (const SpecChunk& c : sp->chunks) {
        PCQ_CUDA(cudaMemcpyAsync(c.cfc, cf + c.k0, (size_t)(c.k1 - c.k0) * sizeof(double), cudaMemcpyDeviceToDevice, st));
        void* args[] = {(void*)&x, (void*)&nn, (void*)&lx, (void*)&y, (void*)&ly};
        PCQ_CUDA(cudaLaunchKernel((const void*)c.kern, dim3(grid), dim3(SPEC_THREADS), args, 0, st));
        g_pcq_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return PCQ_OK;
}
