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

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

constexpr int SPEC_CHUNK = 512;          // exact mode: terms per kernel (k order kept)
constexpr int SPEC_CHUNK_FAST = 1024;    // fast mode: terms per kernel (trie order)
constexpr int SPEC_THREADS = 128;         // exact mode CTA
constexpr int SPEC_THREADS_FAST = 256;    // fast mode CTA (measured best of 64..512, profiles/r02_k2h_summary.md)

// [k0, k1): positions in the coefficient vector the kernel's constant array is filled from -- the caller's cf in
// exact mode, the trie-ordered staging copy (pcq_spec::d_cfperm) in fast mode
struct SpecChunk { cudaKernel_t kern = nullptr; void* cfc = nullptr; int k0 = 0, k1 = 0; };

}  // namespace

struct pcq_spec {
    std::vector<cudaLibrary_t> libs;
    std::vector<SpecChunk> chunks;
    int spt = 1;                    // samples per thread of the generated kernels
    // fast mode (Horner form over the multi-index trie): term order of the generated code
    int32_t* d_perm = nullptr;      // K: trie position -> term index k
    double* d_cfperm = nullptr;     // K: the coefficients in trie order (staging for the constant arrays)
    int nperm = 0;
    // the constant arrays (and d_cfperm) hold ONE coefficient vector at a time: launches on different streams (the
    // host-pointer evalPC alternates its chunks between the plan's two streams) are ordered through this event
    std::mutex mu;
    cudaEvent_t done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool used = false;
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
    // one CTA per block of samples (no grid-stride loop: ptxas would hoist the loop-invariant coefficient loads into
    // registers and lose a resident CTA; PCQ_SPEC_EXACT_LOOP=1 restores the looping form for comparison)
    if (getenv("PCQ_SPEC_EXACT_LOOP")) {
        o += "  const long long nb = (n + " + std::to_string(SPEC_THREADS - 1) + ") / " + std::to_string(SPEC_THREADS) + ";\n";
        o += "  for (long long blk = blockIdx.x; blk < nb; blk += gridDim.x) {\n";
    } else {
        o += "  {\n    const long long blk = blockIdx.x;\n";
    }
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


// ---- fast mode: Horner form over the multi-index trie -----------------------------------------
// y = sum_k c_k prod_i phi_{alpha_ki}(x_i).  Order the non-zero (dimension, order) factors of every term by
// dimension and insert the terms into a trie: a node is a prefix of factors, value(node) = c_node +
// sum_children phi_child * value(child), y = value(root).  Evaluated bottom-up this costs ONE fused multiply-add
// per term (an edge of the trie) instead of the (#factors) multiplies + 1 add of the term-by-term form, the
// coefficient being a constant-bank operand of the DFMA; independent subtrees give the scheduler its ILP.
// The summation order differs from the reference's (fast mode: rounding-level differences only).  A chunk is a
// run of SPEC_CHUNK_FAST terms in trie (lexicographic) order with its own induced trie, so prefixes shared with
// another chunk cost one extra DFMA each.
typedef std::vector<std::pair<int, int>> Factors;     // (dimension, order > 0), ascending dimension

struct TrieNode {
    std::vector<std::pair<std::pair<int, int>, int>> children;   // (factor, node index), insertion order
    std::vector<int> coefs;                                       // slots of the chunk's constant array
};

struct HornerGen {
    std::vector<TrieNode> nodes;
    std::string o;
    std::string name;
    int uid = 0;
    int spt = 1;                      // samples per thread: every per-sample statement is emitted spt times
    const char* sfx[4] = {"a", "b", "c", "d"};

    int child(int nd, std::pair<int, int> key) {
        for (auto& c : nodes[nd].children)
            if (c.first == key) return c.second;
        nodes.emplace_back();
        const int id = (int)nodes.size() - 1;
        nodes[nd].children.push_back({key, id});
        return id;
    }
    std::string cexpr(int nd) const {
        std::string e;
        for (size_t j = 0; j < nodes[nd].coefs.size(); ++j)
            e += (j ? " + " : "") + name + "_cf[" + std::to_string(nodes[nd].coefs[j]) + "]";
        return nodes[nd].coefs.size() > 1 ? "(" + e + ")" : e;
    }
    bool leaf(int nd) const { return nodes[nd].children.empty(); }
    // Emits the statements that compute value(nd) into fresh variables (one per sample) and returns their stem.
    // Every FP64 instruction gets at most ONE constant-bank operand: a DFMA with two (t * c_child + c_node) makes
    // ptxas fetch one of them with a long-latency LDC right in front of its use (the top stall of the first
    // version, profiles/r02_k2h_summary.md).  So a node's own coefficient is the addend of the first child that
    // has a register operand (an inner child), or, when all children are leaves, is added after their sum.
    std::string emit(int nd) {
        const std::string var = "v" + std::to_string(++uid);
        const TrieNode& N = nodes[nd];
        std::vector<int> order;                              // inner children first
        for (size_t c = 0; c < N.children.size(); ++c) if (!leaf(N.children[c].second)) order.push_back((int)c);
        const bool inner_first = !order.empty();
        for (size_t c = 0; c < N.children.size(); ++c) if (leaf(N.children[c].second)) order.push_back((int)c);
        bool started = false;
        for (int ci : order) {
            const auto& c = N.children[ci];
            const std::string t = "t" + std::to_string(c.first.first) + "_" + std::to_string(c.first.second);
            const bool lf = leaf(c.second);
            const std::string operand = lf ? cexpr(c.second) : emit(c.second);
            for (int q = 0; q < spt; ++q) {
                const std::string opq = lf ? operand : operand + sfx[q];
                if (started) o += "    " + var + sfx[q] + " = fma(" + t + sfx[q] + ", " + opq + ", " + var + sfx[q] + ");\n";
                else if (inner_first && !N.coefs.empty())
                    o += "    double " + var + sfx[q] + " = fma(" + t + sfx[q] + ", " + opq + ", " + cexpr(nd) + ");\n";
                else o += "    double " + var + sfx[q] + " = " + t + sfx[q] + " * " + opq + ";\n";
            }
            started = true;
        }
        if (!N.coefs.empty() && !(inner_first && started)) {
            for (int q = 0; q < spt; ++q) {
                if (started) o += "    " + var + sfx[q] + " = " + var + sfx[q] + " + " + cexpr(nd) + ";\n";
                else o += "    double " + var + sfx[q] + " = " + cexpr(nd) + ";\n";
            }
            started = true;
        }
        if (!started)
            for (int q = 0; q < spt; ++q) o += "    double " + var + sfx[q] + " = 0.0;\n";
        return var;
    }
};

Factors term_factors(const pcq_plan* p, int k) {
    Factors f;
    for (int i = 0; i < p->sdim; ++i) {
        const int od = p->h_mi[(size_t)k * p->sdim + i];
        if (od > 0) f.push_back({i, od});
    }
    return f;
}

// trie (lexicographic) order of the terms: perm[j] = term index k
std::vector<int32_t> horner_order(const pcq_plan* p) {
    const int K = p->nterms;
    std::vector<Factors> fs(K);
    for (int k = 0; k < K; ++k) fs[k] = term_factors(p, k);
    std::vector<int32_t> perm(K);
    for (int k = 0; k < K; ++k) perm[k] = k;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return fs[a] < fs[b]; });
    return perm;
}

// source of the fast kernel for trie positions [j0, j1) of `perm`
std::string gen_source_horner(const pcq_plan* p, const std::vector<int32_t>& perm, int j0, int j1, const std::string& name,
                              int spt) {
    const int s = p->sdim;
    HornerGen G;
    G.name = name;
    G.spt = spt;
    G.nodes.emplace_back();                     // root
    std::vector<int> top(s, 0);
    for (int j = j0; j < j1; ++j) {
        const Factors f = term_factors(p, perm[j]);
        int nd = 0;
        for (const auto& key : f) {
            nd = G.child(nd, key);
            top[key.first] = std::max(top[key.first], key.second);
        }
        G.nodes[nd].coefs.push_back(j - j0);
    }
    std::string& o = G.o;
    o.reserve((size_t)(j1 - j0) * 64 * spt + 4096);
    const int per = SPEC_THREADS_FAST * spt;
    o += "__constant__ double " + name + "_cf[" + std::to_string(j1 - j0) + "];\n";
    o += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(SPEC_THREADS_FAST) + ") " + name +
         "(const double* __restrict__ x, long long n, long long ldx, double* __restrict__ y, long long ldy) {\n";
    // One CTA per block of samples, NO grid-stride loop: inside a loop ptxas hoists the (loop-invariant) coefficient
    // loads into registers -- 117 instead of 94 registers, one CTA less per SM, 2.04 -> 1.44 ms at C2.
    o += "  {\n    const long long blk = blockIdx.x;\n";
    for (int q = 0; q < spt; ++q) {
        const std::string z = G.sfx[q];
        o += "    const long long i" + z + " = blk * " + std::to_string(per) + " + " + std::to_string(q * SPEC_THREADS_FAST) + " + threadIdx.x;\n";
        o += "    const bool live" + z + " = i" + z + " < n;\n";
        o += "    const double* xr" + z + " = x + (live" + z + " ? i" + z + " : 0) * ldx;\n";
    }
    for (int i = 0; i < s; ++i) {
        if (top[i] == 0) continue;
        for (int q = 0; q < spt; ++q) {
            const std::string z = G.sfx[q];
            const std::string xi = "x" + std::to_string(i) + z, t = "t" + std::to_string(i) + "_";
            o += "    const double " + xi + " = xr" + z + "[" + std::to_string(i) + "];\n";
            // PC1d.__call__ (rv/pcrv.py:794-802) with the multiply-add fused: p_{n+1} = fma(a_n x, p_n, b_n p_{n-1})
            o += "    const double " + t + "1" + z + " = " + xi + " * " + hexd(p->h_p1[i]) + ";\n";
            for (int nn = 1; nn < top[i]; ++nn) {
                const std::string pm = (nn == 1) ? std::string("1.0") : t + std::to_string(nn - 1) + z;
                o += "    const double " + t + std::to_string(nn + 1) + z + " = fma(" + hexd(p->h_rec_a[(size_t)nn * s + i]) + " * " + xi +
                     ", " + t + std::to_string(nn) + z + ", " + hexd(p->h_rec_b[(size_t)nn * s + i]) + " * " + pm + ");\n";
            }
        }
    }
    const std::string root = G.emit(0);
    for (int q = 0; q < spt; ++q) {
        const std::string z = G.sfx[q];
        o += "    if (live" + z + ") y[i" + z + " * ldy] = " + (j0 == 0 ? std::string() : "y[i" + z + " * ldy] + ") + root + z + ";\n";
    }
    o += "  }\n}\n";
    return o;
}

__global__ void __launch_bounds__(256) pcq_spec_gather_kernel(const double* __restrict__ cf, const int32_t* __restrict__ perm,
                                                              int K, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < K) out[j] = cf[perm[j]];
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
    cudaFree(sp->d_perm);
    cudaFree(sp->d_cfperm);
    if (sp->done) cudaEventDestroy(sp->done);
    delete sp;
}

int pcq_spec_build(pcq_plan* p, int fast) {
    if (p->spec[fast].load(std::memory_order_acquire)) return PCQ_OK;
    if (!nvrtc().ok) {
        pcq_set_error("specialize: libnvrtc.so.12 not found (the generic kernels stay in use)");
        return PCQ_ERR_UNSUPPORTED;
    }
    const int K = p->nterms;
    const bool horner = fast != 0 && !getenv("PCQ_SPEC_NO_HORNER");
    const int CH = horner ? SPEC_CHUNK_FAST : SPEC_CHUNK;
    int spt = 1;
    if (horner) { const char* e = getenv("PCQ_SPEC_SPT"); if (e && atoi(e) >= 1 && atoi(e) <= 4) spt = atoi(e); }
    const int nchunk = (K + CH - 1) / CH;
    std::vector<int32_t> perm;
    if (horner) perm = horner_order(p);
    std::vector<std::vector<char>> cubins(nchunk);
    std::vector<std::string> errs(nchunk);
    std::vector<int> rcs(nchunk, PCQ_OK);
    std::vector<std::string> names(nchunk);
    for (int j = 0; j < nchunk; ++j)
        names[j] = std::string(horner ? "pcq_k2h_fast_" : fast ? "pcq_k2s_fast_" : "pcq_k2s_exact_") + std::to_string(j);
    // NVRTC programs are independent: compile the chunks on a few host threads
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const int nthreads = (int)std::min<unsigned>(hw, (unsigned)nchunk);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t] {
            for (int j = t; j < nchunk; j += nthreads) {
                const int k0 = j * CH, k1 = std::min(K, k0 + CH);
                rcs[j] = compile_chunk(horner ? gen_source_horner(p, perm, k0, k1, names[j], spt)
                                              : gen_source(p, k0, k1, fast != 0, names[j]), names[j], cubins[j], errs[j]);
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
    sp->spt = spt;
    if (horner) {
        sp->nperm = K;
        if (cudaMalloc((void**)&sp->d_perm, (size_t)K * sizeof(int32_t)) != cudaSuccess ||
            cudaMalloc((void**)&sp->d_cfperm, (size_t)K * sizeof(double)) != cudaSuccess ||
            cudaMemcpy(sp->d_perm, perm.data(), (size_t)K * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            pcq_set_error("specialize: device allocation of the trie order failed");
            pcq_spec_free(sp);
            return PCQ_ERR_ALLOC;
        }
    }
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
        c.k0 = j * CH;
        c.k1 = std::min(K, c.k0 + CH);
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
    const bool horner = sp->d_perm != nullptr;
    const int threads = horner ? SPEC_THREADS_FAST : SPEC_THREADS;
    const int64_t nblocks = (n + (int64_t)threads * sp->spt - 1) / ((int64_t)threads * sp->spt);
    const int64_t cap = (int64_t)p->num_sms * 8;
    // Horner kernels: one CTA per sample block (no grid-stride loop in the generated code)
    const int64_t grid64 = (horner || !getenv("PCQ_SPEC_EXACT_LOOP")) ? nblocks : (nblocks < cap ? nblocks : cap);
    if (grid64 > 0x7fffffffLL) { pcq_set_error("specialised evalPC: too many samples in one call"); return PCQ_ERR_ARG; }
    const int grid = (int)grid64;
    long long nn = n, lx = ldx, ly = ldy;
    std::lock_guard<std::mutex> lock(sp->mu);
    if (!sp->done) PCQ_CUDA(cudaEventCreateWithFlags(&sp->done, cudaEventDisableTiming));
    if (sp->used && sp->last_stream != st) PCQ_CUDA(cudaStreamWaitEvent(st, sp->done, 0));
    if (sp->d_perm) {        // Horner kernels take their coefficients in trie order
        pcq_spec_gather_kernel<<<(sp->nperm + 255) / 256, 256, 0, st>>>(cf, sp->d_perm, sp->nperm, sp->d_cfperm);
        PCQ_LAUNCH_CHECK("pcq_spec_gather_kernel");
        cf = sp->d_cfperm;
    }
    for (const SpecChunk& c : sp->chunks) {
        PCQ_CUDA(cudaMemcpyAsync(c.cfc, cf + c.k0, (size_t)(c.k1 - c.k0) * sizeof(double), cudaMemcpyDeviceToDevice, st));
        void* args[] = {(void*)&x, (void*)&nn, (void*)&lx, (void*)&y, (void*)&ly};
        PCQ_CUDA(cudaLaunchKernel((const void*)c.kern, dim3(grid), dim3(threads), args, 0, st));
        g_pcq_launches.fetch_add(1, std::memory_order_relaxed);
    }
    PCQ_CUDA(cudaEventRecord(sp->done, st));
    sp->last_stream = st;
    sp->used = true;
    return PCQ_OK;
}

// Debugging aid (no device needed): the CUDA source pcq_plan_specialize would compile for trie positions [j0, j1) of
// the fast mode (Horner form), for inspection of the generated code / its SASS in a GPU-less container.
extern "C" int64_t pcq_debug_horner_source(int sdim, int nterms, const int32_t* mi, int pmax, const double* rec_a,
                                           const double* rec_b, const double* p1_scale, int j0, int j1, int spt,
                                           char* buf, int64_t bufsize) {
    if (sdim <= 0 || nterms <= 0 || !mi || !rec_a || !rec_b || !p1_scale || j0 < 0 || j1 > nterms || j0 >= j1 || spt < 1 || spt > 4)
        return PCQ_ERR_ARG;
    pcq_plan P;
    P.sdim = sdim; P.nterms = nterms; P.pmax = pmax;
    P.h_mi.assign(mi, mi + (size_t)nterms * sdim);
    P.h_rec_a.assign(rec_a, rec_a + (size_t)(pmax + 1) * sdim);
    P.h_rec_b.assign(rec_b, rec_b + (size_t)(pmax + 1) * sdim);
    P.h_p1.assign(p1_scale, p1_scale + sdim);
    const std::string src = gen_source_horner(&P, horner_order(&P), j0, j1, "dbg", spt);
    if (buf && bufsize > 0) {
        const size_t nb = std::min((size_t)bufsize - 1, src.size());
        memcpy(buf, src.data(), nb);
        buf[nb] = 0;
    }
    return (int64_t)src.size();
}
