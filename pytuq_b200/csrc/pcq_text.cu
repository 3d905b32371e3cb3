// pcq_text.cu — sample files straight into HBM: the text of a numeric table (what numpy.savetxt writes) is copied
// to the device as bytes and converted there, so a fit that reads its samples from files never parses on the host
// and never uploads the binary matrix (reference call sites apps/uqpc/uq_pc.py:232-265).
//
//   scan kernel   one thread per 32 bytes: a blank mask of the span (carry-free word arithmetic), token starts and
//                 first-tokens-of-a-line counted per CTA, '#' at a token start raises a flag (comments are the
//                 host reader's job, pcq_io.h);
//   host          exclusive scan of the per-CTA token counts (a few MB), rows x cols from the totals;
//   parse kernel  same partition, the CTA's text staged in shared memory; a CTA-wide exclusive scan numbers the
//                 tokens, their starts go into one shared list and thread j takes tokens j, j + 256, ...: lexed by
//                 text_convert_token, converted by pcq_decimal_to_double (pcq_decimal.h — the very function the host
//                 library uses) and stored at out[token / cols][token % cols]; the first token of every line must
//                 have index = 0 (mod cols), which together with tokens = rows * cols proves the table is rectangular;
//   fix-up        the few tokens whose rounding one 64 x 128-bit product cannot decide (ties, subnormals, > 19
//                 digits) or that are not decimal literals (inf, nan, junk) are listed by the kernel, converted on
//                 the host with numpy's acceptance rules (libc's strtod) and scattered back.
//
// Both kernels are byte streams over the file: algorithmic traffic = file bytes read once per kernel + 8 bytes
// written per value; the roof is HBM bandwidth (in practice the host -> device copy of the text dominates).
#define PCQIO_POW10_NO_HOST_TABLE
#include "pcq_internal.cuh"
#include "pcq_decimal.h"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <mutex>
#include <new>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <locale.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

__device__ const pcqio_pow10 d_pow10[] = PCQIO_POW10_INIT;

constexpr int TEXT_SPAN = 32;          // bytes per thread
constexpr int TEXT_THREADS = 256;
constexpr int TEXT_CTA_BYTES = TEXT_SPAN * TEXT_THREADS;
constexpr int FIX_CAP = 1 << 20;       // tokens the device may hand to the host per file

struct FixItem {
    long long token;    // index of the value in the table (row-major)
    long long offset;   // byte offset of its text
};

__device__ __forceinline__ bool text_is_space(unsigned char c) {
    return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f';
}
__device__ __forceinline__ bool text_is_eol(unsigned char c) { return c == '\n' || c == '\r'; }
__device__ __forceinline__ bool text_is_numeric(unsigned char c) {
    return (c >= '0' && c <= '9') || c == '.' || c == '-' || c == '+' || c == 'e' || c == 'E';
}

// Is the token that starts at byte i the first one of its line?  (Only blanks between it and the previous line end.)
__device__ __forceinline__ bool text_first_in_line(const unsigned char* __restrict__ text, long long i) {
    long long j = i - 1;
    while (j >= 0 && !text_is_eol(text[j]) && text_is_space(text[j])) --j;
    return j < 0 || text_is_eol(text[j]);
}

// Blank bytes (' ' and 0x09..0x0d) of four characters at a time with plain integer arithmetic: with bit 7 of every
// byte forced on, a byte-wise subtraction cannot borrow from its neighbour, and bit 7 of the difference tells
// "low seven bits >= n".  (The byte-wise __vcmp*4 intrinsics expand to several instructions each on sm_100a;
// profiles/r01_m_text_summary.md.)  Bytes >= 0x80 are never blank.
__device__ __forceinline__ unsigned text_space4(unsigned v) {
    const unsigned a = v | 0x80808080u;
    const unsigned ge9 = a - 0x09090909u, ge14 = a - 0x0e0e0e0eu;
    const unsigned ne20 = ((v ^ 0x20202020u) | 0x80808080u) - 0x01010101u;      // bit 7 clear <=> low bits == 0x20
    const unsigned sp7 = ((ge9 & ~ge14) | ~ne20) & ~v & 0x80808080u;
    return ((sp7 >> 7) * 0x01020408u) >> 24;                                    // byte j -> bit j, no carries
}

// The 32 bytes of this thread -> one bit per blank byte (line ends included).  The buffer is padded with '\n' by a
// whole CTA, so no bounds are needed.
__device__ __forceinline__ unsigned text_space_mask(const uint4 w0, const uint4 w1) {
    return text_space4(w0.x) | (text_space4(w0.y) << 4) | (text_space4(w0.z) << 8) | (text_space4(w0.w) << 12) |
           (text_space4(w1.x) << 16) | (text_space4(w1.y) << 20) | (text_space4(w1.z) << 24) |
           (text_space4(w1.w) << 28);
}

// Token start = non-blank byte behind a blank one.  prev_space: was the byte in front of the span a blank?
__device__ __forceinline__ unsigned text_starts(unsigned space, bool prev_space) {
    return ~space & ((space << 1) | (prev_space ? 1u : 0u));
}

// Device lexer of one token, [+-]digits[.digits][e[+-]digits]: the same grammar and the same (w, q) as the shared
// lexer in pcq_decimal.h, with the up-to-19 significant digits collected nine at a time in 32-bit registers (one IMAD
// per digit instead of a 64-bit multiply-add).  Anything else returns false and goes to the host fix-up, so it may
// decline more than the shared lexer but never accepts differently; the conversion itself is the shared
// pcq_decimal_to_double.
__device__ __forceinline__ bool text_convert_token(const unsigned char* __restrict__ t, int n, double* out) {
    int i = 0;
    unsigned c = t[0];
    const bool neg = c == '-';
    if (c == '-' || c == '+') i = 1;
    unsigned a = 0, b = 0, last = 0;      // significant digits 1-9, 10-18, 19
    int nd = 0, nint = 0, nfrac = 0;
#define TEXT_TAKE_DIGIT(d)                          \
    if (nd | (int)(d)) {                            \
        if (nd < 9) a = a * 10u + (d);              \
        else if (nd < 18) b = b * 10u + (d);        \
        else if (nd == 18) last = (d);              \
        else return false;                          \
        ++nd;                                       \
    }
    for (; i < n; ++i) {
        const unsigned d = (unsigned)t[i] - '0';
        if (d > 9u) break;
        ++nint;
        TEXT_TAKE_DIGIT(d)
    }
    if (i < n && t[i] == '.') {
        for (++i; i < n; ++i) {
            const unsigned d = (unsigned)t[i] - '0';
            if (d > 9u) break;
            ++nfrac;
            TEXT_TAKE_DIGIT(d)
        }
    }
#undef TEXT_TAKE_DIGIT
    if (nint + nfrac == 0) return false;
    int ex = 0;
    if (i < n) {
        if ((t[i] | 0x20u) != 'e') return false;
        ++i;
        bool eneg = false;
        if (i < n && (t[i] == '-' || t[i] == '+')) eneg = t[i++] == '-';
        if (i >= n || n - i > 6) return false;
        for (; i < n; ++i) {
            const unsigned d = (unsigned)t[i] - '0';
            if (d > 9u) return false;
            ex = ex * 10 + (int)d;
        }
        if (eneg) ex = -ex;
    }
    uint64_t w;
    if (nd <= 9) {
        w = a;
    } else {
        unsigned scale = 1;                  // 10^(digits collected in b)
        for (int k = (nd < 18 ? nd : 18) - 9; k > 0; --k) scale *= 10u;
        w = (uint64_t)a * scale + b;
        if (nd == 19) w = w * 10u + last;
    }
    return pcq_decimal_to_double(w, (int64_t)ex - nfrac, neg, out, d_pow10);
}

__global__ void __launch_bounds__(TEXT_THREADS)
pcq_text_scan_kernel(const unsigned char* __restrict__ text, unsigned* __restrict__ cta_tokens,
                     unsigned* __restrict__ cta_rows, int* __restrict__ flags) {
    const long long b0 = ((long long)blockIdx.x * TEXT_THREADS + threadIdx.x) * TEXT_SPAN;
    const uint4* v = reinterpret_cast<const uint4*>(text + b0);
    const unsigned space = text_space_mask(v[0], v[1]);
    const bool prev_space = b0 == 0 || text_is_space(text[b0 - 1]);
    unsigned mask = text_starts(space, prev_space);
    unsigned ntok = __popc(mask), nrow = 0;
    bool comment = false;
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        nrow += text_first_in_line(text, b0 + k) ? 1u : 0u;
        comment |= text[b0 + k] == '#';
    }
    if (comment) atomicOr(&flags[0], 1);
    __shared__ unsigned s_tok[TEXT_THREADS / 32], s_row[TEXT_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ntok += __shfl_xor_sync(0xffffffffu, ntok, o);
        nrow += __shfl_xor_sync(0xffffffffu, nrow, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_tok[threadIdx.x >> 5] = ntok;
        s_row[threadIdx.x >> 5] = nrow;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0, r = 0;
#pragma unroll
        for (int w = 0; w < TEXT_THREADS / 32; ++w) {
            t += s_tok[w];
            r += s_row[w];
        }
        cta_tokens[blockIdx.x] = t;
        cta_rows[blockIdx.x] = r;
    }
}

constexpr int TEXT_TAIL = 64;      // bytes of the next CTA's text staged too: tokens may run over the CTA's end

__global__ void __launch_bounds__(TEXT_THREADS)
pcq_text_parse_kernel(const unsigned char* __restrict__ text, const long long* __restrict__ cta_base, long long cols,
                      double* __restrict__ out, long long ld, FixItem* __restrict__ fix, int* __restrict__ fix_count,
                      int* __restrict__ flags) {
    // the CTA's 8 KB of text (+ tail) in shared memory, and one blank-mask word per 32 bytes of it
    __shared__ __align__(16) unsigned char s_text[TEXT_CTA_BYTES + TEXT_TAIL];
    __shared__ unsigned s_space[TEXT_THREADS + TEXT_TAIL / TEXT_SPAN];
    __shared__ unsigned s_warp[TEXT_THREADS / 32];
    __shared__ unsigned short s_start[TEXT_CTA_BYTES / 2];     // a token and its blank take two bytes at least
    const int tid = threadIdx.x;
    const long long cta0 = (long long)blockIdx.x * TEXT_CTA_BYTES;
    const long long b0 = cta0 + (long long)tid * TEXT_SPAN;
    const uint4* v = reinterpret_cast<const uint4*>(text + b0);
    const uint4 w0 = v[0], w1 = v[1];
    reinterpret_cast<uint4*>(s_text)[2 * tid] = w0;
    reinterpret_cast<uint4*>(s_text)[2 * tid + 1] = w1;
    const unsigned space = text_space_mask(w0, w1);
    s_space[tid] = space;
    if (tid < TEXT_TAIL / TEXT_SPAN) {
        const uint4* tv = reinterpret_cast<const uint4*>(text + cta0 + TEXT_CTA_BYTES + tid * TEXT_SPAN);
        const uint4 t0 = tv[0], t1 = tv[1];
        reinterpret_cast<uint4*>(s_text + TEXT_CTA_BYTES)[2 * tid] = t0;
        reinterpret_cast<uint4*>(s_text + TEXT_CTA_BYTES)[2 * tid + 1] = t1;
        s_space[TEXT_THREADS + tid] = text_space_mask(t0, t1);
    }
    const bool prev_space = b0 == 0 || text_is_space(text[b0 - 1]);
    unsigned mask = text_starts(space, prev_space);
    // exclusive scan of the per-thread token counts over the CTA
    const unsigned ntok = __popc(mask);
    unsigned incl = ntok;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned before = incl - ntok;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    // the CTA's token starts as one list, so that the conversions are dealt out evenly: thread j takes tokens
    // j, j + 256, ... (a span holds one or two tokens; converting them where they were found would make every warp
    // walk its loop twice with most lanes idle the second time) and neighbouring threads store neighbouring values
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        s_start[before++] = (unsigned short)(tid * TEXT_SPAN + k);
    }
    int cta_tokens = 0;
#pragma unroll
    for (int w = 0; w < TEXT_THREADS / 32; ++w) cta_tokens += (int)s_warp[w];
    __syncthreads();
    const long long base = cta_base[blockIdx.x];
    const bool dense = ld == cols;
    for (int t = tid; t < cta_tokens; t += TEXT_THREADS) {
        const int start = s_start[t];
        const long long token = base + t;
        // first token of its line?  Walk back over plain blanks (staged text, then global memory in front of the CTA)
        {
            int j = start - 1;
            while (j >= 0 && !text_is_eol(s_text[j]) && text_is_space(s_text[j])) --j;
            const bool first = j >= 0 ? text_is_eol(s_text[j]) : text_first_in_line(text, cta0);
            if (first && token % cols != 0) atomicOr(&flags[1], 1);
        }
        // end of the token: the next blank, looked up in the mask words of the staged text
        int wi = (start + 1) >> 5;
        unsigned sp = wi < TEXT_THREADS + TEXT_TAIL / TEXT_SPAN ? s_space[wi] & (~0u << ((start + 1) & 31)) : 0u;
        while (!sp && ++wi < TEXT_THREADS + TEXT_TAIL / TEXT_SPAN) sp = s_space[wi];
        double val;
        bool ok;
        if (sp) {
            const int stop = wi * 32 + __ffs(sp) - 1;
            ok = text_convert_token(s_text + start, stop - start, &val);
        } else {        // a token that runs past the staged tail: read it from global memory
            const long long i = cta0 + start;
            size_t n = 1;
            while (!text_is_space(text[i + n])) ++n;       // the '\n' padding behind the file ends the last token
            ok = pcq_parse_decimal_fast(reinterpret_cast<const char*>(text) + i, n, &val, d_pow10);
        }
        if (ok) {
            out[dense ? token : (token / cols) * ld + token % cols] = val;      // dense result: no division
        } else {
            const int slot = atomicAdd(fix_count, 1);
            if (slot < FIX_CAP) fix[slot] = FixItem{token, cta0 + start};
        }
    }
}

__global__ void pcq_text_scatter_kernel(double* __restrict__ out, const long long* __restrict__ index,
                                        const double* __restrict__ value, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[index[j]] = value[j];
}

struct DevGuard {
    int prev = -1;
    bool ok = false;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

constexpr size_t TEXT_STAGE_BYTES = (size_t)64 << 20;

// Reads [off, off + n) of the file into dst with a few host threads (pread: one kernel -> user copy per piece and no
// page faults, unlike copying out of the mapping; one thread moves ~8 GB/s out of the page cache, the link ~50).
bool text_parallel_read(int fd, char* dst, size_t off, size_t n) {
    const int nthr = (int)std::min<size_t>(16, std::max<size_t>(1, std::thread::hardware_concurrency()));
    const size_t part = ((n + nthr - 1) / nthr + 4095) & ~(size_t)4095;
    std::atomic<bool> ok{true};
    auto piece = [&](size_t lo, size_t hi) {
        while (lo < hi) {
            const ssize_t got = pread(fd, dst + lo, hi - lo, (off_t)(off + lo));
            if (got <= 0) {
                if (got < 0 && errno == EINTR) continue;
                ok = false;
                return;
            }
            lo += (size_t)got;
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nthr; ++i) {
        const size_t lo = std::min(n, part * i), hi = std::min(n, part * (i + 1));
        if (hi > lo) pool.emplace_back(piece, lo, hi);
    }
    piece(0, std::min(n, part));
    for (auto& th : pool) th.join();
    return ok.load();
}

// Two pinned staging buffers per process, allocated on first use and kept (cudaHostAlloc costs more than the copy of
// a small file); the mutex serialises uploads of concurrent callers.
struct TextStage {
    std::mutex mu;
    void* pin[2] = {nullptr, nullptr};
};
TextStage g_text_stage;

locale_t text_c_locale() {
    static locale_t loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    return loc;
}

// A token the device handed back -> double with the acceptance rules of numpy's float converter (decimal literals
// through the C locale's correctly rounded strtod; [+-]inf, [+-]infinity, [+-]nan in any case; nothing else: no hex,
// no underscores).  0 = converted, 1 = not a number, 2 = contains '#'.
int text_convert_on_host(const std::string& tok, double* out) {
    if (tok.empty()) return 1;
    bool numeric = true;
    for (char c : tok) {
        if (c == '#') return 2;
        if (!((c >= '0' && c <= '9') || c == '.' || c == '+' || c == '-' || c == 'e' || c == 'E')) numeric = false;
    }
    if (numeric) {
        char* stop = nullptr;
        const double v = strtod_l(tok.c_str(), &stop, text_c_locale());
        if (stop != tok.c_str() + tok.size()) return 1;
        *out = v;
        return 0;
    }
    std::string w;
    for (char c : tok) w.push_back((char)tolower((unsigned char)c));
    bool neg = false;
    if (w[0] == '+' || w[0] == '-') {
        neg = w[0] == '-';
        w.erase(0, 1);
    }
    if (w == "inf" || w == "infinity") {
        *out = neg ? -HUGE_VAL : HUGE_VAL;
        return 0;
    }
    if (w == "nan") {
        *out = neg ? -nan("") : nan("");
        return 0;
    }
    return 1;
}

}  // namespace

struct pcq_text {
    int device = 0;
    const char* map = nullptr;      // the file, mapped on the host (kept for the fix-ups)
    size_t nbytes = 0;
    unsigned char* d_text = nullptr;
    size_t padded = 0;
    long long* d_base = nullptr;    // exclusive scan of the per-CTA token counts
    long long nctas = 0;
    long long rows = 0, cols = 0;
};

extern "C" {

int pcq_text_close(pcq_text* t) {
    if (!t) return PCQ_OK;
    DevGuard guard(t->device);
    if (t->d_text) cudaFree(t->d_text);
    if (t->d_base) cudaFree(t->d_base);
    if (t->map) munmap((void*)t->map, t->nbytes);
    delete t;
    return PCQ_OK;
}

int pcq_text_open(int device, const char* path, pcq_text** out, int64_t* rows, int64_t* cols) {
    if (!path || !out || !rows || !cols || device < 0) {
        pcq_set_error("text_open: bad argument");
        return PCQ_ERR_ARG;
    }
    *out = nullptr;
    DevGuard guard(device);
    if (!guard.ok) { pcq_set_error("text_open: cudaSetDevice(%d) failed", device); return PCQ_ERR_NODEV; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { pcq_set_error("text_open: cannot open %s: %s", path, strerror(errno)); return PCQ_ERR_ARG; }
    struct stat st;
    if (fstat(fd, &st) < 0) { close(fd); pcq_set_error("text_open: cannot stat %s", path); return PCQ_ERR_ARG; }
    pcq_text* t = new (std::nothrow) pcq_text;
    if (!t) { close(fd); pcq_set_error("text_open: out of memory"); return PCQ_ERR_ALLOC; }
    t->device = device;
    t->nbytes = (size_t)st.st_size;
    if (t->nbytes) {
        void* m = mmap(nullptr, t->nbytes, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) {
            close(fd);
            delete t;
            pcq_set_error("text_open: cannot map %s: %s", path, strerror(errno));
            return PCQ_ERR_ARG;
        }
        t->map = (const char*)m;
    }
    *rows = *cols = 0;
    if (!t->nbytes) { close(fd); *out = t; return PCQ_OK; }

    int rc = PCQ_OK;
    unsigned *d_tok = nullptr, *d_row = nullptr;
    int* d_flags = nullptr;
    std::vector<unsigned> h_tok, h_row;
    std::vector<long long> h_base;
    int h_flags[2] = {0, 0};
    long long total = 0, nrows = 0, ncols = 0;
    cudaEvent_t sent[2] = {nullptr, nullptr};
    t->nctas = (long long)((t->nbytes + TEXT_CTA_BYTES - 1) / TEXT_CTA_BYTES);
    t->padded = (size_t)t->nctas * TEXT_CTA_BYTES + TEXT_CTA_BYTES;     // one CTA of '\n' behind the file
#define TEXT_TRY(call)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            pcq_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__);     \
            rc = _e == cudaErrorMemoryAllocation ? PCQ_ERR_ALLOC : PCQ_ERR_CUDA;                           \
            goto done;                                                                                     \
        }                                                                                                  \
    } while (0)
    TEXT_TRY(cudaMalloc(&t->d_text, t->padded));
    TEXT_TRY(cudaMemsetAsync(t->d_text + (t->nbytes / 256) * 256, '\n', t->padded - (t->nbytes / 256) * 256, 0));
    // the text goes over in 64 MB pieces: host threads read piece i+1 from the page cache into one pinned buffer
    // while the DMA engine sends piece i from the other
    {
        std::lock_guard<std::mutex> lock(g_text_stage.mu);
        for (int b = 0; b < 2; ++b) {
            if (!g_text_stage.pin[b]) TEXT_TRY(cudaHostAlloc(&g_text_stage.pin[b], TEXT_STAGE_BYTES, cudaHostAllocDefault));
            if (!sent[b]) TEXT_TRY(cudaEventCreateWithFlags(&sent[b], cudaEventDisableTiming));
        }
        int piece = 0;
        for (size_t off = 0; off < t->nbytes; off += TEXT_STAGE_BYTES, ++piece) {
            const int b = piece & 1;
            const size_t nb = std::min(TEXT_STAGE_BYTES, t->nbytes - off);
            if (piece >= 2) TEXT_TRY(cudaEventSynchronize(sent[b]));
            if (!text_parallel_read(fd, (char*)g_text_stage.pin[b], off, nb)) {
                pcq_set_error("text_open: reading %s failed: %s", path, strerror(errno));
                rc = PCQ_ERR_ARG;
                goto done;
            }
            TEXT_TRY(cudaMemcpyAsync(t->d_text + off, g_text_stage.pin[b], nb, cudaMemcpyHostToDevice, 0));
            TEXT_TRY(cudaEventRecord(sent[b], 0));
        }
        TEXT_TRY(cudaStreamSynchronize(0));      // the staging buffers are free again before the lock goes
    }
    TEXT_TRY(cudaMalloc(&d_tok, (size_t)t->nctas * sizeof(unsigned)));
    TEXT_TRY(cudaMalloc(&d_row, (size_t)t->nctas * sizeof(unsigned)));
    TEXT_TRY(cudaMalloc(&d_flags, 2 * sizeof(int)));
    TEXT_TRY(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), 0));
    pcq_text_scan_kernel<<<(unsigned)t->nctas, TEXT_THREADS>>>(t->d_text, d_tok, d_row, d_flags);
    TEXT_TRY(cudaGetLastError());
    g_pcq_launches.fetch_add(1, std::memory_order_relaxed);
    h_tok.resize((size_t)t->nctas);
    h_row.resize((size_t)t->nctas);
    TEXT_TRY(cudaMemcpy(h_tok.data(), d_tok, (size_t)t->nctas * sizeof(unsigned), cudaMemcpyDeviceToHost));
    TEXT_TRY(cudaMemcpy(h_row.data(), d_row, (size_t)t->nctas * sizeof(unsigned), cudaMemcpyDeviceToHost));
    TEXT_TRY(cudaMemcpy(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost));
    if (h_flags[0]) {
        pcq_set_error("text_open: %s contains '#' comments: use the host reader (pcqio_open)", path);
        rc = PCQ_ERR_UNSUPPORTED;
        goto done;
    }
    h_base.resize((size_t)t->nctas);
    for (long long b = 0; b < t->nctas; ++b) {
        h_base[(size_t)b] = total;
        total += h_tok[(size_t)b];
        nrows += h_row[(size_t)b];
    }
    if (nrows > 0) {
        ncols = total / nrows;
        if (ncols * nrows != total) {
            pcq_set_error("text_open: %s: %lld values in %lld rows: the rows do not all have the same number of "
                          "columns", path, total, nrows);
            rc = PCQ_ERR_ARG;
            goto done;
        }
    }
    TEXT_TRY(cudaMalloc(&t->d_base, (size_t)t->nctas * sizeof(long long)));
    TEXT_TRY(cudaMemcpy(t->d_base, h_base.data(), (size_t)t->nctas * sizeof(long long), cudaMemcpyHostToDevice));
    t->rows = nrows;
    t->cols = ncols;
done:
    close(fd);
    if (rc != PCQ_OK) cudaStreamSynchronize(0);          // nothing may still read the staging buffers
    for (int b = 0; b < 2; ++b)
        if (sent[b]) cudaEventDestroy(sent[b]);
    if (d_tok) cudaFree(d_tok);
    if (d_row) cudaFree(d_row);
    if (d_flags) cudaFree(d_flags);
    if (rc != PCQ_OK) {
        pcq_text_close(t);
        return rc;
    }
    *rows = t->rows;
    *cols = t->cols;
    *out = t;
    return PCQ_OK;
}

int pcq_text_parse_dev(pcq_text* t, double* out_dev, int64_t ld, void* stream) {
    if (!t) { pcq_set_error("text_parse: null handle"); return PCQ_ERR_ARG; }
    if (t->rows == 0) return PCQ_OK;
    if (!out_dev || ld < t->cols) { pcq_set_error("text_parse: bad output buffer"); return PCQ_ERR_ARG; }
    DevGuard guard(t->device);
    if (!guard.ok) { pcq_set_error("text_parse: cudaSetDevice(%d) failed", t->device); return PCQ_ERR_NODEV; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc = PCQ_OK;
    FixItem* d_fix = nullptr;
    int* d_ctl = nullptr;       // [0] fix count, [1..2] flags
    long long* d_index = nullptr;
    double* d_value = nullptr;
    int h_ctl[3] = {0, 0, 0};
    std::vector<FixItem> h_fix;
    std::vector<long long> h_index;
    std::vector<double> h_value;
    TEXT_TRY(cudaMalloc(&d_fix, (size_t)FIX_CAP * sizeof(FixItem)));
    TEXT_TRY(cudaMalloc(&d_ctl, sizeof h_ctl));
    TEXT_TRY(cudaMemsetAsync(d_ctl, 0, sizeof h_ctl, s));
    pcq_text_parse_kernel<<<(unsigned)t->nctas, TEXT_THREADS, 0, s>>>(t->d_text, t->d_base, t->cols, out_dev, ld, d_fix,
                                                                      d_ctl, d_ctl + 1);
    TEXT_TRY(cudaGetLastError());
    g_pcq_launches.fetch_add(1, std::memory_order_relaxed);
    TEXT_TRY(cudaMemcpyAsync(h_ctl, d_ctl, sizeof h_ctl, cudaMemcpyDeviceToHost, s));
    TEXT_TRY(cudaStreamSynchronize(s));
    if (h_ctl[2]) {
        pcq_set_error("text_parse: the rows do not all have the same number of columns (%lld expected)", t->cols);
        rc = PCQ_ERR_ARG;
        goto done;
    }
    if (h_ctl[0] > FIX_CAP) {
        pcq_set_error("text_parse: more than %d values need the host's conversion: use the host reader (pcqio_open)",
                      FIX_CAP);
        rc = PCQ_ERR_UNSUPPORTED;
        goto done;
    }
    if (h_ctl[0] > 0) {
        const int nfix = h_ctl[0];
        h_fix.resize((size_t)nfix);
        TEXT_TRY(cudaMemcpy(h_fix.data(), d_fix, (size_t)nfix * sizeof(FixItem), cudaMemcpyDeviceToHost));
        h_index.resize((size_t)nfix);
        h_value.resize((size_t)nfix);
        for (int j = 0; j < nfix; ++j) {
            const char* p = t->map + h_fix[(size_t)j].offset;
            const char* end = t->map + t->nbytes;
            size_t n = 0;
            while (p + n < end && !isspace((unsigned char)p[n])) ++n;
            std::string tok(p, n);
            double v = 0.0;
            const int verdict = text_convert_on_host(tok, &v);
            if (verdict == 2) {
                pcq_set_error("text_parse: '#' inside the table: comments are the host reader's job (pcqio_open)");
                rc = PCQ_ERR_UNSUPPORTED;
                goto done;
            }
            if (verdict != 0) {
                pcq_set_error("could not convert string '%.60s' to float64 (value %lld of the table)", tok.c_str(),
                              h_fix[(size_t)j].token);
                rc = PCQ_ERR_ARG;
                goto done;
            }
            h_index[(size_t)j] = (h_fix[(size_t)j].token / t->cols) * ld + h_fix[(size_t)j].token % t->cols;
            h_value[(size_t)j] = v;
        }
        TEXT_TRY(cudaMalloc(&d_index, (size_t)nfix * sizeof(long long)));
        TEXT_TRY(cudaMalloc(&d_value, (size_t)nfix * sizeof(double)));
        TEXT_TRY(cudaMemcpyAsync(d_index, h_index.data(), (size_t)nfix * sizeof(long long), cudaMemcpyHostToDevice, s));
        TEXT_TRY(cudaMemcpyAsync(d_value, h_value.data(), (size_t)nfix * sizeof(double), cudaMemcpyHostToDevice, s));
        pcq_text_scatter_kernel<<<(nfix + 255) / 256, 256, 0, s>>>(out_dev, d_index, d_value, nfix);
        TEXT_TRY(cudaGetLastError());
        g_pcq_launches.fetch_add(1, std::memory_order_relaxed);
        TEXT_TRY(cudaStreamSynchronize(s));
    }
done:
    if (d_fix) cudaFree(d_fix);
    if (d_ctl) cudaFree(d_ctl);
    if (d_index) cudaFree(d_index);
    if (d_value) cudaFree(d_value);
    return rc;
}
#undef TEXT_TRY

}  // extern "C"
