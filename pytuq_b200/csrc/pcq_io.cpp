// libpcq_io: numpy.savetxt / numpy.loadtxt (default options) for float64 matrices on all host cores.
// See include/pcq_io.h for the contract and the reference call sites (apps/uqpc/uq_pc.py:202-271).
#include "pcq_io.h"
#include "pcq_io_pow10.h"
#include "pcq_decimal.h"

#include <atomic>
#include <cerrno>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <locale.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

thread_local std::string g_err;

int fail(int status, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return status;
}

int pick_threads(int threads, int64_t units) {
    if (threads <= 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        threads = n > 0 ? (int)n : 1;
    }
    if (threads > 256) threads = 256;
    if (units < threads) threads = units > 0 ? (int)units : 1;
    return threads;
}

// The C locale, whatever LC_NUMERIC the embedding interpreter has switched to.
locale_t c_locale() {
    static locale_t loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    return loc;
}

// ---------------------------------------------------------------------------------------- writing

std::atomic<int64_t> g_slow_formats{0}, g_slow_parses{0};   // conversions that took the libc route

// '%.18e' % v the way libc prints it in the C locale (correctly rounded, exponent of at least two digits),
// which is also what CPython's '%.18e' % v gives, except that Python never signs a nan.
inline int format_value_libc(char* dst, double v) {
    if (std::isnan(v)) {
        memcpy(dst, "nan", 3);
        return 3;
    }
    return snprintf(dst, 40, "%.18e", v);
}

// The same 19 significant digits without libc's multi-precision arithmetic: v * 10^(18-k) from one
// 64 x 128-bit product with the truncated table entry; the product is short by less than 2^64 units of its
// last limb, i.e. by < 2^-61 of the integer it is cut to, so the rounding direction is certain unless the
// fraction lies within 2^-60 of one half — then (exact ties included) the libc route decides.
inline int format_value(char* dst, double v) {
    uint64_t bits;
    memcpy(&bits, &v, 8);
    const int biased = (int)((bits >> 52) & 0x7ff);
    uint64_t frac = bits & ((1ull << 52) - 1);
    if (biased == 0x7ff) return format_value_libc(dst, v);
    char* p = dst;
    if (bits >> 63) *p++ = '-';
    if (biased == 0 && frac == 0) {
        memcpy(p, "0.000000000000000000e+00", 24);
        return (int)(p - dst) + 24;
    }
    int e2 = biased ? biased - 1075 : -1074;
    uint64_t m = biased ? (frac | (1ull << 52)) : frac;
    const int lz = __builtin_clzll(m);
    m <<= lz;
    e2 -= lz;                                   // v = m * 2^e2 with the top bit of m set
    int k = (int)(((int64_t)(e2 + 63) * 315653) >> 20);   // ~ floor(log10 v), corrected below
    uint64_t D = 0, y2 = 0;
    for (int tries = 0;; ++tries) {
        const int s10 = 18 - k;
        if (s10 < PCQIO_Q_MIN || s10 > PCQIO_Q_MAX || tries > 3) {
            g_slow_formats.fetch_add(1, std::memory_order_relaxed);
            return format_value_libc(dst, v);
        }
        const pcqio_pow10& P = PCQIO_POW10[s10 - PCQIO_Q_MIN];
        const pcq_u192 r = pcq_mul_64x128(m, P.hi, P.lo);
        const int sh = -(e2 + P.e);             // v * 10^s10 = r / 2^sh
        if (sh >= 129 && sh <= 191) {
            const int t = sh - 128;
            D = r.r2 >> t;
            y2 = (r.r2 << (64 - t)) | (r.r1 >> t);
        } else if (sh == 128) {
            D = r.r2;
            y2 = r.r1;
        } else if (sh >= 125 && sh < 128) {
            const int u = sh - 64;
            if (r.r2 >> u) { ++k; continue; }   // more than 64 bits: k was too small
            D = (r.r2 << (64 - u)) | (r.r1 >> u);
            y2 = (r.r1 << (64 - u)) | (r.r0 >> u);
        } else if (sh > 191) { --k; continue; }
        else { ++k; continue; }
        if (D >= 10000000000000000000ull) { ++k; continue; }
        if (D < 1000000000000000000ull) { --k; continue; }
        break;
    }
    const uint64_t half = 1ull << 63;
    if (y2 > half + 16) ++D;
    else if (y2 >= half - 16) {
        g_slow_formats.fetch_add(1, std::memory_order_relaxed);
        return format_value_libc(dst, v);
    }
    if (D >= 10000000000000000000ull) {         // 9.99..9 rounded up
        D = 1000000000000000000ull;
        ++k;
    }
    char digits[19];
    for (int i = 18; i >= 0; --i) {
        digits[i] = (char)('0' + D % 10);
        D /= 10;
    }
    *p++ = digits[0];
    *p++ = '.';
    memcpy(p, digits + 1, 18);
    p += 18;
    *p++ = 'e';
    *p++ = k < 0 ? '-' : '+';
    const int ak = k < 0 ? -k : k;
    if (ak >= 100) *p++ = (char)('0' + ak / 100);
    *p++ = (char)('0' + (ak / 10) % 10);
    *p++ = (char)('0' + ak % 10);
    return (int)(p - dst);
}

struct WriteJob {
    const double* a;
    int64_t rows, cols, ld, rows_per_block, nblocks;
    int fd;
    std::atomic<int64_t> next{0};
    std::mutex mu;
    std::condition_variable cv;
    int64_t written = 0;   // blocks appended so far (guarded by mu)
    int error = 0;         // errno of the first failed write
};

void write_worker(WriteJob* job) {
    locale_t old = uselocale(c_locale());
    std::vector<char> buf;
    for (;;) {
        const int64_t b = job->next.fetch_add(1);
        if (b >= job->nblocks) break;
        const int64_t r0 = b * job->rows_per_block;
        const int64_t r1 = std::min(job->rows, r0 + job->rows_per_block);
        buf.resize((size_t)((r1 - r0) * (job->cols * 28 + 1)) + 64);
        char* p = buf.data();
        for (int64_t r = r0; r < r1; ++r) {
            const double* row = job->a + r * job->ld;
            for (int64_t c = 0; c < job->cols; ++c) {
                if (c) *p++ = ' ';
                p += format_value(p, row[c]);
            }
            *p++ = '\n';
        }
        std::unique_lock<std::mutex> lock(job->mu);
        job->cv.wait(lock, [&] { return job->written == b; });
        if (!job->error) {
            const char* q = buf.data();
            size_t left = (size_t)(p - buf.data());
            while (left) {
                ssize_t w = write(job->fd, q, left);
                if (w < 0) {
                    if (errno == EINTR) continue;
                    job->error = errno;
                    break;
                }
                q += w;
                left -= (size_t)w;
            }
        }
        job->written = b + 1;
        lock.unlock();
        job->cv.notify_all();
    }
    uselocale(old);
}

// ---------------------------------------------------------------------------------------- reading

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f'; }
inline bool is_eol(char c) { return c == '\n' || c == '\r'; }

struct Segment {
    const char* begin;
    const char* end;
    int64_t rows = 0;       // data rows in the segment
    int64_t cols = -1;      // tokens of its first data row (-1: none)
    int64_t row0 = 0;       // index of its first data row in the file
};

// Tokens of the line starting at p; returns the position just behind the line terminator.
inline const char* count_line(const char* p, const char* end, int64_t* ntok) {
    int64_t n = 0;
    bool in_tok = false, comment = false;
    for (; p < end; ++p) {
        const char c = *p;
        if (is_eol(c)) { ++p; break; }
        if (comment) continue;
        if (c == '#') { comment = true; in_tok = false; continue; }
        if (is_blank(c)) { in_tok = false; continue; }
        if (!in_tok) { in_tok = true; ++n; }
    }
    *ntok = n;
    return p;
}

void scan_segment(Segment* s) {
    const char* p = s->begin;
    while (p < s->end) {
        // the usual line starts with a value: only its end has to be found (files written by savetxt never use
        // a lone CR, a line that contains one is counted the slow way)
        const char c0 = *p;
        if (s->cols >= 0 && !is_blank(c0) && !is_eol(c0) && c0 != '#') {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(s->end - p));
            const char* stop = nl ? nl + 1 : s->end;
            if (!memchr(p, '\r', (size_t)(stop - p)) || (nl && nl > p && nl[-1] == '\r' &&
                                                          !memchr(p, '\r', (size_t)(nl - 1 - p)))) {
                ++s->rows;
                p = stop;
                continue;
            }
        }
        int64_t ntok;
        p = count_line(p, s->end, &ntok);
        if (ntok) {
            if (s->cols < 0) s->cols = ntok;
            ++s->rows;
        }
    }
}

bool ieq(const char* t, size_t n, const char* word) {
    if (strlen(word) != n) return false;
    for (size_t i = 0; i < n; ++i) {
        char c = t[i];
        if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
        if (c != word[i]) return false;
    }
    return true;
}

// Correctly rounded decimal -> binary64 by libc (the C locale's strtod), for everything the fast route declines.
bool parse_decimal_libc(const char* t, size_t n, double* out, locale_t loc) {
    char small[64];
    std::string big;
    const char* z;
    if (n < sizeof small) {
        memcpy(small, t, n);
        small[n] = 0;
        z = small;
    } else {
        big.assign(t, n);
        z = big.c_str();
    }
    char* stop = nullptr;
    const double v = strtod_l(z, &stop, loc);
    if (stop != z + n || stop == z) return false;
    *out = v;
    return true;
}

// One token -> double with the acceptance rules of numpy's float converter (PyOS_string_to_double:
// decimal literals, [+-]inf, [+-]infinity, [+-]nan, any case; no hex, no underscores).
bool parse_token(const char* t, size_t n, double* out, locale_t loc) {
    if (pcq_parse_decimal_fast(t, n, out, PCQIO_POW10)) return true;
    bool numeric = true;
    for (size_t i = 0; i < n; ++i) {
        const char c = t[i];
        if (!((c >= '0' && c <= '9') || c == '.' || c == '+' || c == '-' || c == 'e' || c == 'E')) {
            numeric = false;
            break;
        }
    }
    if (numeric) {
        g_slow_parses.fetch_add(1, std::memory_order_relaxed);
        return parse_decimal_libc(t, n, out, loc);
    }
    bool neg = false;
    if (n && (t[0] == '+' || t[0] == '-')) {
        neg = t[0] == '-';
        ++t;
        --n;
    }
    if (ieq(t, n, "inf") || ieq(t, n, "infinity")) {
        *out = neg ? -HUGE_VAL : HUGE_VAL;
        return true;
    }
    if (ieq(t, n, "nan")) {
        *out = neg ? -std::nan("") : std::nan("");
        return true;
    }
    return false;
}

struct ParseError {
    int status = 0;
    const char* where = nullptr;
    int64_t found = 0;
};

void parse_segment(const Segment* s, double* out, int64_t ld, int64_t cols, ParseError* err) {
    locale_t loc = c_locale();
    const char* p = s->begin;
    double* row = out + s->row0 * ld;
    while (p < s->end) {
        const char* line = p;
        int64_t c = 0;
        bool comment = false;
        while (p < s->end) {
            const char ch = *p;
            if (is_eol(ch)) { ++p; break; }
            if (comment || is_blank(ch)) { ++p; continue; }
            if (ch == '#') { comment = true; ++p; continue; }
            const char* t = p;
            while (p < s->end && !is_blank(*p) && !is_eol(*p) && *p != '#') ++p;
            if (c < cols) {
                if (!parse_token(t, (size_t)(p - t), row + c, loc)) {
                    *err = ParseError{PCQIO_ERR_PARSE, t, 0};
                    return;
                }
            }
            ++c;
        }
        if (c) {
            if (c != cols) {
                *err = ParseError{PCQIO_ERR_RAGGED, line, c};
                return;
            }
            row += ld;
        }
    }
}

int64_t line_number(const char* base, const char* where) {
    int64_t n = 1;
    for (const char* p = base; p < where; ++p) {
        if (*p == '\n') ++n;
        else if (*p == '\r' && (p + 1 >= where || p[1] != '\n')) ++n;
    }
    return n;
}

template <class F>
void run_parallel(int threads, int64_t nitems, F&& body) {
    std::atomic<int64_t> next{0};
    auto loop = [&] {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= nitems) break;
            body(i);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(loop);
    loop();
    for (auto& th : pool) th.join();
}

}  // namespace

struct pcqio_reader {
    const char* base = nullptr;
    size_t size = 0;
    int threads = 1;
    int64_t rows = 0, cols = 0;
    std::vector<Segment> segs;
};

extern "C" {

int pcqio_version(void) { return PCQIO_VERSION; }

const char* pcqio_last_error(void) { return g_err.c_str(); }

int pcqio_savetxt(const char* path, const double* a, int64_t rows, int64_t cols, int64_t ld, int threads) {
    try {
        if (!path || rows < 0 || cols < 0 || (rows > 0 && cols > 0 && (!a || ld < cols)))
            return fail(PCQIO_ERR_ARG, "pcqio_savetxt: bad argument (rows=%lld cols=%lld ld=%lld)",
                        (long long)rows, (long long)cols, (long long)ld);
        const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) return fail(PCQIO_ERR_IO, "pcqio_savetxt: cannot open %s: %s", path, strerror(errno));
        WriteJob job;
        job.a = a;
        job.rows = rows;
        job.cols = cols;
        job.ld = ld;
        job.fd = fd;
        // ~2 MB of text per block: long enough to amortise the hand-over, short enough to balance
        const int64_t per_row = cols * 26 + 1;
        job.rows_per_block = std::max<int64_t>(1, (2 << 20) / per_row);
        job.nblocks = (rows + job.rows_per_block - 1) / job.rows_per_block;
        const int nthr = pick_threads(threads, job.nblocks);
        std::vector<std::thread> pool;
        for (int t = 1; t < nthr; ++t) pool.emplace_back(write_worker, &job);
        write_worker(&job);
        for (auto& th : pool) th.join();
        const int cerr = close(fd) < 0 ? errno : 0;
        if (job.error || cerr)
            return fail(PCQIO_ERR_IO, "pcqio_savetxt: writing %s failed: %s", path,
                        strerror(job.error ? job.error : cerr));
        return PCQIO_OK;
    } catch (const std::bad_alloc&) {
        return fail(PCQIO_ERR_ALLOC, "pcqio_savetxt: out of memory");
    } catch (const std::exception& e) {
        return fail(PCQIO_ERR_IO, "pcqio_savetxt: %s", e.what());
    }
}

int pcqio_open(const char* path, pcqio_reader** reader, int64_t* rows, int64_t* cols, int threads) {
    try {
        if (!path || !reader || !rows || !cols) return fail(PCQIO_ERR_ARG, "pcqio_open: null argument");
        *reader = nullptr;
        const int fd = open(path, O_RDONLY);
        if (fd < 0) return fail(PCQIO_ERR_IO, "pcqio_open: cannot open %s: %s", path, strerror(errno));
        struct stat st;
        if (fstat(fd, &st) < 0) {
            const int e = errno;
            close(fd);
            return fail(PCQIO_ERR_IO, "pcqio_open: cannot stat %s: %s", path, strerror(e));
        }
        pcqio_reader* rd = new pcqio_reader;
        rd->size = (size_t)st.st_size;
        if (rd->size) {
            void* m = mmap(nullptr, rd->size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) {
                const int e = errno;
                close(fd);
                delete rd;
                return fail(PCQIO_ERR_IO, "pcqio_open: cannot map %s: %s", path, strerror(e));
            }
            madvise(m, rd->size, MADV_SEQUENTIAL);
            rd->base = (const char*)m;
        }
        close(fd);
        // segments of >= 1 MB cut at line starts, a few per thread so that uneven lines balance
        const int64_t want = std::max<int64_t>(1, std::min<int64_t>((int64_t)(rd->size >> 20),
                                                                    4 * (int64_t)pick_threads(threads, 1 << 20)));
        const char* end = rd->base + rd->size;
        const char* p = rd->base;
        for (int64_t i = 0; i < want && p < end; ++i) {
            const char* q = i + 1 == want ? end : rd->base + (rd->size / want) * (i + 1);
            if (q < p) q = p;
            while (q < end && q > rd->base && !is_eol(q[-1])) ++q;   // move to the next line start
            if (q > p) {
                Segment s;
                s.begin = p;
                s.end = q;
                rd->segs.push_back(s);
            }
            p = q;
        }
        rd->threads = pick_threads(threads, (int64_t)rd->segs.size());
        run_parallel(rd->threads, (int64_t)rd->segs.size(), [&](int64_t i) { scan_segment(&rd->segs[i]); });
        int64_t total = 0, ncol = -1;
        for (auto& s : rd->segs) {
            s.row0 = total;
            total += s.rows;
            if (s.cols >= 0) {
                if (ncol < 0) ncol = s.cols;
                else if (ncol != s.cols) {
                    const int64_t ln = line_number(rd->base, s.begin);
                    pcqio_close(rd);
                    return fail(PCQIO_ERR_RAGGED,
                                "pcqio_open: %s: the number of columns changed from %lld to %lld near line %lld",
                                path, (long long)ncol, (long long)s.cols, (long long)ln);
                }
            }
        }
        rd->rows = total;
        rd->cols = ncol < 0 ? 0 : ncol;
        *rows = rd->rows;
        *cols = rd->cols;
        *reader = rd;
        return PCQIO_OK;
    } catch (const std::bad_alloc&) {
        return fail(PCQIO_ERR_ALLOC, "pcqio_open: out of memory");
    } catch (const std::exception& e) {
        return fail(PCQIO_ERR_IO, "pcqio_open: %s", e.what());
    }
}

int pcqio_read(pcqio_reader* rd, double* out, int64_t ld) {
    try {
        if (!rd) return fail(PCQIO_ERR_ARG, "pcqio_read: null reader");
        if (rd->rows == 0) return PCQIO_OK;
        if (!out || ld < rd->cols)
            return fail(PCQIO_ERR_ARG, "pcqio_read: bad output buffer (ld=%lld, cols=%lld)", (long long)ld,
                        (long long)rd->cols);
        std::vector<ParseError> errs(rd->segs.size());
        run_parallel(rd->threads, (int64_t)rd->segs.size(),
                     [&](int64_t i) { parse_segment(&rd->segs[i], out, ld, rd->cols, &errs[i]); });
        for (const auto& e : errs) {
            if (!e.status) continue;
            const int64_t ln = line_number(rd->base, e.where);
            if (e.status == PCQIO_ERR_RAGGED)
                return fail(PCQIO_ERR_RAGGED,
                            "the number of columns changed from %lld to %lld at row %lld; use `usecols` to select "
                            "a subset and avoid this error",
                            (long long)rd->cols, (long long)e.found, (long long)ln);
            const char* t = e.where;
            const char* end = rd->base + rd->size;
            size_t n = 0;
            while (t + n < end && !is_blank(t[n]) && !is_eol(t[n]) && t[n] != '#' && n < 60) ++n;
            return fail(PCQIO_ERR_PARSE, "could not convert string '%.*s' to float64 at row %lld", (int)n, t,
                        (long long)ln);
        }
        return PCQIO_OK;
    } catch (const std::bad_alloc&) {
        return fail(PCQIO_ERR_ALLOC, "pcqio_read: out of memory");
    } catch (const std::exception& e) {
        return fail(PCQIO_ERR_IO, "pcqio_read: %s", e.what());
    }
}

int64_t pcqio_selftest(uint64_t seed, int64_t n, int64_t* slow_formats, int64_t* slow_parses) {
    locale_t old = uselocale(c_locale());
    const int64_t f0 = g_slow_formats.load(), p0 = g_slow_parses.load();
    int64_t bad = 0;
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull;
    auto next = [&] {   // splitmix64
        uint64_t z = (st += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    char a[64], b[64], tok[64];
    for (int64_t i = 0; i < n; ++i) {
        double v;
        const uint64_t r = next();
        switch (i & 7) {
            case 0: case 1: case 2: {           // any bit pattern
                const uint64_t u = next();
                memcpy(&v, &u, 8);
                break;
            }
            case 3: v = (double)(int64_t)(r >> 40) / 1024.0; break;              // short dyadic fractions
            case 4: v = std::ldexp((double)((r >> 11) | 1), (int)(r % 80) - 60); break;
            case 5: v = (double)(r % 2000001) * 1e-6 - 1.0; break;               // decimal-looking values
            case 6: v = std::pow(10.0, (double)((int)(r % 617) - 308)); break;   // near powers of ten
            default: {                          // subnormals and the edges of the range
                const uint64_t u = (r & 1) ? (r >> 12) : (0x7fe0000000000000ull | (r >> 12));
                memcpy(&v, &u, 8);
            }
        }
        const int la = format_value(a, v), lb = format_value_libc(b, v);
        if (la != lb || memcmp(a, b, (size_t)la) != 0) ++bad;
        if (std::isnan(v) || std::isinf(v)) continue;
        // parse what was printed, and a shorter / perturbed decimal of the same magnitude
        double x = 0, y = 0;
        if (!parse_token(b, (size_t)lb, &x, c_locale()) || !parse_decimal_libc(b, (size_t)lb, &y, c_locale()) ||
            memcmp(&x, &y, 8) != 0 || memcmp(&x, &v, 8) != 0)
            ++bad;
        const int prec = (int)(next() % 19);
        int lt = snprintf(tok, sizeof tok, "%.*e", prec, v);
        if ((r >> 7) & 1) {                     // plain notation without an exponent where that is short
            const int l2 = snprintf(tok, sizeof tok, "%.*f", prec, std::fmod(v, 1e6));
            if (l2 > 0 && l2 < 40) lt = l2; else lt = snprintf(tok, sizeof tok, "%.*e", prec, v);
        }
        const bool ok1 = parse_token(tok, (size_t)lt, &x, c_locale());
        const bool ok2 = parse_decimal_libc(tok, (size_t)lt, &y, c_locale());
        if (ok1 != ok2 || (ok1 && memcmp(&x, &y, 8) != 0)) ++bad;
    }
    if (slow_formats) *slow_formats = g_slow_formats.load() - f0;
    if (slow_parses) *slow_parses = g_slow_parses.load() - p0;
    uselocale(old);
    return bad;
}

int pcqio_close(pcqio_reader* rd) {
    if (!rd) return PCQIO_OK;
    if (rd->base) munmap((void*)rd->base, rd->size);
    delete rd;
    return PCQIO_OK;
}

}  // extern "C"
