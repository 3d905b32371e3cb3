// pcq_decimal.h — decimal text -> binary64, the one-product fast route, shared by the host library
// (pcq_io.cpp, g++) and the device parser (pcq_text.cu, nvcc): the same source decides on both sides, so a value
// the GPU accepts is the value the host route (and numpy.loadtxt) would return.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "pcq_io_pow10.h"

#if defined(__CUDACC__)
#define PCQ_HD __host__ __device__ __forceinline__
#else
#define PCQ_HD inline
#endif

struct pcq_u192 {
    uint64_t r2, r1, r0;
};

// m * (hi:lo), all 192 bits
PCQ_HD pcq_u192 pcq_mul_64x128(uint64_t m, uint64_t hi, uint64_t lo) {
#if defined(__CUDA_ARCH__)
    const uint64_t a_lo = m * lo, a_hi = __umul64hi(m, lo);
    const uint64_t b_lo = m * hi, b_hi = __umul64hi(m, hi);
    const uint64_t mid = a_hi + b_lo;
    return pcq_u192{b_hi + (mid < a_hi ? 1ull : 0ull), mid, a_lo};
#else
    const unsigned __int128 a = (unsigned __int128)m * lo;
    const unsigned __int128 b = (unsigned __int128)m * hi;
    const unsigned __int128 mid = (a >> 64) + (uint64_t)b;
    return pcq_u192{(uint64_t)(b >> 64) + (uint64_t)(mid >> 64), (uint64_t)mid, (uint64_t)a};
#endif
}

PCQ_HD int pcq_clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

// (-1)^neg * w * 10^q -> binary64 with one 64 x 128-bit product and the truncated table entry.  The 53-bit result is
// certain unless the bits behind it are within the product's error (< 2^66 of 2^139) of the half-way point; then,
// and for subnormal / overflowing results or q outside the table, it returns false (the libc route decides).
// `table` is the PCQIO_POW10 array of the side that calls (host copy or __device__ copy).
PCQ_HD bool pcq_decimal_to_double(uint64_t w, int64_t q, bool neg, double* out, const pcqio_pow10* table) {
    if (w == 0) {
        *out = neg ? -0.0 : 0.0;
        return true;
    }
    if (q < PCQIO_Q_MIN || q > PCQIO_Q_MAX) return false;
    const int lz = pcq_clz64(w);
    const pcqio_pow10 P = table[q - PCQIO_Q_MIN];
    pcq_u192 r = pcq_mul_64x128(w << lz, P.hi, P.lo);
    int e = P.e - lz;                           // value = r * 2^e, up to the table's truncation
    if (!(r.r2 >> 63)) {
        r.r2 = (r.r2 << 1) | (r.r1 >> 63);
        r.r1 = (r.r1 << 1) | (r.r0 >> 63);
        r.r0 <<= 1;
        --e;
    }
    uint64_t m53 = r.r2 >> 11;
    const uint64_t low = r.r2 & 0x7ff;
    if (low == 0x3ff && r.r1 >= ~0ull - 8) return false;
    if (low == 0x400 && r.r1 <= 8) return false;
    if (low >= 0x400) ++m53;
    int be = e + 139 + 52 + 1023;               // biased exponent of m53 * 2^(e+139)
    if (m53 >> 53) {
        m53 >>= 1;
        ++be;
    }
    if (be < 1 || be > 2046) return false;
    const uint64_t bits = ((uint64_t)neg << 63) | ((uint64_t)be << 52) | (m53 & ((1ull << 52) - 1));
    memcpy(out, &bits, 8);
    return true;
}

// [+-]digits[.digits][e[+-]digits] with at most 19 significant digits -> w * 10^q -> one 64 x 128-bit product
// (pcq_decimal_to_double).  For longer inputs, malformed tokens and whatever the product cannot decide it returns
// false and the libc route decides (or rejects the token).
PCQ_HD bool pcq_parse_decimal_fast(const char* t, size_t n, double* out, const pcqio_pow10* table) {
    const char* p = t;
    const char* end = t + n;
    bool neg = false;
    if (p < end && (*p == '+' || *p == '-')) neg = *p++ == '-';
    uint64_t w = 0;
    int nd = 0, ndig_total = 0;
    int64_t q = 0;
    while (p < end && *p >= '0' && *p <= '9') {
        if (w || *p != '0') {
            if (nd == 19) return false;
            w = w * 10 + (uint64_t)(*p - '0');
            ++nd;
        }
        ++p;
        ++ndig_total;
    }
    if (p < end && *p == '.') {
        ++p;
        while (p < end && *p >= '0' && *p <= '9') {
            if (w || *p != '0') {
                if (nd == 19) return false;
                w = w * 10 + (uint64_t)(*p - '0');
                ++nd;
            }
            --q;
            ++p;
            ++ndig_total;
        }
    }
    if (!ndig_total) return false;
    if (p < end && (*p == 'e' || *p == 'E')) {
        ++p;
        bool eneg = false;
        if (p < end && (*p == '+' || *p == '-')) eneg = *p++ == '-';
        if (p >= end) return false;
        int64_t ex = 0;
        int ne = 0;
        while (p < end && *p >= '0' && *p <= '9') {
            if (++ne > 6) return false;
            ex = ex * 10 + (*p - '0');
            ++p;
        }
        q += eneg ? -ex : ex;
    }
    if (p != end) return false;
    return pcq_decimal_to_double(w, q, neg, out, table);
}
