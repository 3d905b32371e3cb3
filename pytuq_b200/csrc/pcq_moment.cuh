// K3m, "moment form" of the sufficient statistics (pcq_moment.cu): shared host / device geometry.
//
// The Gram matrix of a polynomial-chaos basis is a LINEAR image of far fewer numbers than its K (K + 1) / 2 entries:
// with the linearisation  psi_a psi_b = sum_k c_{abk} psi_k  of each 1-D family,
//     G[i, j] = sum_n Psi_i(x_n) Psi_j(x_n) = sum_gamma ( prod_d c_{alpha_i[d], alpha_j[d], gamma[d]} ) * m_gamma,
//     m_gamma = sum_n prod_d psi_{gamma[d]}(x_n[d]),            |gamma| <= 2 pt  (pt = highest total order of the set).
// For the total-order set d = 20, p = 3 that is 230 230 moments instead of 1 569 106 Gram entries.  The moments are
// themselves a GEMM over the samples once the dimensions are cut in two: gamma = (pi | kappa),
//     m_(pi|kappa) = sum_n P_pi(x_n[A]) Q_kappa(x_n[B]),
// rows = product functions of the dimensions in A, columns = product functions of the dimensions in B, a staircase
// (rows of order a meet the columns of order <= 2 pt - a).  Moments whose support lies in one half are the same problem
// on that half: a binary tree of nodes over the dimensions, each node a staircase GEMM; a leaf (<= 5 dimensions) pairs
// all its functions with the constant 1.
// A^T y, A^T 1 are moments weighted by y (order <= pt).
#pragma once
#include <stdint.h>

constexpr int MOM_MAXORD = 32;     // 2 pt <= MOM_MAXORD
constexpr int MOM_MAXDIM = 64;     // germ dimension (the table of 1-D values must also fit MOM_NT slots: 1 + 2 pt s + m + 1)
constexpr int MOM_WS = 8;          // table slots per sub-function descriptor
constexpr int MOM_NF = 1280;       // functions (rows + columns) of one unit
constexpr int MOM_NSUB = 384;      // distinct sub-functions of one unit
constexpr int MOM_NT = 256;        // table slots: 1 + 2 pt s, the weights, the zero slot
constexpr int MOM_TASKS = 8;       // warp tasks per unit
constexpr int MOM_MAXOUT = 8;      // outputs (weights) the route serves
constexpr int MOM_CSTR = MOM_MAXORD + 1;

#if defined(__CUDACC__)
#define MOM_HD __host__ __device__ __forceinline__
#else
#define MOM_HD inline
#endif

constexpr int MOM_LEAF = 5;        // a node over at most this many dimensions is a leaf
struct MomNode { int lo, mid, hi, ca, cb; };   // dimensions [lo, mid) | [mid, hi); a leaf (ca < 0, mid == hi) keeps all its functions as rows of one column

// class 0: plain moments, budget B0 = 2 pt;  class 1: moments weighted by output w, budget B1 = pt
struct MomGeom {
    int s, pt, m, nnodes, B0, B1;
    const MomNode* nodes;
    const int64_t* base;       // [nnodes][2]
    const int64_t* wsize;      // [nnodes]        size of one weighted copy
    const int64_t* bandoff;    // [nnodes][2][MOM_CSTR]   offset of the rows of order a
    const int32_t* ncols;      // [nnodes][2][MOM_CSTR]   columns met by the rows of order a
    const int64_t* cnt;        // [MOM_MAXDIM + 1][MOM_CSTR]  tuples of m dimensions with total exactly r
    const int64_t* cum;        // same shape, total <= r
};

// number of tuples (fixed total, lexicographically ascending) that precede the sparse tuple (dims, ks)
MOM_HD int64_t mom_lexrank(const int64_t* cnt, const int* dims, const int* ks, int ne, int lo, int m, int total) {
    int64_t rank = 0;
    int r = total;
    for (int e = 0; e < ne; ++e) {
        const int rem = m - (dims[e] - lo) - 1;
        for (int v = 0; v < ks[e]; ++v) rank += cnt[rem * MOM_CSTR + (r - v)];
        r -= ks[e];
    }
    return rank;
}

// location of moment gamma (sparse, ascending dimension, ne >= 1 entries with ks > 0) of class cls / weight w
MOM_HD int64_t mom_loc(const MomGeom& g, const int* dims, const int* ks, int ne, int cls, int w) {
    int node = 0;
    for (;;) {
        const MomNode nd = g.nodes[node];
        const int64_t base = g.base[node * 2 + cls] + (cls ? (int64_t)w * g.wsize[node] : 0);
        const int64_t* bo = g.bandoff + ((size_t)node * 2 + cls) * MOM_CSTR;
        if (nd.ca < 0) {            // leaf: rows = its non-constant functions by (order, lexicographic), one column
            int a = 0;
            for (int e = 0; e < ne; ++e) a += ks[e];
            const int m = nd.hi - nd.lo;
            return base + g.cum[m * MOM_CSTR + (a - 1)] - 1 + mom_lexrank(g.cnt, dims, ks, ne, nd.lo, m, a);
        }
        int na = 0;
        while (na < ne && dims[na] < nd.mid) ++na;
        if (na == ne) { node = nd.ca; continue; }
        if (na == 0) { node = nd.cb; continue; }
        int a = 0, b = 0;
        for (int e = 0; e < na; ++e) a += ks[e];
        for (int e = na; e < ne; ++e) b += ks[e];
        const int mA = nd.mid - nd.lo, mB = nd.hi - nd.mid;
        const int64_t rankA = mom_lexrank(g.cnt, dims, ks, na, nd.lo, mA, a);
        const int64_t col = g.cum[mB * MOM_CSTR + (b - 1)] + mom_lexrank(g.cnt, dims + na, ks + na, ne - na, nd.mid, mB, b) - 1;
        return base + bo[a] + rankA * g.ncols[((size_t)node * 2 + cls) * MOM_CSTR + a] + col;
    }
}

struct MomTask { uint16_t moff, noff; uint8_t fm, fn; uint16_t pad; };       // function offsets inside the unit, 8-row fragments
struct MomTaskOut { int64_t base; int32_t sm, sn, vm, vn; };                  // moment (i, j) of the task -> base + i sm + j sn
struct MomUnit { int32_t func_off, nfunc, sub_off, nsub, task_off, ntasks, wsub, cls; };   // wsub: slots a sub-function multiplies

// Task shapes the kernel is compiled for (row fragments x column fragments of 8): all tasks of a unit run as the unit's
// class, padded with zero functions.  Sub-function 0 of every unit is the zero function, 1 the constant one.
constexpr int MOM_NCLS = 5;
constexpr int MOM_CLS[MOM_NCLS][2] = {{4, 8}, {3, 8}, {2, 16}, {1, 16}, {1, 2}};

// One Gram entry from the moments: terms i and j as sparse (dimension, order) lists in ascending dimension.
//   lin[((d (pt+1) + a) (pt+1) + b) (B0+1) + k] = c^{(d)}_{abk},  psi_a psi_b = sum_k c_{abk} psi_k  in dimension d
MOM_HD double mom_pair(const MomGeom& g, const double* lin, const double* mom, double nsamples,
                       const int16_t* di, const uint8_t* oi, int ni, const int16_t* dj, const uint8_t* oj, int nj) {
    constexpr int ME = 2 * (MOM_MAXORD / 2);
    int dims[ME], av[ME], bv[ME], kk[ME];
    int ne = 0, p = 0, q = 0;
    while (p < ni || q < nj) {
        const int dp = p < ni ? di[p] : 0x7fff, dq = q < nj ? dj[q] : 0x7fff;
        const int d = dp < dq ? dp : dq;
        dims[ne] = d;
        av[ne] = (dp == d) ? oi[p++] : 0;
        bv[ne] = (dq == d) ? oj[q++] : 0;
        kk[ne] = 0;
        ++ne;
    }
    const int P1 = g.pt + 1, KS = g.B0 + 1;
    // odometer over the k_e with a non-zero linearisation coefficient only (Hermite / Legendre products have every other
    // coefficient zero by parity: 2^ne times fewer combinations than the full box)
    const double* le[ME];
    for (int e = 0; e < ne; ++e) {
        le[e] = lin + (((size_t)dims[e] * P1 + av[e]) * P1 + bv[e]) * KS;
        int k = 0;
        while (k <= av[e] + bv[e] && le[e][k] == 0.0) ++k;
        if (k > av[e] + bv[e]) return 0.0;              // cannot happen for a proper family: the top coefficient is not zero
        kk[e] = k;
    }
    double total = 0.0;
    for (;;) {
        double c = 1.0;
        int gd[ME], gk[ME], gn = 0;
        for (int e = 0; e < ne; ++e) {
            c *= le[e][kk[e]];
            if (kk[e] > 0) { gd[gn] = dims[e]; gk[gn] = kk[e]; ++gn; }
        }
        total += c * (gn == 0 ? nsamples : mom[mom_loc(g, gd, gk, gn, 0, 0)]);
        int e = 0;
        while (e < ne) {
            int k = kk[e] + 1;
            while (k <= av[e] + bv[e] && le[e][k] == 0.0) ++k;
            if (k <= av[e] + bv[e]) { kk[e] = k; break; }
            k = 0;
            while (le[e][k] == 0.0) ++k;
            kk[e] = k;
            ++e;
        }
        if (e == ne) break;
    }
    return total;
}
