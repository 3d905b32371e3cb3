// K3m: sufficient statistics S = Z^T Z, Z = [Psi | y | 1], in MOMENT FORM (geometry and derivation: pcq_moment.cuh).
//
// Replaces the N K^2 contraction of K3 (lreg/anl.py:78, lreg/bcs.py:84, lreg/lreg.py:328-339 through the normal
// equations) by  N |Gamma|  multiply-adds on the same FP64 tensor pipe, |Gamma| = number of distinct products
// psi_gamma, |gamma| <= 2 pt, that the entries of Psi^T Psi are linear combinations of (d = 20, p = 3: 230 230 instead of
// 1 569 106 upper-triangle entries), followed by a K x K reconstruction that does not depend on N.
//
//   planner (host, once per plan and output count)
//       tree of nodes over the dimensions -> per node and row order a "band" (rows of exact order a over the node's
//       first half  x  columns of order <= budget - a over its second half) -> warp tasks of <= 32 x 64 moments ->
//       units of 8 tasks whose row / column functions (<= 512) are generated together.  Every function is the product
//       of two sub-functions (its orders on the two halves of its dimension range), every sub-function a product of
//       <= 8 entries of the per-sample table of 1-D values: no dependent chains, two shared-memory gathers per value.
//   pcq_mom_kernel (one persistent CTA per SM, work items = (unit, sample part) from an atomic counter)
//       warps 0-3 and 4-7 alternate between GENERATING the operands of the next 16-sample block in shared memory and
//       issuing the DMMAs of the current one, so that every scheduler's FP64 pipe always has one warp of DMMAs: x and
//       y are the only HBM traffic (8 (s + M) bytes per sample and unit), no design-matrix byte ever exists.
//   pcq_mom_fixup_kernel   sums the sample parts in part order (deterministic) into the moment vector
//   pcq_mom_pairs_kernel / pcq_mom_edges_kernel   S (+)= reconstruction
#include "pcq_mma.cuh"
#include "pcq_moment.cuh"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace {

using Desc = std::array<uint16_t, MOM_WS>;
constexpr int MOM_SUBBATCH = 64;    // sub-function rows per batch of the kernel's pass (32 rows x MOM_VB per 256 threads)
constexpr int MOM_BS_HOST = 16;     // samples per block: descriptors are stored as BYTE offsets of 16-sample rows (x 128)

struct MomHost {
    int s = 0, pt = 0, m = 0, B0 = 0, B1 = 0, nslots = 0, zero_slot = 0, nt = 0, width = 0, nterms = 0;
    std::vector<MomNode> nodes;
    std::vector<int64_t> base, wsize, bandoff, cnt, cum;
    std::vector<int32_t> ncols;
    int64_t nmom = 0;
    std::vector<double> lin, rec_a, rec_b, p1;
    std::vector<int16_t> tdim;      // [K][width] sparse terms, -1 padding
    std::vector<uint8_t> tord;
    std::vector<uint8_t> tnnz;
    std::vector<MomUnit> units;
    std::vector<MomTask> tasks;       // 8 per unit, slot = warp
    std::vector<MomTaskOut> touts;
    std::vector<uint32_t> funcdesc;
    std::vector<uint16_t> subdesc;    // MOM_WS per sub-function
    int nf_max = 0, nsub_max = 0;
    double clk_per_block = 0;         // estimated cycles of one 16-sample block summed over the units (one CTA)
    double mac_slots = 0;             // executed DMMA multiply-adds per sample
    double amp = 0;                   // how much larger than a Gram entry the moments it is summed from can be (see mom_build)
    std::string why;                  // empty when the plan is usable

    MomGeom geom() const {
        MomGeom g;
        g.s = s; g.pt = pt; g.m = m; g.nnodes = (int)nodes.size(); g.B0 = B0; g.B1 = B1;
        g.nodes = nodes.data(); g.base = base.data(); g.wsize = wsize.data(); g.bandoff = bandoff.data();
        g.ncols = ncols.data(); g.cnt = cnt.data(); g.cum = cum.data();
        return g;
    }
};

int build_node(std::vector<MomNode>& nodes, int lo, int hi) {
    const int idx = (int)nodes.size();
    nodes.push_back({lo, hi, hi, -1, -1});
    if (hi - lo > MOM_LEAF) {
        const int mid = lo + (hi - lo + 1) / 2;
        nodes[idx].mid = mid;
        const int ca = build_node(nodes, lo, mid);
        const int cb = build_node(nodes, mid, hi);
        nodes[idx].ca = ca;
        nodes[idx].cb = cb;
    }
    return idx;
}

// psi_a psi_b = sum_k c_k psi_k from the three-term recurrence alone (no orthogonality needed):
//   psi_0 = 1, psi_1 = p1 x, psi_{n+1} = a_n x psi_n + b_n psi_{n-1}  =>  x psi_0 = psi_1 / p1,
//   x psi_n = (psi_{n+1} - b_n psi_{n-1}) / a_n;  psi_a(x) f is built by running the recurrence on coefficient vectors.
void linearise(const double* ra, const double* rb, double p1, int s, int d, int pt, int B0, double* out) {
    const int nk = B0 + 2;
    auto mulx = [&](const std::vector<long double>& u) {
        std::vector<long double> o(nk + 1, 0.0L);
        for (int k = 0; k < nk; ++k) {
            if (u[k] == 0.0L) continue;
            if (k == 0) { o[1] += u[0] / (long double)p1; continue; }
            const long double a = ra[(size_t)k * s + d], b = rb[(size_t)k * s + d];
            o[k + 1] += u[k] / a;
            o[k - 1] -= u[k] * b / a;
        }
        o.resize(nk);
        return o;
    };
    for (int b = 0; b <= pt; ++b) {
        std::vector<long double> um(nk, 0.0L), uc(nk, 0.0L);
        um[b] = 1.0L;                                      // psi_0 psi_b
        for (int a = 0; a <= pt; ++a) {
            if (a == 1) {
                uc = mulx(um);
                for (auto& v : uc) v *= (long double)p1;
            } else if (a >= 2) {
                std::vector<long double> xu = mulx(uc), un(nk, 0.0L);
                const long double an = ra[(size_t)(a - 1) * s + d], bn = rb[(size_t)(a - 1) * s + d];
                for (int k = 0; k < nk; ++k) un[k] = an * xu[k] + bn * um[k];
                um = uc;
                uc = un;
            }
            const std::vector<long double>& r = (a == 0) ? um : uc;
            long double big = 0.0L;
            for (int k = 0; k <= B0; ++k) big = std::max(big, fabsl(r[k]));
            for (int k = 0; k <= B0; ++k) {
                const long double v = (k <= a + b && fabsl(r[k]) > 1e-14L * big) ? r[k] : 0.0L;
                out[(((size_t)d * (pt + 1) + a) * (pt + 1) + b) * (B0 + 1) + k] = (double)v;
            }
        }
    }
}

// tuple number idx (lexicographically ascending) of m dimensions with total exactly `total`
void unrank(const std::vector<int64_t>& cnt, int m, int total, int64_t idx, int* ks) {
    int r = total;
    for (int j = 0; j < m; ++j) {
        const int rem = m - j - 1;
        int v = 0;
        for (; v <= r; ++v) {
            const int64_t c = cnt[(size_t)rem * MOM_CSTR + (r - v)];
            if (idx < c) break;
            idx -= c;
        }
        ks[j] = v;
        r -= v;
    }
}

struct Seg {            // a run of functions generated together: rows (side 0) or columns (side 1) of a node
    int node, side, sel, start, count, n8;     // sel: row order a (side 0) / weight, -1 = none (side 1)
    std::vector<std::pair<Desc, Desc>> fn;     // sub-function descriptors of the 8 * n8 functions
    bool same(const Seg& o) const { return node == o.node && side == o.side && sel == o.sel && start == o.start && n8 == o.n8 && count == o.count; }
};
struct BTask { int segM, segN, fm, fn; MomTaskOut out; };
struct Bundle { std::vector<Seg> segs; std::vector<BTask> tasks; int cls = 0; };

int class_of(int fm, int fn) {
    int best = 0, area = 1 << 30;
    for (int c = 0; c < MOM_NCLS; ++c)
        if (MOM_CLS[c][0] >= fm && MOM_CLS[c][1] >= fn && MOM_CLS[c][0] * MOM_CLS[c][1] < area) { best = c; area = MOM_CLS[c][0] * MOM_CLS[c][1]; }
    return best;
}
// one bundle per task class (the kernel runs all tasks of a unit in one shape)
void push_by_class(std::vector<Bundle>& out, Bundle&& B) {
    for (int c = 0; c < MOM_NCLS; ++c) {
        Bundle Q;
        Q.cls = c;
        std::vector<int> map(B.segs.size(), -1);
        for (const BTask& t : B.tasks) {
            if (class_of(t.fm, t.fn) != c) continue;
            BTask q = t;
            for (int* sg : {&q.segM, &q.segN}) {
                if (map[*sg] < 0) { map[*sg] = (int)Q.segs.size(); Q.segs.push_back(B.segs[*sg]); }
                *sg = map[*sg];
            }
            Q.tasks.push_back(q);
        }
        if (!Q.tasks.empty()) out.push_back(std::move(Q));
    }
}

struct Planner {
    MomHost& H;
    explicit Planner(MomHost& h) : H(h) {}
    bool fail = false;

    Desc trivial() const { Desc d; d.fill(0); return d; }
    Desc zero() const { Desc d; d.fill(0); d[0] = (uint16_t)H.zero_slot; return d; }

    void make_functions(Seg& sg) {
        const MomNode nd = H.nodes[sg.node];
        const bool leaf = (nd.ca < 0);
        int d0, d1;
        if (sg.side == 0) { d0 = nd.lo; d1 = leaf ? nd.hi : nd.mid; }
        else { d0 = nd.mid; d1 = nd.hi; }
        const int md = d1 - d0;
        const int h = d0 + (md + 1) / 2;
        sg.fn.resize((size_t)sg.n8 * 8);
        std::vector<int> ks(std::max(md, 1));
        for (int i = 0; i < sg.n8 * 8; ++i) {
            Desc s1 = trivial(), s2 = trivial();
            if (i >= sg.count) { sg.fn[i] = {zero(), trivial()}; continue; }
            if (md > 0) {
                if (sg.side == 0 && !leaf) {
                    unrank(H.cnt, md, sg.sel, (int64_t)sg.start + i, ks.data());
                } else {                                   // columns, and a leaf's rows: all non-constant functions by (order, lex)
                    const int64_t colidx = (int64_t)sg.start + i + 1;
                    int b = 1;
                    while (H.cum[(size_t)md * MOM_CSTR + b] <= colidx) ++b;
                    unrank(H.cnt, md, b, colidx - H.cum[(size_t)md * MOM_CSTR + (b - 1)], ks.data());
                }
            }
            int n1 = 0, n2 = 0;
            for (int j = 0; j < md; ++j) {
                if (ks[j] == 0) continue;
                const uint16_t slot = (uint16_t)(1 + (ks[j] - 1) * H.s + (d0 + j));
                if (d0 + j < h) { if (n1 < MOM_WS) s1[n1] = slot; ++n1; }
                else { if (n2 < MOM_WS) s2[n2] = slot; ++n2; }
            }
            if (sg.side == 1 && sg.sel >= 0) { if (n2 < MOM_WS) s2[n2] = (uint16_t)(H.nslots + sg.sel); ++n2; }
            if (n1 > MOM_WS || n2 > MOM_WS) fail = true;
            sg.fn[i] = {s1, s2};
        }
    }

    // one band: countM x countN moments at `base`, element (i, j) -> base + i sm + j sn
    void band(std::vector<Bundle>& out, int node, int a, int w, int64_t nrows, int64_t ncols, int64_t base) {
        // orientation: rows on the M side (<= 32 per task) or on the N side (<= 64 per task, <= 128 where the M side
        // has at most 16), fewer tasks wins
        auto tiles = [](int64_t m8_, int64_t n8_, int64_t& TM_, int64_t& TN_, int& ncap) {
            TM_ = (m8_ + 3) / 4;
            const int64_t fmax = (m8_ + TM_ - 1) / TM_;
            ncap = fmax <= 2 ? 16 : 8;
            TN_ = (n8_ + ncap - 1) / ncap;
        };
        const int64_t r8 = (nrows + 7) / 8, c8 = (ncols + 7) / 8;
        int64_t tm1, tn1, tm2, tn2;
        int cap1, cap2;
        tiles(r8, c8, tm1, tn1, cap1);
        tiles(c8, r8, tm2, tn2, cap2);
        const bool rows_on_m = tm1 * tn1 <= tm2 * tn2;
        const int64_t cm = rows_on_m ? nrows : ncols, cn = rows_on_m ? ncols : nrows;
        const int64_t m8 = (cm + 7) / 8, n8 = (cn + 7) / 8;
        const int64_t TM = rows_on_m ? tm1 : tm2, TN = rows_on_m ? tn1 : tn2;
        const int ncap = rows_on_m ? cap1 : cap2;
        const int32_t sm = rows_on_m ? (int32_t)ncols : 1, sn = rows_on_m ? 1 : (int32_t)ncols;
        auto part = [](int64_t total8, int64_t parts, int64_t t, int64_t& off8, int& len8) {
            const int64_t q = total8 / parts, r = total8 % parts;
            off8 = t * q + std::min(t, r);
            len8 = (int)(q + (t < r ? 1 : 0));
        };
        int um = 1, un = 1;
        {
            int64_t best = INT64_MAX;
            for (int a_ = 1; a_ <= 8; ++a_)
                for (int b_ = 1; a_ * b_ <= 8; ++b_) {
                    if (a_ > TM || b_ > TN || 32 * a_ + 8 * ncap * b_ > MOM_NF) continue;
                    const int64_t nb = ((TM + a_ - 1) / a_) * ((TN + b_ - 1) / b_);
                    const int64_t key = nb * 4096 + (32 * a_ + 8 * ncap * b_);
                    if (key < best) { best = key; um = a_; un = b_; }
                }
        }
        for (int64_t bm = 0; bm < TM; bm += um)
            for (int64_t bn = 0; bn < TN; bn += un) {
                Bundle B;
                std::vector<int> segm, segn;
                for (int64_t tm = bm; tm < std::min<int64_t>(bm + um, TM); ++tm) {
                    int64_t off8; int len8;
                    part(m8, TM, tm, off8, len8);
                    Seg sg;
                    sg.node = node; sg.side = rows_on_m ? 0 : 1; sg.sel = rows_on_m ? a : w;
                    sg.start = (int)(off8 * 8); sg.n8 = len8; sg.count = (int)std::min<int64_t>(len8 * 8, cm - off8 * 8);
                    make_functions(sg);
                    segm.push_back((int)B.segs.size());
                    B.segs.push_back(std::move(sg));
                }
                for (int64_t tn = bn; tn < std::min<int64_t>(bn + un, TN); ++tn) {
                    int64_t off8; int len8;
                    part(n8, TN, tn, off8, len8);
                    Seg sg;
                    sg.node = node; sg.side = rows_on_m ? 1 : 0; sg.sel = rows_on_m ? w : a;
                    sg.start = (int)(off8 * 8); sg.n8 = len8; sg.count = (int)std::min<int64_t>(len8 * 8, cn - off8 * 8);
                    make_functions(sg);
                    segn.push_back((int)B.segs.size());
                    B.segs.push_back(std::move(sg));
                }
                for (int im : segm)
                    for (int in : segn) {
                        const Seg& SM_ = B.segs[im];
                        const Seg& SN_ = B.segs[in];
                        BTask t;
                        t.segM = im; t.segN = in; t.fm = SM_.n8; t.fn = SN_.n8;
                        t.out.base = base + (int64_t)SM_.start * sm + (int64_t)SN_.start * sn;
                        t.out.sm = sm; t.out.sn = sn; t.out.vm = SM_.count; t.out.vn = SN_.count;
                        B.tasks.push_back(t);
                    }
                push_by_class(out, std::move(B));
            }
    }
};

struct UnitBuild {
    std::vector<Seg> segs;
    std::vector<int> segoff;          // first function of each segment
    std::vector<BTask> tasks;
    std::map<Desc, int> subs;
    int nfunc = 0, cls = -1;
};

// tries to add a bundle to a unit under the function / sub-function capacities
bool try_add(UnitBuild& U, const Bundle& B, int zero_slot) {
    if (U.tasks.size() + B.tasks.size() > (size_t)MOM_TASKS) return false;
    if (U.cls >= 0 && U.cls != B.cls) return false;
    if (U.subs.empty()) {          // sub-function 0 = the zero function, 1 = the constant one, in every unit
        Desc z, one;
        z.fill(0); one.fill(0);
        z[0] = (uint16_t)zero_slot;
        U.subs.emplace(z, 0);
        U.subs.emplace(one, 1);
    }
    U.cls = B.cls;
    std::vector<int> map(B.segs.size(), -1);
    int nfunc = U.nfunc;
    std::map<Desc, int> extra;
    for (size_t i = 0; i < B.segs.size(); ++i) {
        for (size_t j = 0; j < U.segs.size(); ++j)
            if (U.segs[j].same(B.segs[i])) { map[i] = (int)j; break; }
        if (map[i] >= 0) continue;
        nfunc += B.segs[i].n8 * 8;
        for (const auto& f : B.segs[i].fn) {
            if (!U.subs.count(f.first)) extra.emplace(f.first, 0);
            if (!U.subs.count(f.second)) extra.emplace(f.second, 0);
        }
    }
    if (nfunc > MOM_NF || U.subs.size() + extra.size() > (size_t)MOM_NSUB) return false;
    for (size_t i = 0; i < B.segs.size(); ++i) {
        if (map[i] >= 0) continue;
        map[i] = (int)U.segs.size();
        U.segoff.push_back(U.nfunc);
        U.nfunc += B.segs[i].n8 * 8;
        for (const auto& f : B.segs[i].fn) {
            U.subs.emplace(f.first, (int)U.subs.size());
            U.subs.emplace(f.second, (int)U.subs.size());
        }
        U.segs.push_back(B.segs[i]);
    }
    for (BTask t : B.tasks) {
        t.segM = map[t.segM];
        t.segN = map[t.segN];
        U.tasks.push_back(t);
    }
    return true;
}

// builds the whole host-side plan; H.why explains a refusal
void mom_build(MomHost& H, int s, int nterms, const int32_t* mi, int nord, const double* ra, const double* rb,
               const double* p1, int m) {
    H.s = s; H.m = m; H.nterms = nterms;
    int pt = 0, width = 1;
    for (int k = 0; k < nterms; ++k) {
        int t = 0, nnz = 0;
        for (int i = 0; i < s; ++i) { t += mi[(size_t)k * s + i]; nnz += mi[(size_t)k * s + i] > 0; }
        pt = std::max(pt, t);
        width = std::max(width, nnz);
    }
    H.pt = pt; H.B0 = 2 * pt; H.B1 = pt; H.width = width;
    H.nslots = 1 + H.B0 * s;
    H.zero_slot = H.nslots + m;
    H.nt = H.zero_slot + 1;
    if (pt < 1) { H.why = "constant basis"; return; }
    if (s < 2 || s > MOM_MAXDIM) { H.why = "germ dimension outside [2, 64]"; return; }
    if (H.B0 > MOM_MAXORD || nord < H.B0) { H.why = "order too high for the tabulated recurrence"; return; }
    if (H.nt > MOM_NT) { H.why = "1-D table too large for shared memory"; return; }
    if (m > MOM_MAXOUT) { H.why = "too many outputs"; return; }

    H.rec_a.assign(ra, ra + (size_t)(H.B0 + 1) * s);
    H.rec_b.assign(rb, rb + (size_t)(H.B0 + 1) * s);
    H.p1.assign(p1, p1 + s);
    for (int d = 0; d < s; ++d) {
        if (p1[d] == 0.0) { H.why = "degenerate family"; return; }
        for (int n = 1; n < H.B0; ++n)
            if (!(std::fabs(ra[(size_t)n * s + d]) > 0.0) || !std::isfinite(rb[(size_t)n * s + d])) { H.why = "degenerate recurrence"; return; }
    }
    H.lin.assign((size_t)s * (pt + 1) * (pt + 1) * (H.B0 + 1), 0.0);
    for (int d = 0; d < s; ++d) linearise(ra, rb, p1[d], s, d, pt, H.B0, H.lin.data());
    // Conditioning of the reconstruction: G_ii = sum_k c_aak m_k with m_0 = N h_0 the entry itself and the other m_k
    // sampling noise of size ~ sqrt(N h_k) summed from terms of size ~ sqrt(h_k), h_k = ||psi_k||^2 (from the
    // recurrence: h_n = -b_n h_{n-1} a_{n-1} / a_n).  amp = max_a sum_{k>=1} |c_aak| sqrt(h_k) / c_aa0 measures how far
    // the rounding error of the moments is amplified relative to the entry: Legendre <= 2 a + 1, Hermite 16 / 47 / 134 /
    // 386 at a = 3 .. 6 and super-exponential beyond.  The route rule keeps the dense form above PCQ_MOMENT_MAX_AMP (500).
    for (int d = 0; d < s; ++d) {
        std::vector<long double> h(H.B0 + 1, 1.0L);
        for (int n = 1; n <= H.B0; ++n) {
            const long double an = ra[(size_t)n * s + d], bn = rb[(size_t)n * s + d];
            const long double am = n == 1 ? (long double)p1[d] : (long double)ra[(size_t)(n - 1) * s + d];
            h[n] = fabsl(bn * h[n - 1] * am / an);
        }
        for (int a = 1; a <= pt; ++a) {
            const double* c = H.lin.data() + (((size_t)d * (pt + 1) + a) * (pt + 1) + a) * (H.B0 + 1);
            long double noise = 0.0L;
            for (int k = 1; k <= 2 * a; ++k) noise += fabsl((long double)c[k]) * sqrtl(h[k]);
            if (c[0] > 0.0) H.amp = std::max(H.amp, (double)(noise / (long double)c[0]));
        }
    }

    // sparse terms
    H.tdim.assign((size_t)nterms * width, -1);
    H.tord.assign((size_t)nterms * width, 0);
    H.tnnz.assign(nterms, 0);
    for (int k = 0; k < nterms; ++k) {
        int q = 0;
        for (int i = 0; i < s; ++i)
            if (mi[(size_t)k * s + i] > 0) {
                H.tdim[(size_t)k * width + q] = (int16_t)i;
                H.tord[(size_t)k * width + q] = (uint8_t)mi[(size_t)k * s + i];
                ++q;
            }
        H.tnnz[k] = (uint8_t)q;
    }

    // counting tables
    const int64_t SAT = (int64_t)1 << 50;
    H.cnt.assign((size_t)(MOM_MAXDIM + 1) * MOM_CSTR, 0);
    H.cum.assign((size_t)(MOM_MAXDIM + 1) * MOM_CSTR, 0);
    H.cnt[0] = 1;
    for (int mm = 1; mm <= MOM_MAXDIM; ++mm)
        for (int r = 0; r <= MOM_MAXORD; ++r) {
            int64_t v = H.cnt[(size_t)(mm - 1) * MOM_CSTR + r] + (r > 0 ? H.cnt[(size_t)mm * MOM_CSTR + r - 1] : 0);
            H.cnt[(size_t)mm * MOM_CSTR + r] = std::min(v, SAT);
        }
    for (int mm = 0; mm <= MOM_MAXDIM; ++mm) {
        int64_t c = 0;
        for (int r = 0; r <= MOM_MAXORD; ++r) {
            c = std::min(c + H.cnt[(size_t)mm * MOM_CSTR + r], SAT);
            H.cum[(size_t)mm * MOM_CSTR + r] = c;
        }
    }

    // nodes and moment geometry
    build_node(H.nodes, 0, s);
    const int nn = (int)H.nodes.size();
    H.base.assign((size_t)nn * 2, 0);
    H.wsize.assign(nn, 0);
    H.bandoff.assign((size_t)nn * 2 * MOM_CSTR, 0);
    H.ncols.assign((size_t)nn * 2 * MOM_CSTR, 0);
    int64_t total = 0;
    for (int v = 0; v < nn; ++v) {
        const MomNode nd = H.nodes[v];
        const bool leaf = (nd.ca < 0);
        const int mA = nd.mid - nd.lo, mB = nd.hi - nd.mid;
        for (int cls = 0; cls < 2; ++cls) {
            const int B = cls ? H.B1 : H.B0;
            int64_t off = 0;
            if (leaf) {
                off = H.cum[(size_t)(nd.hi - nd.lo) * MOM_CSTR + B] - 1;
                H.ncols[((size_t)v * 2 + cls) * MOM_CSTR + 1] = 1;
                if (off > ((int64_t)1 << 30)) { H.why = "too many moments"; return; }
            }
            for (int a = 1; !leaf && a <= B - 1; ++a) {
                int64_t nc = H.cum[(size_t)mB * MOM_CSTR + (B - a)] - 1;
                const int64_t nr = H.cnt[(size_t)mA * MOM_CSTR + a];
                if (nc > ((int64_t)1 << 30) || nr > ((int64_t)1 << 30)) { H.why = "too many moments"; return; }
                H.bandoff[((size_t)v * 2 + cls) * MOM_CSTR + a] = off;
                H.ncols[((size_t)v * 2 + cls) * MOM_CSTR + a] = (int32_t)nc;
                off += nr * nc;
                if (off > ((int64_t)1 << 31)) { H.why = "too many moments"; return; }
            }
            H.base[(size_t)v * 2 + cls] = total;
            if (cls == 0) total += off;
            else { H.wsize[v] = off; total += off * m; }
            if (total > ((int64_t)1 << 31)) { H.why = "too many moments"; return; }
        }
    }
    H.nmom = total;

    // bands -> bundles
    Planner P(H);
    std::vector<Bundle> bundles;
    for (int v = 0; v < nn; ++v) {
        const MomNode nd = H.nodes[v];
        const bool leaf = (nd.ca < 0);
        const int mA = nd.mid - nd.lo;
        for (int cls = 0; cls < 2; ++cls) {
            const int B = cls ? H.B1 : H.B0;
            for (int w = 0; w < (cls ? m : 1); ++w)
                for (int a = 1; a <= (leaf ? 1 : B - 1); ++a) {      // a leaf is one band: all its functions, one column
                    const int64_t nr = leaf ? H.cum[(size_t)(nd.hi - nd.lo) * MOM_CSTR + B] - 1 : H.cnt[(size_t)mA * MOM_CSTR + a];
                    const int64_t nc = H.ncols[((size_t)v * 2 + cls) * MOM_CSTR + a];
                    if (nr <= 0 || nc <= 0) continue;
                    const int64_t base = H.base[(size_t)v * 2 + cls] + (cls ? (int64_t)w * H.wsize[v] : 0) +
                                         H.bandoff[((size_t)v * 2 + cls) * MOM_CSTR + a];
                    if (getenv("PCQ_MOMENT_VERBOSE")) fprintf(stderr, "band node %d [%d,%d|%d) leaf %d cls %d a %d rows %lld cols %lld moments %lld\n", v, nd.lo, nd.mid, nd.hi, (int)leaf, cls, a, (long long)nr, (long long)nc, (long long)(nr * nc));
                    P.band(bundles, v, a, cls ? w : -1, nr, nc, base);
                    if (bundles.size() > 200000) { H.why = "too many tasks"; return; }
                }
        }
    }
    if (P.fail) { H.why = "a sub-function has more than 8 factors"; return; }

    // bundles -> units: full bundles alone; partial ones first-fit (largest first) into the few units still open, whole
    // if they fit (their tasks share row / column functions), else task by task, so that only the last unit of a
    // shape is left with idle warps
    std::vector<int> order(bundles.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bundles[a].tasks.size() > bundles[b].tasks.size(); });
    std::vector<UnitBuild> ub;
    std::vector<int> open;
    auto sub_bundle = [](const Bundle& cur, size_t lo, size_t hi) {
        Bundle part;
        part.cls = cur.cls;
        std::vector<int> map(cur.segs.size(), -1);
        for (size_t t = lo; t < hi; ++t) {
            BTask q = cur.tasks[t];
            for (int* sg : {&q.segM, &q.segN}) {
                if (map[*sg] < 0) { map[*sg] = (int)part.segs.size(); part.segs.push_back(cur.segs[*sg]); }
                *sg = map[*sg];
            }
            part.tasks.push_back(q);
        }
        return part;
    };
    auto close_full = [&]() {
        open.erase(std::remove_if(open.begin(), open.end(), [&](int u) { return ub[u].tasks.size() >= (size_t)MOM_TASKS; }), open.end());
    };
    // a new unit for the bundle; halves of it, recursively, where its sub-functions do not fit one unit
    auto place_new = [&](const Bundle& B) -> bool {
        std::vector<Bundle> stack{B};
        while (!stack.empty()) {
            Bundle cur = std::move(stack.back());
            stack.pop_back();
            ub.emplace_back();
            if (try_add(ub.back(), cur, H.zero_slot)) {
                if (ub.back().tasks.size() < (size_t)MOM_TASKS) open.push_back((int)ub.size() - 1);
                continue;
            }
            ub.pop_back();
            if (cur.tasks.size() < 2) return false;
            stack.push_back(sub_bundle(cur, cur.tasks.size() / 2, cur.tasks.size()));
            stack.push_back(sub_bundle(cur, 0, cur.tasks.size() / 2));
        }
        return true;
    };
    auto first_fit = [&](const Bundle& B) -> bool {
        for (size_t q = open.size() > 6 ? open.size() - 6 : 0; q < open.size(); ++q)
            if (try_add(ub[open[q]], B, H.zero_slot)) return true;
        return false;
    };
    for (int bi : order) {
        const Bundle& B = bundles[bi];
        if (B.tasks.size() >= (size_t)MOM_TASKS || !first_fit(B)) {
            bool any_open_same_class = false;
            for (size_t q = open.size() > 6 ? open.size() - 6 : 0; q < open.size(); ++q) any_open_same_class |= ub[open[q]].cls == B.cls;
            if (B.tasks.size() < (size_t)MOM_TASKS && B.tasks.size() > 1 && any_open_same_class) {
                // task by task into the open units; what is left starts a new unit together
                std::vector<size_t> left;
                for (size_t t = 0; t < B.tasks.size(); ++t) {
                    if (!first_fit(sub_bundle(B, t, t + 1))) left.push_back(t);
                    close_full();
                }
                if (!left.empty()) {
                    Bundle rest;
                    rest.cls = B.cls;
                    std::vector<int> map(B.segs.size(), -1);
                    for (size_t t : left) {
                        BTask q = B.tasks[t];
                        for (int* sg : {&q.segM, &q.segN}) {
                            if (map[*sg] < 0) { map[*sg] = (int)rest.segs.size(); rest.segs.push_back(B.segs[*sg]); }
                            *sg = map[*sg];
                        }
                        rest.tasks.push_back(q);
                    }
                    if (!place_new(rest)) { H.why = "a task does not fit a unit"; return; }
                }
            } else if (!place_new(B)) { H.why = "a task does not fit a unit"; return; }
        }
        close_full();
    }

    // finalise: warp slots (four costliest tasks on warps 0-3, the rest on 4-7: the two warp groups issue their
    // DMMAs in different half-phases, so a unit costs max(group 0) + max(group 1)), descriptors, cost, order
    struct Fin { double cost; int idx; int cls; };
    std::vector<Fin> fin(ub.size());
    std::vector<double> macs(ub.size(), 0.0);
    for (size_t u = 0; u < ub.size(); ++u) {
        UnitBuild& U = ub[u];
        const double half = 64.0 * MOM_CLS[U.cls][0] * MOM_CLS[U.cls][1];      // DMMA cycles of one warp per block
        macs[u] = (double)U.tasks.size() * half;          // warps without a task issue no DMMA
        // cycles of one 16-sample block: both warps of a scheduler at the measured DMMA-pipe share of the class
        // (warp-specialised kernels, cycle counters of PCQ_MOMENT_DBG at C2: profiles/r02_moment_notes.md)
        static const double eff[MOM_NCLS] = {0.847, 0.78, 0.77, 0.54, 0.30};
        // a unit with at most four tasks has one warp per scheduler (slots 0-3): half the DMMAs per scheduler
        fin[u] = {(U.tasks.size() <= 4 ? 1.1 : 2.0) * half / eff[U.cls], (int)u, U.cls};

    }
    std::stable_sort(fin.begin(), fin.end(), [](const Fin& a, const Fin& b) { return a.cls != b.cls ? a.cls < b.cls : a.cost > b.cost; });
    for (const Fin& f : fin) {
        const UnitBuild& U = ub[f.idx];
        MomUnit mu;
        memset(&mu, 0, sizeof(mu));
        mu.func_off = (int)H.funcdesc.size();
        mu.nfunc = U.nfunc;
        mu.sub_off = (int)(H.subdesc.size() / MOM_WS);
        mu.nsub = (int)U.subs.size();
        mu.task_off = (int)H.tasks.size();
        mu.ntasks = (int)U.tasks.size();
        // sub-function rows: the zero and the constant function first, the others by falling factor count (stable), so
        // that each batch of 64 rows of the kernel's sub-function pass multiplies only as many factors as its rows have
        // (wsub packs the factor count of batch b in bits 4b .. 4b+3)
        std::vector<Desc> byid(U.subs.size());
        for (const auto& kv : U.subs) byid[kv.second] = kv.first;
        auto nnz_of = [](const Desc& d) { int w = 1; for (int q = 0; q < MOM_WS; ++q) if (d[q] != 0) w = q + 1; return w; };
        std::vector<int> ord(byid.size());
        for (size_t i = 0; i < ord.size(); ++i) ord[i] = (int)i;
        std::stable_sort(ord.begin() + std::min<size_t>(2, ord.size()), ord.end(), [&](int a, int b) { return nnz_of(byid[a]) > nnz_of(byid[b]); });
        std::vector<int> newid(byid.size());
        for (size_t i = 0; i < ord.size(); ++i) newid[ord[i]] = (int)i;
        uint32_t wpack = 0;
        for (size_t i = 0; i < ord.size(); ++i) {
            const Desc& d = byid[ord[i]];
            for (int q = 0; q < MOM_WS; ++q) H.subdesc.push_back((uint16_t)(d[q] * (MOM_BS_HOST * 8)));
            const int b = (int)(i / MOM_SUBBATCH);
            const uint32_t w = (uint32_t)nnz_of(d);
            if (((wpack >> (4 * b)) & 15u) < w) wpack = (wpack & ~(15u << (4 * b))) | (w << (4 * b));
        }
        mu.wsub = (int32_t)wpack;
        mu.cls = U.cls;
        for (size_t sI = 0; sI < U.segs.size(); ++sI)
            for (const auto& fdesc : U.segs[sI].fn)
                H.funcdesc.push_back((uint32_t)newid[U.subs.at(fdesc.first)] * (MOM_BS_HOST * 8) |
                                     ((uint32_t)newid[U.subs.at(fdesc.second)] * (MOM_BS_HOST * 8) << 16));
        for (int t = 0; t < MOM_TASKS; ++t) {
            MomTask mt;
            MomTaskOut to;
            memset(&mt, 0, sizeof(mt));
            memset(&to, 0, sizeof(to));
            if (t < (int)U.tasks.size()) {
                const BTask& bt = U.tasks[t];
                mt.moff = (uint16_t)U.segoff[bt.segM];
                mt.noff = (uint16_t)U.segoff[bt.segN];
                mt.fm = (uint8_t)bt.fm;
                mt.fn = (uint8_t)bt.fn;
                to = bt.out;
            }
            H.tasks.push_back(mt);
            H.touts.push_back(to);
        }
        H.units.push_back(mu);
        H.nf_max = std::max(H.nf_max, mu.nfunc);
        H.nsub_max = std::max(H.nsub_max, mu.nsub);
        H.clk_per_block += f.cost;
        H.mac_slots += macs[f.idx];
        if (getenv("PCQ_MOMENT_VERBOSE")) {
            double useful = 0.0, frag = 0.0;
            for (const BTask& bt : U.tasks) { useful += (double)bt.out.vm * bt.out.vn; frag += 64.0 * bt.fm * bt.fn; }
            fprintf(stderr, "unit %3d class (%d,%2d) tasks %d funcs %4d subs %3d  moments %7.0f  fragment slots %7.0f  executed %7.0f  est. cycles/block %7.0f\n",
                    (int)H.units.size() - 1, MOM_CLS[U.cls][0], MOM_CLS[U.cls][1], (int)U.tasks.size(), U.nfunc, (int)U.subs.size(),
                    useful, frag, macs[f.idx], f.cost);
        }
    }
}

// ------------------------------------------------------------------------------------- host reference
// The unit / task structures executed on the CPU exactly as the kernel does (table -> sub-functions -> functions ->
// task products), then the reconstruction: checks planner, ranking and linearisation without a device.
void host_table(const MomHost& H, const double* x, const double* y, std::vector<double>& T) {
    const int s = H.s;
    T[0] = 1.0;
    for (int d = 0; d < s; ++d) {
        double pm = 1.0, pc = x[d] * H.p1[d];
        T[1 + d] = pc;
        for (int nn = 1; nn < H.B0; ++nn) {
            const double pn = H.rec_a[(size_t)nn * s + d] * x[d] * pc + H.rec_b[(size_t)nn * s + d] * pm;
            T[1 + nn * s + d] = pn;
            pm = pc;
            pc = pn;
        }
    }
    for (int w = 0; w < H.m; ++w) T[H.nslots + w] = y[w];
    T[H.zero_slot] = 0.0;
}

void host_moments(const MomHost& H, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy, std::vector<double>& mom) {
    mom.assign((size_t)H.nmom, 0.0);
    std::vector<double> T(H.nt), V(MOM_NSUB), F(MOM_NF);
    for (int64_t i = 0; i < n; ++i) {
        host_table(H, x + i * ldx, y ? y + i * ldy : nullptr, T);
        for (const MomUnit& U : H.units) {
            for (int q = 0; q < U.nsub; ++q) {
                double v = 1.0;
                const int wq = (int)(((uint32_t)U.wsub >> (4 * (q / MOM_SUBBATCH))) & 15u);
                for (int k = 0; k < wq; ++k) v *= T[H.subdesc[(size_t)(U.sub_off + q) * MOM_WS + k] / (MOM_BS_HOST * 8)];
                V[q] = v;
            }
            for (int f = 0; f < U.nfunc; ++f) {
                const uint32_t d = H.funcdesc[U.func_off + f];
                F[f] = V[(d & 0xffffu) / (MOM_BS_HOST * 8)] * V[(d >> 16) / (MOM_BS_HOST * 8)];
            }
            for (int t = 0; t < MOM_TASKS; ++t) {
                const MomTask& mt = H.tasks[U.task_off + t];
                const MomTaskOut& to = H.touts[U.task_off + t];
                for (int a = 0; a < to.vm; ++a)
                    for (int b = 0; b < to.vn; ++b)
                        mom[(size_t)(to.base + (int64_t)a * to.sm + (int64_t)b * to.sn)] += F[mt.moff + a] * F[mt.noff + b];
            }
        }
    }
}

void host_reconstruct(const MomHost& H, const std::vector<double>& mom, const double* y, int64_t n, int64_t ldy, double* S) {
    const int K = H.nterms, M = H.m, L = K + M + 1, W = H.width;
    const MomGeom g = H.geom();
    for (int i = 0; i < K; ++i)
        for (int j = i; j < K; ++j) {
            const double v = mom_pair(g, H.lin.data(), mom.data(), (double)n, &H.tdim[(size_t)i * W], &H.tord[(size_t)i * W], H.tnnz[i],
                                      &H.tdim[(size_t)j * W], &H.tord[(size_t)j * W], H.tnnz[j]);
            S[(size_t)i * L + j] = v;
            S[(size_t)j * L + i] = v;
        }
    std::vector<double> ysum(M, 0.0), yy((size_t)M * M, 0.0);
    for (int64_t r = 0; r < n; ++r)
        for (int a = 0; a < M; ++a) {
            ysum[a] += y[r * ldy + a];
            for (int b = 0; b < M; ++b) yy[(size_t)a * M + b] += y[r * ldy + a] * y[r * ldy + b];
        }
    for (int k = 0; k < K; ++k) {
        int dims[MOM_MAXORD], ks[MOM_MAXORD];
        const int ne = H.tnnz[k];
        for (int e = 0; e < ne; ++e) { dims[e] = H.tdim[(size_t)k * W + e]; ks[e] = H.tord[(size_t)k * W + e]; }
        for (int w = 0; w < M; ++w) {
            const double v = ne == 0 ? ysum[w] : mom[(size_t)mom_loc(g, dims, ks, ne, 1, w)];
            S[(size_t)k * L + K + w] = v;
            S[(size_t)(K + w) * L + k] = v;
        }
        const double v1 = ne == 0 ? (double)n : mom[(size_t)mom_loc(g, dims, ks, ne, 0, 0)];
        S[(size_t)k * L + L - 1] = v1;
        S[(size_t)(L - 1) * L + k] = v1;
    }
    for (int a = 0; a < M; ++a) {
        for (int b = 0; b < M; ++b) S[(size_t)(K + a) * L + K + b] = yy[(size_t)a * M + b];
        S[(size_t)(K + a) * L + L - 1] = ysum[a];
        S[(size_t)(L - 1) * L + K + a] = ysum[a];
    }
    S[(size_t)(L - 1) * L + L - 1] = (double)n;
}

}  // namespace

extern "C" int pcq_debug_moments_host(int sdim, int nterms, const int32_t* mi, int nord, const double* rec_a_ext,
                                      const double* rec_b_ext, const double* p1_scale, const double* x, int64_t n,
                                      int64_t ldx, const double* y, int64_t ldy, int m, double* s_out, double* info) {
    if (sdim <= 0 || nterms <= 0 || !mi || !rec_a_ext || !rec_b_ext || !p1_scale || n < 0 || m < 0 || (n > 0 && !x) ||
        (n > 0 && m > 0 && !y)) {
        pcq_set_error("debug_moments_host: bad argument");
        return PCQ_ERR_ARG;
    }
    MomHost H;
    mom_build(H, sdim, nterms, mi, nord, rec_a_ext, rec_b_ext, p1_scale, m);
    if (info) {
        info[0] = (double)H.nmom; info[1] = (double)H.units.size(); info[2] = H.clk_per_block; info[3] = H.mac_slots;
        info[4] = H.nf_max; info[5] = H.nsub_max; info[6] = H.why.empty() ? 1.0 : 0.0; info[7] = H.amp;
    }
    if (!H.why.empty()) { pcq_set_error("moment form not applicable: %s", H.why.c_str()); return PCQ_ERR_UNSUPPORTED; }
    if (s_out) {
        std::vector<double> mom;
        host_moments(H, x, n, ldx, y, ldy, mom);
        host_reconstruct(H, mom, y, n, ldy, s_out);
    }
    return PCQ_OK;
}

// ===================================================================================== device side
namespace {

constexpr int MOM_BS = 16;         // samples per block (four DMMA k-steps)
constexpr int MOM_VB = 2;          // sub-function items a thread keeps in flight (the 128 accumulator registers bound them)


constexpr int MOM_NST = 3;         // stages of the table ring

// shared memory: T ring [MOM_NST][nt][16] | V [2][nsubcap][16] | function descriptors | sub-function descriptors |
// tasks | mbarriers | work item.  nsubcap is a multiple of 32 MOM_VB, nfcap of 8.
inline size_t mom_smem_bytes(int nt, int nsubcap, int nfcap) {
    return (size_t)(MOM_NST * nt + 2 * nsubcap) * MOM_BS * 8 + (size_t)nfcap * 4 + (size_t)nsubcap * MOM_WS * 2 +
           MOM_TASKS * sizeof(MomTask) + MOM_NST * 8 + 16;
}

// 32-bit shared-window addresses throughout: with 128 accumulator registers live the compiler has no room for 64-bit
// generic pointers and re-derives them from the parameter bank otherwise
__device__ __forceinline__ double2 lds128(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ uint32_t off_of(const uint4& d, int q) {      // q-th 16-bit byte offset of a sub-function descriptor
    const uint32_t w = q < 2 ? d.x : q < 4 ? d.y : q < 6 ? d.z : d.w;
    return (q & 1) ? (w >> 16) : (w & 0xffffu);
}

// Sub-function rows of a block are 16 samples = two 64-byte halves; the halves of ODD rows are swapped, so that the two
// rows a quarter-warp reads in one fragment load (consecutive functions -> the same or consecutive sub-functions) fall
// into different banks.  Byte offset of (row offset `off` = 128 row, half h):  off + ((h ^ (row & 1)) << 6).
__device__ __forceinline__ uint32_t vrow_half0(uint32_t off) { return off + ((off >> 1) & 64u); }

// Table of 1-D values of a sample chunk, once for all units: tab[block of 16 samples][slot][16]  (slot 0 = 1, slots of
// psi_1 .. psi_B0 per dimension, the outputs, the zero slot; samples past n are all-zero columns).  8 nt bytes per
// sample, read back by the moment kernel with one bulk copy per block.
struct MomTabParams {
    const double* x; int64_t ldx; const double* y; int64_t ldy;
    int64_t n;
    int s, m, B0, nslots, zero_slot, nt;
    const double* rec_a; const double* rec_b; const double* p1;
    double* tab;
};
__global__ void __launch_bounds__(256) pcq_mom_table_kernel(MomTabParams P) {
    const int64_t nb = (P.n + MOM_BS - 1) / MOM_BS;
    const int s = P.s, tid = threadIdx.x;
    for (int64_t b = blockIdx.x; b < nb; b += gridDim.x) {
        double* T = P.tab + (size_t)b * P.nt * MOM_BS;
        const int64_t n0 = b * MOM_BS;
        for (int idx = tid; idx < MOM_BS * s; idx += 256) {
            const int smp = idx & (MOM_BS - 1), dim = idx >> 4;
            const bool live = n0 + smp < P.n;
            const double alive = live ? 1.0 : 0.0;
            const double xv = live ? P.x[(n0 + smp) * P.ldx + dim] : 0.0;
            double pm = alive, pc = xv * P.p1[dim] * alive;
            if (dim == 0) T[smp] = alive;
            T[(size_t)(1 + dim) * MOM_BS + smp] = pc;
            for (int nn = 1; nn < P.B0; ++nn) {
                const double pn = fma(P.rec_a[nn * s + dim] * xv, pc, P.rec_b[nn * s + dim] * pm);
                T[(size_t)(1 + nn * s + dim) * MOM_BS + smp] = pn;
                pm = pc;
                pc = pn;
            }
        }
        for (int idx = tid; idx < MOM_BS * (P.m + 1); idx += 256) {
            const int smp = idx & (MOM_BS - 1), w = idx >> 4;
            T[(size_t)(P.nslots + w) * MOM_BS + smp] = (w < P.m && n0 + smp < P.n) ? P.y[(n0 + smp) * P.ldy + w] : 0.0;     // w == m: the zero slot
        }
    }
}

struct MomParams {
    const double* tab;     // table of this chunk
    int64_t nblocks;       // 16-sample blocks of this chunk
    int nt;
    const MomUnit* units; const MomTask* tasks; const uint32_t* funcdesc; const uint16_t* subdesc;
    int nunits, nsplit, nfcap, nsubcap;      // units of this launch's class, sample parts per unit
    int unit0, item0;                        // first unit of the class; work items (unit, part) before it in ws
    int nvb;                                 // warp-specialised kernel: V buffers in its ring (2 or 3)
    unsigned long long* dbg;                 // PCQ_MOMENT_DBG: cycle counters of one consumer / one producer thread per CTA
    unsigned int* counter;
    double* ws;            // [work item][8 warps][32 x 64]
};

// FM x FN: the task shape (fragments of 8 rows / columns) every warp of the launch runs; units [unit0, unit0 + nunits)
template <int FM, int FN>
__global__ void __launch_bounds__(256, 1) pcq_mom_kernel(MomParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* T = reinterpret_cast<double*>(smem_raw);                 // [MOM_NST][nt][16]
    double* V = T + (size_t)MOM_NST * P.nt * MOM_BS;                  // [2][nsubcap][16]
    uint32_t* fdesc = reinterpret_cast<uint32_t*>(V + (size_t)2 * P.nsubcap * MOM_BS);
    uint16_t* sdesc = reinterpret_cast<uint16_t*>(fdesc + P.nfcap);
    MomTask* tk = reinterpret_cast<MomTask*>(sdesc + (size_t)P.nsubcap * MOM_WS);
    uint64_t* full = reinterpret_cast<uint64_t*>(tk + MOM_TASKS);
    int* sh = reinterpret_cast<int*>(full + MOM_NST);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = warp >> 2;
    const int g = lane >> 2, tig = lane & 3;
    const uint32_t tbytes = (uint32_t)P.nt * MOM_BS * 8, vbytes = (uint32_t)P.nsubcap * MOM_BS * 8;
    if (tid == 0) {
        for (int k = 0; k < MOM_NST; ++k) mbar_init(full + k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int nitems = P.nunits * P.nsplit;
    const uint32_t aT = smem_u32(T), aV = smem_u32(V), aSD = smem_u32(sdesc);
    const int sp = tid & 7;                                           // sample pair of this thread in the sub-function pass
    uint32_t cnt = 0;                                                 // table blocks consumed so far (ring position)

    for (;;) {
        __syncthreads();
        if (tid == 0) sh[0] = (int)atomicAdd(P.counter, 1u);
        __syncthreads();
        const int item = sh[0];
        if (item >= nitems) break;
        const int u = item / P.nsplit, part = item - u * P.nsplit;
        const MomUnit U = P.units[P.unit0 + u];
        const int nsub = U.nsub;
        const uint32_t wpack = (uint32_t)U.wsub;
        const int nsp = (nsub + 127) & ~127;      // the sub-function pass runs whole batches (64 or 128 rows)
        for (int i = tid; i < U.nfunc; i += 256) fdesc[i] = P.funcdesc[U.func_off + i];
        for (int i = tid; i < nsp * MOM_WS; i += 256) sdesc[i] = i < nsub * MOM_WS ? P.subdesc[(size_t)U.sub_off * MOM_WS + i] : (uint16_t)0;
        if (tid < MOM_TASKS) tk[tid] = P.tasks[U.task_off + tid];
        const int64_t b0 = (int64_t)part * P.nblocks / P.nsplit, b1 = (int64_t)(part + 1) * P.nblocks / P.nsplit;
        // table blocks b0 .. of this item into the ring (every stage is free: the last item consumed all it loaded)
        if (tid == 0) {
            for (int k = 0; k < MOM_NST && b0 + k < b1; ++k) {
                const uint32_t st = (cnt + k) % MOM_NST;
                mbar_expect_tx(full + st, tbytes);
                tma_bulk_g2s(reinterpret_cast<unsigned char*>(T) + (size_t)st * tbytes, P.tab + (size_t)(b0 + k) * P.nt * MOM_BS, tbytes, full + st);
            }
        }
        __syncthreads();
        const MomTask mt = tk[warp];
        const int fm = mt.fm, fn = mt.fn;
        // this lane's rows of the task's fragments: shared addresses (buffer 0, half 0, this lane's sample pair) of the
        // two sub-function rows whose product is the function; rows past the task's own fragments read the zero function
        uint32_t a1[FM], a2[FM], c1[FN], c2[FN];
#pragma unroll
        for (int i = 0; i < FM; ++i) {
            const uint32_t d = i < fm ? fdesc[mt.moff + i * 8 + g] : 0u;
            a1[i] = aV + vrow_half0(d & 0xffffu) + tig * 16;
            a2[i] = aV + vrow_half0(d >> 16) + tig * 16;
        }
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const uint32_t d = j < fn ? fdesc[mt.noff + j * 8 + g] : 0u;
            c1[j] = aV + vrow_half0(d & 0xffffu) + tig * 16;
            c2[j] = aV + vrow_half0(d >> 16) + tig * 16;
        }

        double acc[FM][FN][2];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        // sub-functions of block number k of this item (ring position cnt + k): item = (sub-function, sample pair),
        // MOM_VB items in flight; descriptors hold byte offsets of the table rows
        constexpr int VB = (FM * FN >= 32) ? 2 : 4;      // sub-function items a thread keeps in flight: what the accumulators leave room for
        auto gen_v = [&](uint32_t k) {
            const uint32_t c = cnt + k, st = c % MOM_NST;
            mbar_wait(full + st, (c / MOM_NST) & 1u);
            const uint32_t tb = aT + st * tbytes + sp * 16;
            uint32_t sd = aSD + (tid >> 3) * (MOM_WS * 2);
            uint32_t vo = aV + (k & 1u) * vbytes + (tid >> 3) * (MOM_BS * 8) + ((((sp >> 2) ^ (tid >> 3)) & 1) << 6) + (sp & 3) * 16;
            uint32_t wp = wpack;
            for (int i0 = 0; i0 < nsub; i0 += 32 * VB, sd += 32 * VB * MOM_WS * 2, vo += 32 * VB * MOM_BS * 8, wp >>= (VB == 4 ? 8 : 4)) {
                int wsub = (int)(wp & 15u);                  // factors of this batch's rows (planner batches of 64)
                if (VB == 4) { const int w1 = (int)((wp >> 4) & 15u); wsub = w1 > wsub ? w1 : wsub; }
                uint4 dd[VB];
                double2 v[VB];
#pragma unroll
                for (int q = 0; q < VB; ++q) dd[q] = lds128u(sd + q * (32 * MOM_WS * 2));
#pragma unroll
                for (int q = 0; q < VB; ++q) v[q] = lds128(tb + off_of(dd[q], 0));
                // factors two at a time: their product does not wait for v
#pragma unroll
                for (int f = 1; f < MOM_WS; f += 2) {
                    if (f < wsub) {
                        double2 t[VB], w2[VB];
#pragma unroll
                        for (int q = 0; q < VB; ++q) t[q] = lds128(tb + off_of(dd[q], f));
#pragma unroll
                        for (int q = 0; q < VB; ++q) w2[q] = lds128(tb + ((f + 1 < MOM_WS) ? off_of(dd[q], (f + 1 < MOM_WS) ? f + 1 : f) : 0u));
#pragma unroll
                        for (int q = 0; q < VB; ++q) { t[q].x *= w2[q].x; t[q].y *= w2[q].y; }
#pragma unroll
                        for (int q = 0; q < VB; ++q) { v[q].x *= t[q].x; v[q].y *= t[q].y; }
                    }
                }
#pragma unroll
                for (int q = 0; q < VB; ++q) sts128(vo + q * (32 * MOM_BS * 8), v[q].x, v[q].y);     // 32 rows on: same parity
            }
        };
        // the DMMAs of block number k for this warp's task; every function value is the product of its two
        // sub-function values.  No branch in here: mma.sync under a thread-dependent condition costs a WARPSYNC each.
        auto mma = [&](uint32_t k) {
            const uint32_t bo = (k & 1u) * vbytes;
            constexpr int JB = FN < 4 ? FN : 4;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t hx = (uint32_t)h << 6;
                double2 af[FM];
#pragma unroll
                for (int i = 0; i < FM; ++i) {
                    const double2 p = lds128((a1[i] ^ hx) + bo), q = lds128((a2[i] ^ hx) + bo);
                    af[i] = make_double2(p.x * q.x, p.y * q.y);
                }
#pragma unroll
                for (int jb = 0; jb < FN; jb += JB) {
                    double2 bf[JB];
#pragma unroll
                    for (int j = 0; j < JB; ++j) {
                        const double2 p = lds128((c1[jb + j] ^ hx) + bo), q = lds128((c2[jb + j] ^ hx) + bo);
                        bf[j] = make_double2(p.x * q.x, p.y * q.y);
                    }
#pragma unroll
                    for (int i = 0; i < FM; ++i)
#pragma unroll
                        for (int j = 0; j < JB; ++j) dmma(acc[i][jb + j][0], acc[i][jb + j][1], af[i].x, bf[j].x);
#pragma unroll
                    for (int i = 0; i < FM; ++i)
#pragma unroll
                        for (int j = 0; j < JB; ++j) dmma(acc[i][jb + j][0], acc[i][jb + j][1], af[i].y, bf[j].y);
                }
            }
        };

        // per block k: sub-functions of k + 1 and the DMMAs of k, one barrier; then the ring stage block k + 1 came from
        // is refilled.  Warps 0-3 generate first and multiply second, warps 4-7 the other way round, so each
        // scheduler always has one warp of DMMAs in its FP64 pipe.
        const uint32_t nblk = (uint32_t)(b1 - b0);
        if (nblk > 0) {
            gen_v(0);
            __syncthreads();
            if (tid == 0 && MOM_NST < nblk) {
                const uint32_t st = cnt % MOM_NST;
                mbar_expect_tx(full + st, tbytes);
                tma_bulk_g2s(reinterpret_cast<unsigned char*>(T) + (size_t)st * tbytes, P.tab + (size_t)(b0 + MOM_NST) * P.nt * MOM_BS, tbytes, full + st);
            }
        }
        for (uint32_t k = 0; k < nblk; ++k) {
#pragma unroll 1
            for (int ph = 0; ph < 2; ++ph) {
                if (ph == grp && k + 1 < nblk) gen_v(k + 1);
                if (ph == 0) mma(k);
            }
            __syncthreads();
            if (tid == 0 && k + 1 + MOM_NST < nblk) {
                const uint32_t st = (cnt + k + 1) % MOM_NST;
                mbar_expect_tx(full + st, tbytes);
                tma_bulk_g2s(reinterpret_cast<unsigned char*>(T) + (size_t)st * tbytes, P.tab + (size_t)(b0 + k + 1 + MOM_NST) * P.nt * MOM_BS, tbytes, full + st);
            }
        }
        cnt += nblk;

        double* w = P.ws + ((size_t)(P.item0 + item) * MOM_TASKS + warp) * 2048;
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j)
                if (i < fm && j < fn)
                    *reinterpret_cast<double2*>(w + (i * 8 + g) * (FN * 8) + j * 8 + 2 * tig) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// Warp-specialised form of the same kernel (PCQ_MOMENT_WS, default where its shared memory fits): 12 warps.
//   warps 8-11 (one per scheduler, 72 registers after setmaxnreg) GENERATE the sub-function values of block k + 1 .. into
//   a ring of V buffers; warps 0-7 (216 registers) only multiply: both warps of a scheduler issue DMMAs all the time,
//   instead of one generating while the other multiplies.  Hand-over by mbarriers (vfull: 128 producer arrivals,
//   vempty: one arrival per consumer warp); the CTA-wide barrier is only used between work items.
constexpr int MOM_WS_THREADS = 384;
// register budgets after setmaxnreg: the consumers can only take what the producers give back to the CTA's pool,
// 8 x 32 x RC + 4 x 32 x RP <= 384 x 168 (more and the TRY_ALLOC of the consumers never returns)
constexpr int MOM_WS_RC = 216, MOM_WS_RP = 72;
static_assert(2 * MOM_WS_RC + MOM_WS_RP <= 3 * 168 && MOM_WS_RC % 8 == 0 && MOM_WS_RP % 8 == 0, "setmaxnreg budgets exceed the CTA's registers");
constexpr int MOM_PVB = 4;           // sub-function items a producer thread keeps in flight
inline size_t mom_ws_smem_bytes(int nt, int nsubcap, int nfcap, int nvb) {
    return (size_t)(MOM_NST * nt + nvb * nsubcap) * MOM_BS * 8 + (size_t)nfcap * 4 + (size_t)nsubcap * MOM_WS * 2 +
           MOM_TASKS * sizeof(MomTask) + (MOM_NST + 6) * 8 + 16;
}

template <int FM, int FN>
__global__ void __launch_bounds__(MOM_WS_THREADS, 1) pcq_mom_ws_kernel(MomParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* T = reinterpret_cast<double*>(smem_raw);                 // [MOM_NST][nt][16]
    double* V = T + (size_t)MOM_NST * P.nt * MOM_BS;                  // [nvb][nsubcap][16]
    uint32_t* fdesc = reinterpret_cast<uint32_t*>(V + (size_t)P.nvb * P.nsubcap * MOM_BS);
    uint16_t* sdesc = reinterpret_cast<uint16_t*>(fdesc + P.nfcap);
    MomTask* tk = reinterpret_cast<MomTask*>(sdesc + (size_t)P.nsubcap * MOM_WS);
    uint64_t* tfull = reinterpret_cast<uint64_t*>(tk + MOM_TASKS);
    uint64_t* vfull = tfull + MOM_NST;
    uint64_t* vempty = vfull + 3;
    int* sh = reinterpret_cast<int*>(vempty + 3);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp >= 8;
    const uint32_t tbytes = (uint32_t)P.nt * MOM_BS * 8, vbytes = (uint32_t)P.nsubcap * MOM_BS * 8;
    if (tid == 0) {
        for (int k = 0; k < MOM_NST; ++k) mbar_init(tfull + k, 1);
        for (int k = 0; k < 3; ++k) { mbar_init(vfull + k, 128); mbar_init(vempty + k, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nitems = P.nunits * P.nsplit;
    const uint32_t aT = smem_u32(T), aV = smem_u32(V), aSD = smem_u32(sdesc);
    // ring positions, advanced identically by both roles: table stage / its use count, V buffer / its use count
    uint32_t tst = 0, tu = 0, vb = 0, vu = 0;

    // next work item and its descriptors; both roles run this at the same points (CTA barrier 0 counts arrivals, the
    // two role loops below are separate code so that each gets the register budget of its setmaxnreg)
    struct Item { int item, nsub; uint32_t wsub, nblk; int64_t b0; };
    auto next_item = [&]() -> Item {
        Item it;
        asm volatile("bar.sync 0, 384;" ::: "memory");
        if (tid == 0) sh[0] = (int)atomicAdd(P.counter, 1u);
        asm volatile("bar.sync 0, 384;" ::: "memory");
        it.item = sh[0];
        it.nsub = 0; it.wsub = 0; it.nblk = 0; it.b0 = 0;
        if (it.item >= nitems) return it;
        const int u = it.item / P.nsplit, part = it.item - u * P.nsplit;
        const MomUnit U = P.units[P.unit0 + u];
        it.nsub = U.nsub;
        it.wsub = (uint32_t)U.wsub;
        const int nsp = (U.nsub + 127) & ~127;
        for (int i = tid; i < U.nfunc; i += MOM_WS_THREADS) fdesc[i] = P.funcdesc[U.func_off + i];
        for (int i = tid; i < nsp * MOM_WS; i += MOM_WS_THREADS) sdesc[i] = i < U.nsub * MOM_WS ? P.subdesc[(size_t)U.sub_off * MOM_WS + i] : (uint16_t)0;
        if (tid < MOM_TASKS) tk[tid] = P.tasks[U.task_off + tid];
        it.b0 = (int64_t)part * P.nblocks / P.nsplit;
        it.nblk = (uint32_t)((int64_t)(part + 1) * P.nblocks / P.nsplit - it.b0);
        asm volatile("bar.sync 0, 384;" ::: "memory");
        return it;
    };

    if (producer) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MOM_WS_RP));
        const int ptid = tid - 256, sp = ptid & 7, prow = ptid >> 3;      // sample pair, row of a 16-row pass
        for (;;) {
            const Item it = next_item();
            if (it.item >= nitems) break;
            const int nsub = it.nsub;
            const uint32_t nblk = it.nblk;
            const int64_t b0 = it.b0;
            if (ptid == 0) {
                for (uint32_t k = 0; k < MOM_NST && k < nblk; ++k) {
                    const uint32_t st = (tst + k) % MOM_NST;
                    mbar_expect_tx(tfull + st, tbytes);
                    tma_bulk_g2s(reinterpret_cast<unsigned char*>(T) + (size_t)st * tbytes, P.tab + (size_t)(b0 + k) * P.nt * MOM_BS, tbytes, tfull + st);
                }
            }
            long long pe = 0, pt = 0, p0 = P.dbg ? clock64() : 0;
            for (uint32_t k = 0; k < nblk; ++k) {
                if (P.dbg) {
                    const long long a = clock64();
                    if (vu > 0) mbar_wait(vempty + vb, (vu - 1) & 1u);
                    const long long b = clock64();
                    mbar_wait(tfull + tst, tu & 1u);
                    pe += b - a; pt += clock64() - b;
                } else {
                if (vu > 0) mbar_wait(vempty + vb, (vu - 1) & 1u);        // the consumers are done with this buffer's last use
                mbar_wait(tfull + tst, tu & 1u);
                }
                const uint32_t tb = aT + tst * tbytes + sp * 16;
                uint32_t sd = aSD + prow * (MOM_WS * 2);
                uint32_t vo = aV + vb * vbytes + prow * (MOM_BS * 8) + ((((sp >> 2) ^ prow) & 1) << 6) + (sp & 3) * 16;
                // planner batches of 64 rows (one factor count each) = two passes of 16 rows x MOM_PVB
                for (int i0 = 0; i0 < nsub; i0 += 16 * MOM_PVB, sd += 16 * MOM_PVB * MOM_WS * 2, vo += 16 * MOM_PVB * MOM_BS * 8) {
                    const int wsub = (int)((it.wsub >> (4 * (i0 >> 6))) & 15u);
                    uint4 dd[MOM_PVB];
                    double2 v[MOM_PVB];
#pragma unroll
                    for (int q = 0; q < MOM_PVB; ++q) dd[q] = lds128u(sd + q * (16 * MOM_WS * 2));
#pragma unroll
                    for (int q = 0; q < MOM_PVB; ++q) v[q] = lds128(tb + off_of(dd[q], 0));
#pragma unroll
                    for (int f = 1; f < MOM_WS; f += 2) {
                        if (f < wsub) {
                            double2 t[MOM_PVB], w2[MOM_PVB];
#pragma unroll
                            for (int q = 0; q < MOM_PVB; ++q) t[q] = lds128(tb + off_of(dd[q], f));
#pragma unroll
                            for (int q = 0; q < MOM_PVB; ++q) w2[q] = lds128(tb + ((f + 1 < MOM_WS) ? off_of(dd[q], (f + 1 < MOM_WS) ? f + 1 : f) : 0u));
#pragma unroll
                            for (int q = 0; q < MOM_PVB; ++q) { t[q].x *= w2[q].x; t[q].y *= w2[q].y; }
#pragma unroll
                            for (int q = 0; q < MOM_PVB; ++q) { v[q].x *= t[q].x; v[q].y *= t[q].y; }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < MOM_PVB; ++q) sts128(vo + q * (16 * MOM_BS * 8), v[q].x, v[q].y);      // 16 rows on: same parity
                }
                mbar_arrive(vfull + vb);
                // the table stage is free once all four producer warps have read it
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (ptid == 0 && k + MOM_NST < nblk) {
                    mbar_expect_tx(tfull + tst, tbytes);
                    tma_bulk_g2s(reinterpret_cast<unsigned char*>(T) + (size_t)tst * tbytes, P.tab + (size_t)(b0 + k + MOM_NST) * P.nt * MOM_BS, tbytes, tfull + tst);
                }
                if (++tst == MOM_NST) { tst = 0; ++tu; }
                if (++vb == (uint32_t)P.nvb) { vb = 0; ++vu; }
            }
            if (P.dbg && ptid == 0) {
                atomicAdd(P.dbg + 3, (unsigned long long)(clock64() - p0));
                atomicAdd(P.dbg + 4, (unsigned long long)pe);
                atomicAdd(P.dbg + 5, (unsigned long long)pt);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MOM_WS_RC));
        const int g = lane >> 2, tig = lane & 3;
        for (;;) {
            const Item it = next_item();
            if (it.item >= nitems) break;
            const uint32_t nblk = it.nblk;
            const MomTask mt = tk[warp];
            const int fm = mt.fm, fn = mt.fn;
            uint32_t a1[FM], a2[FM], c1[FN], c2[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) {
                const uint32_t d = i < fm ? fdesc[mt.moff + i * 8 + g] : 0u;
                a1[i] = aV + vrow_half0(d & 0xffffu) + tig * 16;
                a2[i] = aV + vrow_half0(d >> 16) + tig * 16;
            }
#pragma unroll
            for (int j = 0; j < FN; ++j) {
                const uint32_t d = j < fn ? fdesc[mt.noff + j * 8 + g] : 0u;
                c1[j] = aV + vrow_half0(d & 0xffffu) + tig * 16;
                c2[j] = aV + vrow_half0(d >> 16) + tig * 16;
            }
            double acc[FM][FN][2];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

            long long tw = 0, t0 = P.dbg ? clock64() : 0;
            for (uint32_t k = 0; k < nblk; ++k) {
                if (P.dbg) { const long long a = clock64(); mbar_wait(vfull + vb, vu & 1u); tw += clock64() - a; }
                else mbar_wait(vfull + vb, vu & 1u);
                const uint32_t bo = vb * vbytes;
                constexpr int JB = FN < 4 ? FN : (FM <= 2 ? 8 : 4);      // column fragments per batch of loads: more where the accumulators leave room
                // a warp without a task (the last unit of a shape) leaves the pipe to the other warp of its scheduler
#pragma unroll
                for (int h = 0; h < (fm ? 2 : 0); ++h) {
                    const uint32_t hx = (uint32_t)h << 6;
                    double2 af[FM];
#pragma unroll
                    for (int i = 0; i < FM; ++i) {
                        const double2 p = lds128((a1[i] ^ hx) + bo), q = lds128((a2[i] ^ hx) + bo);
                        af[i] = make_double2(p.x * q.x, p.y * q.y);
                    }
#pragma unroll
                    for (int jb = 0; jb < FN; jb += JB) {
                        double2 bf[JB];
#pragma unroll
                        for (int j = 0; j < JB; ++j) {
                            const double2 p = lds128((c1[jb + j] ^ hx) + bo), q = lds128((c2[jb + j] ^ hx) + bo);
                            bf[j] = make_double2(p.x * q.x, p.y * q.y);
                        }
#pragma unroll
                        for (int i = 0; i < FM; ++i)
#pragma unroll
                            for (int j = 0; j < JB; ++j) dmma(acc[i][jb + j][0], acc[i][jb + j][1], af[i].x, bf[j].x);
#pragma unroll
                        for (int i = 0; i < FM; ++i)
#pragma unroll
                            for (int j = 0; j < JB; ++j) dmma(acc[i][jb + j][0], acc[i][jb + j][1], af[i].y, bf[j].y);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(vempty + vb);
                if (++vb == (uint32_t)P.nvb) { vb = 0; ++vu; }
            }
            if (P.dbg && tid == 0) {
                atomicAdd(P.dbg + 0, (unsigned long long)(clock64() - t0));
                atomicAdd(P.dbg + 1, (unsigned long long)tw);
                atomicAdd(P.dbg + 2, (unsigned long long)nblk);
            }
            double* w = P.ws + ((size_t)(P.item0 + it.item) * MOM_TASKS + warp) * 2048;
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j)
                    if (i < fm && j < fn)
                        *reinterpret_cast<double2*>(w + (i * 8 + g) * (FN * 8) + j * 8 + 2 * tig) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

// moments of one task = its partial tiles summed over the sample parts, in part order; grid (nunits * 8)
__global__ void __launch_bounds__(256) pcq_mom_fixup_kernel(const MomUnit* units, const MomTaskOut* touts, const double* ws,
                                                            int unit0, int item0, int nsplit, int stride, int accumulate, double* mom) {
    const int u = blockIdx.x >> 3, slot = blockIdx.x & 7;
    const MomTaskOut to = touts[units[unit0 + u].task_off + slot];
    const int total = to.vm * to.vn;
    for (int e = threadIdx.x; e < total; e += 256) {
        const int i = e / to.vn, j = e - i * to.vn;
        const double* w = ws + (((size_t)item0 + (size_t)u * nsplit) * MOM_TASKS + slot) * 2048 + i * stride + j;
        double sum = 0.0;
        for (int q = 0; q < nsplit; ++q) sum += w[(size_t)q * MOM_TASKS * 2048];
        double* dst = mom + to.base + (int64_t)i * to.sm + (int64_t)j * to.sn;
        *dst = (accumulate ? *dst : 0.0) + sum;
    }
}

// sum_n y_a, sum_n y_a y_b (a <= b) of up to 8 outputs: per-block partials, combined in block order by the caller's
// second launch (nblocks == 1 over the partials)
constexpr int MOM_YV = MOM_MAXOUT + MOM_MAXOUT * (MOM_MAXOUT + 1) / 2;
__global__ void __launch_bounds__(256) pcq_mom_yy_kernel(const double* y, int64_t ldy, int64_t n, int m, double* partial) {
    double loc[MOM_YV];
#pragma unroll
    for (int i = 0; i < MOM_YV; ++i) loc[i] = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < n; r += (int64_t)gridDim.x * 256) {
        double yv[MOM_MAXOUT];
#pragma unroll
        for (int a = 0; a < MOM_MAXOUT; ++a) yv[a] = a < m ? y[r * ldy + a] : 0.0;
        int q = MOM_MAXOUT;
#pragma unroll
        for (int a = 0; a < MOM_MAXOUT; ++a) {
            loc[a] += yv[a];
#pragma unroll
            for (int b = a; b < MOM_MAXOUT; ++b) loc[q++] += yv[a] * yv[b];
        }
    }
    __shared__ double red[8][MOM_YV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < MOM_YV; ++i) {
        double v = loc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < MOM_YV) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * MOM_YV + threadIdx.x] = v;
    }
}
__global__ void pcq_mom_yy_final_kernel(const double* partial, int nblocks, double* out) {
    if (threadIdx.x < MOM_YV) {
        double v = 0.0;
        for (int b = 0; b < nblocks; ++b) v += partial[(size_t)b * MOM_YV + threadIdx.x];
        out[threadIdx.x] = v;
    }
}

struct MomRecon {
    MomGeom g;
    const double* lin; const double* mom; const double* yy;    // yy: MOM_YV sums of the outputs
    const int16_t* tdim; const uint8_t* tord; const uint8_t* tnnz;
    int K, M, W;
    int ML;                // outputs S is laid out for (L = K + ML + 1); M of them (0 or ML) are served here
    double nsamples;
    double* S; int accumulate;
};

__global__ void __launch_bounds__(256) pcq_mom_pairs_kernel(MomRecon R) {
    const int j = blockIdx.x * 16 + (threadIdx.x & 15), i = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (i > j || j >= R.K) return;
    const int L = R.K + R.ML + 1;
    const double v = mom_pair(R.g, R.lin, R.mom, R.nsamples, R.tdim + (size_t)i * R.W, R.tord + (size_t)i * R.W, R.tnnz[i],
                              R.tdim + (size_t)j * R.W, R.tord + (size_t)j * R.W, R.tnnz[j]);
    double* a = R.S + (size_t)i * L + j;
    *a = (R.accumulate ? *a : 0.0) + v;
    if (i != j) {
        double* b = R.S + (size_t)j * L + i;
        *b = (R.accumulate ? *b : 0.0) + v;
    }
}

// the y and 1 columns / rows and the (M + 1) x (M + 1) corner
__global__ void __launch_bounds__(256) pcq_mom_edges_kernel(MomRecon R) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    const int K = R.K, M = R.M, L = K + R.ML + 1;
    auto put = [&](int r, int c, double v) {
        double* a = R.S + (size_t)r * L + c;
        *a = (R.accumulate ? *a : 0.0) + v;
    };
    if (k < K) {
        int dims[MOM_MAXORD], ks[MOM_MAXORD];
        const int ne = R.tnnz[k];
        for (int e = 0; e < ne; ++e) { dims[e] = R.tdim[(size_t)k * R.W + e]; ks[e] = R.tord[(size_t)k * R.W + e]; }
        for (int w = 0; w < M; ++w) {
            const double v = ne == 0 ? R.yy[w] : R.mom[mom_loc(R.g, dims, ks, ne, 1, w)];
            put(k, K + w, v);
            put(K + w, k, v);
        }
        const double v1 = ne == 0 ? R.nsamples : R.mom[mom_loc(R.g, dims, ks, ne, 0, 0)];
        put(k, L - 1, v1);
        put(L - 1, k, v1);
    } else if (k == K) {
        int q = MOM_MAXOUT;
        for (int a = 0; a < MOM_MAXOUT; ++a)
            for (int b = a; b < MOM_MAXOUT; ++b, ++q)
                if (b < M) {
                    put(K + a, K + b, R.yy[q]);
                    if (a != b) put(K + b, K + a, R.yy[q]);
                }
        for (int a = 0; a < M; ++a) { put(K + a, L - 1, R.yy[a]); put(L - 1, K + a, R.yy[a]); }
        put(L - 1, L - 1, R.nsamples);
    }
}

// ------------------------------------------------------------------------------------- per-plan state
struct MomDev {
    MomHost H;
    MomNode* nodes = nullptr; int64_t* base = nullptr; int64_t* wsize = nullptr; int64_t* bandoff = nullptr; int32_t* ncols = nullptr;
    int64_t* cnt = nullptr; int64_t* cum = nullptr;
    double* lin = nullptr; double* rec_a = nullptr; double* rec_b = nullptr; double* p1 = nullptr;
    int16_t* tdim = nullptr; uint8_t* tord = nullptr; uint8_t* tnnz = nullptr;
    MomUnit* units = nullptr; MomTask* tasks = nullptr; MomTaskOut* touts = nullptr; uint32_t* funcdesc = nullptr; uint16_t* subdesc = nullptr;
    double* mom = nullptr; double* yy = nullptr; double* yypart = nullptr; unsigned int* counter = nullptr;
    bool ready = false;
    // the class kernels of a chunk run on their own streams (forked from / joined to the caller's), so that the tail of
    // one launch is filled by the CTAs of the next
    cudaStream_t cs[MOM_NCLS] = {};
    cudaEvent_t fork = nullptr, join[MOM_NCLS] = {};
    std::vector<void*> owned;
    ~MomDev() {
        for (void* q : owned) cudaFree(q);
        for (auto st : cs) if (st) cudaStreamDestroy(st);
        if (fork) cudaEventDestroy(fork);
        for (auto e : join) if (e) cudaEventDestroy(e);
    }
};

struct MomState {
    int nord = 0;
    std::vector<double> rec_a, rec_b;        // (nord + 1) x s
    std::mutex mu;
    MomDev* dev[MOM_MAXOUT + 1] = {};
    ~MomState() { for (MomDev* d : dev) delete d; }
};

template <typename T>
int upload(MomDev& D, T** dst, const std::vector<T>& src) {
    *dst = nullptr;
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    PCQ_CUDA(cudaMalloc((void**)dst, bytes));
    D.owned.push_back(*dst);
    if (!src.empty()) PCQ_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PCQ_OK;
}

int mom_upload(MomDev& D) {
    const MomHost& H = D.H;
    int rc;
#define UP(field) if ((rc = upload(D, &D.field, H.field)) != PCQ_OK) return rc
    UP(nodes); UP(base); UP(wsize); UP(bandoff); UP(ncols); UP(cnt); UP(cum); UP(lin); UP(rec_a); UP(rec_b); UP(p1);
    UP(tdim); UP(tord); UP(tnnz); UP(units); UP(tasks); UP(touts); UP(funcdesc); UP(subdesc);
#undef UP
    PCQ_CUDA(cudaMalloc((void**)&D.mom, std::max<size_t>((size_t)H.nmom, 1) * 8)); D.owned.push_back(D.mom);
    PCQ_CUDA(cudaMalloc((void**)&D.yy, MOM_YV * 8)); D.owned.push_back(D.yy);
    PCQ_CUDA(cudaMalloc((void**)&D.yypart, (size_t)1024 * MOM_YV * 8)); D.owned.push_back(D.yypart);
    PCQ_CUDA(cudaMalloc((void**)&D.counter, 64)); D.owned.push_back(D.counter);
    for (int c = 0; c < MOM_NCLS; ++c) {
        PCQ_CUDA(cudaStreamCreateWithFlags(&D.cs[c], cudaStreamNonBlocking));
        PCQ_CUDA(cudaEventCreateWithFlags(&D.join[c], cudaEventDisableTiming));
    }
    PCQ_CUDA(cudaEventCreateWithFlags(&D.fork, cudaEventDisableTiming));
    D.ready = true;
    return PCQ_OK;
}

}  // namespace

void pcq_mom_free(pcq_plan* p) {
    delete static_cast<MomState*>(p->mom);
    p->mom = nullptr;
}

extern "C" int pcq_plan_enable_moments(pcq_plan* p, int nord, const double* rec_a_ext, const double* rec_b_ext) {
    if (!p || nord < 0 || !rec_a_ext || !rec_b_ext) { pcq_set_error("plan_enable_moments: bad argument"); return PCQ_ERR_ARG; }
    if (p->mom) return PCQ_OK;
    MomState* st = new MomState();
    st->nord = nord;
    st->rec_a.assign(rec_a_ext, rec_a_ext + (size_t)(nord + 1) * p->sdim);
    st->rec_b.assign(rec_b_ext, rec_b_ext + (size_t)(nord + 1) * p->sdim);
    p->mom = st;
    return PCQ_OK;
}

namespace {
// device plan of this output count, built on first use
int mom_get_dev(pcq_plan* p, MomState* ms, int m, MomDev** out) {
    std::lock_guard<std::mutex> lock(ms->mu);
    if (!ms->dev[m]) {
        MomDev* D = new MomDev();
        mom_build(D->H, p->sdim, p->nterms, p->h_mi.data(), ms->nord, ms->rec_a.data(), ms->rec_b.data(), p->h_p1.data(), m);
        if (D->H.why.empty()) {
            if (mom_smem_bytes(D->H.nt, (D->H.nsub_max + 127) & ~127, (D->H.nf_max + 7) & ~7) > p->smem_optin) D->H.why = "shared memory";
            else {
                const int rc = mom_upload(*D);
                if (rc != PCQ_OK) { delete D; return rc; }
            }
        }
        ms->dev[m] = D;
    }
    *out = ms->dev[m];
    return PCQ_OK;
}

// cycles per sample and SM of the two forms (pack + SYRK at 0.89 of the DMMA rate) and this form's cost that does not
// depend on n (reconstruction ~0.3 ns per Gram entry, a dozen small launches): true when the moment form is expected
// to be faster for n samples
bool mom_pays(const pcq_plan* p, const MomHost& H, int m, int64_t n) {
    const double L = (double)p->nterms + m + 1;
    const double syrk = (L * L / 2 * 1.04) / 64.0 / 0.89, mine = H.clk_per_block / MOM_BS;
    const double hz = 1.9e9 * p->num_sms;
    const double fixed = 0.31e-9 * (double)p->nterms * p->nterms + 1.5e-4;
    const char* ae = getenv("PCQ_MOMENT_MAX_AMP");
    if (H.amp > (ae ? atof(ae) : 500.0)) return false;          // high-order Hermite: the moments dwarf the entries they rebuild
    return mine * 1.1 <= syrk && (double)n * (syrk - mine) / hz >= 1.25 * fixed;
}
}  // namespace

extern "C" int pcq_moment_info(pcq_plan* p, int m, int64_t n, double* info) {
    if (!p || m < 0 || !info) { pcq_set_error("moment_info: bad argument"); return PCQ_ERR_ARG; }
    for (int i = 0; i < 8; ++i) info[i] = 0.0;
    MomState* ms = static_cast<MomState*>(p->mom);
    if (!ms) return PCQ_OK;
    const int mlay = m;
    if (m > MOM_MAXOUT) m = 0;      // hybrid: G in moment form, the output columns by the SYRK (pcq_launch_gram)
    MomDev* D = nullptr;
    {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(p->device);
        const int rc = mom_get_dev(p, ms, m, &D);
        if (prev >= 0) cudaSetDevice(prev);
        if (rc != PCQ_OK) return rc;
    }
    const MomHost& H = D->H;
    info[0] = (double)H.nmom; info[1] = (double)H.units.size(); info[2] = H.mac_slots; info[3] = H.clk_per_block / MOM_BS;
    info[4] = D->ready ? 1.0 : 0.0;
    const char* env = getenv("PCQ_GRAM_MOMENT");
    const int mode = env ? atoi(env) : -1;
    const char* e2 = getenv("PCQ_MOMENT_MIN_N");
    info[5] = (D->ready && mode != 0 && n > 0 && (mode > 0 || (n >= (e2 ? atoll(e2) : 32768) && mom_pays(p, H, m, n)))) ? 1.0 : 0.0;
    info[6] = 8.0 * H.nt;
    info[7] = mlay > MOM_MAXOUT ? 1.0 : 0.0;
    return PCQ_OK;
}

static unsigned long long* dbg_buf = nullptr;      // PCQ_MOMENT_DBG: managed cycle counters, 8 per task shape

// S (+)= statistics of n samples through the moment form; *handled = false when the route does not apply (the caller
// falls back to the SYRK form).  PCQ_GRAM_MOMENT = 0 never, 1 whenever a plan exists, unset: when it is estimated faster.
int pcq_launch_moments(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* y, int64_t ldy, int m,
                       double* s_out, int accumulate, double** ws, size_t* ws_bytes, double** zbuf, size_t* zbuf_bytes, cudaStream_t st, bool* handled,
                       int layout_m) {
    *handled = false;
    if (layout_m >= 0 && m != 0) { pcq_set_error("moments: layout_m goes with m == 0"); return PCQ_ERR_ARG; }
    MomState* ms = static_cast<MomState*>(p->mom);
    const char* env = getenv("PCQ_GRAM_MOMENT");
    const int mode = env ? atoi(env) : -1;
    if (!ms || mode == 0 || n <= 0 || m > MOM_MAXOUT) return PCQ_OK;
    if (mode < 0) {
        const char* e2 = getenv("PCQ_MOMENT_MIN_N");
        if (n < (e2 ? atoll(e2) : 32768)) return PCQ_OK;
    }
    MomDev* D = nullptr;
    {
        const int rcd = mom_get_dev(p, ms, m, &D);
        if (rcd != PCQ_OK) return rcd;
    }
    if (!D->ready) return PCQ_OK;
    const MomHost& H = D->H;
    if (mode < 0 && !mom_pays(p, H, m, n)) return PCQ_OK;
    *handled = true;

    const int nunits = (int)H.units.size();
    // sample chunks: the table of 1-D values of a chunk (8 nt bytes per sample) lives in the scratch buffer
    const char* cenv = getenv("PCQ_MOMENT_CHUNK_MB");
    const int64_t target = (int64_t)(cenv ? atoi(cenv) : 6144) << 20;
    int64_t rows = std::max<int64_t>(MOM_BS, target / ((int64_t)H.nt * 8) / MOM_BS * MOM_BS);
    {
        const int64_t nchunks = (n + rows - 1) / rows;
        rows = ((n + nchunks - 1) / nchunks + MOM_BS - 1) / MOM_BS * MOM_BS;          // equal chunks
    }
    int rc = pcq_ensure_dev(zbuf, zbuf_bytes, (size_t)rows * H.nt * 8);
    if (rc != PCQ_OK) return rc;
    // units per class, sample parts per class: at least four work items per SM and launch, so that the items dealt out
    // by the atomic counter (costliest first) end together
    int cls_unit0[MOM_NCLS + 1], cls_split[MOM_NCLS], cls_item0[MOM_NCLS + 1];
    {
        int u0 = 0, it0 = 0;
        const int64_t nb_chunk = rows / MOM_BS;
        const char* senv = getenv("PCQ_MOMENT_SPLIT");
        for (int c = 0; c < MOM_NCLS; ++c) {
            int nu = 0;
            while (u0 + nu < nunits && H.units[u0 + nu].cls == c) ++nu;
            int sp = nu ? (6 * p->num_sms + nu - 1) / nu : 1;
            if (senv) sp = atoi(senv);
            sp = (int)std::max<int64_t>(1, std::min<int64_t>(sp, nb_chunk / 24));
            cls_unit0[c] = u0; cls_split[c] = sp; cls_item0[c] = it0;
            u0 += nu; it0 += nu * sp;
        }
        cls_unit0[MOM_NCLS] = u0; cls_item0[MOM_NCLS] = it0;
    }
    rc = pcq_ensure_dev(ws, ws_bytes, (size_t)cls_item0[MOM_NCLS] * MOM_TASKS * 2048 * 8);
    if (rc != PCQ_OK) return rc;
    const int nfcap = (H.nf_max + 7) & ~7, nsubcap = (H.nsub_max + 127) & ~127;
    const size_t smem = mom_smem_bytes(H.nt, nsubcap, nfcap);
    // warp-specialised kernels (default; PCQ_MOMENT_WS=0: the 8-warp kernels, 2: two V buffers): three V buffers where
    // they fit, else two
    bool use_ws = false;
    int nvb = 3;
    size_t smem_ws = 0;
    {
        const char* we = getenv("PCQ_MOMENT_WS");
        const int wmode = we ? atoi(we) : 1;
        if (wmode != 0) {
            if (wmode == 2) nvb = 2;
            if (mom_ws_smem_bytes(H.nt, nsubcap, nfcap, nvb) > p->smem_optin) nvb = 2;
            smem_ws = mom_ws_smem_bytes(H.nt, nsubcap, nfcap, nvb);
            use_ws = smem_ws <= p->smem_optin;
        }
    }

    for (int64_t r0 = 0; r0 < n; r0 += rows) {
        const int64_t nr = std::min(rows, n - r0);
        const int64_t nblocks = (nr + MOM_BS - 1) / MOM_BS;
        MomTabParams TP;
        memset(&TP, 0, sizeof(TP));
        TP.x = x + r0 * ldx; TP.ldx = ldx; TP.y = y ? y + r0 * ldy : nullptr; TP.ldy = ldy; TP.n = nr;
        TP.s = H.s; TP.m = m; TP.B0 = H.B0; TP.nslots = H.nslots; TP.zero_slot = H.zero_slot; TP.nt = H.nt;
        TP.rec_a = D->rec_a; TP.rec_b = D->rec_b; TP.p1 = D->p1; TP.tab = *zbuf;
        pcq_mom_table_kernel<<<(int)std::min<int64_t>(nblocks, (int64_t)p->num_sms * 16), 256, 0, st>>>(TP);
        PCQ_LAUNCH_CHECK("pcq_mom_table_kernel");

        PCQ_CUDA(cudaMemsetAsync(D->counter, 0, 4 * MOM_NCLS, st));
        MomParams P;
        memset(&P, 0, sizeof(P));
        P.tab = *zbuf; P.nblocks = nblocks; P.nt = H.nt;
        P.units = D->units; P.tasks = D->tasks; P.funcdesc = D->funcdesc; P.subdesc = D->subdesc;
        P.nfcap = nfcap; P.nsubcap = nsubcap; P.ws = *ws;
        P.nvb = nvb;
        if (getenv("PCQ_MOMENT_DBG") && !dbg_buf) { cudaMallocManaged(&dbg_buf, 8 * MOM_NCLS * sizeof(unsigned long long)); memset(dbg_buf, 0, 8 * MOM_NCLS * sizeof(unsigned long long)); }
        PCQ_CUDA(cudaEventRecord(D->fork, st));
        for (int c = 0; c < MOM_NCLS; ++c) {
            const int nu = cls_unit0[c + 1] - cls_unit0[c];
            if (nu == 0) continue;
            cudaStream_t sc = D->cs[c];
            PCQ_CUDA(cudaStreamWaitEvent(sc, D->fork, 0));
            P.dbg = getenv("PCQ_MOMENT_DBG") ? dbg_buf + 8 * c : nullptr;
            P.unit0 = cls_unit0[c]; P.nunits = nu; P.nsplit = cls_split[c]; P.item0 = cls_item0[c]; P.counter = D->counter + c;
            const int grid = std::min(p->num_sms, nu * P.nsplit);
            auto launch = [&](auto kern, auto kern_ws) -> int {
                if (use_ws) {
                    PCQ_CUDA(cudaFuncSetAttribute(kern_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
                    kern_ws<<<grid, MOM_WS_THREADS, smem_ws, sc>>>(P);
                    return PCQ_OK;
                }
                PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<grid, 256, smem, sc>>>(P);
                return PCQ_OK;
            };
            switch (c) {
                case 0: rc = launch(pcq_mom_kernel<4, 8>, pcq_mom_ws_kernel<4, 8>); break;
                case 1: rc = launch(pcq_mom_kernel<3, 8>, pcq_mom_ws_kernel<3, 8>); break;
                case 2: rc = launch(pcq_mom_kernel<2, 16>, pcq_mom_ws_kernel<2, 16>); break;
                case 3: rc = launch(pcq_mom_kernel<1, 16>, pcq_mom_ws_kernel<1, 16>); break;
                default: rc = launch(pcq_mom_kernel<1, 2>, pcq_mom_ws_kernel<1, 2>); break;
            }
            if (rc != PCQ_OK) return rc;
            PCQ_LAUNCH_CHECK("pcq_mom_kernel");
            pcq_mom_fixup_kernel<<<nu * MOM_TASKS, 256, 0, sc>>>(D->units, D->touts, *ws, cls_unit0[c], cls_item0[c], cls_split[c],
                                                                  8 * MOM_CLS[c][1], r0 > 0 ? 1 : 0, D->mom);
            PCQ_LAUNCH_CHECK("pcq_mom_fixup_kernel");
            PCQ_CUDA(cudaEventRecord(D->join[c], sc));
            PCQ_CUDA(cudaStreamWaitEvent(st, D->join[c], 0));
        }
    }
    if (getenv("PCQ_MOMENT_DBG")) {
        cudaStreamSynchronize(st);
        for (int c = 0; c < MOM_NCLS && dbg_buf; ++c) {
            unsigned long long* d = dbg_buf + 8 * c;
            if (d[2] == 0) continue;
            fprintf(stderr, "moment dbg class (%d,%2d): blocks %llu  consumer cycles/block %.0f (waiting for V %.0f)  producer cycles/block %.0f "
                            "(waiting for a free V buffer %.0f, for the table %.0f)\n", MOM_CLS[c][0], MOM_CLS[c][1], d[2],
                    (double)d[0] / d[2], (double)d[1] / d[2], (double)d[3] / d[2], (double)d[4] / d[2], (double)d[5] / d[2]);
            for (int q = 0; q < 8; ++q) d[q] = 0;
        }
    }
    if (m > 0) {
        const int nblk = (int)std::min<int64_t>(1024, (n + 2047) / 2048);
        pcq_mom_yy_kernel<<<nblk, 256, 0, st>>>(y, ldy, n, m, D->yypart);
        PCQ_LAUNCH_CHECK("pcq_mom_yy_kernel");
        pcq_mom_yy_final_kernel<<<1, 64, 0, st>>>(D->yypart, nblk, D->yy);
        PCQ_LAUNCH_CHECK("pcq_mom_yy_final_kernel");
    }
    MomRecon R;
    memset(&R, 0, sizeof(R));
    R.g = H.geom();
    R.g.nodes = D->nodes; R.g.base = D->base; R.g.wsize = D->wsize; R.g.bandoff = D->bandoff; R.g.ncols = D->ncols;
    R.g.cnt = D->cnt; R.g.cum = D->cum;
    R.lin = D->lin; R.mom = D->mom; R.yy = D->yy; R.tdim = D->tdim; R.tord = D->tord; R.tnnz = D->tnnz;
    R.K = H.nterms; R.M = m; R.ML = layout_m >= 0 ? layout_m : m; R.W = H.width; R.nsamples = (double)n; R.S = s_out; R.accumulate = accumulate;
    const int nb16 = (H.nterms + 15) / 16;
    pcq_mom_pairs_kernel<<<dim3(nb16, nb16), 256, 0, st>>>(R);
    PCQ_LAUNCH_CHECK("pcq_mom_pairs_kernel");
    pcq_mom_edges_kernel<<<(H.nterms + 1 + 255) / 256, 256, 0, st>>>(R);
    PCQ_LAUNCH_CHECK("pcq_mom_edges_kernel");
    return PCQ_OK;
}
