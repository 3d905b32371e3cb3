// K2h: generic (no run-time compilation) fast-mode evalPC in Horner form over the multi-index trie.
//
// y = sum_k c_k prod_i phi_{alpha_ki}(x_i)  (PCRV.evalPC, rv/pcrv.py:469-498).  The terms' non-zero (dimension, order)
// factors, ordered by dimension, form a trie; value(node) = c_node + sum_children phi_child * value(child), so the
// whole expansion costs ONE fused multiply-add per term.  The NVRTC-specialised kernels (csrc/pcq_spec.cu) unroll
// that trie into straight-line code; this kernel INTERPRETS it, for plans that are not (yet) specialised: the trie is
// flattened once per plan into a record stream
//     ENTER(c_node)   push the accumulator, start the node's value at its own coefficient
//     RUN(n)          the next n records are leaf children: acc = fma(T[slot], c, acc)   (unrolled x4, no decode)
//     POP(slot)       acc = fma(T[slot], acc, popped parent accumulator)
// (16 bytes per record: coefficient + slot / opcode word; control flow is warp-uniform), a tiny kernel fills in the
// caller's coefficients per call, and the evaluation kernel -- one thread per sample, 1-D tables [slot][lane] in
// shared memory as in the other K2 kernels -- walks the stream from shared-memory chunks.  One table gather (8 bytes
// per lane) and one broadcast record per term instead of the W gathers + W multiplies of the term-by-term kernel.
// Fast mode only: the summation order differs from the reference's (exact mode keeps pcq_evalpc2_kernel).
#include "pcq_internal.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct __align__(16) HRec { double c; uint32_t w; uint32_t pad; };
constexpr uint32_t H_RUN = 0u, H_ENTER = 1u, H_POP = 2u;
constexpr int H_CHUNK = 512;            // records staged per shared-memory chunk (8 KB)
constexpr int H_MAXDEPTH = 8;

struct HNode { std::vector<std::pair<std::pair<int, int>, int>> ch; std::vector<int> terms; };

}  // namespace

struct pcq_horner {
    int nrec = 0;
    uint32_t* d_w = nullptr;          // nrec opcode / slot words
    int32_t* d_term = nullptr;        // nrec: term index whose coefficient the record carries, -1 none, -2 - j: also add term j (duplicates)
    int32_t* d_dup = nullptr;         // pairs (record, extra term) for duplicated multi-indices
    int ndup = 0;
    HRec* d_rec = nullptr;            // [mcap][nrec] filled per call
    int mcap = 0;
};

namespace {

void emit_node(const std::vector<HNode>& nodes, int nd, int depth, int nslots_per_order, int sdim,
               std::vector<uint32_t>& w, std::vector<int32_t>& term, std::vector<std::pair<int, int>>& dups, int& maxdepth) {
    maxdepth = std::max(maxdepth, depth);
    const HNode& N = nodes[nd];
    auto carry = [&](const std::vector<int>& ts, int rec) {
        term[rec] = ts.empty() ? -1 : ts[0];
        for (size_t q = 1; q < ts.size(); ++q) dups.push_back({rec, ts[q]});
    };
    // ENTER: own coefficient (0 for a prefix that is not a term itself)
    w.push_back(H_ENTER << 30); term.push_back(-1);
    carry(N.terms, (int)w.size() - 1);
    // leaf children as runs that never straddle a staging chunk
    std::vector<int> leaves, inner;
    for (size_t c = 0; c < N.ch.size(); ++c) (nodes[N.ch[c].second].ch.empty() ? leaves : inner).push_back((int)c);
    size_t done = 0;
    while (done < leaves.size()) {
        int room = H_CHUNK - (int)(w.size() % H_CHUNK) - 1;         // records left in this chunk after the RUN header
        if (room <= 0) {                                             // header would be the chunk's last record: pad
            while (w.size() % H_CHUNK) { w.push_back(H_RUN << 30); term.push_back(-1); }
            room = H_CHUNK - 1;
        }
        const int n = (int)std::min<size_t>(leaves.size() - done, (size_t)room);
        w.push_back((H_RUN << 30) | (uint32_t)n); term.push_back(-1);
        for (int q = 0; q < n; ++q) {
            const auto& c = N.ch[leaves[done + q]];
            const uint32_t slot = (uint32_t)(1 + (c.first.second - 1) * sdim + c.first.first);
            w.push_back(slot); term.push_back(-1);
            carry(nodes[c.second].terms, (int)w.size() - 1);
        }
        done += n;
    }
    for (int ci : inner) {
        const auto& c = N.ch[ci];
        emit_node(nodes, c.second, depth + 1, nslots_per_order, sdim, w, term, dups, maxdepth);
        const uint32_t slot = (uint32_t)(1 + (c.first.second - 1) * sdim + c.first.first);
        w.push_back((H_POP << 30) | slot); term.push_back(-1);
    }
}

__global__ void __launch_bounds__(256) pcq_horner_fill_kernel(const double* __restrict__ cf, int K, int m, int nrec,
                                                              const uint32_t* __restrict__ w, const int32_t* __restrict__ term,
                                                              HRec* __restrict__ rec) {
    const int64_t total = (int64_t)m * nrec;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int mo = (int)(e / nrec), j = (int)(e - (int64_t)mo * nrec);
        const int t = term[j];
        HRec r;
        r.c = t >= 0 ? cf[(size_t)mo * K + t] : 0.0;
        r.w = w[j];
        r.pad = 0;
        rec[e] = r;
    }
}

__global__ void pcq_horner_dup_kernel(const double* __restrict__ cf, int K, int m, int nrec, const int32_t* __restrict__ dup,
                                      int ndup, HRec* __restrict__ rec) {
    // duplicated multi-indices (same term listed more than once): their coefficients add up on one record
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int mo = 0; mo < m; ++mo)
            for (int q = 0; q < ndup; ++q) rec[(size_t)mo * nrec + dup[2 * q]].c += cf[(size_t)mo * K + dup[2 * q + 1]];
}

struct HornerParams {
    const double* x; double* y; int64_t n, ldx, ldy;
    int sdim, pmax, nslots, m, nrec, sumsq;
    const double* rec_a; const double* rec_b; const double* p1;
    const HRec* rec;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) pcq_evalpc_horner_kernel(HornerParams P) {
    extern __shared__ __align__(16) unsigned char smem_h[];
    double* T = reinterpret_cast<double*>(smem_h);                                    // [nslots][THREADS]
    HRec* chunk = reinterpret_cast<HRec*>(smem_h + (size_t)P.nslots * THREADS * sizeof(double));
    const int tid = threadIdx.x;
    const int64_t nblocks = (P.n + THREADS - 1) / THREADS;
    double* Tl = T + tid;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t i = blk * THREADS + tid;
        const bool live = i < P.n;
        const double* xr = P.x + (live ? i : 0) * P.ldx;
        Tl[0] = 1.0;
        for (int c = 0; c < P.sdim; ++c) pcq_fill_1d(xr[c], c, P.sdim, P.pmax, P.rec_a, P.rec_b, P.p1, Tl, THREADS);
        double ss = 0.0;
        for (int mo = 0; mo < P.m; ++mo) {
            const HRec* pr = P.rec + (size_t)mo * P.nrec;
            double cur = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
            int d = 0;
            for (int base = 0; base < P.nrec; base += H_CHUNK) {
                const int cnt = min(H_CHUNK, P.nrec - base);
                __syncthreads();
                for (int q = tid; q < cnt; q += THREADS) chunk[q] = pr[base + q];
                __syncthreads();
                int pc = 0;
                while (pc < cnt) {
                    const HRec h = chunk[pc++];
                    const uint32_t op = h.w >> 30, pay = h.w & 0x3fffffffu;
                    if (op == H_RUN) {
                        const int e = pc + (int)pay;
                        for (; pc + 4 <= e; pc += 4) {
                            const HRec r0 = chunk[pc], r1 = chunk[pc + 1], r2 = chunk[pc + 2], r3 = chunk[pc + 3];
                            const double t0 = Tl[(size_t)r0.w * THREADS], t1 = Tl[(size_t)r1.w * THREADS];
                            const double t2 = Tl[(size_t)r2.w * THREADS], t3 = Tl[(size_t)r3.w * THREADS];
                            cur = fma(t0, r0.c, cur);
                            cur = fma(t1, r1.c, cur);
                            cur = fma(t2, r2.c, cur);
                            cur = fma(t3, r3.c, cur);
                        }
                        for (; pc < e; ++pc) {
                            const HRec r0 = chunk[pc];
                            cur = fma(Tl[(size_t)r0.w * THREADS], r0.c, cur);
                        }
                    } else if (op == H_ENTER) {
                        switch (d) {
                            case 0: s0 = cur; break; case 1: s1 = cur; break; case 2: s2 = cur; break; case 3: s3 = cur; break;
                            case 4: s4 = cur; break; case 5: s5 = cur; break; case 6: s6 = cur; break; default: s7 = cur; break;
                        }
                        ++d;
                        cur = h.c;
                    } else {      // H_POP
                        --d;
                        double par;
                        switch (d) {
                            case 0: par = s0; break; case 1: par = s1; break; case 2: par = s2; break; case 3: par = s3; break;
                            case 4: par = s4; break; case 5: par = s5; break; case 6: par = s6; break; default: par = s7; break;
                        }
                        cur = fma(Tl[(size_t)pay * THREADS], cur, par);
                    }
                }
            }
            if (P.sumsq) ss = fma(cur, cur, ss);
            else if (live) P.y[i * P.ldy + mo] = cur;
        }
        if (P.sumsq && live) P.y[i * P.ldy] = ss;
    }
}

}  // namespace

void pcq_horner_free(pcq_horner* h) {
    if (!h) return;
    cudaFree(h->d_w); cudaFree(h->d_term); cudaFree(h->d_dup); cudaFree(h->d_rec);
    delete h;
}

// builds the record stream of the plan's multi-index (host) and uploads it; null when the trie is deeper than the
// kernel's register stack (more than H_MAXDEPTH - 1 non-zero orders in one term)
static pcq_horner* horner_build(const pcq_plan* p) {
    const int K = p->nterms, s = p->sdim;
    std::vector<std::vector<std::pair<int, int>>> fs(K);
    for (int k = 0; k < K; ++k)
        for (int i = 0; i < s; ++i) {
            const int od = p->h_mi[(size_t)k * s + i];
            if (od > 0) fs[k].push_back({i, od});
        }
    std::vector<int> order(K);
    for (int k = 0; k < K; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return fs[a] < fs[b]; });
    std::vector<HNode> nodes(1);
    for (int k : order) {
        int nd = 0;
        for (const auto& key : fs[k]) {
            int found = -1;
            // children arrive in sorted order: only the last child can match
            if (!nodes[nd].ch.empty() && nodes[nd].ch.back().first == key) found = nodes[nd].ch.back().second;
            if (found < 0) {
                nodes.emplace_back();
                found = (int)nodes.size() - 1;
                nodes[nd].ch.push_back({key, found});
            }
            nd = found;
        }
        nodes[nd].terms.push_back(k);
    }
    std::vector<uint32_t> w;
    std::vector<int32_t> term;
    std::vector<std::pair<int, int>> dups;
    int maxdepth = 0;
    emit_node(nodes, 0, 0, s, s, w, term, dups, maxdepth);
    if (maxdepth + 1 > H_MAXDEPTH) return nullptr;
    pcq_horner* h = new pcq_horner();
    h->nrec = (int)w.size();
    h->ndup = (int)dups.size();
    std::vector<int32_t> dflat;
    for (auto& d : dups) { dflat.push_back(d.first); dflat.push_back(d.second); }
    bool ok = cudaMalloc((void**)&h->d_w, w.size() * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc((void**)&h->d_term, term.size() * sizeof(int32_t)) == cudaSuccess &&
              cudaMemcpy(h->d_w, w.data(), w.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(h->d_term, term.data(), term.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && !dflat.empty())
        ok = cudaMalloc((void**)&h->d_dup, dflat.size() * sizeof(int32_t)) == cudaSuccess &&
             cudaMemcpy(h->d_dup, dflat.data(), dflat.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) { cudaGetLastError(); pcq_horner_free(h); return nullptr; }
    return h;
}

static int horner_threads(const pcq_plan* p) {
    // threads per CTA: the largest of 128 / 64 / 32 whose table + record chunk fits in shared memory at least twice
    const size_t chunk_bytes = (size_t)H_CHUNK * sizeof(HRec);
    for (int t : {128, 64, 32}) {
        const size_t need = (size_t)p->nslots * t * sizeof(double) + chunk_bytes;
        if (need <= p->smem_optin && (t == 32 || 2 * (need + 1024) <= (size_t)228 * 1024)) return t;
    }
    return 0;
}

bool pcq_horner_ready(pcq_plan* p) {
    if (getenv("PCQ_K2_NO_HORNER")) return false;
    if (p->horner_state == 0) {
        p->horner = horner_threads(p) ? horner_build(p) : nullptr;
        p->horner_state = p->horner ? 1 : 2;
    }
    return p->horner_state == 1;
}

// fast-mode evalPC (and its sum-of-squares form) for up to 8 coefficient vectors through the interpreter; returns
// PCQ_ERR_UNSUPPORTED when the plan cannot take this route (the caller falls back to the term-by-term kernel)
int pcq_launch_evalpc_horner(pcq_plan* p, const double* x, int64_t n, int64_t ldx, const double* cf, int m, double* y,
                             int64_t ldy, int sumsq, cudaStream_t st) {
    if (!pcq_horner_ready(p)) return PCQ_ERR_UNSUPPORTED;
    pcq_horner* h = p->horner;
    const size_t chunk_bytes = (size_t)H_CHUNK * sizeof(HRec);
    const int threads = horner_threads(p);
    if (h->mcap < m) {
        cudaFree(h->d_rec);
        h->d_rec = nullptr; h->mcap = 0;
        const int cap = std::max(m, 8);
        if (cudaMalloc((void**)&h->d_rec, (size_t)cap * h->nrec * sizeof(HRec)) != cudaSuccess) {
            cudaGetLastError();
            pcq_set_error("evalPC (Horner): record buffer allocation failed");
            return PCQ_ERR_ALLOC;
        }
        h->mcap = cap;
    }
    const int64_t tot = (int64_t)m * h->nrec;
    pcq_horner_fill_kernel<<<(int)std::min<int64_t>((tot + 255) / 256, p->num_sms * 8), 256, 0, st>>>(cf, p->nterms, m, h->nrec, h->d_w,
                                                                                                      h->d_term, h->d_rec);
    PCQ_LAUNCH_CHECK("pcq_horner_fill_kernel");
    if (h->ndup) {
        pcq_horner_dup_kernel<<<1, 32, 0, st>>>(cf, p->nterms, m, h->nrec, h->d_dup, h->ndup, h->d_rec);
        PCQ_LAUNCH_CHECK("pcq_horner_dup_kernel");
    }
    HornerParams P;
    P.x = x; P.y = y; P.n = n; P.ldx = ldx; P.ldy = ldy;
    P.sdim = p->sdim; P.pmax = p->pmax; P.nslots = p->nslots; P.m = m; P.nrec = h->nrec; P.sumsq = sumsq;
    P.rec_a = p->d_rec_a; P.rec_b = p->d_rec_b; P.p1 = p->d_p1; P.rec = h->d_rec;
    const size_t smem = (size_t)p->nslots * threads * sizeof(double) + chunk_bytes;
    const int64_t nblocks = (n + threads - 1) / threads;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, ((size_t)228 * 1024) / (smem + 1024)));
    const int grid = (int)std::min<int64_t>(nblocks, (int64_t)p->num_sms * per_sm);
    auto launch = [&](auto kern) -> int {
        PCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, st>>>(P);
        return PCQ_OK;
    };
    int rc = threads == 128 ? launch(pcq_evalpc_horner_kernel<128>) : threads == 64 ? launch(pcq_evalpc_horner_kernel<64>)
                                                                                  : launch(pcq_evalpc_horner_kernel<32>);
    if (rc != PCQ_OK) return rc;
    PCQ_LAUNCH_CHECK("pcq_evalpc_horner_kernel");
    return PCQ_OK;
}
