"""Experiment driver for the fast-mode (Horner) specialised evalPC: generates source variants for one shape, compiles
them HERE with nvcc to cubins (tools/exp/var_*.cubin + manifest.json); tools/k2spec/run_variants.py times them on
the GPU box.  Variants differ in where the coefficients come from (constant bank / shared memory), samples per
thread, CTA size and register cap."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pytuq_b200.utils.mindex import get_mi
from pytuq_b200.rv.pcrv import PC1d

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "exp")


class Node:
    __slots__ = ("children", "coefs")
    def __init__(self):
        self.children = {}
        self.coefs = []


def factors(row):
    return tuple((i, int(o)) for i, o in enumerate(row) if o > 0)


def gen_chunk(fam, terms, name, first, coef="const", spt=1, threads=128, minblocks=0, sync_every=0, noloop=False, dsync=0, twoconst=False):
    pc = PC1d(fam)
    root = Node()
    for sl, f in enumerate(terms):
        nd = root
        for key in f:
            nd = nd.children.setdefault(key, Node())
        nd.coefs.append(sl)
    top = {}
    for f in terms:
        for (i, o) in f:
            top[i] = max(top.get(i, 0), o)
    nt = len(terms)
    S = "abcd"[:spt]
    L = []
    cname = "scf" if coef == "smem" else name + "_cf"
    L.append("__constant__ double %s_cf[%d];" % (name, nt))
    lb = "__launch_bounds__(%d%s)" % (threads, ", %d" % minblocks if minblocks else "")
    L.append('extern "C" __global__ void %s %s(const double* __restrict__ x, long long n, long long ldx, double* __restrict__ y, long long ldy, const double* __restrict__ gcf) {' % (lb, name))
    if coef == "smem":
        L.append("  __shared__ __align__(16) double scf[%d];" % (nt + (nt & 1)))
        L.append("  for (int q = threadIdx.x; q < %d; q += %d) scf[q] = gcf[q];" % (nt, threads))
        L.append("  __syncthreads();")
    per = threads * spt
    L.append("  const long long nb = (n + %d - 1) / %d;" % (per, per))
    L.append("  for (long long blk = blockIdx.x; blk < nb; blk += gridDim.x) {" if not noloop else "  { const long long blk = blockIdx.x;")
    for q, z in enumerate(S):
        L.append("    const long long i%s = blk * %d + %d + threadIdx.x; const bool live%s = i%s < n;" % (z, per, q * threads, z, z))
        L.append("    const double* xr%s = x + (live%s ? i%s : 0) * ldx;" % (z, z, z))
    for i in sorted(top):
        for z in S:
            L.append("    const double x%d%s = xr%s[%d];" % (i, z, z, i))
            L.append("    const double t%d_1%s = x%d%s * %s;" % (i, z, i, z, float(pc.p1_scale).hex()))
            for nn in range(1, top[i]):
                a, b = float(pc.a(nn)), float(pc.b(nn))
                pm = "1.0" if nn == 1 else "t%d_%d%s" % (i, nn - 1, z)
                L.append("    const double t%d_%d%s = fma(%s * x%d%s, t%d_%d%s, %s * %s);" % (i, nn + 1, z, a.hex(), i, z, i, nn, z, b.hex(), pm))
    uid = [0]
    cnt = [0]

    def cexpr(nd):
        e = " + ".join("%s[%d]" % (cname, s) for s in nd.coefs)
        return "(%s)" % e if len(nd.coefs) > 1 else e

    def emit(nd):
        uid[0] += 1
        var = "v%d" % uid[0]
        kids = list(nd.children.items())
        inner = [kv for kv in kids if kv[1].children]
        leaves = [kv for kv in kids if not kv[1].children]
        started = False
        for key, ch in inner + leaves:
            t = "t%d_%d" % key
            lf = not ch.children
            operand = cexpr(ch) if lf else emit(ch)
            for z in S:
                op = operand if lf else operand + z
                if started:
                    L.append("    %s%s = fma(%s%s, %s, %s%s);" % (var, z, t, z, op, var, z))
                elif (inner or twoconst) and nd.coefs:
                    L.append("    double %s%s = fma(%s%s, %s, %s);" % (var, z, t, z, op, cexpr(nd)))
                else:
                    L.append("    double %s%s = %s%s * %s;" % (var, z, t, z, op))
            started = True
            cnt[0] += 1
            if sync_every and cnt[0] >= sync_every:
                L.append("    __syncthreads();")
                cnt[0] = 0
            if dsync and cnt[0] >= dsync:
                # data-dependent barrier: cannot be hoisted above the accumulator it reads
                L.append("    if (__syncthreads_or(%sa == 1.2345e300)) y[0] = 0.0;" % var)
                cnt[0] = 0
        if nd.coefs and not ((inner or twoconst) and started):
            for z in S:
                if started:
                    L.append("    %s%s = %s%s + %s;" % (var, z, var, z, cexpr(nd)))
                else:
                    L.append("    double %s%s = %s;" % (var, z, cexpr(nd)))
            started = True
        if not started:
            for z in S:
                L.append("    double %s%s = 0.0;" % (var, z))
        return var

    rv = emit(root)
    for z in S:
        L.append("    if (live%s) y[i%s * ldy] %s %s%s;" % (z, z, "=" if first else "+=", rv, z))
    L.append("  }")
    L.append("}")
    return "\n".join(L)


def build(fam, d, p, tag, chunk=1024, **kw):
    mi = get_mi(p, d)
    fs = [factors(r) for r in mi]
    order = sorted(range(len(fs)), key=lambda k: fs[k])
    kernels = []
    for c, j0 in enumerate(range(0, len(order), chunk)):
        ids = order[j0:j0 + chunk]
        name = "%s_%d" % (tag, c)
        src = gen_chunk(fam, [fs[k] for k in ids], name, c == 0, **kw)
        fn = os.path.join(OUT, name + ".cu")
        open(fn, "w").write(src)
        t0 = time.time()
        r = subprocess.run(["nvcc", "-cubin", "-arch=sm_100a", "-O3", "-Xptxas", "-v", fn, "-o", fn[:-3] + ".cubin"], capture_output=True, text=True)
        if r.returncode:
            print(r.stderr[-2000:]); raise SystemExit(1)
        info = " ".join(l.strip() for l in r.stderr.splitlines() if "registers" in l or "spill" in l)
        print(name, "%.1fs" % (time.time() - t0), info[-150:])
        kernels.append(dict(name=name, cubin=name + ".cubin", j0=j0, n=len(ids)))
    return dict(tag=tag, fam=fam, d=d, p=p, order=[int(k) for k in order], kernels=kernels,
                threads=kw.get("threads", 128), spt=kw.get("spt", 1), coef=kw.get("coef", "const"), noloop=bool(kw.get("noloop")))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    fam, d, p = "HG", 20, 3
    variants = [
        dict(tag="n_t256", noloop=True, threads=256),
        dict(tag="n_t256_2c", noloop=True, threads=256, twoconst=True),
        dict(tag="n_t128_2c", noloop=True, threads=128, twoconst=True),
        dict(tag="n_t192", noloop=True, threads=192),
        dict(tag="n_t384", noloop=True, threads=384),
        dict(tag="n_t256_c512", noloop=True, threads=256, chunk=512),
        dict(tag="n_t256_c768", noloop=True, threads=256, chunk=768),
        dict(tag="n_t256_c896", noloop=True, threads=256, chunk=896),
        dict(tag="n_t256_c1280", noloop=True, threads=256, chunk=1280),
        dict(tag="n_t256_c1536", noloop=True, threads=256, chunk=1536),
        dict(tag="n_t256_2c_c896", noloop=True, threads=256, chunk=896, twoconst=True),
        dict(tag="n_t256_mb3", noloop=True, threads=256, minblocks=3),
        dict(tag="n_t256_mb4", noloop=True, threads=256, minblocks=4),
    ]
    man = [build(fam, d, p, **v) for v in variants]
    json.dump(man, open(os.path.join(OUT, "manifest.json"), "w"))
