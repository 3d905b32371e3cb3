"""Prototype of the fast-mode specialised evalPC in Horner form over the multi-index trie (one DFMA per term).
Generates CUDA source for (family, d, p); used to inspect SASS / registers / compile time locally.  The product
generator is csrc/pcq_spec.cu:gen_source_horner."""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pytuq_b200.utils.mindex import get_mi
from pytuq_b200.rv.pcrv import PC1d


class Node:
    __slots__ = ("children", "coefs", "key")
    def __init__(self, key=None):
        self.children = {}
        self.coefs = []
        self.key = key


def factors(row):
    return tuple((i, int(o)) for i, o in enumerate(row) if o > 0)


def gen_chunk(fam, d, terms, slots, name, sync_every, threads, first, spt=1):
    """terms: list of (factor tuple); slots: their positions in the chunk's constant array"""
    pc = PC1d(fam)
    root = Node()
    for f, sl in zip(terms, slots):
        nd = root
        for key in f:
            nd = nd.children.setdefault(key, Node(key))
        nd.coefs.append(sl)
    top = {}
    for f in terms:
        for (i, o) in f:
            top[i] = max(top.get(i, 0), o)
    L = []
    L.append('__constant__ double %s_cf[%d];' % (name, len(terms)))
    L.append('extern "C" __global__ void __launch_bounds__(%d) %s(const double* __restrict__ x, long long n, long long ldx, double* __restrict__ y, long long ldy) {' % (threads, name))
    per = threads * spt
    L.append('  const long long nb = (n + %d - 1) / %d;' % (per, per))
    L.append('  for (long long blk = blockIdx.x; blk < nb; blk += gridDim.x) {')
    S = 'abcd'[:spt]
    for q, sx in enumerate(S):
        L.append('    const long long i%s = blk * %d + %d + threadIdx.x; const bool live%s = i%s < n;' % (sx, per, q * threads, sx, sx))
        L.append('    const double* xr%s = x + (live%s ? i%s : 0) * ldx;' % (sx, sx, sx))
    for i in sorted(top):
      for sx in S:
        L.append('    const double x%d%s = xr%s[%d];' % (i, sx, sx, i))
        L.append('    const double t%d_1%s = x%d%s * %s;' % (i, sx, i, sx, float(pc.p1_scale).hex()))
        for nn in range(1, top[i]):
            a, b = float(pc.a(nn)), float(pc.b(nn))
            pm = '1.0' if nn == 1 else 't%d_%d%s' % (i, nn - 1, sx)
            L.append('    const double t%d_%d%s = fma(%s * x%d%s, t%d_%d%s, %s * %s);' % (i, nn + 1, sx, a.hex(), i, sx, i, nn, sx, b.hex(), pm))
    state = dict(count=0, uid=0)

    def cexpr(nd):
        return ' + '.join('%s_cf[%d]' % (name, s) for s in nd.coefs)

    def emit(nd):
        """emits code computing value(nd) into a fresh variable; returns its name"""
        state['uid'] += 1
        var = 'v%d' % state['uid']
        started = False
        if nd.coefs:
            for sx in S:
                L.append('    double %s%s = %s;' % (var, sx, cexpr(nd)))
            started = True
            state['count'] += len(nd.coefs)
        for key, ch in nd.children.items():
            t = 't%d_%d' % key
            if not ch.children:
                operand = cexpr(ch) if len(ch.coefs) == 1 else '(%s)' % cexpr(ch)
                state['count'] += len(ch.coefs)
            else:
                operand = emit(ch)
            leafop = not ch.children
            for sx in S:
                opnd = operand if leafop else operand + sx
                if started:
                    L.append('    %s%s = fma(%s%s, %s, %s%s);' % (var, sx, t, sx, opnd, var, sx))
                else:
                    L.append('    double %s%s = %s%s * %s;' % (var, sx, t, sx, opnd))
            started = True
            if sync_every and state['count'] >= sync_every:
                L.append('    __syncthreads();')
                state['count'] = 0
        return var

    rv = emit(root)
    for sx in S:
        if first:
            L.append('    if (live%s) y[i%s * ldy] = %s%s;' % (sx, sx, rv, sx))
        else:
            L.append('    if (live%s) y[i%s * ldy] += %s%s;' % (sx, sx, rv, sx))
    L.append('  }')
    L.append('}')
    return '\n'.join(L)


def gen(fam, d, p, chunk=512, sync_every=64, threads=128, prefix='hk', spt=1):
    mi = get_mi(p, d)
    fs = [factors(r) for r in mi]
    order = sorted(range(len(fs)), key=lambda k: fs[k])
    srcs = []
    for c, j0 in enumerate(range(0, len(order), chunk)):
        ids = order[j0:j0 + chunk]
        srcs.append(gen_chunk(fam, d, [fs[k] for k in ids], list(range(len(ids))), '%s_%d' % (prefix, c), sync_every, threads, c == 0, spt))
    return srcs, order


if __name__ == '__main__':
    fam, d, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 512
    sync = int(sys.argv[5]) if len(sys.argv) > 5 else 64
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'exp')
    spt = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    srcs, order = gen(fam, d, p, chunk, sync, spt=spt)
    for c, s in enumerate(srcs):
        fn = os.path.join(out, 'hk_%d.cu' % c)
        open(fn, 'w').write(s)
        t0 = time.time()
        r = subprocess.run(['nvcc', '-cubin', '-arch=sm_100a', '-O3', '-Xptxas', '-v', fn, '-o', fn[:-3] + '.cubin'], capture_output=True, text=True)
        dt = time.time() - t0
        info = [l for l in r.stderr.splitlines() if 'registers' in l or 'spill' in l]
        sass = subprocess.run(['cuobjdump', '-sass', fn[:-3] + '.cubin'], capture_output=True, text=True).stdout
        ops = {}
        for l in sass.splitlines():
            l = l.strip()
            if l.startswith('/*') and ';' in l:
                parts = l.split()
                # "/*0000*/ OPCODE ..." possibly with predicate
                op = parts[1] if not parts[1].startswith('@') else parts[2]
                op = op.split('.')[0]
                ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:8]
        print('chunk %d: %.1fs' % (c, dt), ' '.join(info)[-120:], top)
