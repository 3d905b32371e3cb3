"""Feasibility experiment: evalPC specialised per multi-index (fully unrolled term list, register-resident
1-D tables).  Generates CUDA source for (family, d, p), exact mode, one output."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pytuq_b200.utils.mindex import get_mi
from pytuq_b200.rv.pcrv import PC1d

def gen(fam, d, p, name, sync_every=128, threads=128):
    mi = get_mi(p, d)
    pc = PC1d(fam)
    L = []
    L.append('extern "C" __global__ void __launch_bounds__(%d) %s(const double* __restrict__ x, long n, long ldx, const double* __restrict__ cf, double* __restrict__ y, long ldy) {' % (threads, name))
    L.append('  const long nb = (n + %d - 1) / %d;' % (threads, threads))
    L.append('  for (long blk = blockIdx.x; blk < nb; blk += gridDim.x) {')
    L.append('    const long i = blk * %d + threadIdx.x; const bool live = i < n;' % threads)
    L.append('    const double* xr = x + (live ? i : 0) * ldx;')
    for i in range(d):
        L.append('    const double x%d = xr[%d];' % (i, i))
        L.append('    const double t%d_1 = __dmul_rn(x%d, %s);' % (i, i, repr(float(pc.p1_scale))))
        for nn in range(1, p):
            a, b = float(pc.a(nn)), float(pc.b(nn))
            pm = '1.0' if nn == 1 else 't%d_%d' % (i, nn - 1)
            L.append('    const double t%d_%d = __dadd_rn(__dmul_rn(__dmul_rn(%s, x%d), t%d_%d), __dmul_rn(%s, %s));' % (i, nn + 1, repr(a), i, i, nn, repr(b), pm))
    L.append('    double acc = 0.0;')
    for k, row in enumerate(mi):
        e = '__ldg(cf + %d)' % k
        for i, o in enumerate(row):
            if o > 0:
                e = '__dmul_rn(%s, t%d_%d)' % (e, i, o)
        L.append('    acc = __dadd_rn(acc, %s);' % e)
        if sync_every and (k + 1) % sync_every == 0:
            L.append('    __syncthreads();')
    L.append('    if (live) y[i * ldy] = acc;')
    L.append('  }')
    L.append('}')
    return '\n'.join(L), mi.shape[0]

if __name__ == '__main__':
    src = ['#include <cuda_runtime.h>']
    meta = []
    for (fam, d, p, nm) in [('HG', 20, 3, 'k2spec_c2'), ('HG', 20, 4, 'k2spec_c4')]:
        s, K = gen(fam, d, p, nm)
        src.append(s); meta.append((nm, d, K))
    src.append('''
extern "C" int k2spec_run(int which, const double* x, long n, long ldx, const double* cf, double* y, long ldy, int grid, void* st) {
  if (which == 0) k2spec_c2<<<grid, 128, 0, (cudaStream_t)st>>>(x, n, ldx, cf, y, ldy);
  else k2spec_c4<<<grid, 128, 0, (cudaStream_t)st>>>(x, n, ldx, cf, y, ldy);
  return (int)cudaGetLastError();
}''')
    open('k2spec.cu', 'w').write('\n'.join(src))
    print(meta)
