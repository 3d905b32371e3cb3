"""Dumps the fast-mode (Horner) source the product generator emits for (family, d, p), compiles every chunk with nvcc
for sm_100a and summarises registers / spills / instruction mix -- a GPU-less check of csrc/pcq_spec.cu."""
import ctypes as C, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pytuq_b200 import _native
from pytuq_b200.rv.pcrv import PCRV
from pytuq_b200.utils.mindex import get_mi

fam, d, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spt = int(sys.argv[4]) if len(sys.argv) > 4 else 1
CH = 1024
mi = get_mi(p, d)
K = mi.shape[0]
pc = PCRV(1, d, fam, mi=mi)
pmax = int(mi.max())
ra = np.zeros((pmax + 1, d)); rb = np.zeros((pmax + 1, d)); p1 = np.empty(d)
for i, q in enumerate(pc.PC1ds):
    ra[:, i], rb[:, i] = q.rec_tables(pmax)
    p1[i] = q.p1_scale
mi32 = np.ascontiguousarray(mi, dtype=np.int32)
lib = _native.lib()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "exp")
os.makedirs(out, exist_ok=True)
tot = {}
for c, j0 in enumerate(range(0, K, CH)):
    j1 = min(K, j0 + CH)
    n = lib.pcq_debug_horner_source(d, K, _native.ptr(mi32), pmax, _native.ptr(ra), _native.ptr(rb), _native.ptr(p1), j0, j1, spt, None, 0)
    buf = C.create_string_buffer(int(n) + 1)
    lib.pcq_debug_horner_source(d, K, _native.ptr(mi32), pmax, _native.ptr(ra), _native.ptr(rb), _native.ptr(p1), j0, j1, spt, buf, n + 1)
    fn = os.path.join(out, f"dbg_{c}.cu")
    open(fn, "wb").write(buf.value)
    t0 = time.time()
    r = subprocess.run(["nvcc", "-cubin", "-arch=sm_100a", "-O3", "-Xptxas", "-v", fn, "-o", fn[:-3] + ".cubin"], capture_output=True, text=True)
    info = " ".join(l.strip() for l in r.stderr.splitlines() if "registers" in l or "spill" in l)
    sass = subprocess.run(["cuobjdump", "-sass", fn[:-3] + ".cubin"], capture_output=True, text=True).stdout
    ops = {}
    for l in sass.splitlines():
        l = l.strip()
        if l.startswith("/*") and ";" in l:
            parts = l.split()
            op = (parts[1] if not parts[1].startswith("@") else parts[2]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
            tot[op] = tot.get(op, 0) + 1
    print(f"chunk {c} [{j0},{j1}): {time.time() - t0:.1f}s", info[-110:], sorted(ops.items(), key=lambda kv: -kv[1])[:7])
print("total", sorted(tot.items(), key=lambda kv: -kv[1])[:8], "terms", K)
