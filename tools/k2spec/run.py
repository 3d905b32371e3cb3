import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pytuq_b200 import _native
from pytuq_b200.rv.pcrv import PCRV
from pytuq_b200.utils.mindex import get_mi
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libk2spec.so"))
lib.k2spec_run.argtypes = [C.c_int, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p]
pl = _native.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for which, (p, n) in enumerate([(3, 1_000_000), (4, 400_000)]):
    d = 20
    mi = get_mi(p, d); K = mi.shape[0]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)
    c = torch.randn(K, dtype=torch.float64, device="cuda", generator=g)
    y0 = torch.empty((n, 1), dtype=torch.float64, device="cuda"); y1 = torch.empty_like(y0)
    pc = PCRV(1, d, "HG", mi=mi); plan = pc._plan(mi)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(pl.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, d, c.data_ptr(), 1, y0.data_ptr(), 1, 0, st))
    for grid in (148, 296, 444, 592, 1184):
        ts = []
        for it in range(5):
            e0.record(); rc = lib.k2spec_run(which, x.data_ptr(), n, d, c.data_ptr(), y1.data_ptr(), 1, grid, st); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ok = torch.equal(y0, y1)
        ms = float(np.median(ts[1:]))
        print(f"p={p} K={K} grid={grid}: {ms:8.3f} ms  {n/ms/1e3:8.1f} Msamples/s  bit-exact={ok} rc={rc}", flush=True)
