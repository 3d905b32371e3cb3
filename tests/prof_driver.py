"""Small driver for ncu: launches each hot kernel a few times on the C2 shape (HG, d=20, p=3)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytuq_b200 import _native
from pytuq_b200.rv.pcrv import PCRV
from pytuq_b200.utils.mindex import get_mi

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
d, p = (20, 3) if len(sys.argv) <= 3 else (int(sys.argv[3]), int(sys.argv[4]))
lib = _native.lib()
mi = get_mi(p, d)
K = mi.shape[0]
pc = PCRV(1, d, "HG", mi=mi)
plan = pc._plan(mi)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)
c = torch.randn(K, dtype=torch.float64, device="cuda", generator=g)
y = torch.empty((n, 1), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
if which == "k2s":
    _native.check(lib.pcq_plan_specialize(plan.handle, 3))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for mode in (0, 1):
        for rep in range(4):
            e0.record()
            _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, d, c.data_ptr(), 1, y.data_ptr(), 1, mode, st))
            e1.record(); e1.synchronize()
            print("k2s mode", mode, "rep", rep, "ms", e0.elapsed_time(e1))
for rep in range(3):
    if which in ("all", "k2"):
        _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, d, c.data_ptr(), 1, y.data_ptr(), 1, 0, st))
        _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, d, c.data_ptr(), 1, y.data_ptr(), 1, 1, st))
    if which in ("all", "k1"):
        A = torch.empty((n, K), dtype=torch.float64, device="cuda")
        _native.check(lib.pcq_eval_bases_dev(plan.handle, x.data_ptr(), n, d, A.data_ptr(), K, st))
        del A
    if which in ("all", "k3"):
        S = torch.empty((K + 2, K + 2), dtype=torch.float64, device="cuda")
        _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), n, d, y.data_ptr(), 1, 1, S.data_ptr(), 0, st))
if which == "k6":
    Cm = torch.randn((K, K), dtype=torch.float64, device="cuda", generator=g)
    Cm = Cm @ Cm.T / K
    v = torch.empty(n, dtype=torch.float64, device="cuda")
    for rep in range(2):
        _native.check(lib.pcq_quadform_dev(plan.handle, x.data_ptr(), n, d, Cm.data_ptr(), K, v.data_ptr(), 1, st))
torch.cuda.synchronize()
print("ok", which, n, K)
