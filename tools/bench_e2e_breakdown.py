"""Where the end-to-end time of lsq.fit(x, y) goes at the headline shape (host arrays): the statistics pass alone, with
resident samples, and the whole fit."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytuq_b200 import parallel, device  # noqa: E402
from pytuq_b200.lreg import lsq, PCBasisEvaluator  # noqa: E402
from pytuq_b200.rv.pcrv import PCRV  # noqa: E402
from pytuq_b200.utils.mindex import get_mi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
device.set_devices([0])
mi = get_mi(3, 20)
K = mi.shape[0]
pc = PCRV(1, 20, "HG", mi=mi)
rng = np.random.default_rng(0)
x = rng.standard_normal((N, 20))
y = rng.standard_normal(N)
xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda().reshape(-1, 1)


def best(f, reps=4):
    b = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        b = min(b, time.perf_counter() - t0)
    return b * 1e3


def fit():
    f = lsq()
    f.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
    f.fit(x, y)


print("statistics, resident samples   %8.2f ms" % best(lambda: parallel.suffstats_pc(pc, xt, yt)))
print("statistics, host samples       %8.2f ms" % best(lambda: parallel.suffstats_pc(pc, x, y)))
print("lsq.fit, host samples          %8.2f ms" % best(fit))
for frac in (2, 4):
    n = N // frac
    print("statistics, host, N/%d          %8.2f ms" % (frac, best(lambda: parallel.suffstats_pc(pc, x[:n], y[:n]))))
    print("statistics, resident, N/%d      %8.2f ms" % (frac, best(lambda: parallel.suffstats_pc(pc, xt[:n], yt[:n]))))
