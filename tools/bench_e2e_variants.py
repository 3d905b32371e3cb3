"""End-to-end lsq.fit(x, y) from host numpy arrays at the headline shape: best of a few runs for the host-copy variants
of parallel.suffstats_pc (PCQ_PY_HOST_COPY, parallel._GROUP_STAGES / _STAGE_BYTES)."""
import os
import sys
import time

import numpy as np

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
for copy in ("threads", "numpy"):
    for stage_mb, group in ((64, 8), (64, 4), (128, 4), (64, 1)):
        os.environ["PCQ_PY_HOST_COPY"] = copy
        parallel._STAGE_BYTES = stage_mb << 20
        parallel._GROUP_STAGES = group
        best = 1e9
        for _ in range(4):
            fit = lsq()
            fit.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
            t0 = time.perf_counter()
            fit.fit(x, y)
            best = min(best, time.perf_counter() - t0)
        print(f"copy={copy:8s} stage={stage_mb:4d} MB group<={group}: {best * 1e3:8.2f} ms  {N * K / best:.4g} basis evals/s", flush=True)
