#!/usr/bin/env python
"""ONE plain Python process, every visible GPU: the headline workload (C2: Hermite d=20 p=3 K=1771, N=10M host numpy
samples) end to end through the unchanged API -- lsq.fit(x, y), PCRV.evalPC(x), lreg.predict(x, msc=1),
PCRV.propagate -- which the C-ABI spreads over the devices (pcq_group, include/pcq.h).  One JSON line; compare `fit`
with bench.py's e2e under torchrun at the same GPU count.

    python tools/bench_group.py [--samples N] [--devices all|1|0,1,..]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pytuq_b200 import _native, device as pdev  # noqa: E402
from pytuq_b200.rv.pcrv import PCRV  # noqa: E402
from pytuq_b200.utils.mindex import get_mi  # noqa: E402
from pytuq_b200.lreg import lsq, anl, PCBasisEvaluator  # noqa: E402


def timed(fn, reps=3):
    ts, out = [], None
    for _ in range(reps):
        for d in pdev.devices():
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return float(np.min(ts)), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=10_000_000)
    ap.add_argument("--devices", default=None)
    args = ap.parse_args()
    if args.devices:
        os.environ["PCQ_DEVICES"] = args.devices
    D, P = 20, 3
    mi = get_mi(P, D)
    K = mi.shape[0]
    n = args.samples
    rng = np.random.default_rng(2)
    x = rng.standard_normal((n, D))
    c_true = rng.standard_normal(K) / (1.0 + mi.sum(axis=1)) ** 2
    pc = PCRV(1, D, "HG", mi=mi, cfs=c_true)
    devs = pdev.devices()
    lib = _native.lib()
    l0 = lib.pcq_launch_count()
    pc.evalPC(x[:100000])                                     # warm-up: plans, scratch buffers, pinned staging
    t_eval, y = timed(lambda: pc.evalPC(x))
    y = y[:, 0] + 1e-2 * rng.standard_normal(n)

    def fit():
        f = lsq()
        f.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        f.fit(x, y)
        return f
    fit()
    t_fit, f = timed(fit)
    err = float(np.max(np.abs(f.cf - c_true)))

    def fit_anl():
        a = anl()
        a.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        a.fit(x, y)
        return a
    t_anl, a = timed(fit_anl, reps=2)
    npred = min(n, 2_000_000)
    a.predict(x[:100000], msc=1)
    t_pred, (mean, var, _) = timed(lambda: a.predict(x[:npred], msc=1), reps=2)
    pc.specialize(exact=False, fast=True)
    t_prop, mom = timed(lambda: pc.propagate(1_000_000_000 // 4, seed=4, specialize=False), reps=2)
    print(json.dumps({
        "tool": "bench_group", "process": "one", "devices": devs, "n_gpus": len(devs),
        "workload": f"C2 Hermite d=20 p=3 K={K}, N={n} host numpy samples",
        "evalPC_s": t_eval, "evalPC_samples_per_s": n / t_eval,
        "lsq_fit_s": t_fit, "fit_basis_evals_per_s": n * K / t_fit, "max_abs_coef_error_vs_truth": err,
        "anl_fit_s": t_anl, "predict_mean_var_rows": npred, "predict_mean_var_s": t_pred,
        "propagate_2.5e8_s": t_prop, "propagate_samples_per_s": 2.5e8 / t_prop,
        "propagate_mean_err_sigmas": float(abs(mom["mean"][0] - pc.computeMean()[0]) / np.sqrt(pc.computeVar()[0] / 2.5e8)),
        "gpu_launches": int(lib.pcq_launch_count() - l0)}), flush=True)


if __name__ == "__main__":
    main()
