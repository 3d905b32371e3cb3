#!/usr/bin/env python
"""The five BASELINE.json configurations on one GPU through the public pytuq_b200 API (host numpy in /
numpy out unless noted), one JSON line each.  Sizes follow BASELINE.json; where the full size needs
more than one GPU's worth of host RAM or minutes of wall time the per-GPU shard of the 8-GPU run is
used and said so in `size_note`.  Not the headline bench (bench.py is); this documents coverage.

    python tools/bench_configs.py [C1 C2 C3 C4 C5]
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pytuq_b200 import _native  # noqa: E402
from pytuq_b200.rv.pcrv import PCRV  # noqa: E402
from pytuq_b200.utils.mindex import get_mi  # noqa: E402
from pytuq_b200.lreg import lsq, anl, bcs, PCBasisEvaluator  # noqa: E402
from pytuq_b200.workflows.fits import pc_fit  # noqa: E402
from pytuq_b200.lreg.stats import SuffStats  # noqa: E402


def timed(fn, reps=2):
    best, out = 1e30, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c1():
    rng = np.random.default_rng(1)
    n, d, p = 100_000, 10, 4
    mi = get_mi(p, d)
    x = rng.uniform(-1, 1, (n, d))
    w = 1.0 / (1 + np.arange(d)) ** 2
    y = np.cos(2 * np.pi * 0.25 + ((x + 1) / 2) @ w) + 1e-2 * rng.standard_normal(n)     # Genz oscillatory
    pc = PCRV(1, d, "LU", mi=mi)
    t_basis, A = timed(lambda: pc.evalBases(x, 0))
    out = {}
    for name, make in (("lsq", lsq), ("anl", anl)):
        def fit():
            f = make()
            f.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
            f.fit(x, y)
            return f
        out[name + "_fit_s"], f = timed(fit)
        out[name + "_rms_resid"] = float(np.sqrt(np.mean((A @ f.cf - y) ** 2)))
    emit(config="C1", desc="LU d=10 p=4 K=1001, N=1e5, Genz oscillatory + noise", evalBases_s=t_basis,
         evalBases_evals_per_s=n * mi.shape[0] / t_basis, **out)


def c2():
    rng = np.random.default_rng(2)
    n, d, p = 1_000_000, 20, 3
    mi = get_mi(p, d)
    K = mi.shape[0]
    x = rng.standard_normal((n, d))
    c_true = rng.standard_normal(K) / (1.0 + mi.sum(axis=1)) ** 2
    pc = PCRV(1, d, "HG", mi=mi, cfs=c_true)
    t_eval, y = timed(lambda: pc.evalPC(x))
    y = y[:, 0] + 1e-2 * rng.standard_normal(n)

    def fit():
        f = lsq()
        f.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        f.fit(x, y)
        return f
    t_fit, f = timed(fit)
    emit(config="C2", desc="HG d=20 p=3 K=1771, N=1e6, basis eval + lstsq fit (fused, host numpy in)",
         evalPC_s=t_eval, fit_s=t_fit, basis_evals_per_s=n * K / t_fit,
         max_abs_coef_err=float(np.max(np.abs(f.cf - c_true))))

    # Bayesian fit with full coefficient covariance + prediction with pushed-forward variance on all points
    # (lreg.predict msc=1: the second N K^2 pass of the reference workflow; K6 here)
    def fit_anl():
        a = anl()
        a.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        a.fit(x, y)
        return a
    t_anl, a = timed(fit_anl)
    t_pred, (mean, var, _) = timed(lambda: a.predict(x, msc=1), reps=2)
    emit(config="C2", desc="anl (full covariance) fit + predict with variance on the 1e6 training points",
         anl_fit_s=t_anl, predict_mean_var_s=t_pred, rms_resid=float(np.sqrt(np.mean((mean - y) ** 2))),
         mean_pred_std=float(np.mean(np.sqrt(var))))


def c3():
    rng = np.random.default_rng(3)
    n, d, p = 1_250_000, 50, 2
    mi = get_mi(p, d)
    K = mi.shape[0]
    x = rng.uniform(-1, 1, (n, d))
    c_true = np.zeros(K)
    sup = rng.choice(K, 40, replace=False)
    c_true[sup] = rng.standard_normal(40)
    pc = PCRV(1, d, "LU", mi=mi, cfs=c_true)
    y = pc.evalPC(x)[:, 0]
    y = y + 0.01 * np.std(y) * rng.standard_normal(n)
    t_stats, st = timed(lambda: SuffStats.from_pc(pc, x, y, 0), reps=2)

    def fit():
        f = bcs(eta=1e-8)
        f._fit_stats(st)
        return f
    t_bcs, f = timed(fit, reps=1)
    emit(config="C3", desc="BCS LU d=50 p=2 K=1326, 40-sparse truth, 1% noise",
         size_note="N=1.25e6 = the per-GPU shard of the 1e7-sample / 8-GPU configuration", stats_s=t_stats,
         bcs_iteration_s=t_bcs, n_used=int(len(f.used)), support_recovered=bool(set(sup) <= set(f.used.tolist())),
         max_abs_coef_err_on_support=float(np.max(np.abs(f.cf[np.argsort(f.used)][np.isin(np.sort(f.used), sup)] -
                                                        c_true[np.sort(f.used)][np.isin(np.sort(f.used), sup)]))))


def c4():
    d, p = 20, 4
    mi = get_mi(p, d)
    K = mi.shape[0]
    rng = np.random.default_rng(4)
    cf = rng.standard_normal(K) / np.arange(1, K + 1)
    pc = PCRV(1, d, "HG", mi=mi, cfs=cf)
    n = 125_000_000
    t_prop, out = timed(lambda: pc.propagate(n, seed=4), reps=1)
    t_sob, sens = timed(lambda: (pc.computeSens(), pc.computeTotSens(), pc.computeJointSens()), reps=1)
    t_comp, ok = timed(lambda: pc.specialize(exact=False, fast=True), reps=1)
    t_spec, outs = timed(lambda: pc.propagate(n, seed=4), reps=1) if ok else (None, None)
    emit(config="C4", desc="HG d=20 p=4 K=10626 forward propagation, germ drawn on device (Philox), fast mode",
         size_note="N=1.25e8 = the per-GPU shard of the 1e9-sample / 8-GPU configuration", propagate_s=t_prop,
         samples_per_s=n / t_prop, mc_mean=float(out["mean"][0]), pc_mean=float(pc.computeMean()[0]),
         mc_var=float(out["var"][0]), pc_var=float(pc.computeVar()[0]), sobol_main_total_joint_s=t_sob,
         specialised={"nvrtc_compile_s": t_comp, "propagate_s": t_spec,
                      "samples_per_s": (n / t_spec) if t_spec else None,
                      "mc_mean": float(outs["mean"][0]) if outs else None})


def c5():
    rng = np.random.default_rng(5)
    n, d, p, M = 100_000, 30, 3, 256
    mi = get_mi(p, d)
    K = mi.shape[0]
    x = rng.uniform(-1, 1, (n, d))
    C_true = rng.standard_normal((M, K)) / (1.0 + np.arange(M))[:, None]
    gen = PCRV(M, d, "LU", mi=mi, cfs=C_true)
    t_gen, Y = timed(lambda: gen.evalPC(x, exact=False), reps=1)
    Y = Y + 1e-3 * rng.standard_normal(Y.shape)
    with contextlib.redirect_stdout(io.StringIO()):
        t_fit, (pc, lregs) = timed(lambda: pc_fit(x, Y, order=p, pctype="LU", method="lsq"), reps=1)
    err = max(float(np.max(np.abs(pc.coefs[j] - C_true[j]))) for j in (0, 100, 255))
    emit(config="C5", desc="KL+PC shape: LU d=30 p=3 K=5456, M=256 outputs sharing one design matrix, pc_fit(lsq)",
         size_note="N=1e5 (host RAM for Y at 1e7 x 256 is 20 GB per GPU; the statistics pass is linear in N)",
         evalPC_256_outputs_s=t_gen, pc_fit_s=t_fit, basis_evals_per_s=n * K / t_fit, max_abs_coef_err=err)


if __name__ == "__main__":
    which = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]
    fns = dict(C1=c1, C2=c2, C3=c3, C4=c4, C5=c5)
    lib = _native.lib()
    for w in which:
        l0 = lib.pcq_launch_count()
        fns[w]()
        emit(config=w, gpu_launches=int(lib.pcq_launch_count() - l0))
