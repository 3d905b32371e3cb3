#!/usr/bin/env python
"""BASELINE.json configs 3, 4, 5 at their STATED sizes on 1/2/4/8 GPUs (one process per GPU, NCCL), one JSON line each.

    python tools/bench_multi.py [C3 C4 C5]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_multi.py [C3 C4 C5] [--scale F]

Strong scaling: the stated sample count is the whole job, every rank owns a contiguous 1/N block.  The synthetic
samples are drawn on the device (the host is not part of these lines; bench.py carries the end-to-end number of the
headline config).  Times are device-side (cuda synchronize + barrier on both sides), max over ranks.

  C3  BCS, Legendre d=50 p=2 (K=1326), N=1e7: sharded statistics pass (K3) -> ONE all-reduce of S (14 MB) -> fast-RVM
      on the K x K statistics (replicated) -> the sharded residual pass (K2 + 2-scalar all-reduce) bcs_fit uses when
      the noise is tiny (lreg/bcs.py:58-265).
  C4  forward propagation of 1e9 Monte Carlo samples through the Hermite d=20 p=4 surrogate (K=10626): device Philox
      germ, disjoint index blocks of one logical stream per rank, no A (rv/pcrv.py:469-498,511-525); moments against
      the closed form, PC Sobol indices (rv/pcrv.py:202-338).
  C5  KL+PC shape: 256 outputs sharing one design matrix, Legendre d=30 p=3 (K=5456), N=1e7: one sharded pass for
      G = A^T A and B = A^T Y (K x 256), one all-reduce of S (261 MB), batched Cholesky solve
      (workflows/fits.py:47-64, linred/klsurr.py:117-118).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from pytuq_b200 import _native, device as pdev, parallel  # noqa: E402
from pytuq_b200.rv.pcrv import PCRV  # noqa: E402
from pytuq_b200.utils.mindex import get_mi  # noqa: E402
from pytuq_b200.lreg import bcs  # noqa: E402
from pytuq_b200.lreg.lreg import solve_normal_device  # noqa: E402

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def barrier():
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=1):
    """max over ranks of the best wall time between device-synchronised barriers"""
    best, out = 1e30, None
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    t = torch.tensor([best], dtype=torch.float64, device="cuda")
    if WORLD > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), out


def emit(**kw):
    if RANK == 0:
        print(json.dumps(dict(n_gpus=WORLD, **kw)), flush=True)


def c3(scale):
    n_total, d, p = int(1e7 * scale), 50, 2
    mi = get_mi(p, d)
    K = mi.shape[0]
    lo, hi = parallel.shard_bounds(n_total, RANK, WORLD)
    n = hi - lo
    rng = np.random.default_rng(3)
    c_true = np.zeros(K)
    sup = np.sort(rng.choice(K, 40, replace=False))
    c_true[sup] = rng.standard_normal(40)
    pc = PCRV(1, d, "LU", mi=mi, cfs=c_true)
    g = torch.Generator(device="cuda").manual_seed(300 + RANK)
    x = torch.rand((n, d), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    y = pc.evalPC(x)                                           # resident tensor in, resident tensor out
    sd = torch.tensor([float((y * y).sum()), float(y.sum()), float(n)], dtype=torch.float64, device="cuda")
    if WORLD > 1:
        dist.all_reduce(sd)
    ystd = float(torch.sqrt(sd[0] / sd[2] - (sd[1] / sd[2]) ** 2))
    y = y + 0.01 * ystd * torch.randn(y.shape, dtype=torch.float64, device="cuda", generator=g)

    with parallel.sharded_rows():
        t_stats, st = timed(lambda: parallel.suffstats_pc(pc, x, y, 0), reps=3)
    S_bytes = int(st.S_dev.numel() * 8)
    ar = torch.empty_like(st.S_dev)
    t_ar, _ = timed(lambda: dist.all_reduce(ar) if WORLD > 1 else None, reps=3)

    def fit():
        f = bcs(eta=1e-8)
        with contextlib.redirect_stdout(io.StringIO()):
            f._fit_stats(st)
        return f
    t_bcs, f = timed(fit)
    full = np.zeros(K)
    full[f.used] = f.cf
    pc._distributed_stats = WORLD > 1
    t_res, (rss, ydr) = timed(lambda: parallel.residual_sums_pc(pc, x, y, 0, 0, full), reps=2)
    rss_identity = st.rss(f.cf, f.used)
    order = np.argsort(f.used)
    used_sorted = np.asarray(f.used)[order]
    on_sup = np.isin(used_sorted, sup)
    flops = float(n_total) * K * (K + 1) + 2.0 * n_total * K + 2.0 * n_total
    emit(config="C3", desc="BCS Legendre d=50 p=2 K=1326, 40-sparse truth, 1% noise; sharded Gram + all-reduce + fast-RVM + sharded residual pass",
         samples=n_total, samples_per_rank=n, stats_pass_s=t_stats, stats_pass_TFLOPs_per_gpu=flops / WORLD / t_stats / 1e12,
         allreduce_bytes=S_bytes, allreduce_s=t_ar, bcs_on_statistics_s=t_bcs, residual_pass_s=t_res,
         total_fit_s=t_stats + t_bcs + t_res, n_used=int(len(f.used)),
         support_recovered=bool(set(sup.tolist()) <= set(np.asarray(f.used).tolist())),
         max_abs_coef_err_on_support=float(np.max(np.abs(f.cf[order][on_sup] - c_true[used_sorted[on_sup]]))),
         residual_ss_pass=rss, residual_ss_identity=rss_identity, sigma2=float(np.reshape(f.datavar, -1)[0]),
         sigma2_truth=(0.01 * ystd) ** 2)


def c4(scale):
    d, p = 20, 4
    mi = get_mi(p, d)
    K = mi.shape[0]
    rng = np.random.default_rng(4)
    cf = rng.standard_normal(K) / np.arange(1, K + 1)
    pc = PCRV(1, d, "HG", mi=mi, cfs=cf)
    n_total = int(1e9 * scale)
    t_comp, ok = timed(lambda: pc.specialize(exact=False, fast=True))
    t_prop, out = timed(lambda: pc.propagate(n_total, seed=4, specialize=False), reps=2)
    t_sob, sens = timed(lambda: (pc.computeSens(), pc.computeTotSens(), pc.computeJointSens()))
    var = float(pc.computeVar()[0])
    flops = 4.0 * d * (p - 1) + 3.0 * K
    emit(config="C4", desc="Hermite d=20 p=4 K=10626 forward propagation of 1e9 samples, Philox germ on device, fast mode, no A",
         samples=n_total, samples_per_rank=-(-n_total // WORLD), specialised=bool(ok), nvrtc_compile_or_cache_s=t_comp,
         propagate_s=t_prop, samples_per_s=n_total / t_prop, TFLOPs_algorithmic_per_gpu=n_total * flops / WORLD / t_prop / 1e12,
         mc_mean=float(out["mean"][0]), pc_mean=float(pc.computeMean()[0]), mc_var=float(out["var"][0]), pc_var=var,
         mean_err_in_sigmas=float(abs(out["mean"][0] - pc.computeMean()[0]) / np.sqrt(var / n_total)),
         var_rel_err=float(abs(out["var"][0] / var - 1.0)), sobol_main_total_joint_s=t_sob,
         sobol_main_sum=float(np.sum(sens[0])), collective="one all-reduce of 4 scalars per output")


def c5(scale):
    n_total, d, p, M = int(1e7 * scale), 30, 3, 256
    mi = get_mi(p, d)
    K = mi.shape[0]
    lo, hi = parallel.shard_bounds(n_total, RANK, WORLD)
    n = hi - lo
    rng = np.random.default_rng(5)
    C_true = rng.standard_normal((M, K)) / (1.0 + np.arange(M))[:, None]
    gen = PCRV(M, d, "LU", mi=mi, cfs=C_true)
    pc = PCRV(1, d, "LU", mi=mi)
    g = torch.Generator(device="cuda").manual_seed(500 + RANK)
    x = torch.rand((n, d), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    t_gen, Y = timed(lambda: gen.evalPC(x, exact=False))       # 256 outputs at once on the tensor pipe (K2m)
    Y += 1e-3 * torch.randn(Y.shape, dtype=torch.float64, device="cuda", generator=g)

    with parallel.sharded_rows():
        t_stats, st = timed(lambda: parallel.suffstats_pc(pc, x, Y, 0), reps=2)
    S_bytes = int(st.S_dev.numel() * 8)
    ar = torch.empty_like(st.S_dev)
    t_ar, _ = timed(lambda: dist.all_reduce(ar) if WORLD > 1 else None, reps=3)
    del ar

    def solve():
        G_t, B_t = st.dev_system()
        return solve_normal_device(G_t, B_t, 1.0e-3)
    t_solve, C = timed(solve, reps=2)      # the first call pays the cuSOLVER handle + workspace
    Cn = C.cpu().numpy()
    err = float(np.max(np.abs(Cn.T - C_true)))
    flops = float(n_total) * K * (K + 1) + 2.0 * n_total * K * M + 2.0 * n_total * M
    emit(config="C5", desc="KL+PC shape: Legendre d=30 p=3 K=5456, 256 outputs sharing one design matrix, batched A^T Y + shared Gram",
         samples=n_total, samples_per_rank=n, outputs=M, evalPC_256_outputs_s=t_gen,
         evalPC_TFLOPs_per_gpu=2.0 * n_total * K * M / WORLD / t_gen / 1e12, stats_pass_s=t_stats,
         stats_pass_TFLOPs_per_gpu=flops / WORLD / t_stats / 1e12, allreduce_bytes=S_bytes, allreduce_s=t_ar,
         batched_cholesky_solve_s=t_solve, total_fit_s=t_stats + t_solve, max_abs_coef_err=err)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C3", "C4", "C5"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the stated sample counts (smoke runs)")
    args = ap.parse_args()
    torch.cuda.set_device(LOCAL)
    pdev.set_device(LOCAL)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                       # NCCL banners etc. go to stderr; JSON lines to the real stdout
    out = os.fdopen(saved, "w")
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    lib = _native.lib()
    fns = dict(C3=c3, C4=c4, C5=c5)
    for w in args.configs:
        l0 = lib.pcq_launch_count()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            fns[w](args.scale)
            emit(config=w, gpu_launches_rank0=int(lib.pcq_launch_count() - l0))
        out.write(buf.getvalue())
        out.flush()
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
