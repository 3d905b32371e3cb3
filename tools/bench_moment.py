"""Moment form of K3 (csrc/pcq_moment.cu) against the pack + SYRK form on resident samples: same S, time of each.

    python tools/bench_moment.py [--cases small|bench|all] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="all")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.utils.mindex import get_mi
    lib = _native.lib()
    dev = torch.device("cuda", 0)

    def stats(pc, mi, xt, yt, M, mode):
        os.environ["PCQ_GRAM_MOMENT"] = mode
        plan = pc._plan(mi, dev=0)
        K = mi.shape[0]
        S = torch.empty((K + M + 1, K + M + 1), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def call():
            _native.check(lib.pcq_suffstats_dev(plan.handle, xt.data_ptr(), xt.shape[0], xt.shape[1],
                                                yt.data_ptr() if M else None, max(M, 1), M, S.data_ptr(), 0, st), "suffstats")
        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(args.reps):
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return S.cpu().numpy(), best

    small = [("LU", 2, 2, 1000, 1), ("HG", 3, 2, 5000, 2), ("LU", 4, 3, 20000, 1), ("HG", 8, 3, 30001, 1),
             ("nLU", 6, 4, 7777, 0), ("HG", 7, 2, 123, 3), ("LU", 10, 4, 100000, 1)]
    bench = [("HG", 20, 3, 1250000, 1), ("HG", 20, 3, 10000000, 1), ("LU", 50, 2, 1250000, 1), ("HG", 30, 3, 200000, 1)]
    cases = small if args.cases == "small" else bench if args.cases == "bench" else small + bench
    if args.cases == "high":      # higher orders: the ratio of moments to Gram entries falls with p
        cases = [("HG", 20, 4, 200000, 1), ("HG", 10, 5, 1000000, 1), ("LU", 6, 8, 1000000, 1), ("nLU", 12, 4, 1000000, 2)]
    if args.cases == "p4":
        cases = [("HG", 20, 4, 200000, 1), ("HG", 20, 3, 1250000, 1)]
    if args.cases == "c2":
        cases = [("HG", 20, 3, 2500000, 1)]
    if args.cases == "prof":
        cases = [("HG", 20, 3, 400000, 1)]
    for fam, d, p, n, M in cases:
        mi = get_mi(p, d)
        pc = PCRV(1, d, fam, mi=mi)
        g = torch.Generator(device=dev).manual_seed(1)
        xt = torch.randn((n, d), dtype=torch.float64, device=dev, generator=g) if fam in ("HG", "nHG") else \
            torch.rand((n, d), dtype=torch.float64, device=dev, generator=g) * 2 - 1
        yt = torch.randn((n, max(M, 1)), dtype=torch.float64, device=dev, generator=g)
        t0 = time.time()
        S1, t1 = stats(pc, mi, xt, yt, M, "1")
        first = time.time() - t0
        S0, t0_ = stats(pc, mi, xt, yt, M, "0")
        dg = np.sqrt(np.abs(np.diag(S0)))
        err = float(np.max(np.abs(S1 - S0) / np.outer(dg, dg)))
        K = mi.shape[0]
        print(json.dumps({"case": f"{fam} d={d} p={p} K={K} N={n} M={M}", "moment_ms": round(t1, 3), "syrk_ms": round(t0_, 3),
                          "speedup": round(t0_ / t1, 2), "max_rel_diff": err, "first_call_s": round(first, 2),
                          "basis_evals_per_s": n * K / (t1 * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
