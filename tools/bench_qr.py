#!/usr/bin/env python
"""Times the QR statistic (csrc/pcq_qr.cu) against cuSOLVER's geqrf (torch.linalg.qr) on the shapes of the path."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pytuq_b200 import _native
from pytuq_b200.lreg import tsqr

lib = _native.lib()
dev = torch.device("cuda", 0)
cap = int(lib.pcq_qr_max_rows(0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (m, L) in [(cap, 1772), (cap, 1002), (20000, 1772), (cap, 257), (3544, 1772), (cap, 71)]:
    Z = torch.randn((m, L), dtype=torch.float64, device=dev)
    ts = []
    for it in range(4):
        W = Z.clone()
        e0.record(); R = tsqr._qr_r(W); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    tl = []
    for it in range(3):
        W = Z.clone()
        e0.record(); R2 = torch.linalg.qr(W, mode="r").R; e1.record(); e1.synchronize()
        tl.append(e0.elapsed_time(e1))
    err = float(((R.T @ R) - (R2.T @ R2)).abs().max() / (R2.T @ R2).abs().max())
    fl = 2.0 * m * L * L - 2.0 / 3.0 * L ** 3
    print(json.dumps(dict(rows=m, cols=L, pcq_qr_ms=min(ts), cusolver_geqrf_ms=min(tl), pcq_TFLOPs=fl / min(ts) / 1e9,
                          rel_diff_RtR=err)), flush=True)

# the K x K SVD that ends the QR route: one-sided Jacobi kernels (csrc/pcq_svd.cu) against cuSOLVER (torch.linalg.svd)
for n in (257, 1002, 1772):
    Z = torch.randn((4 * n, n), dtype=torch.float64, device=dev)
    R = tsqr._qr_r(Z.clone())
    torch.cuda.synchronize()
    t0 = time.perf_counter(); Bt, Vt, sig, sweeps = tsqr.svd_jacobi(R); torch.cuda.synchronize(); tj = time.perf_counter() - t0
    t0 = time.perf_counter(); S2 = torch.linalg.svd(R).S; torch.cuda.synchronize(); tl = time.perf_counter() - t0
    err = float((sig.sort(descending=True).values - S2).abs().max() / S2[0])
    print(json.dumps(dict(svd_n=n, jacobi_ms=1e3 * tj, sweeps=sweeps, cusolver_svd_ms=1e3 * tl, rel_diff_sigma=err)), flush=True)
