#!/usr/bin/env python
"""Randomised soak of the round-2 code against the oracle / numpy (run by hand on a GPU box,
`python tools/soak2.py [seed] [cases]`): the NVRTC-specialised evalPC kernels (exact: bit-identical; fast: Horner form
over the trie, incl. sparse / duplicated / deep multi-indices and chunk boundaries), the Householder QR kernels, the
matrix forms behind lreg.predicta, and the pcq_group calls against the single-plan calls."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                      # noqa: E402
from oracle import pc_oracle as orc              # noqa: E402  (checker only)
from pytuq_b200 import _native, device as pdev   # noqa: E402
from pytuq_b200.rv.pcrv import PCRV              # noqa: E402
from pytuq_b200.lreg import tsqr                 # noqa: E402
from pytuq_b200.lreg.forms import matrix_forms   # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(seed)
os.environ["PCQ_SPEC_CACHE_DIR"] = ""            # always compile: the generator is what is under test
worst = {"fast": 0.0, "qr": 0.0, "forms": 0.0}
lib = _native.lib()
ndev = torch.cuda.device_count()
for it in range(cases):
    fams = [["LU", "HG", "nLU", "nHG"][rng.integers(4)] for _ in range(14)]
    d = int(rng.integers(1, 14))
    kind = rng.integers(4)
    if kind == 0:                                # graded total-order set, possibly thinned (not downward closed)
        mi = orc.get_mi(int(rng.integers(0, 5)), d)
        if mi.shape[0] > 1 and rng.random() < 0.5:
            mi = mi[np.sort(rng.choice(mi.shape[0], max(1, mi.shape[0] // int(rng.integers(2, 5))), replace=False))]
    elif kind == 1:                              # arbitrary rows, duplicates likely
        mi = rng.integers(0, 3, size=(int(rng.integers(1, 300)), d))
    elif kind == 2:                              # deep terms: many non-zero orders
        mi = (rng.random((int(rng.integers(1, 120)), d)) < 0.7).astype(np.int64) * rng.integers(1, 3, size=(1, d))
    else:                                        # more than one 1024-term kernel
        mi = orc.get_mi(3, 12)[: int(rng.integers(1025, 2400))] if d >= 1 else None
        d = 12
    if mi.shape[0] > 2400:
        mi = mi[:2400]
    K = mi.shape[0]
    M = int(rng.choice([1, 1, 2, 5, 8]))
    n = int(rng.choice([1, 33, 255, 256, 257, 1000, 5000]))
    types = fams[:d]
    x = np.stack([rng.uniform(-1, 1, n) if t.endswith("LU") else rng.standard_normal(n) for t in types], axis=1)
    cfs = rng.standard_normal((M, K)) / (1.0 + mi.sum(axis=1))
    pc = PCRV(M, d, types, mi=mi, cfs=cfs)
    ref = orc.eval_pc(x, [mi] * M, list(cfs), types)
    y0 = pc.evalPC(x)
    assert np.array_equal(y0, ref), ("generic exact", it)
    pmax = int(mi.max()) if mi.size else 0
    if d * max(pmax, 1) <= 256 and pc.specialize():
        assert np.array_equal(pc.evalPC(x), ref), ("specialised exact", it, d, K, M)
        yf = pc.evalPC(x, exact=False)
        # the reference sums term by term: compare against the magnitude of the terms, not of the (possibly cancelling) sum
        A = orc.eval_bases(x, mi, types)
        mag = np.abs(A) @ np.abs(cfs).T + 1e-300
        e = float(np.max(np.abs(yf - ref) / mag))
        worst["fast"] = max(worst["fast"], e)
        assert e < 1e-13, ("specialised fast (Horner)", it, d, K, M, e)
    pdev.clear_plans()
    # --- QR kernels
    m, L = int(rng.choice([1, 5, 40, 300, 2000, 9000])), int(rng.integers(1, 200))
    Z = rng.standard_normal((m, L)) * np.logspace(0, -float(rng.integers(0, 10)), L)[None, :]
    if L > 3 and rng.random() < 0.3:
        Z[:, L // 2] = Z[:, 0]
    R = tsqr._qr_r(torch.from_numpy(Z.copy()).cuda()).cpu().numpy()
    G = Z.T @ Z
    sc = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
    e = float(np.max(np.abs(R.T @ R - G) / sc))
    worst["qr"] = max(worst["qr"], e)
    assert e < 1e-12 and np.array_equal(R, np.triu(R)), ("qr", it, m, L, e)
    # --- matrix forms
    N, Kc, Mo = int(rng.choice([1, 17, 300, 2049])), int(rng.integers(1, 260)), int(rng.choice([1, 3, 47, 48, 130]))
    Am = rng.standard_normal((N, Kc + 2))[:, :Kc]
    Cf = rng.standard_normal((Mo, Kc))
    B = rng.standard_normal((Kc, Kc)); Cm = B @ B.T / Kc
    Yf, vf = matrix_forms(Am, Cf, Cm)
    e1 = float(np.max(np.abs(Yf - Am @ Cf.T)) / (np.max(np.abs(Am @ Cf.T)) + 1e-300))
    vr = np.einsum("ij,ij->i", Am @ Cm, Am)
    e2 = float(np.max(np.abs(vf - vr)) / (np.max(np.abs(vr)) + 1e-300))
    worst["forms"] = max(worst["forms"], e1, e2)
    assert e1 < 1e-12 and e2 < 1e-11, ("matrix forms", it, N, Kc, Mo, e1, e2)
print("soak2 ok: seed", seed, "cases", cases, "gpus", ndev, "worst relative errors", worst)
