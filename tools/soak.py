#!/usr/bin/env python
"""Randomised soak of the statistics / quadratic-form / basis kernels against the oracle (not part of the test
suite: run by hand on a GPU box, `python tools/soak.py [seed] [cases]`).  Shapes are drawn so that tile counts,
ragged tails, descriptor widths and chunking (PCQ_SYRK_CHUNK_MB=1 half of the time) all vary."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pc_oracle as orc              # noqa: E402  (checker only)
from pytuq_b200.rv.pcrv import PCRV              # noqa: E402
from pytuq_b200.lreg.stats import SuffStats      # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 120
rng = np.random.default_rng(seed)
worst = {"S": 0.0, "Q": 0.0, "K4": 0.0}
for it in range(cases):
    fam = ["LU", "HG", "nLU", "nHG"][rng.integers(4)]
    d = int(rng.integers(1, 14))
    p = int(rng.integers(0, 5))
    mi = orc.get_mi(p, d)
    if mi.shape[0] > 900:
        mi = mi[np.sort(rng.choice(mi.shape[0], 900, replace=False))]
    if rng.random() < 0.2:                      # arbitrary (non-graded, possibly wide) multi-index
        mi = rng.integers(0, 3, size=(int(rng.integers(1, 400)), d))
    K = mi.shape[0]
    n = int(rng.choice([1, 7, 31, 32, 33, 100, 255, 1000, 4097, 20000]))
    M = int(rng.choice([0, 1, 1, 2, 5, 130]))
    x = rng.uniform(-1, 1, (n, d)) if fam.endswith("LU") else rng.standard_normal((n, d))
    Y = rng.standard_normal((n, M)) if M else None
    for k in ("PCQ_SYRK_CHUNK_MB", "PCQ_SYRK_SPLIT"):
        os.environ.pop(k, None)
    if rng.random() < 0.5:
        os.environ["PCQ_SYRK_CHUNK_MB"] = "1"
    if rng.random() < 0.3:
        os.environ["PCQ_SYRK_SPLIT"] = str(int(rng.integers(1, 9)))
    pc = PCRV(1, d, fam, mi=mi)
    A = orc.eval_bases(x, mi, fam)
    assert np.array_equal(pc.evalBases(x, 0), A), ("K1", it, fam, d, p, n)
    ref = orc.suffstats(A, Y if M else np.zeros((n, 0)))
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref))) + 1e-300
    S = pc.suffStats(x, Y, 0).S
    e = float(np.max(np.abs(S - ref) / scale))
    worst["S"] = max(worst["S"], e)
    assert e < 1e-12 and np.array_equal(S, S.T), ("K3", it, fam, d, p, n, M, K, e)
    S4 = SuffStats.from_matrix(A, Y).S
    e = float(np.max(np.abs(S4 - ref) / scale))
    worst["K4"] = max(worst["K4"], e)
    assert e < 1e-12, ("K4", it, fam, d, p, n, M, K, e)
    B = rng.standard_normal((K, K))
    Cm = B @ B.T / K
    q = pc.evalQuadForm(x, Cm, 0)
    qref = np.einsum("ni,ij,nj->n", A, Cm, A)
    e = float(np.max(np.abs(q - qref)) / (np.max(np.abs(qref)) + 1e-300))
    worst["Q"] = max(worst["Q"], e)
    assert e < 1e-11, ("K6", it, fam, d, p, n, K, e)
print("soak ok: seed", seed, "cases", cases, "worst relative errors", worst)
