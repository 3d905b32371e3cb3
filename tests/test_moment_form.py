"""Moment form of the sufficient statistics (csrc/pcq_moment.cu, include/pcq.h: pcq_plan_enable_moments).

CPU part: the planner's units executed on the host by pcq_debug_moments_host (table -> sub-functions -> functions ->
task products -> reconstruction through the linearisation coefficients) against  Z^T Z  of the oracle's design matrix,
Z = [Psi | y | 1]  (lreg/anl.py:78,83; lreg/bcs.py:83-100).  GPU part: the kernels against the oracle, the pack + SYRK
form and the reference goldens.  Tolerances: 1e-12 relative to sqrt(S_ii S_jj) for the statistics, 1e-10 for fitted
coefficients (BASELINE.json north_star)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import pc_oracle as orc


def _tables(fams, nord):
    from pytuq_b200.rv.pcrv import PC1d
    d = len(fams)
    ra, rb, p1 = np.zeros((nord + 1, d)), np.zeros((nord + 1, d)), np.zeros(d)
    for i, f in enumerate(fams):
        pc = PC1d(f)
        ra[:, i], rb[:, i] = pc.rec_tables(nord)
        p1[i] = pc.p1_scale
    return ra, rb, p1


def _design(x, mi, fams):
    """Oracle design matrix for per-dimension families (products of the oracle's 1-D tables)."""
    n, d = x.shape
    pmax = int(mi.max())
    tabs = [orc.eval_bases(np.ascontiguousarray(x[:, [i]]), np.arange(pmax + 1).reshape(-1, 1), fams[i]) for i in range(d)]
    A = np.ones((n, mi.shape[0]))
    for k in range(mi.shape[0]):
        for i in range(d):
            if mi[k, i]:
                A[:, k] *= tabs[i][:, mi[k, i]]
    return A


def _sample(rng, fams, n):
    from pytuq_b200.rv.pcrv import PC1d
    return np.ascontiguousarray(np.concatenate(
        [rng.standard_normal((n, 1)) if PC1d(f).germ_kind else rng.uniform(-1, 1, (n, 1)) for f in fams], axis=1))


def _host_stats(mi, fams, x, y, m):
    from pytuq_b200 import _native
    lib = _native.lib()
    mi = np.ascontiguousarray(mi, dtype=np.int32)
    K, d = mi.shape
    nord = 2 * int(mi.sum(1).max())
    ra, rb, p1 = _tables(fams, nord)
    L = K + m + 1
    S, info = np.zeros((L, L)), np.zeros(8)
    n = x.shape[0]
    yy = np.ascontiguousarray(y) if m else np.zeros((max(n, 1), 1))
    rc = lib.pcq_debug_moments_host(d, K, _native.ptr(mi), nord, _native.ptr(ra), _native.ptr(rb), _native.ptr(p1),
                                    _native.ptr(x), n, d, _native.ptr(yy), max(m, 1), m, _native.ptr(S), _native.ptr(info))
    return rc, S, info


def _rel(S, Sref):
    dg = np.sqrt(np.abs(np.diag(Sref)))
    return float(np.max(np.abs(S - Sref) / np.outer(dg, dg)))


CASES = [("LU", 2, 2, 1), ("HG", 3, 2, 2), ("LU", 4, 3, 1), (["LU", "HG", "nLU", "nHG", "LU"], 5, 3, 2), ("HG", 7, 2, 0),
         ("nLU", 6, 3, 1), ("nHG", 3, 4, 3), ("HG", 8, 3, 1), ("LU", 2, 7, 1), ("LU", 40, 2, 1), ("HG", 33, 1, 2),
         ("HG", 20, 3, 1), ("LU", 50, 2, 2), ("HG", 10, 4, 1)]      # the last three: C2 (headline), C3 and a p = 4 shape


@pytest.mark.parametrize("fam,d,p,m", CASES)
def test_planned_units_on_the_host_reproduce_the_statistics(fam, d, p, m):
    rng = np.random.default_rng(d * 100 + p)
    fams = fam if isinstance(fam, list) else [fam] * d
    mi = orc.get_mi(p, d)
    n = 57
    x = _sample(rng, fams, n)
    y = rng.standard_normal((n, max(m, 1)))
    rc, S, info = _host_stats(mi, fams, x, y, m)
    assert rc == 0 and info[6] == 1.0
    Z = np.concatenate([_design(x, mi, fams), y[:, :m], np.ones((n, 1))], axis=1)
    assert _rel(S, Z.T @ Z) < 1e-13


def test_sparse_and_anisotropic_multi_indices():
    """Any multi-index set works (the moment set is bounded by twice its highest total order): a hyperbolic-cross-like
    set, a random subset without the constant term, duplicated rows."""
    rng = np.random.default_rng(5)
    full = orc.get_mi(4, 5)
    sets = [full[np.prod(full + 1, axis=1) <= 5],
            full[rng.permutation(full.shape[0])[:40]][1:],
            np.concatenate([full[3:9], full[5:7]])]
    for mi in sets:
        x = _sample(rng, ["LU"] * 5, 41)
        y = rng.standard_normal((41, 1))
        rc, S, _ = _host_stats(mi, ["LU"] * 5, x, y, 1)
        assert rc == 0
        Z = np.concatenate([_design(x, mi, ["LU"] * 5), y, np.ones((41, 1))], axis=1)
        assert _rel(S, Z.T @ Z) < 1e-13


def test_plan_of_the_headline_shape():
    """d = 20, p = 3: 231 999 moments (230 229 plain ones with 1 <= |gamma| <= 6 plus 1 770 weighted by y) instead of
    1 569 106 Gram entries; every unit within the kernel's capacities."""
    from math import comb
    mi = orc.get_mi(3, 20)
    rc, _, info = _host_stats_plan_only(mi, ["HG"] * 20, 1)
    assert rc == 0 and info[6] == 1.0
    assert int(info[0]) == comb(26, 6) - 1 + comb(23, 3) - 1
    assert info[4] <= 1280 and info[5] <= 384
    assert info[3] < 0.3 * mi.shape[0] ** 2 / 2          # executed multiply-adds per sample against K^2 / 2


def _host_stats_plan_only(mi, fams, m):
    from pytuq_b200 import _native
    lib = _native.lib()
    mi = np.ascontiguousarray(mi, dtype=np.int32)
    K, d = mi.shape
    nord = 2 * int(mi.sum(1).max())
    ra, rb, p1 = _tables(fams, nord)
    info = np.zeros(8)
    rc = lib.pcq_debug_moments_host(d, K, _native.ptr(mi), nord, _native.ptr(ra), _native.ptr(rb), _native.ptr(p1),
                                    None, 0, d, None, max(m, 1), m, None, _native.ptr(info))
    return rc, None, info


def test_refusals():
    """Shapes the form does not serve are refused (the caller keeps the pack + SYRK form): one dimension, more than
    64 dimensions, a constant basis."""
    for mi, fams in [(np.arange(4).reshape(-1, 1), ["LU"]), (orc.get_mi(1, 70), ["LU"] * 70), (np.zeros((1, 3), dtype=int), ["HG"] * 3)]:
        rc, _, info = _host_stats_plan_only(mi, fams, 1)
        assert rc == -5 and info[6] == 0.0


# ------------------------------------------------------------------------------------------------ GPU
def _dev_stats(pc, mi, x, y, m, mode, accumulate_in_two=False, ldx=None):
    import torch
    from pytuq_b200 import _native
    lib = _native.lib()
    old = os.environ.get("PCQ_GRAM_MOMENT")
    os.environ["PCQ_GRAM_MOMENT"] = mode
    try:
        dev = torch.device("cuda", 0)
        plan = pc._plan(mi, dev=0)
        K, s = mi.shape
        n = x.shape[0]
        ldx = s if ldx is None else ldx
        xp = np.zeros((n, ldx))
        xp[:, :s] = x
        xt = torch.from_numpy(xp).to(dev)
        yt = torch.from_numpy(np.ascontiguousarray(y if m else np.zeros((n, 1)))).to(dev)
        S = torch.full((K + m + 1, K + m + 1), np.nan, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        cuts = [0, n] if not accumulate_in_two else [0, n // 3, n]
        for c in range(len(cuts) - 1):
            a, b = cuts[c], cuts[c + 1]
            _native.check(lib.pcq_suffstats_dev(plan.handle, xt.data_ptr() + a * ldx * 8, b - a, ldx,
                                                yt.data_ptr() + a * max(m, 1) * 8 if m else None, max(m, 1), m,
                                                S.data_ptr(), 1 if c else 0, st), "pcq_suffstats_dev")
        torch.cuda.synchronize()
        return S.cpu().numpy()
    finally:
        if old is None:
            os.environ.pop("PCQ_GRAM_MOMENT", None)
        else:
            os.environ["PCQ_GRAM_MOMENT"] = old


@pytest.mark.gpu
@pytest.mark.parametrize("fam,d,p,m,n", [("LU", 2, 2, 1, 1), ("LU", 2, 2, 1, 1000), ("HG", 3, 2, 2, 5003), ("LU", 4, 3, 1, 20000),
                                         ("nLU", 6, 3, 0, 7777), ("nHG", 3, 4, 3, 15), ("HG", 8, 3, 1, 30001), ("LU", 50, 2, 1, 9001),
                                         (["LU", "HG", "nLU", "nHG", "LU"], 5, 3, 2, 4097)])
def test_kernels_vs_oracle(fam, d, p, m, n):
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(n + d)
    fams = fam if isinstance(fam, list) else [fam] * d
    mi = orc.get_mi(p, d)
    pc = PCRV(1, d, fam, mi=mi)
    x = _sample(rng, fams, n)
    y = rng.standard_normal((n, max(m, 1)))
    Z = np.concatenate([_design(x, mi, fams), y[:, :m], np.ones((n, 1))], axis=1)
    Sref = Z.T @ Z
    assert _rel(_dev_stats(pc, mi, x, y, m, "1"), Sref) < 1e-12
    assert _rel(_dev_stats(pc, mi, x, y, m, "1", accumulate_in_two=n > 2, ldx=d + 3), Sref) < 1e-12
    assert _rel(_dev_stats(pc, mi, x, y, m, "0"), Sref) < 1e-12


@pytest.mark.gpu
def test_sample_chunks_and_parts(monkeypatch):
    """Several table chunks (PCQ_MOMENT_CHUNK_MB small) and forced numbers of sample parts give the same moments."""
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(11)
    mi = orc.get_mi(3, 6)
    pc = PCRV(1, 6, "HG", mi=mi)
    x = rng.standard_normal((60_011, 6))
    y = rng.standard_normal((60_011, 1))
    base = _dev_stats(pc, mi, x, y, 1, "0")
    for chunk, split in (("1", None), ("4", "1"), ("2048", "7")):
        monkeypatch.setenv("PCQ_MOMENT_CHUNK_MB", chunk)
        if split:
            monkeypatch.setenv("PCQ_MOMENT_SPLIT", split)
        assert _rel(_dev_stats(pc, mi, x, y, 1, "1"), base) < 1e-12


@pytest.mark.gpu
def test_headline_shape_fits_through_the_moment_form_vs_reference(monkeypatch):
    """lsq and anl at the headline shape (HG d=20 p=3, K=1771) with the statistics in moment form, against the
    unmodified reference's coefficients (tests/golden/c2_shape.npz; lreg/lreg.py:328-339, lreg/anl.py:69-105)."""
    from conftest import golden_generator
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import lsq, anl, PCBasisEvaluator
    monkeypatch.setenv("PCQ_GRAM_MOMENT", "1")
    g = load_golden("c2_shape")
    x, _, _ = golden_generator().c2_inputs()
    y = g["y"]
    pc = PCRV(1, 20, "HG", mi=orc.get_mi(3, 20))
    scale = np.max(np.abs(g["cf_lsq"]))
    for name, make in (("lsq", lsq), ("anl", anl)):
        fit = make()
        fit.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        fit.fit(x, y)
        np.testing.assert_allclose(fit.cf, g[f"cf_{name}"], rtol=0, atol=1e-10 * scale)
        if name == "anl":
            np.testing.assert_allclose(fit.datavar, g["anl_datavar"], rtol=1e-9)


@pytest.mark.gpu
def test_full_size_agreement_of_the_two_forms():
    """1.25e6 samples at the headline shape: both forms give the same statistics, and the fit on them recovers the
    generating coefficients."""
    from pytuq_b200.rv.pcrv import PCRV
    mi = orc.get_mi(3, 20)
    K = mi.shape[0]
    pc = PCRV(1, 20, "HG", mi=mi)
    n = 1_250_000
    rng = np.random.default_rng(3)
    x = rng.standard_normal((n, 20))
    cf = rng.standard_normal(K) / (1.0 + np.arange(K))
    pc.setCfs([cf])
    y = (pc.evalPC(x)[:, 0] + 1e-3 * rng.standard_normal(n)).reshape(-1, 1)
    Sm = _dev_stats(pc, mi, x, y, 1, "1")
    Ss = _dev_stats(pc, mi, x, y, 1, "0")
    assert _rel(Sm, Ss) < 1e-12
    sol = np.linalg.solve(Sm[:K, :K], Sm[:K, K])
    assert np.max(np.abs(sol - cf)) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("fam,d,p,m,n", [("HG", 3, 2, 9, 5003), ("LU", 6, 3, 20, 7777), ("HG", 8, 3, 130, 30001),
                                         ("nLU", 5, 4, 300, 4099)])
def test_more_outputs_than_the_moment_form_carries(fam, d, p, m, n, monkeypatch):
    """m > MOM_MAXOUT: G, A^T 1 and N in moment form, the output columns (A^T Y, Y^T Y, Y^T 1) by the SYRK over the
    tile columns that hold them (pcq_launch_gram's hybrid route; workflows/fits.py:47-64 with a field of outputs) --
    against Z^T Z of the oracle's design matrix, in one call, accumulated over two calls, and with the dense form."""
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(n + m)
    fams = [fam] * d
    mi = orc.get_mi(p, d)
    pc = PCRV(1, d, fam, mi=mi)
    x = _sample(rng, fams, n)
    y = rng.standard_normal((n, m))
    Z = np.concatenate([_design(x, mi, fams), y, np.ones((n, 1))], axis=1)
    Sref = Z.T @ Z
    S1 = _dev_stats(pc, mi, x, y, m, "1")
    assert not np.isnan(S1).any()
    assert _rel(S1, Sref) < 1e-12
    np.testing.assert_array_equal(S1, S1.T)
    assert _rel(_dev_stats(pc, mi, x, y, m, "1", accumulate_in_two=True, ldx=d + 1), Sref) < 1e-12
    monkeypatch.setenv("PCQ_SYRK_TILE", "96")
    assert _rel(_dev_stats(pc, mi, x, y, m, "1"), Sref) < 1e-12
    monkeypatch.delenv("PCQ_SYRK_TILE")
    assert _rel(_dev_stats(pc, mi, x, y, m, "0"), Sref) < 1e-12


def test_planner_sweep_capacities_and_refusals():
    """Total-order sets over germ dimensions 2..64 and orders 1..6: the planner serves exactly the shapes whose table of
    1-D values fits (1 + 2 p d + m + 1 <= 256 slots), every unit within the kernel's capacities (<= 1280 functions,
    <= 384 sub-functions), and executes fewer multiply-adds than the dense form wherever K is large."""
    from math import comb
    for d in (2, 5, 9, 16, 33, 50, 64):
        for p in (1, 2, 3, 5, 6):
            K = comb(d + p, p)
            if K > 7000:
                continue
            mi = orc.get_mi(p, d)
            for m in (0, 8):
                rc, _, info = _host_stats_plan_only(mi, ["LU"] * d, m)
                fits = 1 + 2 * p * d + m + 1 <= 256
                assert (rc == 0) == fits, (d, p, m, rc)
                if rc == 0:
                    assert info[4] <= 1280 and info[5] <= 384 and info[1] >= 1
                    if K >= 1000 and m == 0:
                        assert info[3] < 0.7 * K * (K + 1) / 2, (d, p, info[3])


def test_conditioning_estimate_of_the_reconstruction():
    """info[7]: how much larger than a Gram entry the moments it is rebuilt from can be (max_a sum_k |c_aak| sqrt(h_k) / h_a).
    He_3^2 = He_6 + 9 He_4 + 18 He_2 + 6 gives (sqrt(720) + 9 sqrt(24) + 18 sqrt(2)) / 6 = 16.06; Legendre stays below
    2 a + 1; Hermite beyond order 6 exceeds the bound (500) above which the route rule keeps the dense form."""
    amp = {}
    for fam, d, p in [("HG", 20, 3), ("LU", 50, 2), ("LU", 6, 8), ("nLU", 6, 4), ("HG", 4, 6), ("HG", 3, 7), ("nHG", 3, 8)]:
        rc, _, info = _host_stats_plan_only(orc.get_mi(p, d), [fam] * d, 1)
        assert rc == 0
        amp[(fam, p)] = info[7]
    assert abs(amp[("HG", 3)] - (np.sqrt(720.0) + 9 * np.sqrt(24.0) + 18 * np.sqrt(2.0)) / 6.0) < 1e-9
    assert amp[("LU", 2)] < 5 and amp[("LU", 8)] < 17 and amp[("nLU", 4)] < 9
    assert amp[("HG", 6)] < 500 < amp[("HG", 7)] < amp[("nHG", 8)]
