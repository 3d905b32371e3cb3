"""CPU tests of the host-side logic (no GPU arithmetic): fitters on sufficient statistics,
PCRV bookkeeping and Sobol reductions, multi-index utilities, the C-ABI surface.
Statistics are built here with numpy from the golden design matrices, which isolates the
K x K host code from the kernels (those are covered by the -m gpu tests)."""
import copy
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, case_mis_cfs, case_types
from oracle import pc_oracle as orc


def _stats(A, y):
    from pytuq_b200.lreg.stats import SuffStats
    Y = np.asarray(y).reshape(A.shape[0], -1)
    return SuffStats(orc.suffstats(A, Y), A.shape[1], Y.shape[1])


def test_library_exports_every_declared_symbol():
    from pytuq_b200 import _native
    header = open(os.path.join(ROOT, "include", "pcq.h")).read()
    declared = set(re.findall(r"\b(pcq_[a-z0-9_]+)\s*\(", header)) - {"pcq_plan", "pcq_status", "pcq_text"}
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _native.lib().pcq_version() == 100


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the arithmetic entry points must raise, never fall back."""
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    if _native.lib().pcq_device_count() > 0:
        pytest.skip("a GPU is present")
    pc = PCRV(1, 2, "LU", mi=orc.get_mi(2, 2))
    with pytest.raises(_native.PcqError):
        pc.evalBases(np.zeros((3, 2)), 0)
    with pytest.raises(_native.PcqError):
        pc.evalPC(np.zeros((3, 2)))
    with pytest.raises(_native.PcqError):
        pc.suffStats(np.zeros((3, 2)), np.zeros(3))
    with pytest.raises(_native.PcqError):
        pc.evalQuadForm(np.zeros((3, 2)), np.eye(6), 0)
    from pytuq_b200.lreg import lsq, tsqr
    with pytest.raises(_native.PcqError):
        lsq().fita(np.ones((5, 2)), np.ones(5))
    with pytest.raises(_native.PcqError):
        tsqr.r_factor_pc(pc, np.zeros((3, 2)), np.zeros(3))
    from pytuq_b200.linred.kle import KLE
    with pytest.raises(_native.PcqError):
        KLE().build(np.ones((4, 9)), plot=False)
    from pytuq_b200.utils import io
    with pytest.raises(_native.PcqError):          # the device reader does not quietly become the host reader
        io.loadtxt_device(os.path.join(ROOT, "tests", "golden", "uqpc", "qtrain.txt"))


def test_text_entry_points_reject_bad_arguments():
    from pytuq_b200 import _native
    lib = _native.lib()
    rows, cols, h = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_void_p()
    assert lib.pcq_text_open(0, None, ctypes.byref(h), ctypes.byref(rows), ctypes.byref(cols)) == -1
    assert lib.pcq_text_open(-1, b"/nonexistent", ctypes.byref(h), ctypes.byref(rows), ctypes.byref(cols)) == -1
    assert b"bad argument" in lib.pcq_last_error()
    assert lib.pcq_text_parse_dev(None, None, 0, None) == -1
    assert lib.pcq_text_close(None) == 0


def test_c_abi_rejects_bad_arguments_before_touching_the_device():
    """Argument errors are reported through the status code + pcq_last_error(), on any box."""
    from pytuq_b200 import _native
    lib = _native.lib()
    buf = np.zeros(4)
    for call in (lambda: lib.pcq_eval_bases_dev(None, _native.ptr(buf), 1, 2, _native.ptr(buf), 2, None),
                 lambda: lib.pcq_eval_pc_dev(None, _native.ptr(buf), 1, 2, _native.ptr(buf), 1, _native.ptr(buf), 1, 0, None),
                 lambda: lib.pcq_suffstats_dev(None, _native.ptr(buf), 1, 2, _native.ptr(buf), 1, 1, _native.ptr(buf), 0, None),
                 lambda: lib.pcq_quadform_dev(None, _native.ptr(buf), 1, 2, _native.ptr(buf), 2, _native.ptr(buf), 1, None),
                 lambda: lib.pcq_residual_ss_dev(None, _native.ptr(buf), 1, 2, _native.ptr(buf), 1, _native.ptr(buf),
                                                 _native.ptr(buf), 0, None),
                 lambda: lib.pcq_gram_dev(0, _native.ptr(buf), 2, 1, 2, None, 1, 0, _native.ptr(buf), 0, None),   # lda < kcols
                 lambda: lib.pcq_plan_specialize(None, 1)):
        rc = call()
        assert rc < 0
    assert lib.pcq_gram_dev(0, _native.ptr(buf), 2, 1, 2, None, 1, 0, _native.ptr(buf), 0, None) == -1
    assert b"bad argument" in lib.pcq_last_error()
    assert lib.pcq_suffstats_dim(None, 1) < 0 and lib.pcq_plan_info(None, 0) < 0


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import pytuq_b200, pytuq_b200.rv.pcrv, pytuq_b200.lreg, "
            "pytuq_b200.workflows, pytuq_b200.surrogates.pce, pytuq_b200.gsa.gsa, pytuq_b200.parallel, "
            "pytuq_b200.lreg.tsqr, pytuq_b200.linred.kle, pytuq_b200.linred.klsurr; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pytuq_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_mindex_utilities():
    from pytuq_b200.utils import mindex as mx
    g = load_golden("mindex")
    for key in g.files:
        if key.startswith("mi_"):
            _, p, d = key.split("_")
            got = mx.get_mi(int(p), int(d))
            assert got.dtype == g[key].dtype and np.array_equal(got, g[key])
    for p, d, n in g["npc"]:
        assert mx.get_npc(int(p), int(d)) == n
    mi = mx.get_mi(2, 3)
    new, add, front = mx.mi_addfront(mi)
    assert new.shape[0] == mi.shape[0] + add.shape[0] and add.shape[0] == 10
    assert set(map(tuple, new)) == set(map(tuple, mx.get_mi(3, 3)))
    newc, addc, _ = mx.mi_addfront_cons(mi[:7])        # 002 / 020 / 011 missing
    have = set(map(tuple, mi[:7]))
    for row in addc:
        for k in range(3):
            if row[k]:
                par = row.copy(); par[k] -= 1
                assert tuple(par) in have
    mall, call = mx.micf_join([mi[:4], mi[2:6]], [np.arange(4.), 10 + np.arange(4.)])
    assert mall.shape[0] == 6 and call.shape == (2, 6)
    lut = {tuple(r): k for k, r in enumerate(mall)}
    assert call[0, lut[(1, 0, 0)]] == 1.0 and call[1, lut[(1, 1, 0)]] == 13.0 and call[1, lut[(0, 0, 0)]] == 0.0
    enc = mx.encode_mindex(mi[:5]) if True else None
    assert enc[0] == ([0], [0]) and enc[4] == ([1], [2])


def test_pcrv_bookkeeping_and_sobol_vs_reference(pcrv_cases):
    from pytuq_b200.rv.pcrv import PCRV
    for name, c in pcrv_cases.items():
        mis, cfs = case_mis_cfs(c)
        pc = PCRV(len(mis), mis[0].shape[1], case_types(c), mi=mis, cfs=cfs)
        for j in range(pc.pdim):
            assert np.array_equal(pc.evalBasesNormsSq(j), c[f"normsq{j}"]), name
        np.testing.assert_allclose(pc.computeMean(), c["mean"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(pc.computeVar(), c["var"], rtol=1e-12)
        np.testing.assert_allclose(pc.computeSens(), c["main"], rtol=1e-10, atol=1e-15)
        np.testing.assert_allclose(pc.computeTotSens(), c["total"], rtol=1e-10, atol=1e-15)
        if "joint" in c:
            np.testing.assert_allclose(pc.computeJointSens(), c["joint"], rtol=1e-10, atol=1e-15)
        if pc.sdim > 2:
            np.testing.assert_allclose(pc.computeGroupSens([0, 1]), c["group01"], rtol=1e-10, atol=1e-15)
        # objects stay picklable / deep-copyable (reference pickles them: utils/xutils.py:36-55)
        clone = pickle.loads(pickle.dumps(copy.deepcopy(pc)))
        assert clone.pctypes == pc.pctypes and np.array_equal(clone.maxOrd, pc.maxOrd)
        flat = pc.cfsFlatten()
        pc.cfsUnflatten(flat * 2.0)
        assert np.array_equal(pc.coefs[0], cfs[0] * 2.0)


def test_pcrv_misc_host_api():
    from pytuq_b200.rv.pcrv import PCRV, PCRV_iid, PCRV_mvn, PC1d
    pc = PCRV(2, 3, "LU")
    assert pc.mindices[0].shape == (1, 3) and pc.detind == [0, 1] and pc.rndind == []
    assert repr(pc).startswith("['LU', 'LU', 'LU'] PC Random Variable")
    p1 = PC1d("LU")
    assert p1.a(1) == 1.5 and p1.b(1) == -0.5 and p1.normsq(2) == 0.2 and PC1d("HG").normsq(4) == 24.0
    pts, w = p1.quad(5)
    assert abs(np.sum(w) - 1) < 1e-14 and abs(np.sum(w * pts ** 2) - 1 / 3) < 1e-14
    np.random.seed(42)
    g1 = PCRV(1, 3, ["LU", "HG", "nLU"]).sampleGerm(5, seed=7)
    np.random.seed(7)
    ref = np.stack([2. * np.random.rand(5) - 1., np.random.randn(5), 2. * np.random.rand(5) - 1.], axis=1)
    assert np.array_equal(g1, ref)
    q, wq = PCRV(1, 2, "HG").quadGerm([3, 2])
    assert q.shape == (6, 2) and abs(wq.sum() - 1) < 1e-14
    iid = PCRV_iid(3, "HG", orders=np.array([1, 2, 3]))
    assert [m.shape[0] for m in iid.mindices] == [2, 3, 4] and iid.maxOrd.tolist() == [1, 2, 3]
    cov = np.array([[2.0, 0.3], [0.3, 1.0]])
    mvn = PCRV_mvn(2, mean=np.array([1.0, -1.0]), cov=cov)
    np.testing.assert_allclose(mvn.computeMean(), [1, -1])
    np.testing.assert_allclose(mvn.computeVar(), [2.0, 1.0], rtol=1e-14)
    with pytest.raises(SystemExit):
        PC1d("LG")
    c = PCRV(1, 3, "LU", mi=np.array([[0, 0, 0], [2, 0, 1], [2, 0, 1]]), cfs=np.array([1., 2., 3.]))
    c.compressMI()
    assert c.mindices[0].tolist() == [[0, 0, 0], [2, 0, 1]] and c.coefs[0].tolist() == [1., 5.]
    assert c.compressPC().sdim == 2
    s1 = c.slice_1d(0)
    assert s1.sdim == 1 and s1.mindices[0].tolist() == [[0]]


def test_fitters_on_statistics_vs_reference(fit_cases):
    from pytuq_b200.lreg import lsq, anl, bcs
    for name, c in fit_cases.items():
        A = orc.eval_bases(c["x"], c["mi"], str(c["type"]))
        st = _stats(A, c["y"])
        l = lsq(); l._fit_stats(st)
        np.testing.assert_allclose(l.cf, c["lsq_cf"], rtol=1e-10, atol=1e-12)
        a = anl(); a._fit_stats(st)
        np.testing.assert_allclose(a.cf, c["anl_cf"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(a.datavar, c["anl_datavar"], rtol=1e-8)
        np.testing.assert_allclose(a.cf_cov, c["anl_cov"], rtol=1e-8, atol=1e-16)
        ap = anl(datavar=0.01, prior_var=2.0, cov_nugget=1e-8); ap._fit_stats(st)
        np.testing.assert_allclose(ap.cf, c["anlp_cf"], rtol=1e-10, atol=1e-12)
        av = anl(method="vi"); av._fit_stats(st)
        np.testing.assert_allclose(np.diag(av.cf_cov), c["anlvi_cov_diag"], rtol=1e-8)
        for tag in ("bcs0.001", "bcs1e-08"):
            b = bcs(eta=float(tag[3:])); b._fit_stats(st)
            assert np.array_equal(b.used, c[f"{tag}_used"]), (name, tag)
            np.testing.assert_allclose(b.cf, c[f"{tag}_cf"], rtol=1e-10, atol=1e-12)
            assert np.shape(b.datavar) == (1,)
            np.testing.assert_allclose(b.datavar, c[f"{tag}_datavar"], rtol=1e-7)
            np.testing.assert_allclose(b.cf_cov, c[f"{tag}_cov"], rtol=1e-7, atol=1e-18)
    # predictions are O(N K^2) device work (no host fallback): without a GPU they fail loudly
    import torch
    if not torch.cuda.is_available():
        from pytuq_b200._native import PcqError
        with pytest.raises(PcqError):
            a.predicta(orc.eval_bases(c["xt"], c["mi"], str(c["type"])), msc=1)


def test_bcs_gram_form_tracks_matrix_form_through_deletions():
    """Random sparse problems, including ones where the fast-RVM deletes terms: the Gram-form
    iteration must follow the A-based oracle decision for decision."""
    from pytuq_b200.lreg.bcs import bcs_fit_stats
    rng = np.random.default_rng(77)
    ndel = 0
    for trial in range(30):
        n, d = int(rng.integers(60, 300)), int(rng.integers(2, 6))
        mi = orc.get_mi(3, d)
        x = rng.uniform(-1, 1, (n, d))
        A = orc.eval_bases(x, mi, "LU")
        c = np.zeros(mi.shape[0])
        sup = rng.choice(mi.shape[0], 4, replace=False)
        c[sup] = rng.standard_normal(4)
        y = A @ c + 0.3 * rng.standard_normal(n)
        for eta in (1e-2, 1e-9):
            w0, e0, u0, s0, Sig0 = orc.bcs_fit(A, y, eta=eta)
            w1, e1, u1, s1, _, Sig1 = bcs_fit_stats(_stats(A, y), eta=eta)
            assert np.array_equal(u0, u1), (trial, eta)
            np.testing.assert_allclose(w1, w0, rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(s1, s0, rtol=1e-8)
            np.testing.assert_allclose(Sig1, Sig0, rtol=1e-7, atol=1e-14)


def test_solve_normal_rank_deficient():
    from pytuq_b200.lreg.lreg import solve_normal
    rng = np.random.default_rng(1)
    A = rng.standard_normal((50, 6))
    A[:, 5] = A[:, 0] + A[:, 1]
    y = rng.standard_normal(50)
    c = solve_normal(A.T @ A, A.T @ y)
    ref = np.linalg.lstsq(A, y, rcond=None)[0]
    np.testing.assert_allclose(A @ c, A @ ref, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(c, ref, rtol=1e-5, atol=1e-6)


def test_workflow_pieces_on_statistics():
    """pc_fit's per-output slicing of the shared statistics (SuffStats.select)."""
    from pytuq_b200.lreg.stats import SuffStats
    rng = np.random.default_rng(4)
    A = rng.standard_normal((40, 5)); Y = rng.standard_normal((40, 3))
    st = SuffStats(orc.suffstats(A, Y), 5, 3)
    for j in range(3):
        sj = st.select(j)
        ref = orc.suffstats(A, Y[:, j])
        assert sj.M == 1 and sj.K == 5 and sj.S is st.S          # a view, nothing copied
        np.testing.assert_allclose(sj.G, ref[:5, :5], rtol=1e-14)
        np.testing.assert_allclose(sj.b(0), ref[:5, 5], rtol=1e-14)
        np.testing.assert_allclose([sj.yty(0), sj.ysum(0), sj.n], [ref[5, 5], ref[5, 6], ref[6, 6]], rtol=1e-14)
        np.testing.assert_allclose(sj.B()[:, 0], ref[:5, 5], rtol=1e-14)
        assert abs(sj.yvar(0) - np.std(Y[:, j]) ** 2) < 1e-13
        cf = rng.standard_normal(3); used = np.array([0, 2, 4])
        assert abs(sj.rss(cf, used) - np.sum((Y[:, j] - A[:, used] @ cf) ** 2)) < 1e-11


def test_host_copy_rows_flat_and_strided():
    """pcq_host_copy_rows (the pageable -> pinned copy of the staging pipelines): flat and strided blocks, below and above
    the size where it goes multi-threaded, and its argument checks."""
    from pytuq_b200 import _native
    lib = _native.lib()
    rng = np.random.default_rng(0)
    for rows, width, lds, ldd in [(0, 3, 3, 3), (7, 1, 1, 1), (1000, 21, 21, 21), (1000, 20, 23, 26), (60000, 21, 21, 21),
                                  (60000, 20, 21, 27)]:
        src = rng.standard_normal((max(rows, 1), lds))
        dst = np.full((max(rows, 1), ldd), -7.0)
        assert lib.pcq_host_copy_rows(dst.ctypes.data, ldd, src.ctypes.data, lds, width, rows) == 0
        np.testing.assert_array_equal(dst[:rows, :width], src[:rows, :width])
        assert (dst[:rows, width:] == -7.0).all() and (dst[rows:] == -7.0).all()
    a = np.zeros((4, 4))
    assert lib.pcq_host_copy_rows(a.ctypes.data, 3, a.ctypes.data, 4, 4, 4) != 0          # row stride below the width
    assert lib.pcq_host_copy_rows(None, 4, a.ctypes.data, 4, 4, 4) != 0


def test_stage_groups_cover_the_rows_once():
    """Host rows of lreg.fit travel in stages gathered into groups of 1, 2, 4, .. stages (parallel._stage_groups): every row
    once, in order, the first group one stage, no group above the cap."""
    from pytuq_b200 import parallel
    for n, rows, cap in [(0, 100, 8), (1, 100, 8), (100, 100, 8), (101, 100, 8), (12345, 100, 8), (12345, 100, 1), (999, 7, 4)]:
        groups = parallel._stage_groups(n, rows, cap)
        flat = [st for g in groups for st in g]
        assert sum(nr for _, nr in flat) == n
        pos = 0
        for r0, nr in flat:
            assert r0 == pos and 0 < nr <= rows
            pos += nr
        assert all(len(g) <= cap for g in groups)
        if groups:
            assert len(groups[0]) == 1
            assert [len(g) for g in groups[:-1]] == [min(2 ** i, cap) for i in range(len(groups) - 1)]
