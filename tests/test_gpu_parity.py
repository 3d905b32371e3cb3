"""Parity of the CUDA path (through the C-ABI / ctypes shim) against the reference's golden
vectors and the CPU oracle.  Bit-exact for basis values and exact-mode evalPC; 1e-12 for the
tensor-pipe statistics (BLAS-order sums have no defined order in the reference either);
1e-10 for fitted coefficients and Sobol indices (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import load_golden, case_mis_cfs, case_types
from oracle import pc_oracle as orc

pytestmark = pytest.mark.gpu


def _pcrv(case):
    from pytuq_b200.rv.pcrv import PCRV
    mis, cfs = case_mis_cfs(case)
    return PCRV(len(mis), mis[0].shape[1], case_types(case), mi=[m.copy() for m in mis],
                cfs=[c.copy() for c in cfs])


def rel_err(a, b):
    scale = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) / scale


# ----------------------------------------------------------------------------- K1
def test_k1_evalbases_bitexact_vs_reference(pcrv_cases):
    for name, c in pcrv_cases.items():
        pc = _pcrv(c)
        for j in range(pc.pdim):
            A = pc.evalBases(c["x"], j)
            assert A.flags["C_CONTIGUOUS"] and A.dtype == np.float64
            assert np.array_equal(A, c[f"A{j}"]), (name, j)


def test_k1_pc1d_call_bitexact():
    from pytuq_b200.rv.pcrv import PC1d
    g = load_golden("pc1d")
    for fam in orc.FAMILIES:
        vals = PC1d(fam)(g[f"{fam}_x"], 7)
        assert len(vals) == 8 and np.array_equal(np.array(vals), g[f"{fam}_vals"]), fam
        assert np.array_equal(np.array(PC1d(fam)(g[f"{fam}_x"], 0)), g[f"{fam}_vals"][:1])
        assert np.array_equal(np.array(PC1d(fam)(g[f"{fam}_x"], 1)), g[f"{fam}_vals"][:2])


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 1000, 4099])
def test_k1_ragged_sizes_vs_oracle(n):
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(n)
    for (fam, p, d) in [("LU", 4, 10), ("HG", 3, 20), ("nLU", 2, 50), ("nHG", 5, 3)]:
        mi = orc.get_mi(p, d)
        x = rng.uniform(-1, 1, (n, d)) if fam.endswith("LU") else rng.standard_normal((n, d))
        nn = min(n, 64)   # the oracle's Python loop is slow: check head and tail rows
        A = PCRV(1, d, fam, mi=mi).evalBases(x, 0)
        assert np.array_equal(A[:nn], orc.eval_bases(x[:nn], mi, fam))
        assert np.array_equal(A[-nn:], orc.eval_bases(x[-nn:], mi, fam))


def test_k1_input_forms():
    """F-order, strided and integer inputs behave like the reference."""
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(3)
    mi = orc.get_mi(3, 4)
    pc = PCRV(1, 4, "LU", mi=mi)
    x = rng.uniform(-1, 1, (50, 4))
    ref = orc.eval_bases(x, mi, "LU")
    assert np.array_equal(pc.evalBases(np.asfortranarray(x), 0), ref)
    big = rng.uniform(-1, 1, (100, 8))
    big[::2, ::2] = x
    assert np.array_equal(pc.evalBases(big[::2, ::2], 0), ref)
    xi = rng.integers(-1, 2, (20, 4))
    assert np.array_equal(pc.evalBases(xi, 0), orc.eval_bases(xi.astype(float), mi, "LU"))
    with pytest.raises(AssertionError):
        pc.evalBases(x[:, :3], 0)
    with pytest.raises(AssertionError):
        pc.evalBases(x, 1)
    assert pc.evalBases(x[:0], 0).shape == (0, mi.shape[0])


def test_k1_stale_maxord_raises_indexerror():
    from pytuq_b200.rv.pcrv import PCRV
    pc = PCRV(1, 3, "LU", mi=orc.get_mi(2, 3))
    pc.setMiCfs(orc.get_mi(3, 3))          # grown after construction: reference raises IndexError
    with pytest.raises(IndexError):
        pc.evalBases(np.zeros((2, 3)), 0)
    pc.setMiCfs(orc.get_mi(1, 3))          # shrinking is fine
    assert pc.evalBases(np.zeros((2, 3)), 0).shape == (2, 4)


def test_k1_device_pointers_with_leading_dimensions():
    """_dev entry point, ldx > s and lda > K (also exercises the unaligned-row store path)."""
    import torch
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(5)
    for (p, d) in [(3, 5), (2, 6), (4, 3), (3, 9)]:
        mi = orc.get_mi(p, d)
        K = mi.shape[0]
        pc = PCRV(1, d, "HG", mi=mi)
        plan = pc._plan(mi)
        n, ldx, lda = 77, d + 3, K + 5
        xh = rng.standard_normal((n, ldx))
        xd = torch.from_numpy(xh).cuda()
        ad = torch.full((n, lda), -7.0, dtype=torch.float64, device="cuda")
        st = _native.lib().pcq_eval_bases_dev(plan.handle, xd.data_ptr(), n, ldx, ad.data_ptr(), lda,
                                              torch.cuda.current_stream().cuda_stream)
        _native.check(st)
        torch.cuda.synchronize()
        got = ad.cpu().numpy()
        assert np.array_equal(got[:, :K], orc.eval_bases(xh[:, :d].copy(), mi, "HG"))
        assert np.all(got[:, K:] == -7.0)
        # output shifted by one double (8-byte aligned only), even and odd leading dimensions, K > 64 so that
        # the shifted 16-byte stores cross lanes and chunk boundaries; canaries all around
        for extra in (0, 1):
            lda2 = K + 4 + extra
            flat = torch.full((n * lda2 + 3,), -7.0, dtype=torch.float64, device="cuda")
            base = flat[1:]
            _native.check(_native.lib().pcq_eval_bases_dev(plan.handle, xd.data_ptr(), n, ldx, base.data_ptr(), lda2,
                                                           torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            fh = flat.cpu().numpy()
            got2 = fh[1:1 + n * lda2].reshape(n, lda2)
            assert np.array_equal(got2[:, :K], got[:, :K]), (p, d, extra)
            assert np.all(got2[:, K:] == -7.0) and fh[0] == -7.0 and np.all(fh[1 + n * lda2:] == -7.0)


# ----------------------------------------------------------------------------- K2
def test_k2_evalpc_bitexact_vs_reference(pcrv_cases):
    for name, c in pcrv_cases.items():
        pc = _pcrv(c)
        y = pc.evalPC(c["x"])
        assert y.shape == c["y"].shape
        assert np.array_equal(y, c["y"]), name
        yf = pc.evalPC(c["x"], exact=False)
        assert rel_err(yf, c["y"]) < 1e-13, name
        pc.setFunction()
        assert np.array_equal(pc.function(c["x"]), c["y"]), name


def test_k2_many_outputs_and_sizes():
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(11)
    mi = orc.get_mi(3, 6)
    K = mi.shape[0]
    for M in (1, 3, 4, 9):
        cfs = rng.standard_normal((M, K))
        pc = PCRV(M, 6, "LU", mi=mi, cfs=cfs)
        for n in (1, 31, 129, 700):
            x = rng.uniform(-1, 1, (n, 6))
            ref = orc.eval_pc(x, [mi] * M, list(cfs), "LU")
            assert np.array_equal(pc.evalPC(x), ref), (M, n)
            assert rel_err(pc.evalPC(x, exact=False), ref) < 1e-13


# ----------------------------------------------------------------------------- K3 / K4
@pytest.mark.parametrize("fam,p,d,n,M", [("LU", 3, 5, 5000, 1), ("HG", 3, 5, 4097, 3), ("nLU", 2, 8, 333, 0),
                                         ("LU", 4, 10, 2500, 1), ("HG", 3, 20, 3000, 2), ("LU", 2, 50, 1500, 1),
                                         ("nHG", 4, 3, 63, 1), ("LU", 1, 2, 1, 1), ("LU", 3, 5, 777, 40),
                                         ("HG", 2, 6, 1234, 70), ("LU", 6, 1, 100, 1), ("LU", 0, 3, 50, 2)])
def test_k3_suffstats_vs_oracle(fam, p, d, n, M):
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(p * 100 + d)
    mi = orc.get_mi(p, d)
    x = rng.uniform(-1, 1, (n, d)) if fam.endswith("LU") else rng.standard_normal((n, d))
    Y = rng.standard_normal((n, M)) if M else None
    pc = PCRV(1, d, fam, mi=mi)
    st = pc.suffStats(x, Y, 0)
    A = pc.evalBases(x, 0)
    ref = orc.suffstats(A, Y if M else np.zeros((n, 0)))
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    assert np.max(np.abs(st.S - ref) / np.maximum(scale, 1e-300)) < 1e-12
    assert np.array_equal(st.S, st.S.T)
    assert st.n == n


def test_k3_general_multiindex_and_accumulate(pcrv_cases):
    import torch
    from pytuq_b200 import _native
    c = pcrv_cases["general_dense"]
    pc = _pcrv(c)
    x = np.tile(c["x"], (9, 1))
    y = np.arange(x.shape[0], dtype=float) / 7.0
    st = pc.suffStats(x, y, 0)
    ref = orc.suffstats(np.tile(c["A0"], (9, 1)), y)
    assert rel_err(st.S, ref) < 1e-12
    # accumulate flag of the device entry point: two halves add up to the whole
    plan = pc._plan(pc.mindices[0])
    L = int(_native.lib().pcq_suffstats_dim(plan.handle, 1))
    assert L == st.S.shape[0]
    xd = torch.from_numpy(x).cuda()
    yd = torch.from_numpy(y).cuda()
    sd = torch.empty((L, L), dtype=torch.float64, device="cuda")
    h = x.shape[0] // 2 + 3
    stream = torch.cuda.current_stream().cuda_stream
    _native.check(_native.lib().pcq_suffstats_dev(plan.handle, xd.data_ptr(), h, x.shape[1], yd.data_ptr(), 1, 1,
                                                  sd.data_ptr(), 0, stream))
    _native.check(_native.lib().pcq_suffstats_dev(plan.handle, xd[h:].data_ptr(), x.shape[0] - h, x.shape[1],
                                                  yd[h:].data_ptr(), 1, 1, sd.data_ptr(), 1, stream))
    torch.cuda.synchronize()
    assert rel_err(sd.cpu().numpy(), ref) < 1e-12


def test_empty_inputs():
    from pytuq_b200.rv.pcrv import PCRV
    mi = orc.get_mi(2, 3)
    pc = PCRV(2, 3, "LU", mi=mi, cfs=np.ones((2, mi.shape[0])))
    assert pc.evalPC(np.zeros((0, 3))).shape == (0, 2)
    st = pc.suffStats(np.zeros((0, 3)), np.zeros((0, 1)), 0)
    assert st.S.shape == (12, 12) and not st.S.any()


def test_k3_large_orders_and_dims_fall_back_cleanly():
    """Shapes outside the packed fast path: width > 4 (generic descriptors) and a table too large for
    the pipelined kernel's double buffers."""
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(21)
    mi = rng.integers(0, 3, size=(50, 7))          # up to 7 non-zero orders per term
    x = rng.uniform(-1, 1, (300, 7))
    y = rng.standard_normal(300)
    pc = PCRV(1, 7, "LU", mi=mi)
    st = pc.suffStats(x, y, 0)
    assert rel_err(st.S, orc.suffstats(orc.eval_bases(x, mi, "LU"), y)) < 1e-12
    d = 120                                         # 1 + 2*120 = 241 slots
    mi = orc.get_mi(2, d)[:400]
    x = rng.uniform(-1, 1, (200, d))
    pc = PCRV(1, d, "LU", mi=mi)
    A = pc.evalBases(x, 0)
    assert np.array_equal(A[:20], orc.eval_bases(x[:20], mi, "LU", np.full(d, 2)))
    st = pc.suffStats(x, None, 0)
    assert rel_err(st.S, orc.suffstats(A, np.zeros((200, 0)))) < 1e-12
    cf = rng.standard_normal(400)
    pc.setCfs(cf)
    assert rel_err(pc.evalPC(x)[:, 0], A @ cf) < 1e-12


def test_wide_germ_takes_the_global_table_route():
    """Germ dimensions whose slot tables do not fit shared memory (d = 1000 and up): tables in global memory,
    same arithmetic -- evalBases / evalPC stay bit-identical, statistics, fits and propagation work."""
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import lsq, anl, PCBasisEvaluator
    rng = np.random.default_rng(91)
    for fam, d, p, n, kmax in [("LU", 1000, 1, 700, None), ("HG", 1200, 2, 333, 1500), ("LU", 3000, 1, 150, 2000)]:
        mi = orc.get_mi(p, d)
        if kmax:
            keep = np.sort(rng.choice(mi.shape[0] - 1, kmax - 1, replace=False)) + 1
            mi = np.vstack([mi[:1], mi[keep]])
        K = mi.shape[0]
        x = rng.uniform(-1, 1, (n, d)) if fam == "LU" else rng.standard_normal((n, d))
        cf = rng.standard_normal(K) / np.sqrt(K)
        pc = PCRV(1, d, fam, mi=mi, cfs=[cf])
        A = pc.evalBases(x, 0)
        Aref = orc.eval_bases(x[:40], mi, fam, np.full(d, p))
        assert np.array_equal(A[:40], Aref), (fam, d, p)
        y = pc.evalPC(x)
        assert np.array_equal(y[:40], orc.eval_pc(x[:40], [mi], [cf], fam, np.full(d, p))), (fam, d, p)
        Y = (A @ cf + 0.01 * rng.standard_normal(n))
        st = pc.suffStats(x, Y, 0)
        assert rel_err(st.S, orc.suffstats(A, Y)) < 1e-12, (fam, d, p)
        v = pc.evalLinearForms(x, np.stack([cf, 2 * cf]), 0, sumsq=True)
        assert rel_err(v, 5 * (A @ cf) ** 2) < 1e-12
    # an overdetermined fit through lreg (statistics route) and a propagation, d = 1000
    d, n = 1000, 5000
    mi = orc.get_mi(1, d)
    x = rng.uniform(-1, 1, (n, d))
    ctrue = rng.standard_normal(d + 1)
    pc = PCRV(1, d, "LU", mi=mi)
    yv = orc.eval_bases(x, mi, "LU", np.full(d, 1)) @ ctrue
    f = lsq()
    f.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
    f.fit(x, yv)
    assert rel_err(f.cf, ctrue) < 1e-9
    m1, v1, _ = f.predict(x[:50], msc=1)
    assert rel_err(m1, yv[:50]) < 1e-9 and not v1.any()
    pc.setCfs([ctrue])
    res = pc.propagate(200000, seed=3)
    assert abs(res["mean"][0] - ctrue[0]) < 5 * np.sqrt(pc.computeVar()[0] / 200000)
    assert abs(res["var"][0] / pc.computeVar()[0] - 1) < 0.05


def test_wide_germ_across_host_chunks_on_both_plan_streams():
    """The global-table route when N spans several host chunks: the host-pointer entry points alternate their
    chunks between the plan's two streams, so each stream needs its own slot-table scratch (round-1 advisor
    finding: one shared table was overwritten by chunk c+1 while chunk c was still reading it).  Rows of every
    chunk, the last ones included, against the oracle."""
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(92)
    d = 1000
    mi = orc.get_mi(1, d)
    K = mi.shape[0]
    n_pc, n_a = 70_000, 26_000                # evalPC: 33520-row chunks; evalBases: 8380-row chunks
    x = rng.uniform(-1, 1, (n_pc, d))
    cf = rng.standard_normal(K) / np.sqrt(K)
    pc = PCRV(1, d, "LU", mi=mi, cfs=[cf])
    for _ in range(2):                        # twice: the second call starts with warm buffers
        A = pc.evalBases(x[:n_a], 0)
        rows = np.unique(np.concatenate([np.arange(0, n_a, 1733), np.arange(8370, 8390), np.arange(n_a - 48, n_a)]))
        assert np.array_equal(A[rows], orc.eval_bases(x[rows], mi, "LU", np.full(d, 1)))
        del A
        y = pc.evalPC(x)
        rows = np.unique(np.concatenate([np.arange(0, n_pc, 4999), np.arange(33500, 33540), np.arange(67030, 67050),
                                         np.arange(n_pc - 48, n_pc)]))
        assert np.array_equal(y[rows], orc.eval_pc(x[rows], [mi], [cf], "LU", np.full(d, 1)))


def test_k4_gram_of_materialised_matrix():
    from pytuq_b200.lreg.stats import SuffStats
    rng = np.random.default_rng(8)
    for (n, K, M) in [(1000, 37, 1), (513, 130, 2), (64, 300, 0), (5, 3, 1)]:
        A = rng.standard_normal((n, K))
        Y = rng.standard_normal((n, M)) if M else None
        st = SuffStats.from_matrix(A, Y)
        ref = orc.suffstats(A, Y if M else np.zeros((n, 0)))
        assert rel_err(st.S, ref) < 1e-12, (n, K, M)


def test_k3_k4_syrk_chunks_splits_and_fused_fallback(monkeypatch):
    """The pack + SYRK path under forced chunking (accumulation across chunk launches), forced numbers of
    sample parts (direct flush vs workspace + fix-up), and the fused-generator path (PCQ_GRAM_SYRK=0):
    all must agree with the oracle."""
    import torch
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(77)
    d, p, n, M = 6, 3, 9001, 2            # K = 84 -> one tile; n is not a multiple of 8/16/32
    mi = orc.get_mi(p, d)
    x = rng.standard_normal((n, d))
    Y = rng.standard_normal((n, M))
    pc = PCRV(1, d, "HG", mi=mi)
    ref = orc.suffstats(orc.eval_bases(x, mi, "HG"), Y)
    d2, n2 = 12, 2111                      # K = 455 -> 4 tiles (10 upper-triangle tiles)
    mi2 = orc.get_mi(3, d2)
    x2 = rng.uniform(-1, 1, (n2, d2))
    y2 = rng.standard_normal(n2)
    pc2 = PCRV(1, d2, "LU", mi=mi2)
    ref2 = orc.suffstats(orc.eval_bases(x2, mi2, "LU"), y2)
    for env in ({}, {"PCQ_SYRK_CHUNK_MB": "1"}, {"PCQ_SYRK_SPLIT": "1"}, {"PCQ_SYRK_SPLIT": "5"},
                {"PCQ_SYRK_CHUNK_MB": "1", "PCQ_SYRK_SPLIT": "3"}, {"PCQ_SYRK_TILE": "96"}, {"PCQ_SYRK_DIAG": "0"},
                {"PCQ_SYRK_DIAG": "0", "PCQ_SYRK_SPLIT": "2"},
                {"PCQ_SYRK_TILE": "96", "PCQ_SYRK_SPLIT": "4", "PCQ_SYRK_CHUNK_MB": "1"}, {"PCQ_GRAM_SYRK": "0"},
                {"PCQ_GRAM_SYRK": "0", "PCQ_GRAM_V1": "1"}):
        for k in ("PCQ_SYRK_CHUNK_MB", "PCQ_SYRK_SPLIT", "PCQ_SYRK_TILE", "PCQ_SYRK_DIAG", "PCQ_GRAM_SYRK", "PCQ_GRAM_V1"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        st = pc.suffStats(x, Y, 0)
        assert rel_err(st.S, ref) < 1e-12, env
        assert np.array_equal(st.S, st.S.T), env
        st2 = pc2.suffStats(x2, y2, 0)
        assert rel_err(st2.S, ref2) < 1e-12, env
        # K4 from a device matrix with an odd leading dimension and an offset y column
        A = rng.standard_normal((700, 131))
        buf = torch.zeros((700, 137), dtype=torch.float64, device="cuda")
        buf[:, :131] = torch.from_numpy(A).cuda()
        yb = torch.from_numpy(rng.standard_normal((700, 3))).cuda()
        L = 131 + 1 + 1
        sd = torch.full((L, L), np.nan, dtype=torch.float64, device="cuda")
        _native.check(_native.lib().pcq_gram_dev(0, buf.data_ptr(), 700, 137, 131, yb[:, 1:].data_ptr(), 3, 1,
                                                 sd.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert rel_err(sd.cpu().numpy(), orc.suffstats(A, yb[:, 1].cpu().numpy())) < 1e-12, env


def test_k2_fast_mode_many_outputs_on_the_tensor_pipe():
    """evalPC fast mode with >= 64 outputs runs as a packed GEMM on DMMA (pcq_qform_kernel<true>): same values as
    the design-matrix product to rounding, ragged N / K / M, both entry points."""
    import torch
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(8)
    for fam, p, d, n, M in [("LU", 3, 6, 3000, 64), ("HG", 3, 9, 2047, 130), ("LU", 4, 5, 1500, 257)]:
        mi = orc.get_mi(p, d)
        K = mi.shape[0]
        x = rng.uniform(-1, 1, (n, d)) if fam == "LU" else rng.standard_normal((n, d))
        cf = rng.standard_normal((M, K))
        pc = PCRV(1, d, fam, mi=mi)
        A = orc.eval_bases(x, mi, fam)
        ref = A @ cf.T
        got = pc.evalLinearForms(x, cf, 0, exact=False)
        assert got.shape == (n, M) and rel_err(got, ref) < 1e-12, (fam, p, d, n, M)
        # device entry point with a leading dimension and canary columns
        plan = pc._plan(mi)
        xd = torch.from_numpy(x).cuda(); cd = torch.from_numpy(cf).cuda()
        yd = torch.full((n, M + 3), -7.0, dtype=torch.float64, device="cuda")
        _native.check(_native.lib().pcq_eval_pc_dev(plan.handle, xd.data_ptr(), n, d, cd.data_ptr(), M, yd.data_ptr(),
                                                    M + 3, 1, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        yh = yd.cpu().numpy()
        assert rel_err(yh[:, :M], ref) < 1e-12 and np.all(yh[:, M:] == -7.0)
        # exact mode is untouched by the route: still bit-identical
        assert np.array_equal(pc.evalLinearForms(x[:50], cf[:3], 0, exact=True),
                              orc.eval_pc(x[:50], [mi] * 3, list(cf[:3]), fam))


# ----------------------------------------------------------------------------- K6
def test_k6_quadratic_forms_on_the_tensor_pipe():
    """v_n = psi_n^T C psi_n (pcq_quadform_dev) against the design-matrix formula, over shapes that exercise
    the padding of terms (K not a multiple of 128) and samples (N not a multiple of 128 / 16), generic
    descriptor widths, several column tiles, chunking, and lreg.predict's use of it."""
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import anl, PCBasisEvaluator
    rng = np.random.default_rng(66)
    for fam, p, d, n in [("LU", 3, 6, 1000), ("HG", 3, 12, 777), ("LU", 2, 3, 5), ("LU", 4, 7, 2049)]:
        mi = orc.get_mi(p, d)
        K = mi.shape[0]
        x = rng.uniform(-1, 1, (n, d)) if fam == "LU" else rng.standard_normal((n, d))
        B = rng.standard_normal((K, K))
        Cm = B @ B.T / K
        pc = PCRV(1, d, fam, mi=mi)
        A = orc.eval_bases(x, mi, fam)
        ref = np.einsum("ni,ij,nj->n", A, Cm, A)
        v = pc.evalQuadForm(x, Cm, 0)
        assert np.max(np.abs(v - ref)) < 1e-12 * np.max(np.abs(ref)), (fam, p, d, n)
    mi = rng.integers(0, 3, size=(150, 6))                     # generic width (up to 6 non-zero orders)
    x = rng.uniform(-1, 1, (300, 6))
    Cm = np.diag(rng.uniform(0.5, 2.0, 150))
    pc = PCRV(1, 6, "LU", mi=mi)
    A = orc.eval_bases(x, mi, "LU")
    assert rel_err(pc.evalQuadForm(x, Cm, 0), np.einsum("ni,ij,nj->n", A, Cm, A)) < 1e-12
    # through lreg.predict (anl, full covariance): same mean / variance as predicta on the materialised matrix
    d, p, n = 8, 3, 9000
    mi = orc.get_mi(p, d)
    x = rng.uniform(-1, 1, (n, d))
    pc = PCRV(1, d, "LU", mi=mi)
    A = orc.eval_bases(x, mi, "LU")
    y = A @ (rng.standard_normal(mi.shape[0]) / (1 + mi.sum(1))) + 0.05 * rng.standard_normal(n)
    a = anl()
    a.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
    a.fit(x, y)
    assert float(n) * mi.shape[0] ** 2 >= 2.0e8            # takes the K6 route
    m1, v1, _ = a.predict(x, msc=1)
    m2, v2, _ = a.predicta(A, msc=1)
    assert rel_err(m1, m2) < 1e-12 and rel_err(v1, v2) < 1e-9


# ----------------------------------------------------------------------------- K5
def test_k5_germ_stream_and_propagation():
    import torch
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    mi = orc.get_mi(3, 5)
    types = ["LU", "HG", "HG", "nLU", "nHG"]
    rng = np.random.default_rng(2)
    cfs = rng.standard_normal((2, mi.shape[0]))
    pc = PCRV(2, 5, types, mi=mi, cfs=cfs)
    kinds = pc._germ_kinds()
    assert kinds.tolist() == [0, 1, 1, 0, 1]
    n, seed, first = 4001, 12345, 1000
    plan = pc._plan(mi)
    xd = torch.empty((n, 5), dtype=torch.float64, device="cuda")
    _native.check(_native.lib().pcq_sample_germ_dev(plan.handle, seed, first, n, _native.ptr(kinds),
                                                    xd.data_ptr(), 5, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    x = xd.cpu().numpy()
    ref = orc.germ_philox(seed, first, n, kinds)
    assert np.array_equal(x[:, [0, 3]], ref[:, [0, 3]])          # uniform dims: exact
    assert np.max(np.abs(x - ref)) < 1e-13                          # normal dims: libm-level
    out = pc.propagate(n, seed=seed, exact=True, return_samples=True, first_index=first)
    yref = orc.eval_pc(x, [mi, mi], list(cfs), types)
    assert np.array_equal(out["samples"], pc.evalPC(x))
    assert rel_err(out["samples"], yref) < 1e-13
    np.testing.assert_allclose(out["sum"], yref.sum(axis=0), rtol=1e-12)
    np.testing.assert_allclose(out["sumsq"], (yref ** 2).sum(axis=0), rtol=1e-12)
    assert np.array_equal(out["min"], out["samples"].min(axis=0))
    assert np.array_equal(out["max"], out["samples"].max(axis=0))
    fast = pc.propagate(n, seed=seed, first_index=first)
    np.testing.assert_allclose(fast["mean"], out["mean"], rtol=1e-11, atol=1e-13)


def test_k5_host_pointer_propagate_matches_device_form():
    """pcq_propagate (host coefficients in, host sums out) draws the same Philox stream as pcq_propagate_dev."""
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(4)
    mi = orc.get_mi(3, 5)
    cfs = rng.standard_normal((2, mi.shape[0]))
    pc = PCRV(2, 5, "HG", mi=mi, cfs=[cfs[0], cfs[1]])
    ref = pc.propagate(70001, seed=11)
    kinds = np.ones(5, dtype=np.int32)
    sums = np.zeros((2, 4))
    plan = pc._plan(mi)
    _native.check(_native.lib().pcq_propagate(plan.handle, 11, 0, 70001, _native.ptr(kinds),
                                              _native.ptr(np.ascontiguousarray(cfs)), 2, _native.ptr(sums), 1))
    assert np.allclose(sums[:, 0] / 70001, ref["mean"], rtol=1e-13, atol=0)
    assert np.array_equal(sums[:, 2], ref["min"]) and np.array_equal(sums[:, 3], ref["max"])


def test_k2_specialised_kernels_bitexact():
    """Opt-in NVRTC-specialised evalPC: bit-identical to the generic kernel (and so to the reference) in
    exact mode, rounding-level in fast mode, and consistent through propagate."""
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(99)
    for (types, p, d, n) in [("HG", 3, 6, 3001), (["LU", "HG", "nLU", "nHG"], 4, 4, 777), ("LU", 2, 30, 257)]:
        mi = orc.get_mi(p, d)                      # (2, 30): 496 terms; (3, 6): 84; all < 512 -> one chunk
        if d == 6:
            mi = np.vstack([mi] * 8)[: 600]       # 600 terms (duplicates allowed) -> two chunks
        K = mi.shape[0]
        cfs = rng.standard_normal((2, K))
        tlist = [types] * d if isinstance(types, str) else types
        x = np.stack([rng.uniform(-1, 1, n) if t.endswith("LU") else rng.standard_normal(n) for t in tlist], axis=1)
        gen = PCRV(2, d, types, mi=mi, cfs=cfs)
        y_exact, y_fast = gen.evalPC(x), gen.evalPC(x, exact=False)
        spec = PCRV(2, d, types, mi=mi.copy(), cfs=cfs)     # a different plan (content key differs by identity? no: same content)
        if not spec.specialize():
            pytest.skip("NVRTC not available on this box")
        from pytuq_b200 import _native
        assert _native.lib().pcq_plan_info(spec._plan(spec.mindices[0]).handle, 7) == 3
        ys = spec.evalPC(x)
        assert np.array_equal(ys, y_exact), (types, p, d)
        assert np.array_equal(ys, orc.eval_pc(x, [mi, mi], list(cfs), types)), (types, p, d)
        assert rel_err(spec.evalPC(x, exact=False), y_fast) < 1e-13
        a = spec.propagate(5000, seed=3, exact=True, return_samples=True)
        assert np.array_equal(a["samples"].shape, (5000, 2))
        np.testing.assert_allclose(a["sum"], a["samples"].sum(axis=0), rtol=1e-12)
        np.testing.assert_allclose(a["sumsq"], (a["samples"] ** 2).sum(axis=0), rtol=1e-12)
        assert np.array_equal(a["min"], a["samples"].min(axis=0)) and np.array_equal(a["max"], a["samples"].max(axis=0))
        b = spec.propagate(5000, seed=3, exact=False)
        np.testing.assert_allclose(b["mean"], a["mean"], rtol=1e-11, atol=1e-13)


def test_k2_fast_specialised_kernels_in_horner_form():
    """Fast-mode specialised evalPC = Horner form over the multi-index trie (csrc/pcq_spec.cu:gen_source_horner): one
    DFMA per term, terms re-ordered -- values differ from the reference's summation order by rounding only
    (rv/pcrv.py:486-497).  Checked against the bit-exact mode on downward-closed, sparse (not downward closed),
    duplicated and multi-chunk multi-indices, with two outputs across several host chunks (both plan streams)."""
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200 import _native
    rng = np.random.default_rng(123)
    cases = [("HG", orc.get_mi(3, 20), 5000),                                   # C2 shape: 1771 terms, 2 chunks
             ("LU", orc.get_mi(4, 12)[np.sort(rng.choice(1820, 700, replace=False))], 3000),   # sparse subset, no constant term guaranteed
             (["LU", "HG", "nLU", "nHG", "LU"], np.vstack([orc.get_mi(3, 5)] * 2), 1000),     # duplicates, mixed families
             ("HG", orc.get_mi(4, 20), 1500)]                                   # C4 shape: 10626 terms, 11 chunks
    for types, mi, n in cases:
        d = mi.shape[1]
        K = mi.shape[0]
        cfs = rng.standard_normal((2, K)) / (1.0 + mi.sum(axis=1)) ** 2
        tlist = [types] * d if isinstance(types, str) else types
        x = np.stack([rng.uniform(-1, 1, n) if t.endswith("LU") else rng.standard_normal(n) for t in tlist], axis=1)
        pc = PCRV(2, d, types, mi=mi, cfs=cfs)
        y_exact = pc.evalPC(x)
        if not pc.specialize(exact=False, fast=True):
            pytest.skip("NVRTC not available on this box")
        assert _native.lib().pcq_plan_info(pc._plan(pc.mindices[0]).handle, 7) & 2
        y_fast = pc.evalPC(x, exact=False)
        scale = np.abs(cfs).sum(axis=1) * 0 + np.max(np.abs(y_exact), axis=0)
        assert np.max(np.abs(y_fast - y_exact) / scale) < 1e-12, (types, K)
        assert np.array_equal(pc.evalPC(x), y_exact)                            # exact mode untouched
    # two outputs, N spanning three host chunks of the host-pointer route (alternating plan streams)
    mi = orc.get_mi(3, 20)
    cfs = rng.standard_normal((2, 1771)) / (1.0 + mi.sum(axis=1)) ** 2
    pc = PCRV(2, 20, "HG", mi=mi, cfs=cfs)
    pc.specialize(exact=False, fast=True)
    n = 3_600_000
    x = rng.standard_normal((n, 20))
    y = pc.evalPC(x, exact=False)
    rows = np.unique(np.concatenate([np.arange(0, n, 100_003), np.arange(1_597_000, 1_599_000, 7), np.arange(n - 100, n)]))
    ref = orc.eval_pc(x[rows], [mi, mi], list(cfs), "HG")
    assert np.max(np.abs(y[rows] - ref) / np.max(np.abs(ref), axis=0)) < 1e-12


# ----------------------------------------------------------------------------- fitters and workflows
def test_k2_specialised_cubins_are_cached_on_disk(monkeypatch, tmp_path):
    """Second build of the same plan (new process in real life; here: plan cache cleared) loads the cubins from
    PCQ_SPEC_CACHE_DIR instead of compiling, and evaluates to the same bits."""
    import time
    from pytuq_b200 import device as _device
    from pytuq_b200.rv.pcrv import PCRV
    monkeypatch.setenv("PCQ_SPEC_CACHE_DIR", str(tmp_path / "k2s"))
    rng = np.random.default_rng(13)
    mi = orc.get_mi(3, 9)[:211]
    cf = rng.standard_normal(211)
    x = rng.standard_normal((500, 9))
    times, ys = [], []
    for rep in range(2):
        _device.clear_plans()
        pc = PCRV(1, 9, "HG", mi=mi, cfs=[cf])
        t0 = time.perf_counter()
        if not pc.specialize(exact=True, fast=False):
            pytest.skip("NVRTC not available on this box")
        times.append(time.perf_counter() - t0)
        ys.append(pc.evalPC(x))
        if rep == 0:
            files = sorted(p.name for p in (tmp_path / "k2s").iterdir())
            assert len(files) == 1 and files[0].endswith(".cubin")
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], orc.eval_pc(x, [mi], [cf], "HG"))
    assert times[1] < 0.5 * times[0], times
    _device.clear_plans()


def test_k2_hot_plan_is_specialised_in_the_background(monkeypatch):
    """A plan that keeps being evaluated gets its specialised kernels built on a background thread once the
    accumulated work passes the threshold; evaluations never block on the compiler and stay bit-identical
    across the switch."""
    import time
    from pytuq_b200 import device as _device
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(12)
    mi = orc.get_mi(3, 7)[:97]            # a plan no other test uses
    pc = PCRV(1, 7, "LU", mi=mi, cfs=[rng.standard_normal(97)])
    x = rng.uniform(-1, 1, (3000, 7))
    monkeypatch.setattr(_device, "_AUTO_SPECIALIZE", 2.5 * 3000 * 97)
    plan = pc._plan(pc.mindices[0])
    y1 = pc.evalPC(x)
    y2 = pc.evalPC(x)
    assert plan.info(7) == 0 and plan.info(8) == 0
    t0 = time.perf_counter()
    y3 = pc.evalPC(x)                     # crosses the threshold: build starts, this call does not wait for it
    dt = time.perf_counter() - t0
    ys = [y1, y2, y3]
    for _ in range(400):                  # keep evaluating while the build runs; at some point the kernels switch
        ys.append(pc.evalPC(x))
        if plan.info(7) == 1:
            break
        time.sleep(0.01)
    if plan.info(7) == 0 and plan.info(8) == 0:
        pytest.skip("NVRTC not available on this box")
    assert plan.info(7) == 1 and plan.info(8) == 0
    assert dt < 0.2, dt
    ys.append(pc.evalPC(x))               # certainly through the specialised kernel
    ref = orc.eval_pc(x, [mi], [pc.coefs[0]], "LU")
    assert all(np.array_equal(y, ref) for y in ys)
    _device.clear_plans()                 # destroying the plan joins the (finished) build thread


def test_fitters_vs_reference(fit_cases):
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import lsq, anl, bcs, PCBasisEvaluator
    for name, c in fit_cases.items():
        fam = str(c["type"])
        mi = c["mi"]
        pc = PCRV(1, mi.shape[1], fam, mi=mi)
        x, y = c["x"], c["y"]

        def fused(fitter):
            fitter.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
            fitter.fit(x, y)
            return fitter

        A = pc.evalBases(x, 0)
        for make in (lambda f: fused(f), lambda f: (f.fita(A, y), f)[1]):
            l = make(lsq())
            np.testing.assert_allclose(l.cf, c["lsq_cf"], rtol=1e-10, atol=1e-12)
            assert np.array_equal(l.used, np.arange(mi.shape[0])) and not l.cf_cov.any()
            a = make(anl())
            np.testing.assert_allclose(a.cf, c["anl_cf"], rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(a.datavar, c["anl_datavar"], rtol=1e-8)
            np.testing.assert_allclose(a.cf_cov, c["anl_cov"], rtol=1e-8, atol=1e-16)
            ap = make(anl(datavar=0.01, prior_var=2.0, cov_nugget=1e-8))
            np.testing.assert_allclose(ap.cf, c["anlp_cf"], rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(ap.cf_cov, c["anlp_cov"], rtol=1e-9, atol=1e-18)
            av = make(anl(method="vi"))
            np.testing.assert_allclose(np.diag(av.cf_cov), c["anlvi_cov_diag"], rtol=1e-8)
            for tag in ("bcs0.001", "bcs1e-08"):
                b = make(bcs(eta=float(tag[3:])))
                assert np.array_equal(b.used, c[f"{tag}_used"]), (name, tag)
                np.testing.assert_allclose(b.cf, c[f"{tag}_cf"], rtol=1e-10, atol=1e-12)
                assert np.shape(b.datavar) == (1,)
                np.testing.assert_allclose(b.datavar, c[f"{tag}_datavar"], rtol=1e-7)
                np.testing.assert_allclose(b.cf_cov, c[f"{tag}_cov"], rtol=1e-7, atol=1e-18)
        a = fused(anl())
        mean, var, _ = a.predict(c["xt"], msc=1)
        np.testing.assert_allclose(mean, c["anl_pred_mean"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(var, c["anl_pred_var"], rtol=1e-7)
        # fused predict (no design matrix) agrees with predicta on the evaluated basis, incl. pp
        At = pc.evalBases(c["xt"], 0)
        m2, v2, _ = a.predicta(At, msc=1, pp=True)
        m1, v1, _ = a.predict(c["xt"], msc=1, pp=True)
        np.testing.assert_allclose(m1, m2, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(v1, v2, rtol=1e-9)
        l = fused(lsq())
        m0, v0, c0 = l.predict(c["xt"], msc=1)
        assert not v0.any() and c0 is None
        np.testing.assert_allclose(m0, At @ l.cf, rtol=1e-12, atol=1e-13)
        assert l.predict(c["xt"])[1] is None


def test_predictions_on_a_given_design_matrix_run_on_the_device(fit_cases):
    """lreg.predicta / compute_stdev / predicta_samples on a caller-supplied Amat (lreg/lreg.py:112-198; the call
    pattern of apps/uqpc/uq_pc.py:313-315) and the reference idiom of a plain function as basis evaluator
    (examples/ex_lreg_basiseval.py:55, surrogates/pce.py:378-382): contracted on the device, compared with the
    reference's numpy expressions evaluated here on the oracle's design matrix."""
    import torch
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import anl, lsq, PCBasisEvaluator
    from pytuq_b200 import _native
    name, c = sorted(fit_cases.items())[0]
    fam, mi = str(c["type"]), c["mi"]
    pc = PCRV(1, mi.shape[1], fam, mi=mi)
    a = anl()
    a.setBasisEvaluator(lambda xx, pars: pars[0].evalBases(xx, 0), (pc,))      # plain closure: recognised
    n0 = _native.lib().pcq_launch_count()
    a.fit(c["x"], c["y"])
    assert a._pc_route() == (pc, 0)
    np.testing.assert_allclose(a.cf, c["anl_cf"], rtol=1e-10, atol=1e-12)
    mean, var, _ = a.predict(c["xt"], msc=1)
    np.testing.assert_allclose(mean, c["anl_pred_mean"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(var, c["anl_pred_var"], rtol=1e-7)
    # an evaluator that is NOT the PC design matrix must not be mistaken for one
    b = anl()
    b.setBasisEvaluator(lambda xx, pars: 2.0 * pars[0].evalBases(xx, 0), (pc,))
    assert b._pc_route() is None
    b.fit(c["x"], c["y"])
    np.testing.assert_allclose(2.0 * b.cf, c["anl_cf"], rtol=1e-9, atol=1e-12)

    At = orc.eval_bases(c["xt"], mi, fam)
    C, cf = a.cf_cov, a.cf
    Lc = np.linalg.cholesky(C)
    for msc in (0, 1, 2):
        for pp in (False, True):
            m, v, cov = a.predicta(At, msc=msc, pp=pp)
            np.testing.assert_allclose(m, At @ cf, rtol=1e-12, atol=1e-13)
            if msc == 0:
                assert v is None and cov is None
            elif msc == 1:
                np.testing.assert_allclose(v, np.linalg.norm(At @ Lc, axis=1) ** 2 + int(pp) * a.datavar, rtol=1e-10)
                assert cov is None
            else:
                ref = (At @ C) @ At.T + int(pp) * a.datavar * np.eye(At.shape[0])
                np.testing.assert_allclose(cov, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())
                np.testing.assert_allclose(v, np.diag(ref), rtol=1e-10)
    for method in ("chol", "choleye", "svd", "loop", "fullcov"):
        Cm = C + (abs(np.linalg.eigvalsh(C)[0]) + 1e-14) * np.eye(C.shape[0]) if method == "choleye" else C
        np.testing.assert_allclose(a.compute_stdev(At, method=method), np.sqrt(np.einsum("ij,ij->i", At @ Cm, At)),
                                   rtol=1e-7, err_msg=method)
    assert not a.compute_stdev(At, method="none").any()
    # a covariance that is not positive definite: 'chol' raises like numpy, predicta takes the reference's SVD route
    Cbad = C.copy(); Cbad[0, 0] = -abs(Cbad[0, 0])
    a2 = anl(); a2.cf, a2.cf_cov, a2.fitted, a2.datavar = cf, Cbad, True, 0.0
    with pytest.raises(np.linalg.LinAlgError):
        a2.compute_stdev(At, method="chol")
    u, sv, _ = np.linalg.svd(Cbad, hermitian=True)
    np.testing.assert_allclose(a2.predicta(At, msc=1)[1], np.linalg.norm((At @ u) @ np.sqrt(np.diag(sv)), axis=1) ** 2, rtol=1e-9)
    # predictive samples: numpy's stream for the coefficient draws, the product on the device
    np.random.seed(5); s_dev = a.predicta_samples(At, nsamples=70)
    np.random.seed(5); s_ref = At @ np.random.multivariate_normal(cf, C, 70).T
    np.testing.assert_allclose(s_dev, s_ref, rtol=1e-10, atol=1e-12)
    np.random.seed(6); s_pc = a.predict_samples(c["xt"], nsamples=5)
    np.random.seed(6); s_ref = At @ np.random.multivariate_normal(cf, C, 5).T
    np.testing.assert_allclose(s_pc, s_ref, rtol=1e-10, atol=1e-12)
    # larger, odd-shaped problem incl. leading dimension, many vectors (tensor-pipe route) and a device-resident matrix
    rng = np.random.default_rng(8)
    N, K, M = 3001, 203, 131
    Abig = rng.standard_normal((N, K + 3))[:, :K]
    Cb = rng.standard_normal((K, K)); Cb = Cb @ Cb.T / K
    cfm = rng.standard_normal((M, K))
    from pytuq_b200.lreg.forms import matrix_forms
    Y, v = matrix_forms(Abig, cfm, Cb)
    np.testing.assert_allclose(Y, Abig @ cfm.T, rtol=0, atol=1e-12 * np.abs(Abig @ cfm.T).max())
    np.testing.assert_allclose(v, np.einsum("ij,ij->i", Abig @ Cb, Abig), rtol=1e-11)
    Yd, vd = matrix_forms(torch.from_numpy(np.ascontiguousarray(Abig)).cuda(), cfm[:3], Cb)
    np.testing.assert_allclose(Yd.cpu().numpy(), Abig @ cfm[:3].T, rtol=0, atol=1e-12 * np.abs(Abig @ cfm.T).max())
    np.testing.assert_allclose(vd.cpu().numpy(), v, rtol=1e-13)
    assert _native.lib().pcq_launch_count() > n0
    m0, v0, _ = lsq_zero_cov_predicta(At, cf)
    assert not v0.any()
    np.testing.assert_allclose(m0, At @ cf, rtol=1e-12, atol=1e-13)


def lsq_zero_cov_predicta(At, cf):
    from pytuq_b200.lreg import lsq
    l = lsq(); l.cf, l.cf_cov, l.fitted = cf, np.zeros((cf.shape[0],) * 2), True
    return l.predicta(At, msc=1)


def test_householder_qr_kernels_give_the_r_factor():
    """pcq_qr_r_dev (csrc/pcq_qr.cu): hand-written blocked Householder QR.  R^T R = Z^T Z, |R| equals LAPACK's R up
    to row signs, the singular values of R are those of Z even when Z is ill-conditioned or rank deficient (the
    reason this route exists: lreg/lreg.py:335), for tall / square / wide shapes, ragged panel widths, rows beyond
    one cooperative launch, leading dimensions, and the tree that combines chunk triangles."""
    import torch
    from pytuq_b200 import _native
    from pytuq_b200.lreg import tsqr
    lib = _native.lib()
    dev = torch.device("cuda", 0)
    cap = int(lib.pcq_qr_max_rows(0))
    assert cap >= 256
    rng = np.random.default_rng(2024)
    n0 = lib.pcq_launch_count()
    for (m, L, kind) in [(1, 1, "rand"), (5, 9, "rand"), (40, 40, "rand"), (70, 33, "rand"), (300, 64, "rand"),
                         (1000, 97, "rand"), (5000, 130, "graded"), (3000, 200, "rankdef"), (cap + 1000, 70, "rand"),
                         (20000, 333, "graded")]:
        Z = rng.standard_normal((m, L))
        if kind == "graded":                       # condition number ~1e12
            Z = Z * np.logspace(0, -12, L)[None, :]
            Z[:, 1::7] += Z[:, 0:1]
        if kind == "rankdef":
            Z[:, 5] = Z[:, 3]
            Z[:, L - 1] = 2.0 * Z[:, 0] - Z[:, 1]
        Zt = torch.from_numpy(Z).to(dev)
        pad = torch.full((m, L + 3), 7.0, dtype=torch.float64, device=dev)      # a leading dimension > L
        pad[:, :L] = Zt
        for src in (Zt.clone(), pad[:, :L]):
            R = tsqr._qr_r(src).cpu().numpy()
            assert R.shape == (L, L) and np.array_equal(R, np.triu(R))
            G = Z.T @ Z
            sc = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
            assert np.max(np.abs(R.T @ R - G) / sc) < 1e-13, (m, L, kind)
            sv = np.linalg.svd(Z, compute_uv=False)
            svr = np.linalg.svd(R, compute_uv=False)
            np.testing.assert_allclose(svr[:len(sv)], sv, rtol=0, atol=1e-13 * sv[0], err_msg=str((m, L, kind)))
            if kind == "rand" and m >= L:
                Rref = np.linalg.qr(Z, mode="r")
                np.testing.assert_allclose(np.abs(R), np.abs(Rref), rtol=1e-9, atol=1e-12 * np.abs(Rref).max())
            if m < L:
                assert not R[m:].any()
        assert torch.equal(pad[:, L:], torch.full((m, 3), 7.0, dtype=torch.float64, device=dev))   # nothing outside
    # chunked factorisation through the balanced tree equals the one-shot factor (up to signs)
    Z = rng.standard_normal((9000, 60))
    chunks = [torch.from_numpy(Z[i:i + 1300]).to(dev) for i in range(0, 9000, 1300)]       # 7 triangles: tree with byes
    Rt = tsqr.combine_r([tsqr._qr_r(c) for c in chunks]).cpu().numpy()
    R1 = tsqr._qr_r(torch.from_numpy(Z).to(dev)).cpu().numpy()
    np.testing.assert_allclose(np.abs(Rt), np.abs(R1), rtol=1e-10, atol=1e-12)
    tree = tsqr._Tree()
    for c in [torch.from_numpy(Z[i:i + 1300]).to(dev) for i in range(0, 9000, 1300)]:
        tree.push(tsqr._qr_r(c))
    np.testing.assert_allclose(np.abs(tree.result().cpu().numpy()), np.abs(R1), rtol=1e-10, atol=1e-12)
    assert lib.pcq_launch_count() > n0


def test_jacobi_svd_kernels():
    """pcq_svd_jacobi_dev (csrc/pcq_svd.cu): one-sided Jacobi SVD of the small factor of the QR route.  Singular values
    against LAPACK to high RELATIVE accuracy on well-conditioned, graded (cond 1e14), rank-deficient and odd-sized
    matrices; A = U S V^T reconstructed; V orthogonal."""
    import torch
    from pytuq_b200.lreg import tsqr
    rng = np.random.default_rng(77)
    for n, kind in [(1, "rand"), (2, "rand"), (7, "rand"), (64, "rand"), (131, "graded"), (200, "rankdef"), (333, "triu"), (600, "rand")]:
        A = rng.standard_normal((n, n))
        if kind == "graded":
            A = A * np.logspace(0, -14, n)[None, :]
        if kind == "rankdef":
            A[:, 5] = A[:, 3]; A[:, 17] = 0.0
        if kind == "triu":
            A = np.triu(A) * np.logspace(0, -6, n)[None, :]
        Bt, Vt, sig, sweeps = tsqr.svd_jacobi(torch.from_numpy(A).cuda())
        Bt, Vt, sig = Bt.cpu().numpy(), Vt.cpu().numpy(), sig.cpu().numpy()
        assert 0 <= sweeps < 60 or n == 1
        ref = np.linalg.svd(A, compute_uv=False)
        got = np.sort(sig)[::-1]
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13 * ref[0], err_msg=kind)
        if kind == "graded":          # column scaling: Jacobi keeps the small ones accurate relative to THEMSELVES
            big = ref > 1e-13 * ref[0]
            np.testing.assert_allclose(got[big], ref[big], rtol=1e-9)
        np.testing.assert_allclose(Vt @ Vt.T, np.eye(n), rtol=0, atol=1e-12)
        np.testing.assert_allclose(Bt.T @ Vt, A, rtol=0, atol=1e-12 * max(1.0, np.abs(A).max()))      # A = (U S) V^T
        # the columns of U S are orthogonal
        Gm = Bt @ Bt.T
        off = Gm - np.diag(np.diag(Gm))
        assert np.max(np.abs(off)) <= 1e-12 * max(np.max(np.diag(Gm)), 1e-300), kind


def test_lsq_qr_route_on_illconditioned_systems():
    """8(f2): where the normal equations cannot follow the reference's SVD least squares, lsq switches to the
    QR statistic.  Golden coefficients come from the unmodified reference (tests/golden/lsq_illcond.npz)."""
    from conftest import load_golden, golden_cases
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import lsq, PCBasisEvaluator
    for name, c in golden_cases(load_golden("lsq_illcond")).items():
        fam, mi, x, y = str(c["type"]), c["mi"], c["x"], c["y"]
        pc = PCRV(1, x.shape[1], fam, mi=mi)
        A = pc.evalBases(x, 0)
        ref = c["lsq_cf"]
        for route in ("fit", "fita"):
            f = lsq()
            if route == "fit":
                f.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
                f.fit(x, y)
            else:
                f.fita(A, y)
            scale = np.abs(y).max()
            # same fitted values as the reference (the minimum-norm solution is what makes them unique)
            assert np.max(np.abs(A @ f.cf - A @ ref)) < 1e-7 * scale, (name, route)
            if name == "collinear_hg_d3p5":      # cond(A) = 2.5e8: forward error of any stable solver ~ cond * eps
                assert rel_err(f.cf, ref) < 1e-4, (name, route)
            else:
                assert rel_err(f.cf, ref) < 1e-8, (name, route, rel_err(f.cf, ref))
            expect_rank = {"rankdef_lu_d3p2": 10, "under_lu_d4p4": 40}.get(name, mi.shape[0])
            assert f.rank == expect_rank, (name, route, f.rank)
            assert f.cf_cov.shape == (mi.shape[0],) * 2 and not f.cf_cov.any()
    # chunked R accumulation equals the one-shot factorisation (forced tiny chunks), M = 2
    from pytuq_b200.lreg import tsqr
    rng = np.random.default_rng(3)
    mi = orc.get_mi(3, 4)
    x = rng.uniform(-1, 1, (1500, 4))
    Y = rng.standard_normal((1500, 2))
    pc = PCRV(1, 4, "LU", mi=mi)
    R1 = tsqr.r_factor_pc(pc, x, Y, 0).cpu().numpy()
    R2 = tsqr.r_factor_pc(pc, x, Y, 0, chunk_rows=100).cpu().numpy()
    assert rel_err(R2.T @ R2, R1.T @ R1) < 1e-12
    Z = np.hstack([orc.eval_bases(x, mi, "LU"), Y])
    assert rel_err(R1.T @ R1, Z.T @ Z) < 1e-12
    C, sv, rank = tsqr.lstsq_from_r(tsqr.r_factor_pc(pc, x, Y, 0), mi.shape[0])
    assert rank == 35 and rel_err(C, np.linalg.lstsq(Z[:, :35], Y, rcond=None)[0]) < 1e-11
    # a well-conditioned tall problem still takes the Gram route and agrees with the QR route
    g, q = lsq(), lsq(solver="qr")
    for f in (g, q):
        f.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        f.fit(x, Y[:, 0])
    assert rel_err(g.cf, q.cf) < 1e-12
    # pc_fit(lsq) with fewer than 4 K samples: one batched QR solve for all outputs
    import io, contextlib
    from pytuq_b200.workflows.fits import pc_fit
    xs, Ys = x[:90], np.stack([np.sin(x[:90, 0]) + x[:90, 1] * x[:90, 2], np.exp(0.3 * x[:90, 3])], axis=1)
    with contextlib.redirect_stdout(io.StringIO()):
        pcf, regs = pc_fit(xs, Ys, order=3, pctype="LU", method="lsq")
    As = orc.eval_bases(xs, mi, "LU")
    for j in range(2):
        assert rel_err(regs[j].cf, orc.lsq_fit(As, Ys[:, j])) < 1e-9


def test_small_residuals_take_the_exact_residual_pass():
    """Noise far below the data scale: the statistics identity for the residual sum of squares cancels
    (relative size 1e-18 of y^T y at noise 1e-9), so bcs's re-estimated sigma2 (lreg/bcs.py:229) must come
    from the residual pass over the data (K2) -- compared with the reference algorithm on a materialised A.
    anl's estimate uses y.(y - A cf) (lreg/anl.py:61), which at such noise levels is dominated by the rounding
    of the K x K solve in the reference itself; there only the order of magnitude is comparable."""
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import anl, bcs, PCBasisEvaluator
    from pytuq_b200.lreg.stats import SuffStats
    rng = np.random.default_rng(17)
    d, p, n = 4, 3, 4000
    mi = orc.get_mi(p, d)
    K = mi.shape[0]
    x = rng.uniform(-1, 1, (n, d))
    A = orc.eval_bases(x, mi, "LU")
    c = np.zeros(K); c[[0, 2, 7, 19]] = [1.0, -0.5, 0.25, 0.8]
    pc = PCRV(1, d, "LU", mi=mi)
    for noise in (1e-6, 1e-9):
        y = A @ c + noise * rng.standard_normal(n)
        w_ref, _, used_ref, s2_ref, _ = orc.bcs_fit(A, y, eta=1e-8)
        s2_ref = float(np.ravel(s2_ref)[0])
        assert s2_ref < 1e-9          # far below RESIDUAL_REFINE * y^T y / N: the identity alone would cancel
        for route in ("fit", "fita"):
            b = bcs(eta=1e-8)
            if route == "fit":
                b.setBasisEvaluator(PCBasisEvaluator(0), (pc,)); b.fit(x, y)
            else:
                b.fita(A, y)
            assert np.array_equal(np.sort(b.used), np.sort(used_ref)), (noise, route)
            s2 = float(np.ravel(b.datavar)[0])
            assert abs(s2 - s2_ref) < 1e-4 * s2_ref, (noise, route, s2, s2_ref)
        # the residual accessor itself, against a direct evaluation
        st = SuffStats.from_pc(pc, x, y, 0)
        r = y - A[:, used_ref] @ w_ref
        assert abs(st.rss(w_ref, used_ref, 0) - r @ r) < 1e-6 * (r @ r)
        a = anl(); a.setBasisEvaluator(PCBasisEvaluator(0), (pc,)); a.fit(x, y)
        cf_ref, _, dv_ref = orc.anl_fit(A, y)
        assert rel_err(a.cf, cf_ref) < 1e-9
        assert a.datavar >= 0.0 and a.datavar < 30 * noise ** 2 + 1e-13
    # the identity is kept whenever it is accurate: no residual pass for ordinary noise levels
    y = A @ c + 1e-2 * rng.standard_normal(n)
    st = SuffStats.from_pc(pc, x, y, 0)
    calls = []
    inner = st._residual
    st._residual = lambda cf, col: (calls.append(1), inner(cf, col))[1]
    a = anl(); a._fit_stats(st)
    assert not calls and abs(a.datavar - orc.anl_fit(A, y)[2]) < 1e-9 * a.datavar


def test_kle_and_klsurr_vs_reference():
    """8(f3): KL expansion + PC surrogate of a field output (BASELINE config 5 shape) against vectors generated
    with the reference's KLE / pc_fit / PCRV (tests/golden/klsurr.npz).  Eigenvector signs are arbitrary, so
    only sign-invariant quantities are compared."""
    import io, contextlib
    from conftest import load_golden
    from pytuq_b200.linred.kle import KLE
    from pytuq_b200.linred.klsurr import KLSurr
    g = load_golden("klsurr")
    x, y = g["x"], g["y"]
    kl = KLE()
    kl.build(y.T, plot=False)
    big = g["eigval"] > 1e-12 * g["eigval"][0]
    assert rel_err(kl.eigval[big], g["eigval"][big]) < 1e-9
    assert rel_err(kl.mean, g["mean"]) < 1e-13
    neig = int(g["neig"])
    assert rel_err(np.abs(kl.xi[:, :neig]), g["abs_xi"]) < 1e-8
    assert rel_err(kl.compute_relvar()[:, :neig], g["relvar"]) < 1e-9
    assert rel_err(kl.eval(neig=neig), g["ytrain_kl"]) < 1e-10
    assert rel_err(kl.eval(xi=kl.project(y.T), neig=neig), g["ytrain_kl"]) < 1e-10
    ks = KLSurr()
    with contextlib.redirect_stdout(io.StringIO()):
        ks.build(x, y, pc_order=3, bcs_eta=1.e-8)
        sens = ks.compute_sens_pc(totmain='main')
        sens_eig = ks.sens_eig
        sens_tot = ks.compute_sens_pc(totmain='total')
    assert ks.neig == neig and ks.built
    assert rel_err(ks.ytrain_kl, g["ytrain_kl"]) < 1e-10
    assert rel_err(ks.ytrain_klsurr, g["ytrain_klsurr"]) < 1e-8
    assert rel_err(ks.eval(g["xt"]), g["ytest"]) < 1e-8
    assert rel_err(ks.function(g["xt"]), g["ytest"].T) < 1e-8
    from pytuq_b200.utils.mindex import micf_join
    mi_all, cf_all = micf_join(ks.smodel.mindices, ks.smodel.coefs)
    assert np.array_equal(mi_all, g["mindex_all"])
    assert rel_err(np.abs(cf_all), g["abs_cfs_all"]) < 1e-8
    assert rel_err(sens_eig, g["sens_eig_main"]) < 1e-8
    assert rel_err(sens, g["sens_main"]) < 1e-8
    assert rel_err(sens_tot, g["sens_total"]) < 1e-8


def test_pc_fit_workflow_vs_reference():
    import contextlib, io
    from pytuq_b200.workflows.fits import pc_fit
    g = load_golden("workflow")
    x, y = g["x"], g["y"]
    for method in ("lsq", "anl", "bcs"):
        with contextlib.redirect_stdout(io.StringIO()):
            pc, lregs = pc_fit(x, y, order=3, pctype="LU", method=method)
        assert len(lregs) == 2
        for j in range(2):
            assert np.array_equal(pc.mindices[j], g[f"{method}__mi{j}"]), method
            np.testing.assert_allclose(pc.coefs[j], g[f"{method}__cf{j}"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(pc.computeSens(), g[f"{method}__main"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(pc.computeTotSens(), g[f"{method}__total"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(pc.function(x[:16]), g[f"{method}__eval"], rtol=1e-10, atol=1e-12)


def test_model_sens_pcsobol_batched_vs_reference():
    """gsa.model_sens(method='PCSobol') (reference gsa/gsa.py:21-66): same germ from numpy's stream, all outputs fitted
    in ONE statistics pass, against the unmodified reference's per-output loop (tests/golden/gsa.npz)."""
    import contextlib, io
    from conftest import golden_generator
    from pytuq_b200.gsa.gsa import model_sens, PCSobol
    gen = golden_generator()
    g = load_golden("gsa")
    dom = np.array([[-np.pi, np.pi]] * 3)
    np.random.seed(31)
    with contextlib.redirect_stdout(io.StringIO()):
        main, tot = model_sens(gen.gsa_model, (7.0, 0.1), dom, method="PCSobol", nsam=4000, plot=False, pctype="LU", order=5)
    assert main.shape == (2, 3) and tot.shape == (2, 3)
    np.testing.assert_allclose(main, g["main"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(tot, g["total"], rtol=1e-9, atol=1e-12)
    np.random.seed(31)
    with contextlib.redirect_stdout(io.StringIO()):
        sm = PCSobol(dom, pctype="LU", order=5)
        xs = sm.sample(4000)
        sens = sm.compute(gen.gsa_model(xs, (7.0, 0.1))[:, 0])
    np.testing.assert_allclose(xs[:8], g["xsam_head"], rtol=1e-14)
    np.testing.assert_allclose(xs.sum(), g["xsam_sum"], rtol=1e-12)
    np.testing.assert_allclose(sens["main"], g["main0"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(sens["jointt"], g["jointt0"], rtol=1e-9, atol=1e-12)
    with pytest.raises(NotImplementedError):
        model_sens(gen.gsa_model, (7.0, 0.1), dom, method="SamSobol", nsam=10, plot=False)


def test_pce_and_pcsobol():
    import contextlib, io
    from pytuq_b200.surrogates.pce import PCE
    from pytuq_b200.gsa.gsa import PCSobol
    rng = np.random.default_rng(17)
    x = rng.uniform(-1, 1, (300, 3))
    y = 1.0 + 2.0 * x[:, 0] + 0.5 * x[:, 1] * x[:, 2] + x[:, 0] ** 2
    pce = PCE(3, 2, "LU")
    pce.set_training_data(x, y)
    cf = pce.build(regression="lsq")
    assert pce.get_pc_terms() == [10]
    out = pce.evaluate(x)
    np.testing.assert_allclose(out["Y_eval"], y, atol=1e-10)
    assert out["Y_eval_std"] is None
    pce.build(regression="anl")
    assert pce.evaluate(x[:7])["Y_eval_std"].shape == (7,)
    cfb = pce.build(regression="bcs", eta=1e-6)
    assert len(cfb) == len(pce.lreg.used) <= 10
    mi2 = orc.get_mi(2, 3)
    wref, _, uref, _, _ = orc.bcs_fit(orc.eval_bases(x, mi2, "LU"), y, eta=1e-6)
    assert np.array_equal(pce.lreg.used, uref)
    np.testing.assert_allclose(cfb, wref, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(pce.evaluate(x)["Y_eval"], orc.eval_bases(x, mi2[uref], "LU") @ wref, rtol=1e-9, atol=1e-10)
    eta = pce.build(regression="bcs", eta=[1e-2, 1e-6], nfolds=2)
    assert pce.lreg.eta in (1e-2, 1e-6)
    # PC Sobol on a seeded germ: same numpy stream as the reference, indices vs the oracle fit
    with contextlib.redirect_stdout(io.StringIO()):
        dom = np.array([[-1., 1.], [0., 2.], [-3., 3.]])
        sob = PCSobol(dom, pctype="LU", order=3)
        np.random.seed(5)
        xs = sob.sample(500)
        ys = xs[:, 0] + 0.3 * xs[:, 1] ** 2 + 0.1 * xs[:, 0] * xs[:, 2]
        sens = sob.compute(ys)
    mi = orc.get_mi(3, 3)
    cf_ref = orc.lsq_fit(orc.eval_bases(sob.germ_sam, mi, "LU"), ys)
    np.testing.assert_allclose(sens["main"], orc.sobol_main(mi, cf_ref, "LU"), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(sens["total"], orc.sobol_total(mi, cf_ref, "LU"), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(sens["jointt"], orc.sobol_joint(mi, cf_ref, "LU"), rtol=1e-10, atol=1e-13)


def test_randomized_shapes_all_kernels():
    """Random (families, dimension, order, N, outputs) combinations, including truncated / shuffled
    multi-index sets and ragged N: K1 and K2 bit-exact, K3 / K4 1e-12 against the oracle."""
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg.stats import SuffStats
    rng = np.random.default_rng(2024)
    fams = list(orc.FAMILIES)
    for trial in range(36):
        d = int(rng.integers(1, 9))
        p = int(rng.integers(0, 5 if d > 4 else 7))
        mi = orc.get_mi(p, d)
        if mi.shape[0] > 3 and trial % 3 == 0:          # general sets: subset, shuffled, with a duplicate
            keep = rng.permutation(mi.shape[0])[: max(2, mi.shape[0] // 2)]
            mi = mi[keep]
            mi = np.vstack([mi, mi[:1]])
        types = [fams[int(rng.integers(0, 4))] for _ in range(d)]
        n = int(rng.choice([1, 3, 15, 16, 17, 31, 64, 65, 255, 1000, 2049]))
        M = int(rng.choice([0, 1, 1, 2, 5]))
        x = np.stack([rng.uniform(-1, 1, n) if t.endswith("LU") else rng.standard_normal(n) for t in types], axis=1)
        Y = rng.standard_normal((n, M)) if M else None
        pdim = max(M, 1)
        cfs = rng.standard_normal((pdim, mi.shape[0]))
        pc = PCRV(pdim, d, types, mi=mi, cfs=cfs)
        max_ord = mi.max(axis=0)
        A = pc.evalBases(x, 0)
        ref = orc.eval_bases(x, mi, types, max_ord)
        assert np.array_equal(A, ref), (trial, d, p, n)
        assert np.array_equal(pc.evalPC(x), orc.eval_pc(x, [mi] * pdim, list(cfs), types, max_ord)), (trial, d, p, n)
        Sref = orc.suffstats(ref, Y if M else np.zeros((n, 0)))
        scale = np.sqrt(np.outer(np.diag(Sref), np.diag(Sref))) + 1e-300
        S3 = pc.suffStats(x, Y, 0).S
        assert np.max(np.abs(S3 - Sref) / scale) < 1e-12, (trial, d, p, n, M)
        S4 = SuffStats.from_matrix(ref, Y).S
        assert np.max(np.abs(S4 - Sref) / scale) < 1e-12, (trial, d, p, n, M)


def test_uprop_projection_and_regression():
    """workflows/uprop.py:6-69 on the fused statistics: Galerkin projection on a tensor Gauss grid and
    least-squares regression of a model pushed through an input PC, against the A-based formulas."""
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.workflows.uprop import uprop_proj, uprop_regr
    d = 3
    in_pc = PCRV(d, d, "LU", mi=[orc.get_mi(1, d)] * d,
                 cfs=[np.array([0.5, 1.0, 0.0, 0.0]), np.array([-1.0, 0.0, 2.0, 0.0]), np.array([0.0, 0.0, 0.0, 0.5])])

    def model(z):
        return np.stack([z[:, 0] * z[:, 1] + z[:, 2] ** 2, np.sin(z[:, 0]) + z[:, 2]], axis=1)

    mi = orc.get_mi(3, d)
    # --- projection
    out_pc = PCRV(2, d, "LU", mi=mi)
    uprop_proj(in_pc, model, 5, out_pc)
    q, w = in_pc.quadGerm([5] * d)
    yq = model(orc.eval_pc(q, in_pc.mindices, in_pc.coefs, "LU"))
    Aq = orc.eval_bases(q, mi, "LU")
    nsq = orc.norms_sq(mi, "LU")
    for j in range(2):
        ref = np.array([np.dot(yq[:, j] * w, Aq[:, i]) / nsq[i] for i in range(mi.shape[0])])
        np.testing.assert_allclose(out_pc.coefs[j], ref, rtol=1e-10, atol=1e-13)
    # first output is a polynomial of degree 2 in the germ: projection is exact, so is the PC mean
    zq = orc.eval_pc(q, in_pc.mindices, in_pc.coefs, "LU")
    assert abs(out_pc.computeMean()[0] - np.sum(w * (zq[:, 0] * zq[:, 1] + zq[:, 2] ** 2))) < 1e-12
    # --- regression (same numpy germ stream as the reference: sampleGerm is host-side)
    out_pc2 = PCRV(2, d, "LU", mi=mi)
    np.random.seed(11)
    uprop_regr(in_pc, model, 400, out_pc2)
    np.random.seed(11)
    g = in_pc.sampleGerm(400)
    ys = model(orc.eval_pc(g, in_pc.mindices, in_pc.coefs, "LU"))
    Ag = orc.eval_bases(g, mi, "LU")
    for j in range(2):
        ref = np.linalg.lstsq(Ag, ys[:, j], rcond=None)[0]
        np.testing.assert_allclose(out_pc2.coefs[j], ref, rtol=1e-9, atol=1e-11)


def test_pce_eta_search_and_single_sample_latency_path():
    from pytuq_b200.surrogates.pce import PCE
    from pytuq_b200.rv.pcrv import PCRV
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (240, 4))
    y = 2.0 * x[:, 0] - x[:, 1] * x[:, 2] + 0.05 * rng.standard_normal(240)
    pce = PCE(4, 3, "LU")
    pce.set_training_data(x, y)
    etas = [1e-1, 1e-3, 1e-6]
    pce.build(regression="bcs", eta=etas, nfolds=3)
    assert pce.lreg.eta in etas and pce.eta_cv["rmse_val"].shape == (3, 3)
    assert {0 * 0 + k for k in pce.lreg.used} >= set()      # used is an index array into the full basis
    out = pce.evaluate(x[:5])
    assert out["Y_eval"].shape == (5,) and out["Y_eval_std"].shape == (5,)
    # N = 1 evaluation (the MCMC call pattern, apps/iuq/run_infer.py:147): same bits as a batch
    pc = PCRV(1, 4, "LU", mi=orc.get_mi(3, 4), cfs=rng.standard_normal(35))
    pc.setFunction()
    batch = pc.function(x[:7])
    for i in range(7):
        assert np.array_equal(pc.function(x[i:i + 1]), batch[i:i + 1])


# ----------------------------------------------------------------------------- full-size properties
def test_full_size_c2_properties():
    """BASELINE config 2 at full size (HG, d=20, p=3, K=1771, N=1e6): size-independent checks."""
    import torch
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    d, p, n = 20, 3, 1_000_000
    mi = orc.get_mi(p, d)
    K = mi.shape[0]
    pc = PCRV(1, d, "HG", mi=mi)
    plan = pc._plan(mi)
    lib = _native.lib()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)
    c = torch.randn(K, dtype=torch.float64, device="cuda", generator=g) / (1.0 + torch.from_numpy(mi.sum(1)).cuda()) ** 2
    stream = torch.cuda.current_stream().cuda_stream
    y = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, d, c.data_ptr(), 1, y.data_ptr(), 1, 0, stream))
    L = K + 2
    S = torch.empty((L, L), dtype=torch.float64, device="cuda")
    _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), n, d, y.data_ptr(), 1, 1, S.data_ptr(), 0, stream))
    # (1) linearity over the sample axis: halves accumulate to the whole
    S2 = torch.empty_like(S)
    h = 499_999
    _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), h, d, y.data_ptr(), 1, 1, S2.data_ptr(), 0, stream))
    _native.check(lib.pcq_suffstats_dev(plan.handle, x[h:].data_ptr(), n - h, d, y[h:].data_ptr(), 1, 1, S2.data_ptr(), 1, stream))
    torch.cuda.synchronize()
    Sh, S2h = S.cpu().numpy(), S2.cpu().numpy()
    dg = np.sqrt(np.outer(np.diag(Sh), np.diag(Sh)))
    assert np.max(np.abs(Sh - S2h) / dg) < 1e-12
    assert Sh[-1, -1] == n and Sh[0, 0] == n and np.array_equal(Sh, Sh.T)
    # (2) noiseless data: the normal equations recover the coefficients that generated y
    from pytuq_b200.lreg.lreg import solve_normal
    cf = solve_normal(Sh[:K, :K], Sh[:K, K])
    np.testing.assert_allclose(cf, c.cpu().numpy(), rtol=0, atol=1e-10)
    # (3) K1 rows agree with the statistics: A^T 1 column and a block of A^T A from a 64k-row slab
    m = 65536
    A = torch.empty((m, K), dtype=torch.float64, device="cuda")
    _native.check(lib.pcq_eval_bases_dev(plan.handle, x.data_ptr(), m, d, A.data_ptr(), K, stream))
    Sm = torch.empty_like(S)
    _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), m, d, y.data_ptr(), 1, 1, Sm.data_ptr(), 0, stream))
    torch.cuda.synchronize()
    Gm = (A.T @ A).cpu().numpy()
    Smh = Sm.cpu().numpy()
    dgm = np.sqrt(np.outer(np.diag(Gm), np.diag(Gm)))
    assert np.max(np.abs(Smh[:K, :K] - Gm) / dgm) < 1e-12
    np.testing.assert_allclose(Smh[:K, K], (A.T @ y[:m, 0]).cpu().numpy(), rtol=1e-10, atol=1e-7)
    # (4) evalPC equals A @ c on the slab (sequential vs BLAS order: tolerance)
    np.testing.assert_allclose((A @ c).cpu().numpy(), y[:m, 0].cpu().numpy(), rtol=1e-11, atol=1e-11)
    # oracle spot check on rows from the far end of the array
    idx = [0, 1, m - 1]
    assert np.array_equal(A[idx].cpu().numpy(), orc.eval_bases(x[idx].cpu().numpy(), mi, "HG"))


def uqpc_model(p):
    # the stand-in model of the replay; same expression as tests/golden/make_golden.py:uqpc_model
    return np.stack([np.exp(0.3 * p[:, 0]) + p[:, 1] * p[:, 2], np.sin(p[:, 0]) + 0.1 * p[:, 2] ** 2 + p[:, 1]], axis=1)


def test_uqpc_app_replay_through_sample_files(tmp_path):
    """apps/uqpc/uq_pc.py:120-345 (parameter-domain input, regime online_example) replayed with this package and its
    sample-file I/O against the same replay done with the reference (tests/golden/make_golden.py:uqpc): the files
    the app writes are reproduced byte for byte (evalPC is bit-identical), fits / predictions / variances /
    Sobol indices to the north-star tolerances."""
    import contextlib, io as _io, os
    from conftest import GOLDEN
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.utils import io
    from pytuq_b200.utils.mindex import get_mi
    from pytuq_b200.workflows.fits import pc_fit
    gold = os.path.join(GOLDEN, "uqpc")
    g = load_golden("uqpc")
    pdom = io.loadtxt(os.path.join(gold, "pdomain.txt")).reshape(-1, 2)
    pcf_all = np.vstack((0.5 * (pdom[:, 1] + pdom[:, 0]), np.diag(0.5 * (pdom[:, 1] - pdom[:, 0]))))
    in_pcdim, npar = pdom.shape[0], pcf_all.shape[1]
    pc = PCRV(in_pcdim, in_pcdim, "LU", mi=get_mi(1, in_pcdim), cfs=pcf_all.T)

    def same_file(name):
        return open(tmp_path / name, "rb").read() == open(os.path.join(gold, name), "rb").read()

    for q, p, n in (("qtrain.txt", "ptrain.txt", 240), ("qtest.txt", "ptest.txt", 50)):
        germ = pc.sampleGerm(n, seed=7)
        io.savetxt(str(tmp_path / q), germ)
        io.savetxt(str(tmp_path / p), pc.evalPC(germ))
        assert same_file(q) and same_file(p), (q, p)
    ptrain = io.loadtxt(str(tmp_path / "ptrain.txt")).reshape(-1, npar)
    germ_train = io.loadtxt(str(tmp_path / "qtrain.txt")).reshape(-1, in_pcdim)
    ptest = io.loadtxt(str(tmp_path / "ptest.txt")).reshape(-1, npar)
    germ_test = io.loadtxt(str(tmp_path / "qtest.txt")).reshape(-1, in_pcdim)
    # model outputs: read from the files as the app's offline regime does (uq_pc.py:262-265) -- exp / sin of the
    # host's numpy may differ in the last bit from box to box, so the model itself is only checked to 1e-14
    ytrain = io.loadtxt(os.path.join(gold, "ytrain.txt")).reshape(ptrain.shape[0], -1)
    ytest = io.loadtxt(os.path.join(gold, "ytest.txt")).reshape(ptest.shape[0], -1)
    np.testing.assert_allclose(uqpc_model(ptrain), ytrain, rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(uqpc_model(ptest), ytest, rtol=1e-14, atol=1e-15)
    io.savetxt(str(tmp_path / "ytrain.txt"), ytrain)
    assert same_file("ytrain.txt")

    for method in ("lsq", "anl", "bcs"):
        with contextlib.redirect_stdout(_io.StringIO()):
            opc, lregs = pc_fit(germ_train, ytrain, order=3, pctype="LU", method=method, eta=1e-6)
        ytrain_pc, ytest_pc = opc.function(germ_train), opc.function(germ_test)
        vtr, vte = np.empty(ytrain.shape), np.empty(ytest.shape)
        for iout, lreg in enumerate(lregs):
            assert np.array_equal(opc.mindices[iout], g[f"{method}__mi{iout}"]), (method, iout)
            np.testing.assert_allclose(opc.coefs[iout], g[f"{method}__cf{iout}"], rtol=1e-10, atol=1e-12)
            vtr[:, iout] = lreg.predicta(opc.evalBases(germ_train, iout), msc=1)[1]
            vte[:, iout] = lreg.predicta(opc.evalBases(germ_test, iout), msc=1)[1]
        np.testing.assert_allclose(ytrain_pc, g[f"{method}__ytrain_pc"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(ytest_pc, g[f"{method}__ytest_pc"], rtol=1e-10, atol=1e-12)
        vscale = max(np.max(g[f"{method}__ytrain_pc_var"]), 1e-300)
        np.testing.assert_allclose(vtr, g[f"{method}__ytrain_pc_var"], rtol=1e-7, atol=1e-9 * vscale)
        np.testing.assert_allclose(vte, g[f"{method}__ytest_pc_var"], rtol=1e-7, atol=1e-9 * vscale)
        relerr = np.linalg.norm(ytrain - ytrain_pc, axis=0) / np.linalg.norm(ytrain, axis=0)
        np.testing.assert_allclose(relerr, g[f"{method}__relerr_train"], rtol=1e-7)
        np.testing.assert_allclose(opc.computeSens(), g[f"{method}__main"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(opc.computeTotSens(), g[f"{method}__total"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(opc.computeJointSens(), g[f"{method}__joint"], rtol=1e-10, atol=1e-13)


def test_full_size_baseline_config_1_vs_reference():
    """BASELINE.json configs[0] at FULL size — Legendre PCE of total order 4 in d=10 (K=1001), lsq and anl fits of
    the Genz oscillatory function on 1e5 uniform samples — against the unmodified reference run on the same inputs
    (tests/golden/c1_full.npz, make_golden.py:c1_full).  Bars of the north star: basis values bit-identical,
    coefficients and Sobol indices within 1e-10."""
    from conftest import c1_inputs
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import lsq, anl, PCBasisEvaluator
    g = load_golden("c1_full")
    x, y = c1_inputs()
    np.testing.assert_allclose(y[:16], g["y_head"], rtol=0, atol=1e-14)
    mi = orc.get_mi(4, 10)
    pc = PCRV(1, 10, "LU", mi=mi)
    A = pc.evalBases(x, 0)
    assert A.shape == (100_000, 1001) and np.array_equal(A[g["rows"]], g["A_rows"])
    scale = np.max(np.abs(g["cf_lsq"]))
    fits = {}
    for name, make in (("lsq", lsq), ("anl", anl)):
        fused = make()
        fused.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        fused.fit(x, y)                       # K3: statistics without a design matrix
        given = make()
        given.fita(A, y)                      # K4: statistics of the caller's matrix
        for f in (fused, given):
            np.testing.assert_allclose(f.cf, g[f"cf_{name}"], rtol=0, atol=1e-10 * scale)
            assert np.array_equal(f.used, np.arange(1001))
        fits[name] = fused
    np.testing.assert_allclose(fits["anl"].datavar, g["anl_datavar"], rtol=1e-9)
    np.testing.assert_allclose(np.diag(fits["anl"].cf_cov), g["anl_cov_diag"], rtol=1e-8)
    pc.setCfs([fits["lsq"].cf])
    np.testing.assert_allclose(pc.computeMean(), g["mean"], rtol=1e-10)
    np.testing.assert_allclose(pc.computeVar(), g["var"], rtol=1e-10)
    np.testing.assert_allclose(pc.computeSens(), g["main"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(pc.computeTotSens(), g["total"], rtol=1e-10, atol=1e-13)
    # surrogate evaluation equals the design matrix times the coefficients (sequential vs BLAS order)
    np.testing.assert_allclose(pc.evalPC(x[:4096])[:, 0], A[:4096] @ fits["lsq"].cf, rtol=1e-12, atol=1e-13)


def test_headline_shape_c2_fit_vs_reference():
    """The headline shape — Hermite PCE d=20 p=3 (K=1771), lsq and anl — against the unmodified reference run on the
    same 2e4 Gaussian samples (tests/golden/c2_shape.npz, make_golden.py:c2_shape; lreg/lreg.py:328-339,
    lreg/anl.py:69-105).  Basis values bit-identical, coefficients and Sobol indices within 1e-10."""
    from conftest import golden_generator
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import lsq, anl, PCBasisEvaluator
    g = load_golden("c2_shape")
    x, _, _ = golden_generator().c2_inputs()
    y = g["y"]
    mi = orc.get_mi(3, 20)
    pc = PCRV(1, 20, "HG", mi=mi)
    A = pc.evalBases(x, 0)
    assert A.shape == (20_000, 1771) and np.array_equal(A[g["rows"]], g["A_rows"])
    scale = np.max(np.abs(g["cf_lsq"]))
    fits = {}
    for name, make in (("lsq", lsq), ("anl", anl)):
        fused = make()
        fused.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        fused.fit(x, y)                       # K3: statistics without a design matrix
        plain = make()                        # the reference idiom: a plain closure as the evaluator
        plain.setBasisEvaluator(lambda xx, pars: pars[0].evalBases(xx, 0), (pc,))
        plain.fit(x, y)
        given = make()
        given.fita(A, y)                      # K4: statistics of the caller's matrix
        for f in (fused, plain, given):
            np.testing.assert_allclose(f.cf, g[f"cf_{name}"], rtol=0, atol=1e-10 * scale)
            assert np.array_equal(f.used, np.arange(1771))
        fits[name] = fused
    np.testing.assert_allclose(fits["anl"].datavar, g["anl_datavar"], rtol=1e-9)
    np.testing.assert_allclose(np.diag(fits["anl"].cf_cov), g["anl_cov_diag"], rtol=1e-8)
    pc.setCfs([fits["lsq"].cf])
    np.testing.assert_allclose(pc.computeMean(), g["mean"], rtol=1e-10)
    np.testing.assert_allclose(pc.computeVar(), g["var"], rtol=1e-10)
    np.testing.assert_allclose(pc.computeSens(), g["main"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(pc.computeTotSens(), g["total"], rtol=1e-10, atol=1e-13)


def test_shapes_of_baseline_configs_3_4_5_vs_reference():
    """The CUDA path at the (d, p, K) of BASELINE.json configs 3, 4, 5 (sample counts reduced to what the unmodified
    reference finishes in about a minute) against that reference (tests/golden/c{3,4,5}_shape.npz,
    make_golden.py:config_shapes).  Basis values and exact-mode evalPC bit-identical; coefficients, moments and
    Sobol indices within 1e-10; BCS keeps the identical term set."""
    import contextlib, io as _io
    from conftest import golden_generator
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import bcs, PCBasisEvaluator
    from pytuq_b200.workflows.fits import pc_fit
    gen = golden_generator()

    # --- C3: BCS sparse regression, LU d=50 p=2 (K=1326), 40-sparse truth
    g, c = load_golden("c3_shape"), gen.shape_inputs(3)
    mi = orc.get_mi(c["p"], c["d"])
    pc = PCRV(1, c["d"], c["fam"], mi=mi)
    A = pc.evalBases(c["x"], 0)
    assert np.array_equal(A[g["rows"]], g["A_rows"])
    y = gen.sparse_target(A, c["idx"], c["val"], c["noise"])
    assert np.array_equal(y[:16], g["y_head"])
    fused = bcs(eta=1e-8)
    fused.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
    given = bcs(eta=1e-8)
    with contextlib.redirect_stdout(_io.StringIO()):
        fused.fit(c["x"], y)                  # K3, Gram-form iteration
        given.fita(A, y)                      # K4
    for f in (fused, given):
        assert np.array_equal(f.used, g["used"]) and set(f.used) == set(c["idx"])
        np.testing.assert_allclose(f.cf, g["cf"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(np.reshape(f.datavar, -1), g["datavar"], rtol=1e-7)
        np.testing.assert_allclose(np.diag(f.cf_cov), g["cov_diag"], rtol=1e-7)

    # --- C4: the d=20 p=4 surrogate (K=10626): evalPC, moments, Sobol indices
    g, c = load_golden("c4_shape"), gen.shape_inputs(4)
    mi = orc.get_mi(c["p"], c["d"])
    pc = PCRV(1, c["d"], c["fam"], mi=mi, cfs=c["cf"])
    assert np.array_equal(pc.evalPC(c["x"]), g["y"])
    np.testing.assert_allclose(pc.evalPC(c["x"], exact=False), g["y"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(pc.computeMean(), g["mean"], rtol=1e-14)
    np.testing.assert_allclose(pc.computeVar(), g["var"], rtol=1e-10)
    np.testing.assert_allclose(pc.computeSens(), g["main"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(pc.computeTotSens(), g["total"], rtol=1e-10, atol=1e-14)
    # forward propagation with the device germ reproduces the closed-form moments statistically
    mom = pc.propagate(2_000_000, seed=4)
    sd = np.sqrt(g["var"][0])
    assert abs(mom["mean"][0] - g["mean"][0]) < 6 * sd / np.sqrt(2e6)
    assert abs(mom["var"][0] / g["var"][0] - 1.0) < 0.05

    # --- C5: multi-output fit sharing one design matrix, LU d=30 p=3 (K=5456), anl
    g, c = load_golden("c5_shape"), gen.shape_inputs(5)
    with contextlib.redirect_stdout(_io.StringIO()):
        opc, lregs = pc_fit(c["x"], c["y"], order=c["p"], pctype=c["fam"], method="anl")
    cf = np.stack(opc.coefs, axis=1)
    assert cf.shape == (5456, 4)
    np.testing.assert_allclose(cf, g["cf"], rtol=0, atol=1e-10 * np.max(np.abs(g["cf"])))
    np.testing.assert_allclose([float(l.datavar) for l in lregs], g["datavar"], rtol=1e-7)
    np.testing.assert_allclose(np.stack([np.diag(l.cf_cov) for l in lregs], axis=1), g["cov_diag"], rtol=1e-7)
    np.testing.assert_allclose(opc.computeSens(), g["main"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(opc.computeTotSens(), g["total"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(opc.function(c["x"][:64]), g["eval"], rtol=1e-10, atol=1e-12)


# ----------------------------------------------------------------------------- sample files -> HBM (pcq_text.cu)
def _write(path, text):
    with open(path, "wb") as f:
        f.write(text if isinstance(text, bytes) else text.encode())
    return str(path)


def test_text_tables_are_converted_on_the_device_bit_for_bit(tmp_path):
    """utils.io.loadtxt_device (pcq_text_open / pcq_text_parse_dev) against np.loadtxt: same shape, same bits, for
    files written by np.savetxt and for hand-written layouts; what the device route declines or rejects."""
    from pytuq_b200.utils import io
    rng = np.random.default_rng(11)
    tables = {
        "samples": rng.standard_normal((5000, 20)),
        "wide_range": rng.standard_normal((700, 7)) * 10.0 ** rng.integers(-300, 300, (700, 7)),   # subnormals -> fix-ups
        "column": rng.uniform(-1, 1, (3000, 1)),
        "one_value": np.array([[0.1]]),
        "ties": ((2.0 * rng.integers(10 ** 8, 10 ** 9, 64) + 1) * 5.0 ** 10 / 2.0 ** 11 / 10.0 ** 10).reshape(16, 4),
        "integers": np.arange(-50, 50).reshape(20, 5),
        "edges": np.array([[5e-324, 1.7976931348623157e308, 2.2250738585072014e-308, -0.0, 1e22, 1e23]]),
    }
    for name, X in tables.items():
        f = str(tmp_path / (name + ".txt"))
        np.savetxt(f, X)
        want = np.loadtxt(f, ndmin=2)
        got = io.loadtxt_device(f)
        assert got.is_cuda and tuple(got.shape) == want.shape, name
        assert np.array_equal(got.cpu().numpy().view(np.uint64), want.view(np.uint64)), name
    # sizes around the 8 KB CTA partition and the 32-byte thread span: every cut position of a token
    for n in (1, 2, 327, 328, 329, 655, 656, 1311):
        X = rng.standard_normal((n, 1))
        f = str(tmp_path / "cut.txt")
        np.savetxt(f, X)
        assert np.array_equal(io.loadtxt_device(f).cpu().numpy(), X), n
    for shift in range(33):
        f = _write(tmp_path / "shift.txt", " " * shift + "1.5 -2.25e3\n" + "\n" * (shift % 3) + "3 4")
        assert np.array_equal(io.loadtxt_device(f).cpu().numpy(), [[1.5, -2250.0], [3.0, 4.0]]), shift
    # hand-written layout: tabs, CRLF, blank lines, leading blanks, short forms, no newline at the end
    f = _write(tmp_path / "layout.txt", "1 2\t3.5\r\n\r\n   \t \n-4e0 +5E-1 .25\n1. 00012 1e-400\n7 8 9")
    want = np.loadtxt(f, ndmin=2)
    got = io.loadtxt_device(f).cpu().numpy()
    assert got.shape == (4, 3) and np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # tokens longer than the 64 bytes staged behind a CTA's 8 KB, lying across that boundary: read from global memory
    long_small = "0." + "0" * 140 + "125"           # few significant digits: converted on the device
    long_digits = "1" + "234567890" * 11            # 100 digits: handed to the host
    for pad in range(8192 - 80, 8192 - 20, 7):
        rows = ["1.5 2.5"] * (pad // 8)
        text = "\n".join(rows) + "\n" + " " * (pad - 8 * (pad // 8)) + long_small + " " + long_digits + "\n3 4\n"
        f = _write(tmp_path / "long.txt", text)
        want = np.loadtxt(f, ndmin=2)
        got = io.loadtxt_device(f).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got.view(np.uint64), want.view(np.uint64)), pad
    f = _write(tmp_path / "empty.txt", "\n\n  \n")
    assert tuple(io.loadtxt_device(f).shape) == (0, 0)
    f = _write(tmp_path / "zero.txt", b"")
    assert tuple(io.loadtxt_device(f).shape) == (0, 0)
    # inf / nan: the device hands them back, the host converts them with numpy's rules, the result is patched in
    f = _write(tmp_path / "special.txt", "1 nan -inf\n+Infinity 2 NaN\n-0 INF 3e0\n")
    want = np.loadtxt(f, ndmin=2)
    got = io.loadtxt_device(f).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got, nan=7.0).view(np.uint64), np.nan_to_num(want, nan=7.0).view(np.uint64))
    # declined: comments are the host reader's job
    for text in ("# header\n1 2\n", "1 2 # tail\n3 4\n", "1 2\n#3 4\n"):
        with pytest.raises(ValueError, match="host reader"):
            io.loadtxt_device(_write(tmp_path / "declined.txt", text))
        assert io.loadtxt(str(tmp_path / "declined.txt"), ndmin=2).shape[1] == 2
    # rejected, as numpy rejects them
    for text in ("1 2 3\n4 5\n", "1 2\n3 4 5 6\n", "1 2\n3 1e\n", "1 2\n3 1.2.3\n", "1 2\n3 --4\n", "1,2\n3,4\n",
                 "1 2\n3 0x10\n", "1 2\n3 1_0\n", "1 2\n3 abc\n", "1 2\n3 na\n"):
        with pytest.raises(ValueError):
            np.loadtxt(_write(tmp_path / "bad.txt", text))
        with pytest.raises(ValueError):
            io.loadtxt_device(str(tmp_path / "bad.txt"))
    with pytest.raises(FileNotFoundError):
        io.loadtxt_device(str(tmp_path / "missing.txt"))


def test_fit_from_sample_files_resident_in_hbm():
    """The uq_pc application's offline regime with the samples converted on the device and the fit reading them in
    place (no host matrix at any point): same goldens as test_uqpc_app_replay_through_sample_files."""
    import contextlib, io as _io, os
    from conftest import GOLDEN
    from pytuq_b200.utils import io
    from pytuq_b200.workflows.fits import pc_fit
    from pytuq_b200.lreg import anl, PCBasisEvaluator
    gold = os.path.join(GOLDEN, "uqpc")
    g = load_golden("uqpc")
    germ = io.loadtxt_device(os.path.join(gold, "qtrain.txt"))
    ytrain = io.loadtxt_device(os.path.join(gold, "ytrain.txt"))
    assert germ.is_cuda and tuple(germ.shape) == (240, 3) and tuple(ytrain.shape) == (240, 2)
    import tempfile
    with tempfile.TemporaryDirectory() as d:           # and back: a resident tensor written in the same format
        io.savetxt(os.path.join(d, "q.txt"), germ)
        assert open(os.path.join(d, "q.txt"), "rb").read() == open(os.path.join(gold, "qtrain.txt"), "rb").read()
    assert np.array_equal(germ.cpu().numpy(), np.loadtxt(os.path.join(gold, "qtrain.txt")))
    for method in ("lsq", "anl", "bcs"):
        with contextlib.redirect_stdout(_io.StringIO()):
            opc, lregs = pc_fit(germ, ytrain, order=3, pctype="LU", method=method, eta=1e-6)
        for iout in range(2):
            assert np.array_equal(opc.mindices[iout], g[f"{method}__mi{iout}"]), (method, iout)
            np.testing.assert_allclose(opc.coefs[iout], g[f"{method}__cf{iout}"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(opc.computeSens(), g[f"{method}__main"], rtol=1e-10, atol=1e-13)
        # evaluation on resident samples: same kernels through the device-pointer entry points, results stay in HBM
        germ_h = germ.cpu().numpy()
        for exact in (True, False):
            yd = opc.evalPC(germ, exact=exact)
            assert yd.is_cuda and np.array_equal(yd.cpu().numpy(), opc.evalPC(germ_h, exact=exact)), (method, exact)
        np.testing.assert_allclose(opc.function(germ).cpu().numpy(), g[f"{method}__ytrain_pc"], rtol=1e-10, atol=1e-12)
        Ad = opc.evalBases(germ, 1)
        assert Ad.is_cuda and np.array_equal(Ad.cpu().numpy(), opc.evalBases(germ_h, 1))
        lregs[0].setBasisEvaluator(PCBasisEvaluator(0), (opc,))
        m_d, v_d, _ = lregs[0].predict(germ, msc=1)
        m_h, v_h, _ = lregs[0].predict(germ_h, msc=1)
        assert np.array_equal(m_d, m_h) and np.array_equal(v_d, v_h)
    # single-output fitter on resident samples, and noise-free data (the residual pass reads them in place too)
    fit = anl()
    fit.setBasisEvaluator(PCBasisEvaluator(0), (_pc_of_order(3, 3),))
    fit.fit(germ, ytrain[:, 0])
    np.testing.assert_allclose(fit.cf, g["anl__cf0"], rtol=1e-10, atol=1e-12)
    import torch
    pc = _pc_of_order(3, 3)
    c = np.random.default_rng(3).standard_normal(20)
    pc.setCfs([c])
    yexact = torch.from_numpy(pc.evalPC(germ.cpu().numpy())).cuda()
    fit = anl()
    fit.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
    fit.fit(germ, yexact[:, 0])
    np.testing.assert_allclose(fit.cf, c, rtol=0, atol=1e-11)
    assert 0.0 <= float(fit.datavar) < 1e-12


def _pc_of_order(order, dim, fam="LU"):
    from pytuq_b200.rv.pcrv import PCRV
    return PCRV(1, dim, fam, mi=orc.get_mi(order, dim))
