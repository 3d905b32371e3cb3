"""pcq_group: one process, several GPUs behind the host-array API (include/pcq.h).  The C-ABI calls are checked against
the single-plan calls (bit-identical rows, statistics to rounding, the same logical Philox stream) with whatever
number of GPUs the box has (a one-member group on a single-GPU box still runs the group's own pipeline: pinned
staging, per-member threads, reduction on member 0); with two or more GPUs the public API is driven through the
fan-out policy (device.devices / fan_out) and compared with the one-GPU result and the reference goldens."""
import contextlib
import ctypes as C
import io

import numpy as np
import pytest

from conftest import load_golden, golden_generator
from oracle import pc_oracle as orc

pytestmark = pytest.mark.gpu


def _members():
    import torch
    return list(range(min(torch.cuda.device_count(), 8)))


def test_group_calls_match_single_plan_calls():
    from pytuq_b200 import _native
    from pytuq_b200.rv.pcrv import PCRV
    lib = _native.lib()
    rng = np.random.default_rng(17)
    d, p, n, M = 6, 3, 70_001, 2
    mi = orc.get_mi(p, d)
    K = mi.shape[0]
    cfs = rng.standard_normal((M, K))
    pc = PCRV(M, d, "LU", mi=mi, cfs=cfs)
    x = rng.uniform(-1, 1, (n, d))
    devs = _members()
    from pytuq_b200 import device as pdev
    pdev.set_devices(devs)
    try:
        grp = pc._group(mi)
        assert lib.pcq_group_size(grp.handle) == len(devs)
        plan = pc._plan(mi)
        # rows are independent: evalBases / evalPC bit-identical to the single-plan calls, for every active count
        A1 = np.empty((n, K)); y1 = np.empty((n, M))
        _native.check(lib.pcq_eval_bases(plan.handle, _native.ptr(x), n, d, _native.ptr(A1), K))
        _native.check(lib.pcq_eval_pc(plan.handle, _native.ptr(x), n, d, _native.ptr(cfs), M, _native.ptr(y1), M, 0))
        assert np.array_equal(A1[:50], orc.eval_bases(x[:50], mi, "LU"))
        Y = y1 + 1e-3 * rng.standard_normal((n, M))
        L = K + M + 1
        S1 = np.empty((L, L))
        _native.check(lib.pcq_suffstats(plan.handle, _native.ptr(x), n, d, _native.ptr(Y), M, M, _native.ptr(S1)))
        Cq = rng.standard_normal((K, K)); Cq = Cq @ Cq.T
        v1 = np.empty(n)
        _native.check(lib.pcq_quadform(plan.handle, _native.ptr(x), n, d, _native.ptr(Cq), K, _native.ptr(v1), 1))
        r1 = np.zeros(2)
        ycol = np.ascontiguousarray(Y[:, 1])
        _native.check(lib.pcq_residual_ss(plan.handle, _native.ptr(x), n, d, _native.ptr(ycol), 1, _native.ptr(cfs[1].copy()), _native.ptr(r1)))
        kinds = pc._germ_kinds()
        p1 = np.zeros((M, 4))
        _native.check(lib.pcq_propagate(plan.handle, 5, 0, 300_001, _native.ptr(kinds), _native.ptr(cfs), M, _native.ptr(p1), 0))
        for nact in sorted({1, len(devs), max(1, len(devs) // 2)}):
            _native.check(lib.pcq_group_set_active(grp.handle, nact))
            A2 = np.empty((n, K)); y2 = np.empty((n, M))
            _native.check(lib.pcq_group_eval_bases(grp.handle, _native.ptr(x), n, d, _native.ptr(A2), K))
            _native.check(lib.pcq_group_eval_pc(grp.handle, _native.ptr(x), n, d, _native.ptr(cfs), M, _native.ptr(y2), M, 0))
            assert np.array_equal(A1, A2) and np.array_equal(y1, y2), nact
            S2 = np.empty((L, L))
            _native.check(lib.pcq_group_suffstats(grp.handle, _native.ptr(x), n, d, _native.ptr(Y), M, M, None, _native.ptr(S2)))
            scale = np.sqrt(np.outer(np.diag(S1), np.diag(S1)))
            assert np.max(np.abs(S2 - S1) / scale) < 1e-13, nact
            assert S2[-1, -1] == n
            v2 = np.empty(n)
            _native.check(lib.pcq_group_quadform(grp.handle, _native.ptr(x), n, d, _native.ptr(Cq), K, _native.ptr(v2), 1))
            np.testing.assert_allclose(v2, v1, rtol=1e-12)
            r2 = np.zeros(2)
            _native.check(lib.pcq_group_residual_ss(grp.handle, _native.ptr(x), n, d, _native.ptr(ycol), 1, _native.ptr(cfs[1].copy()), _native.ptr(r2)))
            np.testing.assert_allclose(r2, r1, rtol=1e-10)
            p2 = np.zeros((M, 4))
            _native.check(lib.pcq_group_propagate(grp.handle, 5, 0, 300_001, _native.ptr(kinds), _native.ptr(cfs), M, _native.ptr(p2), 0))
            np.testing.assert_allclose(p2[:, :2], p1[:, :2], rtol=1e-12)
            assert np.array_equal(p2[:, 2:], p1[:, 2:]), nact            # min / max of the same stream
        # strided inputs and empty blocks (fewer rows than members)
        xs = np.ascontiguousarray(rng.uniform(-1, 1, (5, d + 3)))
        _native.check(lib.pcq_group_set_active(grp.handle, len(devs)))
        A3 = np.empty((5, K + 2))
        _native.check(lib.pcq_group_eval_bases(grp.handle, _native.ptr(xs), 5, d + 3, _native.ptr(A3), K + 2))
        assert np.array_equal(A3[:, :K], orc.eval_bases(xs[:, :d], mi, "LU"))
        S3 = np.empty((K + 1, K + 1))
        _native.check(lib.pcq_group_suffstats(grp.handle, _native.ptr(xs), 5, d + 3, None, 1, 0, None, _native.ptr(S3)))
        Aref = orc.eval_bases(xs[:, :d], mi, "LU")
        np.testing.assert_allclose(S3[:K, :K], Aref.T @ Aref, rtol=1e-12, atol=1e-13)
        assert lib.pcq_group_set_active(grp.handle, len(devs) + 1) < 0
    finally:
        pdev.set_devices(None)


def test_public_api_spreads_over_all_gpus_of_the_process(monkeypatch):
    """lsq / anl / bcs fits, pc_fit, evalPC, predict and propagate in ONE plain process with every visible GPU
    (lreg/lreg.py:62-71, workflows/fits.py:16-81, rv/pcrv.py:469-525) against the reference goldens and the one-GPU
    result.  Needs two GPUs; on one GPU the policy yields the single-plan path (checked)."""
    import torch
    from pytuq_b200 import device as pdev, _native
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.lreg import lsq, anl, bcs, PCBasisEvaluator
    from pytuq_b200.workflows.fits import pc_fit
    ndev = torch.cuda.device_count()
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("PCQ_DEVICES", raising=False)
    pdev.set_devices(None)
    assert pdev.devices() == [pdev.current_device()] + [d for d in range(ndev) if d != pdev.current_device()]
    monkeypatch.setenv("LOCAL_RANK", "0")
    assert pdev.devices() == [pdev.current_device()]          # a torchrun rank keeps to its own device
    monkeypatch.delenv("LOCAL_RANK")
    if ndev < 2:
        assert pdev.fan_out(1e30, 1.0, 10**9) == 1
        pytest.skip("needs two GPUs for the fan-out itself (gpurun --gpus 2)")
    monkeypatch.setenv("PCQ_FANOUT_SCALE", "1e-9")             # every call is "large": forces the fan-out
    gen = golden_generator()
    g = load_golden("c2_shape")
    x, _, _ = gen.c2_inputs()
    y = g["y"]
    mi = orc.get_mi(3, 20)
    pc = PCRV(1, 20, "HG", mi=mi)
    scale = np.max(np.abs(g["cf_lsq"]))
    l0 = _native.lib().pcq_launch_count()
    for name, make in (("lsq", lsq), ("anl", anl)):
        f = make()
        f.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
        f.fit(x, y)
        np.testing.assert_allclose(f.cf, g[f"cf_{name}"], rtol=0, atol=1e-10 * scale)
    assert np.array_equal(pc.evalBases(x, 0)[g["rows"]], g["A_rows"])
    pc.setCfs([f.cf])
    ym = pc.evalPC(x)
    m1, v1, _ = f.predict(x[:9000], msc=1)
    monkeypatch.setenv("PCQ_DEVICES", "1")
    assert pdev.devices() == [pdev.current_device()]
    assert np.array_equal(pc.evalPC(x), ym)
    m0, v0, _ = f.predict(x[:9000], msc=1)
    np.testing.assert_allclose(m1, m0, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(v1, v0, rtol=1e-10)
    one = pc.propagate(3_000_001, seed=9)
    monkeypatch.delenv("PCQ_DEVICES")
    many = pc.propagate(3_000_001, seed=9)
    np.testing.assert_allclose(many["sum"], one["sum"], rtol=1e-12)
    np.testing.assert_allclose(many["sumsq"], one["sumsq"], rtol=1e-12)
    assert many["min"][0] == one["min"][0] and many["max"][0] == one["max"][0]
    # BCS incl. its residual pass, and the multi-output workflow
    g3, c = load_golden("c3_shape"), gen.shape_inputs(3)
    mi3 = orc.get_mi(c["p"], c["d"])
    pc3 = PCRV(1, c["d"], c["fam"], mi=mi3)
    y3 = gen.sparse_target(pc3.evalBases(c["x"], 0), c["idx"], c["val"], c["noise"])
    b = bcs(eta=1e-8)
    b.setBasisEvaluator(PCBasisEvaluator(0), (pc3,))
    with contextlib.redirect_stdout(io.StringIO()):
        b.fit(c["x"], y3)
    assert np.array_equal(b.used, g3["used"])
    np.testing.assert_allclose(b.cf, g3["cf"], rtol=1e-10, atol=1e-13)
    wf = load_golden("workflow")
    for method in ("lsq", "anl", "bcs"):
        with contextlib.redirect_stdout(io.StringIO()):
            opc, _ = pc_fit(wf["x"], wf["y"], order=3, pctype="LU", method=method)
        for j in range(2):
            np.testing.assert_allclose(opc.coefs[j], wf[f"{method}__cf{j}"], rtol=1e-10, atol=1e-12)
    assert _native.lib().pcq_launch_count() > l0
