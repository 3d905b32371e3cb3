"""The behaviours the reference's own unit tests pin for this path, asserted on the drop-in.

Each test names the reference test it mirrors (paths under /root/reference/tests/); inputs use the same
seeds / constructions, the assertions are the same tolerances.  They run on the GPU path (the reference
runs them on numpy)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


# --- tests/test_pcrv.py ------------------------------------------------------------------------------
def test_pc1d_values_norms_quadrature():
    """test_pcrv.py:8-48 (P0(0)=1, P1(0)=0; normsq LU = 1/(2n+1), HG = n!; 5-point Gauss rule)."""
    from pytuq_b200.rv.pcrv import PC1d
    leg = PC1d('LU')
    b = leg(np.array([0.0]), 2)
    assert np.isclose(b[0][0], 1.0) and np.isclose(b[1][0], 0.0)
    for n in range(5):
        assert np.isclose(leg.normsq(n), 1.0 / (2 * n + 1))
        assert np.isclose(PC1d('HG').normsq(n), float(math.factorial(n)))
    pts, wts = leg.quad(5)
    assert np.isclose(np.sum(wts), 1.0)
    assert np.isclose(np.sum(wts * pts ** 2), 1.0 / 3.0, atol=1.e-10)


def test_pcrv_moments_and_sampling():
    """test_pcrv.py:51-120 (mean / variance of a 1-D PC; sample shape; PCRV_mvn moments, 50k samples)."""
    from pytuq_b200.rv.pcrv import PC1d, PCRV, PCRV_mvn
    pc = PCRV(1, 1, 'LU', mi=np.array([[0], [1], [2]]), cfs=np.array([[3.0, 2.0, 1.0]]))
    assert np.isclose(pc.computeMean()[0], 3.0)
    leg = PC1d('LU')
    assert np.isclose(pc.computeVar()[0], 2.0 ** 2 * leg.normsq(1) + 1.0 ** 2 * leg.normsq(2))
    mi = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    pc2 = PCRV(2, 3, 'LU', mi=mi, cfs=np.random.randn(2, 4))
    assert pc2.sample(100, seed=42).shape == (100, 2)
    mean = np.array([1.0, 2.0])
    cov = np.array([[1.0, 0.3], [0.3, 2.0]])
    mvn = PCRV_mvn(2, mean=mean, cov=cov)
    assert np.allclose(mvn.computeMean(), mean)
    assert np.allclose(mvn.computeVar(), np.diag(cov), atol=1.e-10)
    sam = mvn.sample(50000, seed=42)
    assert np.allclose(np.mean(sam, axis=0), mean, atol=0.05)
    assert np.allclose(np.cov(sam.T), cov, atol=0.1)


# --- tests/test_lreg.py ------------------------------------------------------------------------------
def test_lsq_exact_recoveries():
    """test_lreg.py:8-95: exact linear models are recovered to 1e-10 through fita(Amat, y) (K4 + solve)."""
    from pytuq_b200.lreg.lreg import lsq
    np.random.seed(42)
    x = np.random.rand(50, 1)
    A = np.column_stack([np.ones(50), x[:, 0]])
    reg = lsq()
    reg.fita(A, 3.0 + 2.0 * x[:, 0])
    assert reg.fitted and np.allclose(reg.cf, [3.0, 2.0], atol=1.e-10)
    xt = np.random.rand(20, 1)
    pred, var, cov = reg.predicta(np.column_stack([np.ones(20), xt[:, 0]]))
    assert np.allclose(pred, 3.0 + 2.0 * xt[:, 0], atol=1.e-10) and var is None and cov is None

    np.random.seed(42)
    x = np.random.rand(100)
    reg = lsq()
    reg.fita(np.column_stack([np.ones(100), x]), 1.0 + 2.0 * x + 0.01 * np.random.randn(100))
    assert np.allclose(reg.cf, [1.0, 2.0], atol=0.05)

    np.random.seed(42)
    x = np.random.rand(100, 2)
    reg = lsq()
    reg.fita(np.column_stack([np.ones(100), x]), 1.0 + 2.0 * x[:, 0] + 3.0 * x[:, 1])
    assert np.allclose(reg.cf, [1.0, 2.0, 3.0], atol=1.e-10)

    np.random.seed(42)
    A = np.random.rand(50, 4)
    reg = lsq()
    reg.fita(A, A @ np.array([1., 2., 3., 4.]))
    assert np.array_equal(reg.used, np.arange(4)) and np.allclose(reg.cf, [1., 2., 3., 4.], atol=1e-9)


# --- tests/test_pce.py -------------------------------------------------------------------------------
def test_pce_builds_and_reproduces_polynomials():
    """test_pce.py:6-100: lsq builds, polynomial targets reproduced to 1e-8, result keys, term count."""
    from pytuq_b200.surrogates.pce import PCE
    np.random.seed(42)
    x = np.random.rand(100, 2) * 2 - 1
    pce = PCE(2, 3, 'LU', verbose=0)
    pce.set_training_data(x, 1.0 + 2.0 * x[:, 0] + 3.0 * x[:, 1])
    assert pce.build(regression='lsq') is not None

    np.random.seed(42)
    x = np.random.rand(50, 1) * 2 - 1
    y = 1.0 + 2.0 * x[:, 0] + 0.5 * (3 * x[:, 0] ** 2 - 1) / 2
    pce = PCE(1, 3, 'LU', verbose=0)
    pce.set_training_data(x, y)
    pce.build(regression='lsq')
    res = pce.evaluate(x)
    assert np.allclose(res['Y_eval'], y, atol=1.e-8)
    assert 'Y_eval' in res and 'Y_eval_std' in res

    np.random.seed(42)
    x = np.random.rand(50, 2) * 2 - 1
    pce = PCE(2, 1, 'LU', verbose=0)
    pce.set_training_data(x, 3.0 + 2.0 * x[:, 0] - 1.0 * x[:, 1])
    pce.build(regression='lsq')
    xt = np.random.rand(30, 2) * 2 - 1
    assert np.allclose(pce.evaluate(xt)['Y_eval'], 3.0 + 2.0 * xt[:, 0] - 1.0 * xt[:, 1], atol=1.e-8)
    assert PCE(3, 2, 'LU', verbose=0).get_pc_terms()[0] == 10
    with pytest.raises(RuntimeError):
        PCE(2, 2, 'LU').build()
    with pytest.raises(ValueError):
        PCE(2, 2, 'LU').set_training_data(np.zeros(5), np.zeros(5))


# --- tests/test_mindex.py, test_mi.py ------------------------------------------------------------------
def test_mindex_shapes_and_growth():
    """test_mindex.py / test_mi.py: shapes, zero first row, total order bound, known counts, front growth."""
    from pytuq_b200.utils.mindex import get_mi, get_npc, mi_addfront
    for p, d in [(2, 3), (3, 2), (4, 4), (1, 6)]:
        mi = get_mi(p, d)
        assert mi.shape == (get_npc(p, d), d) and not mi[0].any() and mi.sum(axis=1).max() <= p
    assert get_npc(2, 2) == 6 and get_npc(3, 3) == 20 and get_npc(0, 5) == 1
    grown = mi_addfront(get_mi(1, 2))[0]
    assert grown.shape[0] > 3
