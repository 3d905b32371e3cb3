"""Pins the CPU oracle (oracle/pc_oracle.py) to the reference: golden vectors produced by the
unmodified PyTUQ (tests/golden/make_golden.py) and the known-answer vectors of SURVEY.md
Appendix A.  Bit-exact where the arithmetic order is defined, 1e-12 for LAPACK solves."""
import numpy as np

from conftest import load_golden, case_mis_cfs, case_types
from oracle import pc_oracle as orc


def test_pc1d_families_bitexact():
    g = load_golden("pc1d")
    for fam in orc.FAMILIES:
        x = g[f"{fam}_x"]
        vals = np.array(orc.pc1d(fam, x, 7))
        assert np.array_equal(vals, g[f"{fam}_vals"]), fam
        assert np.array_equal([orc.normsq_1d(fam, n) for n in range(8)], g[f"{fam}_normsq"])
        assert np.array_equal([float(orc.rec_a(fam, n)) for n in range(1, 8)], g[f"{fam}_a"])
        assert np.array_equal([float(orc.rec_b(fam, n)) for n in range(1, 8)], g[f"{fam}_b"])


def test_appendix_a_known_answers():
    x = np.array([0.45])
    expect = {
        "LU": [1, 0.45, -0.19624999999999998, -0.44718749999999996, -0.20497265625, 0.19172214843750002],
        "nLU": [1, 0.7794228634059948, -0.43882834058433373, -1.1831469144166966, -0.61491796875, 0.63587043036801],
        "HG": [1, 0.45, -0.7975, -1.258875, 1.82600625, 5.8572028125],
        "nHG": [1, 0.45, -0.39875, -0.20981249999999999, 0.07608359375, 0.0488100234375],
    }
    for fam, ref in expect.items():
        got = [float(v[0]) for v in orc.pc1d(fam, x, 5)]
        np.testing.assert_allclose(got, ref, rtol=1e-15, atol=0)
    mi = orc.get_mi(2, 3)
    assert ["".join(map(str, r)) for r in mi] == "000,100,010,001,200,110,101,020,011,002".split(",")
    assert orc.get_mi(3, 2).T.tolist() == [[0, 1, 0, 2, 1, 0, 3, 2, 1, 0], [0, 0, 1, 0, 1, 2, 0, 1, 2, 3]]
    X = np.array([[0.1, -0.2, 0.3], [0.5, 0.6, -0.7]])
    A = orc.eval_bases(X, mi, "LU")
    np.testing.assert_allclose(A[0], [1, 0.1, -0.2, 0.3, -0.485, -0.02, 0.03, -0.44, -0.06, -0.365], rtol=2e-15)
    A = orc.eval_bases(X, mi, "HG")
    np.testing.assert_allclose(A[1], [1, 0.5, 0.6, -0.7, -0.75, 0.3, -0.35, -0.64, -0.42, -0.51], rtol=2e-15)
    types = ["LU", "HG", "nLU"]
    cfs = np.arange(20.).reshape(2, 10) / 10
    y = orc.eval_pc(X, [mi, mi], [cfs[0], cfs[1]], types)
    np.testing.assert_allclose(y, [[-1.5366252821545485, -2.4501363758983747],
                                   [-1.0145020332808654, -1.400140745694665]], rtol=1e-15)
    np.testing.assert_allclose(orc.norms_sq(mi, types), [1, 1 / 3, 1, 1, 0.2, 1 / 3, 1 / 3, 2, 1, 1], rtol=1e-15)
    np.testing.assert_allclose([orc.pc_var(mi, c, types) for c in cfs], [2.7986666666666666, 18.158666666666665], rtol=1e-15)
    np.testing.assert_allclose(orc.sobol_main(mi, cfs[0], types),
                               [0.0126250595521677, 0.36445926631729386, 0.3215817055740829], rtol=1e-14)
    np.testing.assert_allclose(orc.sobol_total(mi, cfs[1], types),
                               [0.13209486746457158, 0.6173360746016594, 0.5172920185035613], rtol=1e-14)
    for (p, d, n) in [(4, 10, 1001), (3, 20, 1771), (2, 50, 1326), (4, 20, 10626), (3, 30, 5456)]:
        assert orc.get_npc(p, d) == n


def test_get_mi_matches_reference():
    g = load_golden("mindex")
    for key in g.files:
        if key.startswith("mi_"):
            _, p, d = key.split("_")
            got = orc.get_mi(int(p), int(d))
            assert got.dtype == g[key].dtype and np.array_equal(got, g[key]), key
    for p, d, n in g["npc"]:
        assert orc.get_npc(int(p), int(d)) == n


def test_eval_bases_and_pc_bitexact(pcrv_cases):
    for name, c in pcrv_cases.items():
        mis, cfs = case_mis_cfs(c)
        types = case_types(c)
        max_ord = np.max([m.max(axis=0) for m in mis], axis=0)
        for j, mi in enumerate(mis):
            A = orc.eval_bases(c["x"], mi, types, max_ord)
            assert np.array_equal(A, c[f"A{j}"]), (name, j)
            assert np.array_equal(orc.norms_sq(mi, types), c[f"normsq{j}"]), (name, j)
        y = orc.eval_pc(c["x"], mis, cfs, types, max_ord)
        assert np.array_equal(y, c["y"]), name


def test_moments_and_sobol_bitexact(pcrv_cases):
    for name, c in pcrv_cases.items():
        mis, cfs = case_mis_cfs(c)
        types = case_types(c)
        for j, (mi, cf) in enumerate(zip(mis, cfs)):
            assert orc.pc_mean(mi, cf) == c["mean"][j], name
            assert orc.pc_var(mi, cf, types) == c["var"][j], name
            assert np.array_equal(orc.sobol_main(mi, cf, types), c["main"][j]), name
            assert np.array_equal(orc.sobol_total(mi, cf, types), c["total"][j]), name
            if "joint" in c:
                assert np.array_equal(orc.sobol_joint(mi, cf, types), c["joint"][j]), name
            if mi.shape[1] > 2:
                assert orc.sobol_group(mi, cf, types, [0, 1]) == c["group01"][j], name


def test_fitters_match_reference(fit_cases):
    for name, c in fit_cases.items():
        A = orc.eval_bases(c["x"], c["mi"], str(c["type"]))
        y = c["y"]
        np.testing.assert_allclose(orc.lsq_fit(A, y), c["lsq_cf"], rtol=1e-12, atol=1e-14)
        cf, cov, dv = orc.anl_fit(A, y)
        np.testing.assert_allclose(cf, c["anl_cf"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(cov, c["anl_cov"], rtol=1e-10, atol=1e-18)
        np.testing.assert_allclose(dv, c["anl_datavar"], rtol=1e-12)
        cf, cov, _ = orc.anl_fit(A, y, datavar=0.01, prior_var=2.0, cov_nugget=1e-8)
        np.testing.assert_allclose(cf, c["anlp_cf"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(cov, c["anlp_cov"], rtol=1e-10, atol=1e-18)
        _, cov, _ = orc.anl_fit(A, y, method="vi")
        np.testing.assert_allclose(np.diag(cov), c["anlvi_cov_diag"], rtol=1e-12)
        for tag in ("bcs0.001", "bcs1e-08"):
            eta = float(tag[3:])
            w, err, used, s2, Sig = orc.bcs_fit(A, y, eta=eta)
            assert np.array_equal(used, c[f"{tag}_used"]), (name, tag)
            np.testing.assert_allclose(w, c[f"{tag}_cf"], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(s2, c[f"{tag}_datavar"], rtol=1e-12)
            np.testing.assert_allclose(Sig, c[f"{tag}_cov"], rtol=1e-10, atol=1e-20)


def test_lsq_on_illconditioned_systems_matches_reference():
    """Square, underdetermined, nearly collinear and rank-deficient systems (lsq_illcond.npz, generated from the
    unmodified reference): the oracle's gelsd restatement reproduces the reference's coefficients."""
    from conftest import load_golden, golden_cases
    cases = golden_cases(load_golden("lsq_illcond"))
    assert set(cases) == {"square_lu_d4p4", "under_lu_d4p4", "collinear_hg_d3p5", "rankdef_lu_d3p2"}
    for name, c in cases.items():
        A = orc.eval_bases(c["x"], c["mi"], str(c["type"]))
        np.testing.assert_allclose(np.linalg.svd(A, compute_uv=False), c["sv"], rtol=1e-9, atol=1e-12 * c["sv"][0])
        cf = orc.lsq_fit(A, c["y"])
        # two runs of the same LAPACK driver on bit-identical A: equal up to threading effects
        np.testing.assert_allclose(A @ cf, A @ c["lsq_cf"], rtol=0, atol=1e-9 * np.abs(c["y"]).max(), err_msg=name)
        if name != "collinear_hg_d3p5":
            np.testing.assert_allclose(cf, c["lsq_cf"], rtol=1e-8, atol=1e-10, err_msg=name)


def test_philox_known_answer():
    # Random123 known-answer test: counter = key = 0 and the all-ones / pi vectors
    out = orc.philox4x32_10((0, 0, 0, 0), (0, 0))
    assert [int(v) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = orc.philox4x32_10((0xffffffff,) * 4, (0xffffffff, 0xffffffff))
    assert [int(v) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = orc.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))
    assert [int(v) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    x = orc.germ_philox(4, 0, 20000, [0, 1, 1, 0, 1])
    assert x.shape == (20000, 5)
    assert abs(x[:, 0].mean()) < 0.02 and abs(x[:, 0].var() - 1 / 3) < 0.01
    assert abs(x[:, 1].mean()) < 0.03 and abs(x[:, 1].var() - 1) < 0.04
    assert abs(np.corrcoef(x[:, 1], x[:, 2])[0, 1]) < 0.03


def test_oracle_at_full_size_baseline_config_1():
    """BASELINE.json configs[0] at FULL size (LU, order 4, d=10, K=1001, lsq + anl fit of the Genz oscillatory
    function on 1e5 samples): the oracle against the unmodified reference (tests/golden/c1_full.npz)."""
    from conftest import c1_inputs
    g = load_golden("c1_full")
    x, y = c1_inputs()
    np.testing.assert_allclose(y[:16], g["y_head"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(y.sum(), g["y_sum"], rtol=1e-13)
    mi = orc.get_mi(4, 10)
    A = orc.eval_bases(x, mi, "LU")
    assert A.shape == (100_000, 1001)
    assert np.array_equal(A[g["rows"]], g["A_rows"])
    scale = np.max(np.abs(g["cf_lsq"]))
    np.testing.assert_allclose(orc.lsq_fit(A, y), g["cf_lsq"], rtol=0, atol=1e-12 * scale)
    cf, cov, sig = orc.anl_fit(A, y)
    np.testing.assert_allclose(cf, g["cf_anl"], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(sig, g["anl_datavar"], rtol=1e-12)
    np.testing.assert_allclose(np.diag(cov), g["anl_cov_diag"], rtol=1e-11)
    np.testing.assert_allclose(orc.sobol_main(mi, g["cf_lsq"], "LU"), g["main"][0], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(orc.sobol_total(mi, g["cf_lsq"], "LU"), g["total"][0], rtol=1e-12, atol=1e-15)


def test_oracle_at_the_headline_shape_c2():
    """BASELINE.json configs[1] shape (HG d=20 p=3 K=1771, lsq + anl on 2e4 Gaussian samples): the oracle against
    the unmodified reference (tests/golden/c2_shape.npz, make_golden.py:c2_shape)."""
    from conftest import golden_generator
    g = load_golden("c2_shape")
    x, cdraw, noise = golden_generator().c2_inputs()
    mi = orc.get_mi(3, 20)
    assert mi.shape == (1771, 20)
    np.testing.assert_array_equal(g["c_true"], cdraw / (1.0 + mi.sum(axis=1)) ** 2)
    A = orc.eval_bases(x, mi, "HG")
    assert np.array_equal(A[g["rows"]], g["A_rows"])
    y = np.zeros(x.shape[0])
    for k in range(mi.shape[0]):
        y = y + g["c_true"][k] * A[:, k]
    assert np.array_equal(y + noise, g["y"])
    scale = np.max(np.abs(g["cf_lsq"]))
    np.testing.assert_allclose(orc.lsq_fit(A, g["y"]), g["cf_lsq"], rtol=0, atol=1e-12 * scale)
    cf, cov, sig = orc.anl_fit(A, g["y"])
    np.testing.assert_allclose(cf, g["cf_anl"], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(sig, g["anl_datavar"], rtol=1e-11)
    np.testing.assert_allclose(np.diag(cov), g["anl_cov_diag"], rtol=1e-10)
    np.testing.assert_allclose(orc.sobol_main(mi, g["cf_lsq"], "HG"), g["main"][0], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(orc.sobol_total(mi, g["cf_lsq"], "HG"), g["total"][0], rtol=1e-12, atol=1e-15)


def test_oracle_at_the_shapes_of_baseline_configs_3_4_5():
    """Oracle vs the unmodified reference at the (d, p, K) of BASELINE.json configs 3, 4, 5 with reduced sample
    counts (tests/golden/c{3,4,5}_shape.npz, make_golden.py:config_shapes).  C5's 25000 x 5456 design matrix is
    checked on its statistics-free parts only here (rows + Sobol); its fit is the GPU test's job."""
    from conftest import golden_generator
    gen = golden_generator()
    # C3: BCS, LU d=50 p=2
    g, c = load_golden("c3_shape"), gen.shape_inputs(3)
    mi = orc.get_mi(c["p"], c["d"])
    A = orc.eval_bases(c["x"], mi, c["fam"])
    assert np.array_equal(A[g["rows"]], g["A_rows"])
    y = gen.sparse_target(A, c["idx"], c["val"], c["noise"])
    assert np.array_equal(y[:16], g["y_head"])
    weights, errbars, used, sigma2, Sig = orc.bcs_fit(A, y, eta=1e-8)
    assert np.array_equal(used, g["used"]) and set(used) == set(c["idx"])
    np.testing.assert_allclose(weights, g["cf"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(np.reshape(sigma2, -1), g["datavar"], rtol=1e-10)
    # C4: evalPC of the d=20 p=4 surrogate (K=10626), moments and Sobol indices
    g, c = load_golden("c4_shape"), gen.shape_inputs(4)
    mi = orc.get_mi(c["p"], c["d"])
    assert mi.shape == (10626, 20)
    assert np.array_equal(orc.eval_pc(c["x"][:256], [mi], [c["cf"]], c["fam"]), g["y"][:256])
    np.testing.assert_allclose(orc.pc_mean(mi, c["cf"]), g["mean"][0], rtol=1e-14)
    np.testing.assert_allclose(orc.pc_var(mi, c["cf"], c["fam"]), g["var"][0], rtol=1e-12)
    np.testing.assert_allclose(orc.sobol_main(mi, c["cf"], c["fam"]), g["main"][0], rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(orc.sobol_total(mi, c["cf"], c["fam"]), g["total"][0], rtol=1e-11, atol=1e-15)
    # C5: Sobol indices of the fitted multi-output surrogate from its coefficients
    g, c = load_golden("c5_shape"), gen.shape_inputs(5)
    mi = orc.get_mi(c["p"], c["d"])
    assert g["cf"].shape == (5456, 4)
    for j in range(4):
        np.testing.assert_allclose(orc.sobol_main(mi, g["cf"][:, j], c["fam"]), g["main"][j], rtol=1e-11, atol=1e-15)
        np.testing.assert_allclose(orc.sobol_total(mi, g["cf"][:, j], c["fam"]), g["total"][j], rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(orc.eval_pc(c["x"][:64], [mi] * 4, list(g["cf"].T), c["fam"]), g["eval"], rtol=1e-13, atol=1e-14)
