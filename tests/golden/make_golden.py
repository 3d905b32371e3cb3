#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (PyTUQ) in this container.

Run once, here (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

matplotlib is absent from the image and PyTUQ imports it at module import time
(func/func.py:7, utils/plotting.py:13-16, lreg/lreg.py:6), so it is stubbed in
sys.modules before the import.  Nothing on the arithmetic path touches matplotlib.
The committed .npz files are what tests/ compares the oracle and the CUDA path with.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = os.environ.get("PYTUQ_REF", "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.lines",
                 "matplotlib.patches", "matplotlib.ticker", "matplotlib.cm", "matplotlib.gridspec",
                 "matplotlib.collections", "matplotlib.tri", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "mpl_toolkits.axes_grid1"):
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, REF)
    from pytuq.rv.pcrv import PCRV, PC1d                      # noqa
    from pytuq.utils.mindex import get_mi, get_npc            # noqa
    from pytuq.lreg.lreg import lsq                           # noqa
    from pytuq.lreg.anl import anl                            # noqa
    from pytuq.lreg.bcs import bcs, bcs_fit                   # noqa
    from pytuq.workflows.fits import pc_fit                   # noqa
    from pytuq.surrogates.pce import PCE                      # noqa
    return dict(PCRV=PCRV, PC1d=PC1d, get_mi=get_mi, get_npc=get_npc, lsq=lsq, anl=anl,
                bcs=bcs, bcs_fit=bcs_fit, pc_fit=pc_fit, PCE=PCE)


def germ(rng, fam, shape):
    return rng.uniform(-1, 1, shape) if fam in ("LU", "nLU") else rng.standard_normal(shape)


def illcond(R):
    """lsq on ill-conditioned / square / underdetermined / rank-deficient systems: the cases where the
    reference's SVD least squares (scipy gelsd, cond=1e-13, lreg/lreg.py:335) and a normal-equation solve
    part ways.  Written to lsq_illcond.npz."""
    PCRV, get_mi = R["PCRV"], R["get_mi"]
    out = {}

    def case(name, pctype, mi, x, y):
        pc = PCRV(1, x.shape[1], pctype, mi=mi)
        A = pc.evalBases(x, 0)
        l = R["lsq"](); l.fita(A, y)
        sv = np.linalg.svd(A, compute_uv=False)
        out[f"{name}__x"] = x; out[f"{name}__y"] = y; out[f"{name}__mi"] = mi
        out[f"{name}__type"] = np.array(pctype); out[f"{name}__lsq_cf"] = l.cf; out[f"{name}__sv"] = sv
        print(f"  {name}: N={x.shape[0]} K={mi.shape[0]} cond={sv[0] / max(sv[-1], 1e-300):.3g}")

    rng = np.random.default_rng(41)
    f = lambda x: np.cos(1.0 + x @ (1.0 / (1.0 + np.arange(x.shape[1])) ** 2)) + 0.1 * x[:, 0] ** 3
    mi = get_mi(4, 4)                                    # K = 70
    x = rng.uniform(-1, 1, (70, 4));  case("square_lu_d4p4", "LU", mi, x, f(x))
    x = rng.uniform(-1, 1, (40, 4));  case("under_lu_d4p4", "LU", mi, x, f(x))
    mi = get_mi(5, 3)                                    # K = 56, narrow germ: nearly collinear columns
    x = 0.05 * rng.standard_normal((300, 3));  case("collinear_hg_d3p5", "HG", mi, x, f(x))
    mi = np.vstack([get_mi(2, 3), get_mi(2, 3)[3:5]])    # two duplicated columns: rank deficient
    x = rng.uniform(-1, 1, (200, 3));  case("rankdef_lu_d3p2", "LU", mi, x, f(x) + 1e-3 * rng.standard_normal(200))
    np.savez_compressed(os.path.join(HERE, "lsq_illcond.npz"), **out)
    print(f"lsq_illcond.npz  {os.path.getsize(os.path.join(HERE, 'lsq_illcond.npz')) / 1024:.1f} KiB")


def klsurr(R):
    """KLE and the PC branch of KLSurr (linred/kle.py, linred/klsurr.py:61-136,247-306).  klsurr.py itself cannot
    be imported without QUiNN (hard sys.exit at import), so its PC branch is replayed here line by line with the
    reference's own KLE, pc_fit, micf_join and PCRV.  Written to klsurr.npz; only quantities that do not depend
    on the (arbitrary) signs of the eigenvectors are compared by the tests."""
    import io, contextlib
    from pytuq.linred.kle import KLE
    from pytuq.utils.mindex import micf_join
    PCRV = R["PCRV"]
    rng = np.random.default_rng(51)
    N, d, nt = 300, 3, 25
    x = rng.uniform(-1, 1, (N, d))
    t = np.linspace(0, 1, nt)
    y = (np.sin(x[:, :1] + 2 * t[None, :]) + x[:, 1:2] * t[None, :] ** 2 + 0.5 * x[:, 2:3] * np.cos(5 * t[None, :])
         + 0.3 * x[:, :1] * x[:, 1:2] * np.sin(3 * t[None, :]))
    with contextlib.redirect_stdout(io.StringIO()):
        kl = KLE(); kl.build(y.T, plot=False)
        neig = kl.get_neig(0.99)
        xitrain = kl.xi[:, :neig]
        pcrv, _ = R["pc_fit"](x, xitrain, order=3, pctype='LU', method='bcs', eta=1.e-8)
        ytrain_kl = kl.eval(neig=neig)
        xi_pred = pcrv.function(x)
        ytrain_klsurr = kl.eval(xi=xi_pred, neig=neig)
        xt = rng.uniform(-1, 1, (17, d))
        ytest = kl.eval(xi=pcrv.function(xt), neig=neig)
        mindex_all, cfs_all = micf_join(pcrv.mindices, pcrv.coefs)
        pe = PCRV(neig, d, pcrv.pctypes, mi=mindex_all, cfs=cfs_all)
        cfs_glob = np.dot(np.dot(cfs_all.T, np.diag(np.sqrt(kl.eigval[:neig]))), kl.modes[:, :neig].T)
        cfs_glob[0, :] += kl.mean
        pp = PCRV(nt, d, pcrv.pctypes, mi=mindex_all, cfs=cfs_glob.T)
        out = dict(x=x, y=y, xt=xt, eigval=kl.eigval, mean=kl.mean, neig=np.array(neig), abs_xi=np.abs(xitrain),
                   relvar=kl.compute_relvar()[:, :neig], ytrain_kl=ytrain_kl, ytrain_klsurr=ytrain_klsurr,
                   ytest=ytest, sens_eig_main=pe.computeSens(), sens_eig_total=pe.computeTotSens(),
                   sens_main=pp.computeSens(), sens_total=pp.computeTotSens(), mindex_all=mindex_all,
                   abs_cfs_all=np.abs(cfs_all))
    np.savez_compressed(os.path.join(HERE, "klsurr.npz"), **out)
    print(f"klsurr.npz  neig={neig} K_union={mindex_all.shape[0]}  {os.path.getsize(os.path.join(HERE, 'klsurr.npz')) / 1024:.1f} KiB")


def uqpc_model(p):
    """The 'online_example' stand-in model of the replay below: two observables of three parameters."""
    return np.stack([np.exp(0.3 * p[:, 0]) + p[:, 1] * p[:, 2], np.sin(p[:, 0]) + 0.1 * p[:, 2] ** 2 + p[:, 1]], axis=1)


def uqpc(R):
    """apps/uqpc/uq_pc.py:120-345 (parameter-domain input, regime online_example, sampling 'rand') replayed step by
    step with the reference's own objects.  The sample files are written with np.savetxt exactly as the app does
    and committed as fixtures (tests/golden/uqpc/*.txt); everything computed from them goes to uqpc.npz."""
    import io, contextlib
    d = os.path.join(HERE, "uqpc")
    os.makedirs(d, exist_ok=True)
    f = lambda name: os.path.join(d, name)                                  # noqa: E731
    np.savetxt(f("pdomain.txt"), np.array([[-1.0, 2.0], [0.5, 1.5], [-3.0, -1.0]]))
    pdom = np.loadtxt(f("pdomain.txt")).reshape(-1, 2)                      # uq_pc.py:139
    pcf_all = np.vstack((0.5 * (pdom[:, 1] + pdom[:, 0]), np.diag(0.5 * (pdom[:, 1] - pdom[:, 0]))))
    in_pcdim, npar = pdom.shape[0], pcf_all.shape[1]
    pc = R["PCRV"](in_pcdim, in_pcdim, "LU", mi=R["get_mi"](1, in_pcdim), cfs=pcf_all.T)   # uq_pc.py:191-192
    germ_train = pc.sampleGerm(240, seed=7)                                 # uq_pc.py:206-207
    np.savetxt(f("qtrain.txt"), germ_train)
    np.savetxt(f("ptrain.txt"), pc.evalPC(germ_train))                      # uq_pc.py:217-218
    germ_test = pc.sampleGerm(50, seed=7)                                   # uq_pc.py:224-229
    np.savetxt(f("qtest.txt"), germ_test)
    np.savetxt(f("ptest.txt"), pc.evalPC(germ_test))
    ptrain = np.loadtxt(f("ptrain.txt")).reshape(-1, npar)                  # uq_pc.py:234-240
    germ_train = np.loadtxt(f("qtrain.txt")).reshape(-1, in_pcdim)
    ptest = np.loadtxt(f("ptest.txt")).reshape(-1, npar)
    germ_test = np.loadtxt(f("qtest.txt")).reshape(-1, in_pcdim)
    ytrain, ytest = uqpc_model(ptrain), uqpc_model(ptest)                   # uq_pc.py:252-258
    np.savetxt(f("ytrain.txt"), ytrain)
    np.savetxt(f("ytest.txt"), ytest)
    out = {}
    for method in ("lsq", "anl", "bcs"):
        with contextlib.redirect_stdout(io.StringIO()):
            opc, lregs = R["pc_fit"](germ_train, ytrain, order=3, pctype="LU", method=method, eta=1e-6)
        out[f"{method}__ytrain_pc"] = opc.function(germ_train)              # uq_pc.py:306-308
        out[f"{method}__ytest_pc"] = opc.function(germ_test)
        vtr, vte = np.empty(ytrain.shape), np.empty(ytest.shape)
        for iout, lreg in enumerate(lregs):                                 # uq_pc.py:311-314
            vtr[:, iout] = lreg.predicta(opc.evalBases(germ_train, iout), msc=1)[1]
            vte[:, iout] = lreg.predicta(opc.evalBases(germ_test, iout), msc=1)[1]
            out[f"{method}__mi{iout}"] = opc.mindices[iout]
            out[f"{method}__cf{iout}"] = opc.coefs[iout]
        out[f"{method}__ytrain_pc_var"], out[f"{method}__ytest_pc_var"] = vtr, vte
        out[f"{method}__relerr_train"] = np.linalg.norm(ytrain - out[f"{method}__ytrain_pc"], axis=0) / np.linalg.norm(ytrain, axis=0)
        out[f"{method}__relerr_test"] = np.linalg.norm(ytest - out[f"{method}__ytest_pc"], axis=0) / np.linalg.norm(ytest, axis=0)
        out[f"{method}__main"] = opc.computeSens()                          # uq_pc.py:332-334
        out[f"{method}__total"] = opc.computeTotSens()
        out[f"{method}__joint"] = opc.computeJointSens()
    np.savez_compressed(os.path.join(HERE, "uqpc.npz"), **out)
    print(f"uqpc.npz  {os.path.getsize(os.path.join(HERE, 'uqpc.npz')) / 1024:.1f} KiB  (+ uqpc/*.txt)")


def c1_inputs(n=100_000, d=10):
    """BASELINE.json configs[0] / SURVEY 8(d) C1 synthetic inputs (seed = config number)."""
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (n, d))
    w = 1.0 / (1 + np.arange(d)) ** 2
    noise = 1e-2 * rng.standard_normal(n)
    return x, w, noise


def c1_full(R):
    """BASELINE.json configs[0] at FULL size with the unmodified reference: Legendre PCE, total order 4, d=10
    (K=1001), lsq and anl fits of the Genz oscillatory function (func/genz.py:32-54, shift 0.25, weights
    1/(1+i)^2 on (x+1)/2) + 1e-2 noise on 1e5 uniform samples.  Stores the fitted coefficients, Sobol indices
    and three design-matrix rows; the inputs are regenerated from the seed by the test."""
    import time
    from pytuq.func.genz import GenzOscillatory
    x, w, noise = c1_inputs()
    n, d = x.shape
    genz = GenzOscillatory(shift=0.25, weights=w)
    y = genz((x + 1) / 2)[:, 0] + noise
    y_formula = np.cos(2 * np.pi * 0.25 + ((x + 1) / 2) @ w) + noise
    assert np.max(np.abs(y - y_formula)) < 1e-14
    mi = R["get_mi"](4, d)
    pc = R["PCRV"](1, d, "LU", mi=mi)
    t0 = time.perf_counter()
    A = pc.evalBases(x, 0)
    t1 = time.perf_counter()
    f_lsq = R["lsq"]()
    f_lsq.fita(A, y)
    t2 = time.perf_counter()
    f_anl = R["anl"]()
    f_anl.fita(A, y)
    t3 = time.perf_counter()
    pc.setCfs([f_lsq.cf])
    out = dict(y_head=y[:16], y_sum=np.array(y.sum()), rows=np.array([0, 1, n - 1]), A_rows=A[[0, 1, n - 1]],
               cf_lsq=f_lsq.cf, cf_anl=f_anl.cf, anl_datavar=np.array(f_anl.datavar),
               anl_cov_diag=np.diag(f_anl.cf_cov).copy(), mean=pc.computeMean(), var=pc.computeVar(),
               main=pc.computeSens(), total=pc.computeTotSens(),
               ref_seconds=np.array([t1 - t0, t2 - t1, t3 - t2]), ref_cores=np.array(os.cpu_count()))
    np.savez_compressed(os.path.join(HERE, "c1_full.npz"), **out)
    print(f"c1_full.npz  {os.path.getsize(os.path.join(HERE, 'c1_full.npz')) / 1024:.1f} KiB; reference: evalBases "
          f"{t1 - t0:.2f} s, lsq.fita {t2 - t1:.2f} s, anl.fita {t3 - t2:.2f} s on {os.cpu_count()} cores")


def gsa_model(x, params):
    """Two-output test model of the PCSobol golden (an Ishigami-like function and a polynomial)."""
    a, b = params
    y0 = np.sin(x[:, 0]) + a * np.sin(x[:, 1]) ** 2 + b * x[:, 2] ** 4 * np.sin(x[:, 0])
    y1 = x[:, 0] * x[:, 1] + 0.5 * x[:, 2] ** 2 - 0.3 * x[:, 1]
    return np.stack([y0, y1], axis=1)


def gsa_golden(R):
    """model_sens(method='PCSobol') of the unmodified reference (gsa/gsa.py:21-66, 383-457) on a 3-D, two-output model;
    numpy's global stream seeded, so the drop-in draws the same germ.  Written to gsa.npz."""
    import io, contextlib
    from pytuq.gsa.gsa import model_sens, PCSobol
    dom = np.array([[-np.pi, np.pi]] * 3)
    np.random.seed(31)
    with contextlib.redirect_stdout(io.StringIO()):
        main, tot = model_sens(gsa_model, (7.0, 0.1), dom, method="PCSobol", nsam=4000, plot=False, pctype="LU", order=5)
    np.random.seed(31)
    with contextlib.redirect_stdout(io.StringIO()):
        sm = PCSobol(dom, pctype="LU", order=5)
        xs = sm.sample(4000)
        sens = sm.compute(gsa_model(xs, (7.0, 0.1))[:, 0])
    np.savez_compressed(os.path.join(HERE, "gsa.npz"), main=main, total=tot, xsam_head=xs[:8], xsam_sum=np.array(xs.sum()),
                        jointt0=sens["jointt"], main0=sens["main"])
    print(f"gsa.npz  {os.path.getsize(os.path.join(HERE, 'gsa.npz')) / 1024:.1f} KiB")


def c2_inputs(n=20_000):
    """BASELINE.json configs[1] / SURVEY 8(d) C2 synthetic inputs at reduced N (seed = config number):
    Hermite d=20 p=3 (K=1771), Gaussian germ samples, decaying true coefficients, 1e-2 noise."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((n, 20))
    cdraw = rng.standard_normal(1771)
    noise = 1e-2 * rng.standard_normal(n)
    return x, cdraw, noise


def c2_shape(R):
    """The headline shape (HG d=20 p=3 K=1771) with the unmodified reference: evalBases, lsq.fita
    (lreg/lreg.py:328-339) and anl.fita (lreg/anl.py:69-105) on 2e4 samples.  y is stored (it is built from the
    reference's own design matrix); x is regenerated from the seed by the test."""
    import time
    x, cdraw, noise = c2_inputs()
    n = x.shape[0]
    mi = R["get_mi"](3, 20)
    c_true = cdraw / (1.0 + mi.sum(axis=1)) ** 2
    pc = R["PCRV"](1, 20, "HG", mi=mi)
    t0 = time.perf_counter()
    A = pc.evalBases(x, 0)
    t1 = time.perf_counter()
    y = np.zeros(n)
    for k in range(mi.shape[0]):               # term by term: no BLAS in the target
        y = y + c_true[k] * A[:, k]
    y = y + noise
    t1b = time.perf_counter()
    f_lsq = R["lsq"]()
    f_lsq.fita(A, y)
    t2 = time.perf_counter()
    f_anl = R["anl"]()
    f_anl.fita(A, y)
    t3 = time.perf_counter()
    pc.setCfs([f_lsq.cf])
    rows = np.array([0, 1, n - 1])
    out = dict(y=y, rows=rows, A_rows=A[rows], c_true=c_true, cf_lsq=f_lsq.cf, cf_anl=f_anl.cf,
               anl_datavar=np.array(f_anl.datavar), anl_cov_diag=np.diag(f_anl.cf_cov).copy(),
               mean=pc.computeMean(), var=pc.computeVar(), main=pc.computeSens(), total=pc.computeTotSens(),
               ref_seconds=np.array([t1 - t0, t2 - t1b, t3 - t2]), ref_cores=np.array(os.cpu_count()))
    np.savez_compressed(os.path.join(HERE, "c2_shape.npz"), **out)
    print(f"c2_shape.npz  {os.path.getsize(os.path.join(HERE, 'c2_shape.npz')) / 1024:.1f} KiB; reference: evalBases "
          f"{t1 - t0:.2f} s, lsq.fita {t2 - t1b:.2f} s, anl.fita {t3 - t2:.2f} s on {os.cpu_count()} cores")


def shape_inputs(config):
    """Reduced-N synthetic inputs at the shapes of BASELINE.json configs 3, 4, 5; only elementwise
    numpy arithmetic, so the test regenerates them bit for bit from the seeds.  Needs get_mi from the caller."""
    rng = np.random.default_rng(config)
    if config == 3:      # BCS: LU d=50 p=2 (K=1326), 40-sparse truth, 1 % noise
        n, d, p, fam = 20_000, 50, 2, "LU"
        x = rng.uniform(-1, 1, (n, d))
        idx = np.sort(rng.choice(1326, 40, replace=False))
        val = rng.standard_normal(40)
        noise = rng.standard_normal(n)
        return dict(n=n, d=d, p=p, fam=fam, x=x, idx=idx, val=val, noise=noise)
    if config == 4:      # forward propagation surrogate: HG d=20 p=4 (K=10626)
        n, d, p, fam = 2048, 20, 4, "HG"
        cf = rng.standard_normal(10626) / np.arange(1, 10627)
        x = rng.standard_normal((n, d))
        return dict(n=n, d=d, p=p, fam=fam, x=x, cf=cf)
    if config == 5:      # KL+PC multi-output: LU d=30 p=3 (K=5456), M outputs sharing one design matrix
        n, d, p, fam, m = 25_000, 30, 3, "LU", 4
        x = rng.uniform(-1, 1, (n, d))
        y = np.stack([np.exp(0.2 * (j + 1) * x[:, j]) * (1.0 + 0.5 * x[:, j + 1] * x[:, j + 2]) / (1 + j)
                      for j in range(m)], axis=1) + 1e-3 * rng.standard_normal((n, m))
        return dict(n=n, d=d, p=p, fam=fam, x=x, y=y)
    raise ValueError(config)


def sparse_target(A, idx, val, noise):
    """y = sum_j val_j A[:, idx_j] + 1 % noise, accumulated term by term (no BLAS: reproducible anywhere)."""
    y = np.zeros(A.shape[0])
    for j, v in zip(idx, val):
        y = y + v * A[:, j]
    return y + 0.01 * np.std(y) * noise


def config_shapes(R):
    """BASELINE.json configs 3, 4, 5 at their own (d, p, K) with the sample count reduced to what the reference
    finishes in about a minute here: BCS (C3), evalPC + moments + Sobol of the big surrogate (C4), multi-output
    pc_fit with anl (C5)."""
    import io, contextlib, time
    rows = np.array([0, 1, 777])
    # --- C3
    c = shape_inputs(3)
    mi = R["get_mi"](c["p"], c["d"])
    pc = R["PCRV"](1, c["d"], c["fam"], mi=mi)
    t0 = time.perf_counter()
    A = pc.evalBases(c["x"], 0)
    y = sparse_target(A, c["idx"], c["val"], c["noise"])
    fit = R["bcs"](eta=1e-8)
    t1 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        fit.fita(A, y)
    t2 = time.perf_counter()
    out = dict(rows=rows, A_rows=A[rows], y_head=y[:16], used=np.asarray(fit.used), cf=fit.cf,
               datavar=np.asarray(fit.datavar, dtype=float).reshape(-1), cov_diag=np.diag(fit.cf_cov).copy(),
               ref_seconds=np.array([t1 - t0, t2 - t1]))
    np.savez_compressed(os.path.join(HERE, "c3_shape.npz"), **out)
    print(f"c3_shape.npz  {os.path.getsize(os.path.join(HERE, 'c3_shape.npz')) / 1024:.1f} KiB; {len(fit.used)} terms kept "
          f"({len(set(fit.used) & set(c['idx']))} of the 40 true ones); bcs {t2 - t1:.1f} s")
    # --- C4
    c = shape_inputs(4)
    mi = R["get_mi"](c["p"], c["d"])
    pc = R["PCRV"](1, c["d"], c["fam"], mi=mi, cfs=c["cf"])
    t0 = time.perf_counter()
    yv = pc.evalPC(c["x"])
    t1 = time.perf_counter()
    out = dict(y=yv, mean=pc.computeMean(), var=pc.computeVar(), main=pc.computeSens(), total=pc.computeTotSens(),
               ref_seconds=np.array([t1 - t0]))
    np.savez_compressed(os.path.join(HERE, "c4_shape.npz"), **out)
    print(f"c4_shape.npz  {os.path.getsize(os.path.join(HERE, 'c4_shape.npz')) / 1024:.1f} KiB; evalPC of {c['n']} rows {t1 - t0:.1f} s")
    # --- C5
    c = shape_inputs(5)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pc, lregs = R["pc_fit"](c["x"], c["y"], order=c["p"], pctype=c["fam"], method="anl")
    t1 = time.perf_counter()
    out = dict(cf=np.stack(pc.coefs, axis=1), datavar=np.array([float(l.datavar) for l in lregs]),
               cov_diag=np.stack([np.diag(l.cf_cov) for l in lregs], axis=1),
               main=pc.computeSens(), total=pc.computeTotSens(), eval=pc.function(c["x"][:64]),
               ref_seconds=np.array([t1 - t0]))
    np.savez_compressed(os.path.join(HERE, "c5_shape.npz"), **out)
    print(f"c5_shape.npz  {os.path.getsize(os.path.join(HERE, 'c5_shape.npz')) / 1024:.1f} KiB; pc_fit(anl, {c['y'].shape[1]} outputs) {t1 - t0:.1f} s")


def main():
    R = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "illcond":
        return illcond(R)
    if len(sys.argv) > 1 and sys.argv[1] == "klsurr":
        return klsurr(R)
    if len(sys.argv) > 1 and sys.argv[1] == "c2":
        return c2_shape(R)
    if len(sys.argv) > 1 and sys.argv[1] == "gsa":
        return gsa_golden(R)
    PCRV, PC1d, get_mi = R["PCRV"], R["PC1d"], R["get_mi"]
    out = {}

    # --- 1-D families -------------------------------------------------------------------
    rng = np.random.default_rng(101)
    d1 = {}
    for fam in ("LU", "nLU", "HG", "nHG"):
        x = germ(rng, fam, 48)
        x[:3] = [0.0, 0.45, -0.45]
        vals = PC1d(fam)(x, 7)
        d1[f"{fam}_x"] = x
        d1[f"{fam}_vals"] = np.array(vals)
        d1[f"{fam}_normsq"] = np.array([PC1d(fam).normsq(n) for n in range(8)])
        d1[f"{fam}_a"] = np.array([float(PC1d(fam).a(n)) for n in range(1, 8)])
        d1[f"{fam}_b"] = np.array([float(PC1d(fam).b(n)) for n in range(1, 8)])
    np.savez_compressed(os.path.join(HERE, "pc1d.npz"), **d1)

    # --- multi-indices -------------------------------------------------------------------
    dmi = {}
    for (p, d) in [(0, 3), (1, 4), (2, 3), (3, 2), (3, 5), (4, 3), (2, 10), (5, 2), (3, 7)]:
        dmi[f"mi_{p}_{d}"] = get_mi(p, d)
    dmi["npc"] = np.array([[p, d, R["get_npc"](p, d)] for p in range(6) for d in range(1, 52, 7)])
    np.savez_compressed(os.path.join(HERE, "mindex.npz"), **dmi)

    # --- evalBases / evalPC / moments / Sobol ---------------------------------------------
    cases = {}

    def add_case(name, pctype, mis, nx, seed, strided=False):
        rng = np.random.default_rng(seed)
        sdim = mis[0].shape[1]
        types = [pctype] * sdim if isinstance(pctype, str) else pctype
        cols = [germ(rng, t, nx) for t in types]
        x = np.stack(cols, axis=1)
        if strided:
            x = np.asfortranarray(x)
        cfs = [rng.standard_normal(m.shape[0]) / (1.0 + m.sum(axis=1)) ** 2 for m in mis]
        pc = PCRV(len(mis), sdim, pctype, mi=[m.copy() for m in mis], cfs=[c.copy() for c in cfs])
        cases[f"{name}__x"] = np.ascontiguousarray(x)
        cases[f"{name}__types"] = np.array(types)
        cases[f"{name}__npdim"] = np.array(len(mis))
        for j, (m, c) in enumerate(zip(mis, cfs)):
            cases[f"{name}__mi{j}"] = m
            cases[f"{name}__cf{j}"] = c
            cases[f"{name}__A{j}"] = pc.evalBases(x, j)
            cases[f"{name}__normsq{j}"] = pc.evalBasesNormsSq(j)
        cases[f"{name}__y"] = pc.evalPC(x)
        cases[f"{name}__mean"] = pc.computeMean()
        cases[f"{name}__var"] = pc.computeVar()
        cases[f"{name}__main"] = pc.computeSens()
        cases[f"{name}__total"] = pc.computeTotSens()
        if sdim <= 12:
            cases[f"{name}__joint"] = pc.computeJointSens()
        cases[f"{name}__group01"] = pc.computeGroupSens([0, 1]) if sdim > 2 else np.zeros(len(mis))

    for fam in ("LU", "nLU", "HG", "nHG"):
        add_case(f"fam_{fam}", fam, [get_mi(4, 3)], 40, 7)
    add_case("mixed", ["LU", "HG", "nLU", "nHG"], [get_mi(3, 4), get_mi(2, 4)], 33, 8, strided=True)
    add_case("c1_lu_d10p4", "LU", [get_mi(4, 10)], 24, 1)
    add_case("c2_hg_d20p3", "HG", [get_mi(3, 20)], 20, 2)
    add_case("c3_lu_d50p2", "LU", [get_mi(2, 50)], 12, 3)
    add_case("c4_hg_d20p4", "HG", [get_mi(4, 20)], 6, 4)
    add_case("d1_p6", "LU", [get_mi(6, 1)], 17, 5)
    add_case("p0_const", "HG", [get_mi(0, 3)], 9, 6)
    # general multi-index: not downward closed, duplicate rows, dense rows (width 6 > 4)
    rng = np.random.default_rng(9)
    gen = rng.integers(0, 4, size=(37, 6))
    gen[5] = gen[11]
    gen[0] = 0
    add_case("general_dense", ["LU", "HG", "LU", "nLU", "HG", "nHG"], [gen, gen[::2].copy()], 29, 10)
    # single-dimension growth only (width 1)
    w1 = np.zeros((9, 3), dtype=int)
    w1[1:4, 0] = [1, 2, 3]; w1[4:7, 1] = [1, 2, 3]; w1[7:9, 2] = [1, 2]
    add_case("width1", "HG", [w1], 21, 11)
    np.savez_compressed(os.path.join(HERE, "pcrv_cases.npz"), **cases)

    # --- fitters ---------------------------------------------------------------------------
    fits = {}

    def fit_case(name, pctype, order, dim, N, seed, noise, sparse=0):
        rng = np.random.default_rng(seed)
        mi = get_mi(order, dim)
        K = mi.shape[0]
        x = germ(rng, pctype, (N, dim))
        pc = PCRV(1, dim, pctype, mi=mi)
        A = pc.evalBases(x, 0)
        if sparse:
            c = np.zeros(K)
            sup = rng.choice(K, sparse, replace=False)
            c[sup] = rng.standard_normal(sparse)
        else:
            c = rng.standard_normal(K) / (1.0 + mi.sum(axis=1)) ** 2
        y = A @ c + noise * rng.standard_normal(N)
        fits[f"{name}__x"] = x
        fits[f"{name}__y"] = y
        fits[f"{name}__mi"] = mi
        fits[f"{name}__type"] = np.array(pctype)
        fits[f"{name}__ctrue"] = c
        l = R["lsq"](); l.fita(A, y)
        fits[f"{name}__lsq_cf"] = l.cf
        a = R["anl"](); a.fita(A, y)
        fits[f"{name}__anl_cf"] = a.cf
        fits[f"{name}__anl_cov"] = a.cf_cov
        fits[f"{name}__anl_datavar"] = np.array(a.datavar)
        a2 = R["anl"](datavar=0.01, prior_var=2.0, cov_nugget=1e-8); a2.fita(A, y)
        fits[f"{name}__anlp_cf"] = a2.cf
        fits[f"{name}__anlp_cov"] = a2.cf_cov
        a3 = R["anl"](method="vi"); a3.fita(A, y)
        fits[f"{name}__anlvi_cov_diag"] = np.diag(a3.cf_cov)
        for eta in (1e-3, 1e-8):
            b = R["bcs"](eta=eta); b.fita(A, y)
            tag = f"{name}__bcs{eta:g}"
            fits[f"{tag}_cf"] = b.cf
            fits[f"{tag}_used"] = np.asarray(b.used)
            fits[f"{tag}_datavar"] = np.asarray(b.datavar)
            fits[f"{tag}_cov"] = b.cf_cov
        # prediction with variance on a few fresh points (lreg.predicta msc=1)
        xt = germ(rng, pctype, (11, dim))
        At = pc.evalBases(xt, 0)
        mean, var, _ = a.predicta(At, msc=1)
        fits[f"{name}__xt"] = xt
        fits[f"{name}__anl_pred_mean"] = mean
        fits[f"{name}__anl_pred_var"] = var

    fit_case("lu_d3p3", "LU", 3, 3, 400, 21, 1e-2)
    fit_case("hg_d5p3", "HG", 3, 5, 900, 22, 1e-2)
    fit_case("lu_d8p2_sparse", "LU", 2, 8, 700, 23, 1e-2, sparse=6)
    fit_case("nlu_d4p4", "nLU", 4, 4, 1200, 24, 5e-3)
    np.savez_compressed(os.path.join(HERE, "fits.npz"), **fits)

    # --- workflow: pc_fit with two outputs + Sobol -------------------------------------------
    wf = {}
    rng = np.random.default_rng(31)
    x = rng.uniform(-1, 1, (600, 4))
    ymat = np.stack([np.sin(x[:, 0]) + 0.3 * x[:, 1] * x[:, 2], x[:, 3] ** 2 + 0.1 * x[:, 0]], axis=1)
    import io, contextlib
    for method in ("lsq", "anl", "bcs"):
        with contextlib.redirect_stdout(io.StringIO()):
            pc, lregs = R["pc_fit"](x, ymat, order=3, pctype="LU", method=method)
        for j in range(2):
            wf[f"{method}__cf{j}"] = pc.coefs[j]
            wf[f"{method}__mi{j}"] = pc.mindices[j]
        wf[f"{method}__main"] = pc.computeSens()
        wf[f"{method}__total"] = pc.computeTotSens()
        wf[f"{method}__eval"] = pc.function(x[:16])
    wf["x"] = x
    wf["y"] = ymat
    np.savez_compressed(os.path.join(HERE, "workflow.npz"), **wf)

    for f in ("pc1d", "mindex", "pcrv_cases", "fits", "workflow"):
        p = os.path.join(HERE, f + ".npz")
        print(f"{f}.npz  {os.path.getsize(p) / 1024:.1f} KiB")
    illcond(R)
    klsurr(R)
    uqpc(R)
    c1_full(R)
    c2_shape(R)
    gsa_golden(R)
    config_shapes(R)


if __name__ == "__main__":
    main()
