"""Linear-regression base class and least squares on GPU-computed sufficient statistics.

Drop-in for the reference's ``lreg/lreg.py`` (``lreg`` 12-306, ``lsq`` 310-339): same
attributes after a fit (``cf, cf_cov, used, datavar, fitted``) and same methods.  What
changes is where the O(N K^2) work happens:

* ``fit(x, y)`` with a PC basis evaluator never materialises the design matrix: the fused
  kernel K3 reduces (x, y) to S = [A|y|1]^T [A|y|1] on the device (include/pcq.h) and the
  fitter works on the K x K statistics;
* ``fita(Amat, y)`` with a caller-supplied matrix reduces it with the DMMA Gram kernel K4.

The K x K factorisation itself is host LAPACK (it is not sample-parallel).
"""
import sys

import numpy as np
from scipy.linalg.lapack import dpotrf, dpotrs

from ..fit.fit import fitbase
from .stats import SuffStats
from .. import _native


class PCBasisEvaluator:
    """Basis evaluator ``(x, pars) -> pars[0].evalBases(x, jdim)`` that ``lreg`` recognises:
    fits through it take the fused no-A path.  Same call convention as the closure the
    reference installs in surrogates/pce.py:378-382."""

    def __init__(self, jdim=0):
        self.jdim = jdim

    def __call__(self, x, pars):
        return pars[0].evalBases(x, self.jdim)


def _current_device():
    from .. import device as _device
    return _device.current_device()


def solve_normal_device(G_t, B_t, pivot_tol=1.0e-6):
    """Cholesky solve of G C = B on the GPU (torch tensors, float64).  The K x K factorisation is not
    sample-parallel; it is a plain library call (cuSOLVER potrf/potrs through
    torch.linalg), kept on the device so that only K coefficients travel back to the host.
    Returns None when G is not numerically positive definite (callers then take the QR / eigen path).
    One host synchronisation decides acceptance: factorisation status, positive diagonal and the scaled
    pivots L_ii / sqrt(G_ii) (cond of the equilibrated G well below 1 / tol^2) are combined on the device."""
    import torch
    diag = torch.diagonal(G_t)
    L, info = torch.linalg.cholesky_ex(G_t)
    ratio = torch.diagonal(L).abs() / diag.clamp_min(0.0).sqrt()          # NaN / inf where diag <= 0
    ok = (info == 0) & (diag.min() > 0.0) & (torch.nan_to_num(ratio, nan=0.0).min() > pivot_tol)
    if not bool(ok):
        return None
    C = torch.cholesky_solve(B_t if B_t.ndim == 2 else B_t[:, None], L)
    if not bool(torch.isfinite(C.sum())):
        return None
    return C if B_t.ndim == 2 else C[:, 0]


def solve_normal(G, b):
    """Solves G c = b for symmetric positive (semi)definite G (numpy in / numpy out).

    Cholesky on the GPU; its backward error is invariant under diagonal scaling, so no explicit
    equilibration pass is needed and conditioning is judged on the scaled pivots.  When G is
    numerically singular the minimum-norm solution on the retained eigen-space is returned instead
    (the closest normal-equation analogue of the reference's SVD least squares with cond=1e-13,
    lreg/lreg.py:335 -- see DESIGN.md for the conditioning limit of that equivalence)."""
    import torch
    from .. import device as _device
    if torch.cuda.is_available():
        dev = torch.device("cuda", _device.current_device())
        G_t = torch.from_numpy(np.ascontiguousarray(G, dtype=np.float64)).to(dev)
        B_t = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).to(dev)
        C = solve_normal_device(G_t, B_t)
        if C is not None:
            return C.cpu().numpy()
    else:
        # no device (host-logic tests): LAPACK dpotrf/dpotrs with the same acceptance test
        diag = np.diag(G)
        if diag.size and np.all(diag > 0.0) and np.all(np.isfinite(diag)):
            fac = np.array(G, dtype=np.float64, order="C").T    # symmetric: same matrix, Fortran order
            L, info = dpotrf(fac, lower=0, overwrite_a=1, clean=0)
            if info == 0 and (np.abs(np.diag(L)) / np.sqrt(diag)).min() > 1.0e-6:
                c, info2 = dpotrs(L, np.asfortranarray(b), lower=0)
                if info2 == 0 and np.all(np.isfinite(c)):
                    return np.ascontiguousarray(c) if c.ndim == 2 else c
    # (near-)singular: minimum-norm solution in the original variables, like an SVD solve
    lam, V = np.linalg.eigh(G)
    keep = lam > lam[-1] * max(G.shape[0], 10) * 1.0e-14
    return (V[:, keep] / lam[keep]) @ (V[:, keep].T @ b)


class lreg(fitbase):
    """Base class for linear regression (attributes as in lreg/lreg.py:12-34)."""

    def __init__(self):
        super().__init__()
        self.cf = None
        self.cf_cov = None
        self.datavar = 0.0
        self.basisEvaluatorSet = False
        self.used = None

    def setBasisEvaluator(self, basiseval, basisevalpars):
        self.basisEval = basiseval
        self.basisEvalPars = basisevalpars
        self.basisEvaluatorSet = True
        self._route = None

    def _pc_route(self):
        """(pcrv, jdim) when the basis evaluator is a PC design-matrix evaluator, else None.  Recognised are
        PCBasisEvaluator and the reference idiom -- any callable (a plain function, the closure of
        surrogates/pce.py:378-382, examples/ex_lreg_basiseval.py:55) whose ``pars[0]`` is a PCRV and whose result on
        two probe points is bit-identical to ``pars[0].evalBases(x, j)`` for some output dimension j.  The probe is
        repeated when the evaluator, the PCRV object or its number of terms changes."""
        from ..rv.pcrv import PCRV
        if isinstance(self.basisEval, PCBasisEvaluator):
            return self.basisEvalPars[0], self.basisEval.jdim
        pars = self.basisEvalPars
        if not isinstance(pars, (tuple, list)) or len(pars) < 1 or not isinstance(pars[0], PCRV):
            return None
        pcrv = pars[0]
        key = (id(self.basisEval), id(pcrv), tuple(m.shape[0] for m in pcrv.mindices))
        cached = getattr(self, "_route", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        route = None
        try:
            xp = np.array([[0.3 + 0.05 * np.sin(1.0 + i) for i in range(pcrv.sdim)],
                           [-0.45 + 0.07 * np.cos(2.0 + 3.0 * i) for i in range(pcrv.sdim)]])
            out = np.asarray(self.basisEval(xp, pars))
            for j in range(pcrv.pdim):
                if out.shape == (2, pcrv.mindices[j].shape[0]) and np.array_equal(out, pcrv.evalBases(xp, j)):
                    route = (pcrv, j)
                    break
        except Exception:
            route = None
        self._route = (key, route)
        return route

    # -- fitting ------------------------------------------------------------------------
    def _fit_stats(self, stats):
        """Fit from sufficient statistics of one output (SuffStats with M = 1)."""
        raise NotImplementedError("Fitting not implemented in the base class.")

    def fita(self, Amat, y):
        r"""Fit given a materialised design matrix (N, K) and data (N,): K4 + _fit_stats."""
        self._fit_stats(SuffStats.from_matrix(Amat, y))

    def fit(self, x, y):
        r"""Fit with (x, y) pairs through the basis evaluator (lreg/lreg.py:62-71)."""
        assert(self.basisEvaluatorSet)
        route = self._pc_route()
        if route is not None:
            stats = SuffStats.from_pc(route[0], x, y, route[1])
            self._fit_stats(stats)
        else:
            Amat = self.basisEval(x, self.basisEvalPars)
            self.fita(Amat, y)

    def print_coefs(self):
        assert(self.fitted)
        print(self.cf)

    # -- prediction ---------------------------------------------------------------------------
    def _cov_factor(self):
        """F with cf_cov = F F^T: Cholesky, or the SVD factor when cf_cov is only semi-definite (the same
        two routes lreg.predicta takes, lreg/lreg.py:130-134).  None for an all-zero covariance."""
        C = self.cf_cov
        if C is None or not np.any(C):
            return None
        try:
            return np.linalg.cholesky(C)
        except np.linalg.LinAlgError:
            u, sv, _ = np.linalg.svd(C, hermitian=True)
            return u * np.sqrt(sv)

    def predict(self, x, msc=0, pp=False):
        r"""Mean (and variance) of the fit at x (lreg/lreg.py:80-94).  With a PC basis evaluator and
        msc in (0, 1) nothing of size N x K is formed: the mean is an evalPC (K2) and the variance is the
        quadratic form psi^T cf_cov psi -- on the tensor pipe (K6) for dense covariances with enough work,
        else as || Psi(x_n) F ||^2, cf_cov = F F^T, through K2's quadratic-form mode."""
        assert(self.basisEvaluatorSet)
        route = self._pc_route()
        if route is not None and msc in (0, 1):
            assert(self.fitted)
            pcrv, jdim = route
            if pcrv.mindices[jdim].shape[0] == self.cf.shape[0]:
                ypred = pcrv.evalLinearForms(x, self.cf[None, :], jdim)[:, 0]
                if msc == 0:
                    return ypred, None, None
                K = self.cf.shape[0]
                if self.cf_cov is None or not np.any(self.cf_cov):
                    ypred_var = np.zeros(ypred.shape[0])
                else:
                    ypred_var = None
                    if K >= 96 and float(ypred.shape[0]) * K * K >= 2.0e8:
                        # dense covariance, enough work: psi^T C psi as a triangular GEMM on the tensor pipe (K6)
                        try:
                            ypred_var = np.maximum(pcrv.evalQuadForm(x, self.cf_cov, jdim), 0.0)
                        except _native.PcqError as e:
                            if e.status != -5:         # PCQ_ERR_UNSUPPORTED (germ too wide for K6's pack kernel)
                                raise
                    if ypred_var is None:
                        F = self._cov_factor()
                        ypred_var = pcrv.evalLinearForms(x, np.ascontiguousarray(F.T), jdim, sumsq=True)
                ypred_var += int(pp) * self.datavar
                return ypred, ypred_var, None
        Amat = self.basisEval(x, self.basisEvalPars)
        return self.predicta(Amat, msc=msc, pp=pp)

    def predict_samples(self, x, nsamples=100):
        r"""Predictive samples at x (lreg/lreg.py:96-110).  With a PC basis evaluator the coefficient draws (numpy's
        global stream, as in the reference) are pushed through evalPC's many-output mode: no design matrix."""
        assert(self.basisEvaluatorSet)
        route = self._pc_route()
        if route is not None and self.fitted and route[0].mindices[route[1]].shape[0] == self.cf.shape[0]:
            assert(self.cf_cov is not None)
            draws = np.random.multivariate_normal(self.cf, self.cf_cov, nsamples)
            return route[0].evalLinearForms(x, draws, route[1], exact=False)
        Amat = self.basisEval(x, self.basisEvalPars)
        return self.predicta_samples(Amat, nsamples=nsamples)

    def predicta(self, Amat, msc=0, pp=False):
        r"""(mean, variance | None, covariance | None) for an evaluated basis (lreg/lreg.py:112-143).  The matrix is
        contracted on the device (``forms.matrix_forms``): mean = HBM-bound row dot products, variance = quadratic
        form on the tensor pipe (K6 on the packed matrix), msc=2 = two linear-forms GEMMs."""
        from .forms import matrix_forms, psd_part
        from .. import device as _device
        assert(self.fitted)
        _device.require_device()
        if msc not in (0, 1, 2):
            print(f"msc={msc}, but needs to be 0,1, or 2. Exiting.")
            sys.exit()
        ypred_var = ypred_cov = None
        N = Amat.shape[0]
        if msc == 1:
            assert(self.cf_cov is not None)
            if not np.any(self.cf_cov):
                ypred = matrix_forms(Amat, self.cf[None, :])[0][:, 0]
                ypred_var = np.zeros(N)
            else:
                Y, v = matrix_forms(Amat, self.cf[None, :], psd_part(self.cf_cov)[0])
                ypred, ypred_var = Y[:, 0], np.maximum(v, 0.0)
            ypred_var += int(pp) * self.datavar
        else:
            ypred = matrix_forms(Amat, self.cf[None, :])[0][:, 0]
            if msc == 2:
                T = matrix_forms(Amat, self.cf_cov.T)[0]              # Amat @ cf_cov       (N, K)
                ypred_cov = matrix_forms(T, Amat)[0]                  # (Amat cf_cov) Amat^T (N, N)
                ypred_cov += int(pp) * self.datavar * np.eye(N)
                ypred_var = np.diag(ypred_cov)
        return ypred, ypred_var, ypred_cov

    def predict_plot(self, aa_list, yy_list, labels=None, colors=None):
        r"""Diagonal plots of fit vs data and |residual| vs predictive standard deviation for several (A, y) sets
        (lreg/lreg.py:201-240): the predictions run on the device (``predicta(msc=1)``); drawing needs matplotlib,
        which this package does not depend on otherwise."""
        nlist = len(aa_list)
        assert(nlist == len(yy_list))
        preds, stds = [], []
        for aa in aa_list:
            yy_pred, yy_pred_var, _ = self.predicta(aa, msc=1)
            preds.append(yy_pred)
            stds.append(np.sqrt(yy_pred_var))
        labels = [f'Set {i+1}' for i in range(nlist)] if labels is None else labels
        colors = (['b', 'g', 'r', 'c', 'm', 'y'] * nlist)[:nlist] if colors is None else colors
        assert(len(labels) == nlist and len(colors) == nlist)
        try:
            import matplotlib.pyplot as plt
        except ImportError as exc:
            raise ImportError("lreg.predict_plot draws with matplotlib, which is not installed") from exc
        for fname, xs, ys, errs, axl in (
                ('fitdiag.png', yy_list, preds, stds, ('Data', 'Fit')),
                ('resid_vs_std.png', [np.abs(p - y) for y, p in zip(yy_list, preds)], stds, None, ('|Residual|', 'Fit StDev'))):
            fig, ax = plt.subplots(figsize=(8, 8))
            for i in range(nlist):
                ax.errorbar(xs[i], ys[i], yerr=None if errs is None else errs[i], fmt='o', color=colors[i], label=labels[i])
            lo = min(float(np.min(v)) for v in list(xs) + list(ys))
            hi = max(float(np.max(v)) for v in list(xs) + list(ys))
            ax.plot([lo, hi], [lo, hi], 'k--', lw=1)
            ax.set_xlabel(axl[0]); ax.set_ylabel(axl[1]); ax.legend()
            fig.savefig(fname)
            plt.close(fig)

    def predicta_samples(self, Amat, nsamples=100):
        r"""Amat @ (coefficient samples)^T (lreg/lreg.py:145-162): the draws come from numpy's global stream as in the
        reference, the (N, K) x (K, nsamples) product runs on the device."""
        from .forms import matrix_forms
        assert(self.fitted)
        assert(self.cf_cov is not None)
        draws = np.random.multivariate_normal(self.cf, self.cf_cov, nsamples)
        return matrix_forms(Amat, draws)[0]

    def compute_stdev(self, Amat, method="chol"):
        r"""Pushed-forward standard deviation sqrt(diag(A C A^T)) (lreg/lreg.py:164-198).  Every method of the
        reference is the square root of a quadratic form a^T C' a (C' = C; C + shift I for 'choleye'; u diag(s) u^T
        for 'svd'), evaluated on the tensor pipe; 'chol' raises numpy's LinAlgError for a non-positive-definite
        covariance exactly as np.linalg.cholesky does there."""
        from .forms import matrix_forms, psd_part
        from .. import device as _device
        assert(self.cf_cov is not None)
        _device.require_device()
        C = self.cf_cov
        if method == "chol":
            Cq, pd = psd_part(C)
            if not pd:
                raise np.linalg.LinAlgError("Matrix is not positive definite")
        elif method == "choleye":
            shift = abs(np.linalg.eigvalsh(C)[0]) + 1e-14
            Cq = C + shift * np.eye(C.shape[0])
        elif method == "svd":
            u, sv, _ = np.linalg.svd(C, hermitian=True)
            Cq = (u * sv) @ u.T
        elif method in ("loop", "fullcov"):
            return np.sqrt(matrix_forms(Amat, None, C)[1])
        else:
            return np.zeros(Amat.shape[0])
        return np.sqrt(np.maximum(matrix_forms(Amat, None, Cq)[1], 0.0))


class lsq(lreg):
    """Least squares.  The reference solves min ||A c - y|| with an SVD of the N x K matrix
    (scipy gelsd, cond=1e-13, lreg/lreg.py:328-339).  Here the sample axis is reduced on the GPU:

    * well-conditioned systems (every PC basis sampled from its own measure: cond(G) ~ 10..100) go through
      the sufficient statistics (K3/K4) and a Cholesky solve of the K x K normal equations, which agrees
      with the SVD solution to ~1e-14;
    * when N < 4 K, or the scaled Cholesky pivots say cond(G) is beyond ~1e6, the fit switches to the QR
      statistic (``tsqr``): same singular values as A and the same truncated minimum-norm solution as the
      reference, for square, underdetermined, nearly collinear and rank-deficient systems.

    ``solver`` ('auto' | 'gram' | 'qr') is an extension of the reference's argument-free constructor."""

    def __init__(self, solver='auto'):
        super().__init__()
        assert solver in ('auto', 'gram', 'qr')
        self.solver = solver
        self.rank = None

    def _finish(self, cf):
        K = cf.shape[0]
        self.cf = cf
        self.cf_cov = np.zeros((K, K))
        self.fitted = True
        self.used = np.arange(K)

    def _try_gram(self, stats, strict):
        """Cholesky of the normal equations on the device; False when G is not comfortably positive
        definite (strict: pivot test at 1e-3, i.e. cond(G) ~ 1e6, so that the answer matches the SVD
        solution to ~1e-10)."""
        sysdev = stats.dev_system()
        if sysdev is None:
            import torch
            if not torch.cuda.is_available():
                return False
            dev = torch.device("cuda", _current_device())
            sysdev = (torch.from_numpy(np.ascontiguousarray(stats.G)).to(dev),
                      torch.from_numpy(np.ascontiguousarray(stats.B())).to(dev))
        C = solve_normal_device(sysdev[0], sysdev[1][:, 0], 1.0e-3 if strict else 1.0e-6)
        if C is None:
            return False
        self._finish(C.cpu().numpy())
        self.rank = stats.K
        return True

    def _fit_r(self, R, K):
        from . import tsqr
        C, _, rank = tsqr.lstsq_from_r(R, K)
        self._finish(np.ascontiguousarray(C[:, 0]))
        self.rank = rank

    def _fit_stats(self, stats):
        if not self._try_gram(stats, strict=False):
            self._finish(solve_normal(stats.G, stats.b(0)))

    def fit(self, x, y):
        assert(self.basisEvaluatorSet)
        route = self._pc_route()
        if route is None:
            return self.fita(self.basisEval(x, self.basisEvalPars), y)
        from . import tsqr
        pc, jdim = route
        K = pc.mindices[jdim].shape[0]
        N = np.shape(x)[0]
        if self.solver != 'qr' and (self.solver == 'gram' or N >= 4 * K):
            stats = SuffStats.from_pc(pc, x, y, jdim)
            if self._try_gram(stats, strict=self.solver == 'auto'):
                return
            if self.solver == 'gram':
                return self._fit_stats(stats)
        self._fit_r(tsqr.shared_r(tsqr.r_factor_pc(pc, x, y, jdim)), K)

    def fita(self, Amat, y):
        from . import tsqr
        N, K = np.shape(Amat)
        if self.solver != 'qr' and (self.solver == 'gram' or N >= 4 * K):
            stats = SuffStats.from_matrix(Amat, y)
            if self._try_gram(stats, strict=self.solver == 'auto'):
                return
            if self.solver == 'gram':
                return self._fit_stats(stats)
        self._fit_r(tsqr.r_factor_matrix(Amat, y), K)
