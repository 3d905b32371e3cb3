"""Analytical Bayesian linear regression on sufficient statistics.

Drop-in for the reference's ``lreg/anl.py``: ``anl(datavar, prior_var, cov_nugget, method)``
with ``fita`` (69-105), ``_compute_sigmahatsq`` (46-67) and ``compute_evidence`` (109-138).
Every quantity there is a function of (G = A^T A, b = A^T y, y^T y, N), which K3 / K4
deliver without a second pass over the samples:
    y . (y - A cf) = y^T y - b . cf .
"""
import sys

import numpy as np

from .lreg import lreg
from .stats import SuffStats


def _inverse(mat):
    """K x K inverse (lreg/anl.py:82).  cuSOLVER through torch.linalg on the device when one is present
    (a plain library call for the step that is not sample-parallel), numpy otherwise (host-logic tests);
    singular matrices raise np.linalg.LinAlgError like the reference."""
    import torch
    if torch.cuda.is_available():
        from .. import device as _device
        t = torch.from_numpy(np.ascontiguousarray(mat)).to(torch.device("cuda", _device.current_device()))
        inv, info = torch.linalg.inv_ex(t)
        if int(info) != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        return inv.cpu().numpy()
    return np.linalg.inv(mat)


class anl(lreg):
    def __init__(self, datavar=None, prior_var=None, cov_nugget=0.0, method='full'):
        super().__init__()
        self.prior_var = prior_var
        self.datavar = datavar
        self.cov_nugget = cov_nugget
        self.method = method
        if self.prior_var is not None:
            assert(self.datavar is not None)
            assert(self.method == 'full')

    def _sigmahatsq(self, stats):
        # lreg/anl.py:57-66 with y.(y - A cf) written on the statistics
        bp = stats.y_dot_residual(self.cf, 0) / 2.
        ap = (stats.n - stats.K) / 2.
        return max(bp / (ap - 1.), 0.0)

    def _compute_sigmahatsq(self, Amat, y):
        r"""Best variance estimate for a materialised A (same signature as the reference)."""
        return self._sigmahatsq(SuffStats.from_matrix(Amat, y))

    def _fit_stats(self, stats):
        nbas = stats.K
        ptp = np.array(stats.G, dtype=np.float64, copy=True)
        if self.prior_var is not None:
            ptp += (self.datavar / self.prior_var) * np.eye(nbas)
        ptp += self.cov_nugget * np.eye(nbas)
        invptp = _inverse(ptp)               # LinAlgError on singular G, as in the reference
        self.cf = np.dot(invptp, stats.b(0))
        sigmahatsq = self._sigmahatsq(stats) if self.datavar is None else self.datavar
        self.datavar = sigmahatsq + 0.0
        if self.method == 'full':
            self.cf_cov = sigmahatsq * invptp
        elif self.method == 'vi':
            self.cf_cov = np.diag(sigmahatsq / (np.diag(stats.G) + self.cov_nugget))
        else:
            print(f"Method {self.method} unknown. Exiting.")
            sys.exit()
        self.fitted = True
        self.used = np.arange(nbas)

    def compute_evidence(self, Amat, ydata):
        """Log-evidence (lreg/anl.py:109-138) from the statistics of (Amat, ydata)."""
        assert(self.prior_var is not None)
        assert(self.fitted)
        stats = SuffStats.from_matrix(Amat, ydata)
        npt, nbas = int(stats.n), stats.K
        det_cf_cov = np.linalg.det(self.cf_cov)
        if det_cf_cov <= 0.0 or self.datavar <= 0.0:
            return -sys.float_info.max
        quad = np.dot(self.cf, np.dot((1. / self.datavar) * stats.G + (1. / self.prior_var) * np.eye(nbas), self.cf))
        evid = 0.5 * np.log(det_cf_cov)
        evid += -0.5 * nbas * np.log(self.prior_var)
        evid += -0.5 * npt * np.log(2. * np.pi * self.datavar)
        evid += -(stats.yty(0) - self.datavar * quad) / (2. * self.datavar)
        return evid
