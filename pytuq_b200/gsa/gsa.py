"""PC-based Sobol sensitivity analysis: drop-in for ``PCSobol`` (reference gsa/gsa.py:383-457).

Workflow of the reference class, kept call for call: ``sample(N)`` draws the germ of a total-order PC and maps it to
the input box, the caller evaluates the model there, ``compute(y)`` regresses the PC on (germ, y) and reads the
main / total / joint indices off the coefficients.  Here the regression is the fused statistics pass (K3) -- the
design matrix the reference builds in ``compute`` never exists -- and the indices come from the vectorised
``PCRV.compute*Sens``.  The sampling-based estimators that share the reference's file (SamSobol, Moat, Linreg,
model_sens) never touch the PC path and are out of scope."""
import sys

import numpy as np

from ..lreg.lreg import lsq, PCBasisEvaluator, solve_normal_device, solve_normal
from ..lreg.stats import SuffStats
from ..rv.pcrv import PCRV
from ..utils.mindex import get_mi


def model_sens(model, model_params, domain, method='SamSobol', nsam=100, plot=True, **kwargs):
    r"""Main and total sensitivities of a multi-output model ``model(x, params) -> (N, o)``, each (o, d)
    (reference gsa/gsa.py:21-66).  Only ``method='PCSobol'`` -- the PC path -- is provided here; the sampling estimators
    (SamSobol, LinReg, Moat) are outside this package.  Where the reference regresses output by output on the same
    design matrix, all outputs share ONE statistics pass here (``PCSobol.compute_all``).  ``plot`` is accepted and
    ignored (no matplotlib dependency)."""
    if method != "PCSobol":
        if method in ("SamSobol", "LinReg", "Moat"):
            raise NotImplementedError(f"model_sens: method '{method}' is a sampling estimator outside the PC path of this package")
        print(f'Sensitivity method {method} is unknown. Exiting.')
        sys.exit()
    estimator = PCSobol(domain, **kwargs)
    xsam = estimator.sample(nsam)
    ysam = model(xsam, model_params)
    return estimator.compute_all(ysam)


def scale01ToDom(xx, dom):
    """Affine map of points of the unit cube onto the box ``dom`` (d, 2) (reference: utils/maps.py:7-22)."""
    unit = np.asarray(xx)
    if (unit < 0.0).any() or (unit > 1.0).any():
        print("Warning: some elements are outside the [0,1] range.")
    lower = np.min(dom, axis=1)
    width = np.abs(dom[:, 1] - dom[:, 0])
    return lower + unit * width


class SensMethod():
    """Common state of the sensitivity estimators: the input box and one result slot per index family."""

    def __init__(self, dom, sens_names):
        self.dom = dom
        self.dim = len(dom)
        self.sens_names = sens_names
        self.sens = {name: [None] * self.dim for name in sens_names}

    def sample(self, nsam):
        raise NotImplementedError("Sampling for sensitivity is not implemented in the base class.")

    def compute(self, ysam):
        raise NotImplementedError("Computing sensitivity is not implemented in the base class.")


class PCSobol(SensMethod):
    r"""Sobol indices from a regression PC surrogate (Sudret 2008).  Attributes after ``sample`` / ``compute`` as in the
    reference: ``pcrv``, ``nsam``, ``germ_sam (N, d)``, ``xsam (N, d)``, ``sens`` with keys main, total, jointt."""

    def __init__(self, dom, pctype='LU', order=3):
        print("Initializing PCSobol Sensitivity Method")
        self.sens_names = ['main', 'total', 'jointt']
        SensMethod.__init__(self, dom, self.sens_names)
        self.pctype, self.order = pctype, order

    def _inputs_of(self, germ):
        """Germ samples -> model inputs: each coordinate through its own germ CDF onto [0, 1], then onto the box."""
        unit = np.column_stack([pc1.germCdf(germ[:, i]) for i, pc1 in enumerate(self.pcrv.PC1ds)])
        return scale01ToDom(unit, self.dom)

    def sample(self, nsam):
        print("Sampling PCSobol Sensitivity Method")
        self.nsam = nsam
        self.pcrv = PCRV(1, self.dim, self.pctype, mi=get_mi(self.order, self.dim))
        self.germ_sam = self.pcrv.sampleGerm(nsam=nsam)
        self.xsam = self._inputs_of(self.germ_sam)
        return self.xsam

    def compute(self, ysam):
        assert(ysam.shape[0] == self.nsam)
        regression = lsq()
        regression.setBasisEvaluator(PCBasisEvaluator(0), (self.pcrv,))
        regression.fit(self.germ_sam, ysam)                    # statistics pass on the device, no design matrix
        self.pcrv.setCfs([regression.cf])
        for key, indices in (('main', self.pcrv.computeSens), ('total', self.pcrv.computeTotSens),
                             ('jointt', self.pcrv.computeJointSens)):
            self.sens[key] = indices()[0]
        return self.sens

    def compute_all(self, ysam):
        r"""Main and total indices of ALL outputs ``ysam (N, o)`` at once, each (o, d): one fused statistics pass with
        the o outputs as right-hand sides, one batched K x K solve, the vectorised index formulas on the (o, K)
        coefficient matrix.  Same numbers as calling ``compute`` per column (what the reference's model_sens does)."""
        ysam = np.asarray(ysam, dtype=np.float64).reshape(self.nsam, -1)
        nout = ysam.shape[1]
        stats = SuffStats.from_pc(self.pcrv, self.germ_sam, ysam, 0)
        system = stats.dev_system()
        coef = solve_normal_device(system[0], system[1], 1.0e-6) if system is not None else None
        coef = solve_normal(stats.G, stats.B()) if coef is None else coef.cpu().numpy()          # (K, o)
        for i in range(nout):
            print(f'Output # {i+1}')
        mindex = self.pcrv.mindices[0]
        multi = PCRV(nout, self.dim, self.pctype, mi=mindex, cfs=np.ascontiguousarray(coef.T))
        self.pcrv.setCfs([np.ascontiguousarray(coef[:, -1])])      # state after the reference's loop: the last output
        self.sens['main'], self.sens['total'] = multi.computeSens()[-1], multi.computeTotSens()[-1]
        self.sens['jointt'] = self.pcrv.computeJointSens()[0]
        return multi.computeSens(), multi.computeTotSens()
