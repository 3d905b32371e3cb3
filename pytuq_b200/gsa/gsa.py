"""PC-based Sobol sensitivity analysis (drop-in for ``PCSobol``, reference gsa/gsa.py:383-457).
The sampling estimators that share the reference's file (SamSobol, Moat, Linreg) do not
touch the PC path and are out of scope."""
import numpy as np

from ..lreg.lreg import lsq, PCBasisEvaluator
from ..rv.pcrv import PCRV
from ..utils.mindex import get_mi


def scale01ToDom(xx, dom):
    """[0,1]^d -> domain (reference: utils/maps.py:7-22)."""
    if np.any(xx < 0.0) or np.any(xx > 1.0):
        print("Warning: some elements are outside the [0,1] range.")
    return xx * np.abs(dom[:, 1] - dom[:, 0]) + np.min(dom, axis=1)


class SensMethod():
    def __init__(self, dom, sens_names):
        self.dom = dom
        self.dim = dom.shape[0]
        self.sens_names = sens_names
        self.sens = dict((k, [None] * self.dim) for k in self.sens_names)

    def sample(self, nsam):
        raise NotImplementedError("Sampling for sensitivity is not implemented in the base class.")

    def compute(self, ysam):
        raise NotImplementedError("Computing sensitivity is not implemented in the base class.")


class PCSobol(SensMethod):
    r"""Samples the germ, fits a PC by least squares (fused statistics, no design matrix)
    and reads main / total / joint Sobol indices off the coefficients."""

    def __init__(self, dom, pctype='LU', order=3):
        print("Initializing PCSobol Sensitivity Method")
        self.sens_names = ['main', 'total', 'jointt']
        super().__init__(dom, self.sens_names)
        self.pctype = pctype
        self.order = order

    def sample(self, nsam):
        print("Sampling PCSobol Sensitivity Method")
        self.pcrv = PCRV(1, self.dim, self.pctype, mi=get_mi(self.order, self.dim))
        self.nsam = nsam
        self.germ_sam = self.pcrv.sampleGerm(nsam=self.nsam)
        unif_sam = np.zeros((self.nsam, self.dim))
        for idim in range(self.dim):
            unif_sam[:, idim] = self.pcrv.PC1ds[idim].germCdf(self.germ_sam[:, idim])
        self.xsam = scale01ToDom(unif_sam, self.dom)
        return self.xsam

    def compute(self, ysam):
        assert(ysam.shape[0] == self.nsam)
        fitter = lsq()
        fitter.setBasisEvaluator(PCBasisEvaluator(0), (self.pcrv,))
        fitter.fit(self.germ_sam, ysam)
        self.pcrv.setCfs([fitter.cf])
        self.sens['main'] = self.pcrv.computeSens()[0]
        self.sens['total'] = self.pcrv.computeTotSens()[0]
        self.sens['jointt'] = self.pcrv.computeJointSens()[0]
        return self.sens
