"""KL + PC surrogate of a field output (reference: linred/klsurr.py:27-306, PC branch) -- BASELINE configuration 5.

The N_t outputs are compressed to n_eig latent features by a KLE; ALL features are fitted in one pass over the
samples (``pc_fit``: one shared Gram through K3, per-feature BCS on the K x K statistics) and evaluated together through
K2; Sobol indices of the physical outputs follow from ONE GEMM that pushes the PC coefficients of the features through
the modes, and the vectorised ``computeSens``.  The reference's NN branch (QUiNN) and its plots are out of scope."""
import sys

import numpy as np

from .kle import KLE
from .linred import _mm
from ..workflows.fits import pc_fit
from ..func.func import Function
from ..utils.mindex import micf_join
from ..rv.pcrv import PCRV


class KLSurr():
    """Attributes as in the reference: ``ndim``, ``built``, ``kl`` (KLE), ``neig``, ``smodel`` (the PCRV of the latent
    features), ``xisurr`` (callable inputs -> features), ``function``, and the training diagnostics ``ytrain_kl``,
    ``xitrain_kl``, ``ytrain_klsurr``."""

    def __init__(self):
        self.ndim = self.neig = None
        self.xisurr = self.smodel = self.function = None
        self.built = False

    def setDim(self, ndim):
        self.ndim = ndim

    def _fit_features(self, inputs, features, surr, pc_order, bcs_eta):
        """Surrogate of the latent features as a function of the inputs (all features in one statistics pass)."""
        if surr == 'NN':
            raise NotImplementedError("KLSurr: the NN surrogate (QUiNN) is outside the PC path of this package.")
        if surr != 'PC':
            print(f"Surrogate type {surr} is unknown. Please use NN or PC. Exiting.")
            sys.exit()
        expansion, _ = pc_fit(inputs, features, order=pc_order, pctype='LU', method='bcs', eta=bcs_eta)
        return expansion

    def build(self, ptrain, ytrain, neig=None, surr='PC', tr_frac=0.9, pc_order=3, bcs_eta=1.e-8):
        r"""ptrain (M, d) inputs, ytrain (M, N) field outputs."""
        assert(ptrain.shape[0] == ytrain.shape[0])
        self.setDim(ndim=ptrain.shape[1])
        self.kl = KLE()
        self.kl.build(ytrain.T, plot=False)
        keep = self.kl.get_neig(0.99) if neig is None else neig

        self.smodel = self._fit_features(ptrain, self.kl.xi[:, :keep], surr, pc_order, bcs_eta)
        self.xisurr = self.smodel.function
        self.built = True
        self.neig = keep
        # training-set diagnostics the reference keeps: truncated KL, surrogate features, KL of the surrogate features
        self.ytrain_kl = self.kl.eval(neig=keep)
        self.xitrain_kl = self.evalxi(ptrain)
        self.ytrain_klsurr = self.kl.eval(xi=self.xitrain_kl, neig=keep)
        self.setFunction()

    def setFunction(self):
        wrapped = Function(name='KLSurr Function')
        wrapped.setDimDom(dimension=self.ndim)
        wrapped.setCall(lambda x: self.eval(x).T)
        self.function = wrapped

    def evalxi(self, ptrain, otherpars=None):
        assert(self.built)
        return self.xisurr(ptrain)

    def eval(self, ptrain, neig=None):
        r"""(M, d) inputs -> (N, M) fields through the leading ``neig`` modes (default: all fitted ones)."""
        assert(self.built)
        feats = self.evalxi(ptrain)
        return self.kl.eval(xi=feats, neig=feats.shape[1] if neig is None else neig)

    def compute_sens_pc(self, colors=None, pnames=None, totmain='main'):
        r"""PC-based Sobol indices of the physical outputs, (N, d) (plots of the reference left out).  The indices of
        the latent features themselves are kept in ``self.sens_eig`` (n_eig, d)."""
        assert(isinstance(self.smodel, PCRV))
        families = self.smodel.pctypes
        assert(len(set(families)) == 1)
        union_mi, union_cf = micf_join(self.smodel.mindices, self.smodel.coefs)      # one multi-index for all features
        assert(not union_mi[0].any())
        assert(union_cf.shape[0] == self.neig)
        if totmain not in ('main', 'total'):
            print("Please indicate main or total sensititvity. Exiting.")
            sys.exit()

        def indices_of(expansion):
            return expansion.computeSens() if totmain == 'main' else expansion.computeTotSens()

        self.sens_eig = indices_of(PCRV(self.neig, self.ndim, families, mi=union_mi, cfs=union_cf))
        # coefficients of the physical outputs: features' coefficients pushed through sqrt(eigval) * modes, plus the mean
        amplitudes = np.sqrt(self.kl.eigval[:self.neig])
        phys_cf = _mm(self.kl.modes[:, :self.neig] * amplitudes, union_cf)           # (N, npc)
        phys_cf[:, 0] += self.kl.mean
        print("Multiindex shape :", union_mi.shape)
        print("PC coefficients' shape :", phys_cf.T.shape)
        return indices_of(PCRV(phys_cf.shape[0], self.ndim, families, mi=union_mi, cfs=phys_cf))
