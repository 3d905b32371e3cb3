"""KL + PC surrogate of a field output (reference: linred/klsurr.py:27-306, PC branch).

This is BASELINE configuration 5 end to end: the N_t outputs are compressed to n_eig latent features, ALL of
which are fitted in one pass over the samples (``pc_fit``: one shared Gram through K3, per-output BCS on the
K x K statistics), evaluated together through K2, and their Sobol indices pushed back to the physical outputs
with one GEMM on the coefficients.  The reference's NN branch (QUiNN) and its plots are out of scope."""
import sys

import numpy as np

from .kle import KLE
from .linred import _mm
from ..workflows.fits import pc_fit
from ..func.func import Function
from ..utils.mindex import micf_join
from ..rv.pcrv import PCRV


class KLSurr():
    def __init__(self):
        self.ndim = None
        self.built = False
        self.xisurr = None
        self.smodel = None
        self.function = None
        self.neig = None

    def setDim(self, ndim):
        self.ndim = ndim

    def build(self, ptrain, ytrain, neig=None, surr='PC', tr_frac=0.9, pc_order=3, bcs_eta=1.e-8):
        r"""ptrain (M, d), ytrain (M, N)  (klsurr.py:61-136)."""
        nens, ndim = ptrain.shape
        nens_, nt = ytrain.shape
        assert(nens == nens_)
        self.setDim(ndim=ndim)

        self.kl = KLE()
        self.kl.build(ytrain.T, plot=False)
        if neig is None:
            neig = self.kl.get_neig(0.99)

        xitrain = self.kl.xi[:, :neig]
        if surr == 'PC':
            pcrv, _ = pc_fit(ptrain, xitrain, order=pc_order, pctype='LU', method='bcs', eta=bcs_eta)
            self.xisurr = lambda x: pcrv.function(x)
            self.smodel = pcrv
        elif surr == 'NN':
            raise NotImplementedError("KLSurr: the NN surrogate (QUiNN) is outside the PC path of this package.")
        else:
            print(f"Surrogate type {surr} is unknown. Please use NN or PC. Exiting.")
            sys.exit()

        self.built = True
        self.ytrain_kl = self.kl.eval(neig=neig)
        self.xitrain_kl = self.evalxi(ptrain)
        self.ytrain_klsurr = self.kl.eval(xi=self.xitrain_kl, neig=neig)
        self.neig = neig
        self.setFunction()

    def setFunction(self):
        self.function = Function(name='KLSurr Function')
        self.function.setDimDom(dimension=self.ndim)
        self.function.setCall(lambda x: self.eval(x).T)

    def evalxi(self, ptrain, otherpars=None):
        assert(self.built)
        return self.xisurr(ptrain)

    def eval(self, ptrain, neig=None):
        r"""(M, d) -> (N, M)  (klsurr.py:159-177)."""
        assert(self.built)
        xi = self.evalxi(ptrain)
        if neig is None:
            neig = xi.shape[1]
        return self.kl.eval(xi=xi, neig=neig)

    def compute_sens_pc(self, colors=None, pnames=None, totmain='main'):
        r"""PC-based Sobol indices of the physical outputs, (N, d)  (klsurr.py:247-306, plots left out).
        Also keeps the indices of the latent features in ``self.sens_eig`` (n_eig, d)."""
        assert(isinstance(self.smodel, PCRV))
        mindex_all, cfs_all = micf_join(self.smodel.mindices, self.smodel.coefs)
        assert(np.sum(mindex_all[0, :]) == 0)
        assert(self.neig == cfs_all.shape[0])
        assert(len(set(self.smodel.pctypes)) == 1)
        if totmain not in ('main', 'total'):
            print("Please indicate main or total sensititvity. Exiting.")
            sys.exit()

        pcrv_eig = PCRV(self.neig, self.ndim, self.smodel.pctypes, mi=mindex_all, cfs=cfs_all)
        self.sens_eig = pcrv_eig.computeSens() if totmain == 'main' else pcrv_eig.computeTotSens()

        cfs_glob = _mm(cfs_all.T * np.sqrt(self.kl.eigval[:self.neig])[np.newaxis, :],
                       self.kl.modes[:, :self.neig].T)                       # npc, nx
        cfs_glob[0, :] += self.kl.mean
        print("Multiindex shape :", mindex_all.shape)
        print("PC coefficients' shape :", cfs_glob.shape)
        nx = cfs_glob.shape[1]
        pcrv_phys = PCRV(nx, self.ndim, self.smodel.pctypes, mi=mindex_all, cfs=cfs_glob.T)
        return pcrv_phys.computeSens() if totmain == 'main' else pcrv_phys.computeTotSens()
