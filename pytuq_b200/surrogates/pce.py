"""``PCE`` surrogate wrapper (drop-in for the reference's ``surrogates/pce.py:30-427``):
``PCE(dim, order, type, verbose=0, **kw)`` with ``set_training_data / build / evaluate /
optimize_eta / kfold_split / kfold_cv / get_pc_terms``.

``build`` installs a :class:`~pytuq_b200.lreg.lreg.PCBasisEvaluator`, so ``lreg.fit`` takes
the fused no-A statistics path (K3); the k-fold eta search re-reduces each training fold
once and sweeps all etas on the resulting K x K statistics instead of rebuilding the
design matrix nfolds x len(etas) times (pce.py:244-262).
"""
import copy
import math

import numpy as np

from ..rv.pcrv import PCRV
from ..utils.mindex import get_mi
from ..lreg.lreg import lsq, PCBasisEvaluator
from ..lreg.anl import anl
from ..lreg.bcs import bcs
from ..lreg.stats import SuffStats


class PCE:
    def __init__(self, pce_dim, pce_order, pce_type, verbose=0, **kwargs):
        self.sdim = pce_dim
        self.order = pce_order
        self.pctype = pce_type
        self.outdim = kwargs.get('pce_outdim', 1)
        self.verbose = verbose
        self.lreg = None
        self._x_train = None
        self._y_train = None
        self.regression_method = None
        self.mindex = kwargs.get('mi', get_mi(self.order, self.sdim))
        self.pcrv = PCRV(self.outdim, self.sdim, self.pctype, mi=self.mindex, cfs=kwargs.get('cfs'))
        if self.verbose > 0:
            print("Constructed PC Surrogate with the following attributes:")
            print(self.pcrv, end='\n\n')

    def set_training_data(self, x_train, y_train):
        if not (isinstance(x_train, np.ndarray) and x_train.ndim == 2):
            raise ValueError("x_train must be a 2D numpy array.")
        if not (isinstance(y_train, np.ndarray) and y_train.ndim == 1):
            raise ValueError("y_train must be a 1D numpy array.")
        if x_train.shape[0] != y_train.shape[0]:
            raise ValueError("The number of samples in x and y must be the same.")
        self._x_train = x_train
        self._y_train = y_train

    def get_pc_terms(self):
        return [mi.shape[0] for mi in self.pcrv.mindices]

    # ---------------------------------------------------------------- cross-validation
    def kfold_split(self, nsamples, nfolds, seed=13):
        """{fold: {'train index', 'val index'}} from one seeded permutation (pce.py:128-166)."""
        parts = np.array_split(np.random.RandomState(seed).permutation(nsamples), nfolds)
        folds = {}
        for j in range(nfolds):
            rest = [parts[i] for i in range(nfolds) if i != j]
            train = np.concatenate(rest) if rest else np.array([], dtype='int64')
            folds[j] = {'train index': train.astype('int64', copy=False), 'val index': parts[j]}
        return folds

    def kfold_cv(self, x, y, nfolds=3, seed=13):
        """Training / validation arrays per fold (pce.py:169-204)."""
        n = x.shape[0]
        ycol = np.atleast_2d(y)
        if len(ycol) == 1:
            ycol = ycol.T
        data = {}
        for k, idx in self.kfold_split(n, nfolds, seed).items():
            data[k] = {'xtrain': x[idx['train index']], 'xval': x[idx['val index']],
                       'ytrain': np.squeeze(ycol[idx['train index']]),
                       'yval': np.squeeze(ycol[idx['val index']])}
            if nfolds == 1:
                data[k]['xtrain'], data[k]['ytrain'] = data[k]['xval'], data[k]['yval']
        return data

    def optimize_eta(self, etas, verbose=0, nfolds=3, plot=False):
        """Eta with the lowest fold-averaged validation RMSE (pce.py:207-314)."""
        folds = self.kfold_cv(self._x_train, self._y_train, nfolds)
        work = PCRV(self.outdim, self.sdim, self.pctype, mi=self.mindex)
        rmse_tr = np.zeros((nfolds, len(etas)))
        rmse_val = np.zeros((nfolds, len(etas)))
        for i in range(nfolds):
            x_tr, y_tr = folds[i]['xtrain'], folds[i]['ytrain']
            x_val, y_val = folds[i]['xval'], folds[i]['yval']
            stats = SuffStats.from_pc(work, x_tr, y_tr, 0)       # once per fold, reused by every eta
            for e, eta in enumerate(etas):
                fitter = bcs(eta=eta)
                fitter._fit_stats(stats)
                sparse = PCRV(1, self.sdim, self.pctype, mi=self.mindex[fitter.used, :], cfs=fitter.cf)
                if verbose > 0:
                    print("Fold " + str(i + 1) + ", eta " + str(eta) + ", " + str(len(fitter.cf)) +
                          " terms retained out of a full basis of size " + str(len(self.mindex)))
                rmse_tr[i, e] = math.sqrt(np.square(y_tr - sparse.evalPC(x_tr)[:, 0]).mean())
                rmse_val[i, e] = math.sqrt(np.square(y_val - sparse.evalPC(x_val)[:, 0]).mean())
        self.eta_cv = {'etas': np.asarray(etas), 'rmse_train': rmse_tr, 'rmse_val': rmse_val}
        return etas[int(np.argmin(rmse_val.mean(axis=0)))]

    # ---------------------------------------------------------------- build / evaluate
    def build(self, **kwargs):
        """Fits the PC coefficients; keyword arguments as in the reference (pce.py:317-389)."""
        if self._x_train is None or self._y_train is None:
            raise RuntimeError("Training data must be set using set_training_data() before calling build().")
        self.pcrv.setMiCfs(self.mindex, cfs=None)
        regression = kwargs.get('regression', 'lsq')
        if regression == 'lsq':
            self.lreg = lsq()
        elif regression == 'anl':
            if kwargs.get('method', None) == 'vi':
                self.lreg = anl(method='vi', datavar=kwargs.get('datavar'), cov_nugget=kwargs.get('cov_nugget', 0.0))
            else:
                self.lreg = anl(datavar=kwargs.get('datavar'), prior_var=kwargs.get('prior_var'),
                                cov_nugget=kwargs.get('cov_nugget', 0.0))
        elif regression == 'bcs':
            eta = kwargs.get('eta', 1.e-8)
            if isinstance(eta, list) or (isinstance(eta, np.ndarray) and eta.ndim == 1):
                opt_eta = self.optimize_eta(eta, nfolds=kwargs.get('nfolds', 3),
                                            verbose=kwargs.get('eta_verbose', False),
                                            plot=kwargs.get('eta_plot', False))
                self.lreg = bcs(eta=opt_eta, datavar_init=kwargs.get('datavar_init'))
            elif isinstance(eta, float) and eta > 0:
                self.lreg = bcs(eta=eta, datavar_init=kwargs.get('datavar_init'))
            else:
                raise ValueError("You may provide either a positive float (defaulting to 1.e-8) or 1D list/numpy array for the value of eta. If a list/numpy array is provided, the most optimal eta from the array will be chosen.")
        else:
            raise ValueError(f"Regression method '{regression}' is not recognized and/or supported yet.")
        self.regression_method = type(self.lreg).__name__
        if self.verbose > 0:
            print("Regression method:", self.regression_method)

        self.lreg.setBasisEvaluator(PCBasisEvaluator(0), (self.pcrv,))
        self.lreg.fit(self._x_train, self._y_train)
        if regression == 'bcs':
            self.pcrv.setMiCfs([self.mindex[self.lreg.used, :]], [self.lreg.cf])
        return self.lreg.cf

    def evaluate(self, x_eval, **kwargs):
        """{'Y_eval', 'Y_eval_std', 'Y_eval_cov', 'Y_eval_var'} at x_eval (pce.py:391-427)."""
        if self.outdim == 1:
            y_eval, y_eval_var, _ = self.lreg.predict(x_eval, msc=1, pp=kwargs.get('data_variance', False))
            y_eval_std = np.sqrt(y_eval_var)
            y_eval_cov = None
        else:
            y_eval, y_eval_var, y_eval_cov = self.lreg.predict(x_eval, msc=2)
            y_eval_std = np.sqrt(np.diag(y_eval_cov))
        if self.regression_method == 'lsq':
            y_eval_std = y_eval_var = y_eval_cov = None
        return {'Y_eval': y_eval, 'Y_eval_std': y_eval_std, 'Y_eval_cov': y_eval_cov, 'Y_eval_var': y_eval_var}

    def __deepcopy__(self, memo):
        new = PCE.__new__(PCE)
        for k, v in self.__dict__.items():
            setattr(new, k, copy.deepcopy(v, memo))
        return new
