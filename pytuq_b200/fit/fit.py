"""Base class of the fitters (reference: fit/fit.py:6-54)."""
import numpy as np


class fitbase(object):
    def __init__(self):
        self.fitted = False

    def fit(self, x, y):
        raise NotImplementedError

    def predict(self, x, msc=0, pp=False):
        raise NotImplementedError

    def predict_wstd(self, xc, pp=False):
        yc, yc_var, _ = self.predict(xc, msc=1, pp=pp)
        return yc, np.sqrt(yc_var)
