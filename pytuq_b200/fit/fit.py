"""Interface every fitter of the package implements (the role of the reference's fit/fit.py:6-54): a ``fitted`` flag,
``fit(x, y)``, ``predict(x, msc, pp) -> (mean, variance | None, covariance | None)`` and the mean / standard-deviation
shorthand built on it."""


class fitbase(object):
    fitted = False

    def __init__(self):
        self.fitted = False

    def fit(self, x, y):
        raise NotImplementedError(f"{type(self).__name__} does not implement fit(x, y)")

    def predict(self, x, msc=0, pp=False):
        raise NotImplementedError(f"{type(self).__name__} does not implement predict(x, msc, pp)")

    def predict_wstd(self, xc, pp=False):
        """(mean, standard deviation) at xc; ``pp`` adds the data variance (posterior predictive)."""
        mean, variance = self.predict(xc, msc=1, pp=pp)[:2]
        return mean, variance ** 0.5
