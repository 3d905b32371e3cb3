"""Base class PCRV derives from (reference: rv/mrv.py:18-47; the distributions that live in
the reference's mrv.py are outside the PC path)."""
import numpy as np


class MRV():
    def __init__(self, pdim):
        self.pdim = pdim

    def __repr__(self):
        return f"Multivariate Random Variable(dim={self.pdim})"

    def sample(self, nsam):
        raise NotImplementedError("Sampling is not implemented in the base class.")

    def pdf(self, x):
        raise NotImplementedError("PDF function is not implemented in the base class.")

    def logpdf(self, x):
        return np.log(self.pdf(x))

    def cdf(self, x):
        raise NotImplementedError("CDF function is not implemented in the base class.")
