"""Minimal ``Function`` wrapper PCRV.setFunction hands out (reference: func/func.py:14-56,
82-99, 185-208).  Only the members the PC path and its callers use are provided; the
reference's function algebra and test-function zoo are out of scope."""
import numpy as np


class Function():
    def __init__(self, name='Base'):
        self.dim = None
        self.outdim = None
        self.domain = None
        self.name = name
        self.eval = None
        self.dmax = 10.0

    def __call__(self, x):
        if self.eval is None:
            raise NotImplementedError("Function evaluation is not implemented in the base class.")
        return self.eval(x)

    def __repr__(self):
        return f"Function({self.name}, dim={self.dim}, outdim={self.outdim})"

    def setCall(self, eval):
        self.eval = eval

    def call1d(self, odim=0):
        return lambda x: self.__call__(x)[:, odim]

    def setDimDom(self, domain=None, dimension=None):
        if domain is None:
            assert dimension is not None
            self.dim = int(dimension)
            self.domain = np.tile(np.array([-self.dmax, self.dmax]), (self.dim, 1))
        else:
            assert dimension is None
            self.dim = domain.shape[0]
            self.domain = domain

    def inDomain(self, x):
        return bool(np.all(x >= self.domain[:, 0]) and np.all(x <= self.domain[:, 1]))
