"""The small ``Function`` object ``PCRV.setFunction`` / ``KLSurr.setFunction`` hand out: a named callable on a box domain
(the members of the reference's func/func.py:14-56, 82-99, 185-208 that the PC path and its callers use).  The
reference's function algebra, gradients, plotting and test-function zoo are out of scope."""
import numpy as np


class Function():
    dmax = 10.0            # half-width of the default domain

    def __init__(self, name='Base'):
        self.name = name
        self.eval = None
        self.dim = self.outdim = self.domain = None
        self.dmax = Function.dmax

    def __repr__(self):
        return f"Function({self.name}, dim={self.dim}, outdim={self.outdim})"

    def __call__(self, x):
        if self.eval is None:
            raise NotImplementedError("Function evaluation is not implemented in the base class.")
        return self.eval(x)

    def setCall(self, eval):
        """Installs the callable x (N, d) -> (N, o)."""
        self.eval = eval

    def call1d(self, odim=0):
        """The odim-th output as a function of its own."""
        return lambda x: self(x)[:, odim]

    def setDimDom(self, domain=None, dimension=None):
        """Either a (d, 2) box, or a dimension (the box is then [-dmax, dmax]^d)."""
        assert (domain is None) != (dimension is None)
        if domain is not None:
            self.domain = domain
            self.dim = domain.shape[0]
        else:
            self.dim = int(dimension)
            self.domain = np.repeat(np.array([[-self.dmax, self.dmax]]), self.dim, axis=0)

    def inDomain(self, x):
        lo, hi = self.domain[:, 0], self.domain[:, 1]
        return bool((x >= lo).all() and (x <= hi).all())
