"""Sufficient statistics of a linear regression, computed on the GPU (K3 / K4).

Every fitter on the path needs only  G = A^T A,  b = A^T y,  y^T y,  sum(y)  and N
(anl: lreg/anl.py:78-99; bcs_fit: lreg/bcs.py:79-100,166-193,229; lsq through the normal
equations).  libpcq returns them packed in one symmetric matrix

    S = sum_n z_n z_n^T ,   z_n = [ A[n,:] (K) | y_n (M) | 1 ]

so the sample axis is reduced in one streaming pass, shards over samples with a single
all-reduce of S, and A is never materialised on the fused path.
"""
import numpy as np

from .. import _native, device as _device


class SuffStats:
    """View of the packed statistics matrix S ((K+M+1) x (K+M+1), float64).

    The matrix may live on the device (``S_dev``, a torch tensor produced by the fused path): the
    least-squares solve then runs on the GPU and only K coefficients come back; the host copy ``S`` is
    made lazily (once, shared by all per-output views) for the fitters that work on K x K slices with
    numpy (anl, bcs).  ``select(j)`` is a zero-copy view exposing one output."""

    # when the identity-based residual falls below this fraction of y^T y its leading digits have
    # cancelled; the exact residual pass (if the data are still reachable) takes over
    RESIDUAL_REFINE = 1.0e-8

    def __init__(self, S, nbasis, nout, S_dev=None, cols=None, parent=None, residual=None):
        self._residual = residual     # callable (cf (K,), output column) -> (sum r^2, sum y r), or None
        self._S = S
        self.S_dev = S_dev
        self.K = int(nbasis)
        self._Mall = int(nout)                       # outputs held by the packed matrix
        self.cols = list(range(self._Mall)) if cols is None else list(cols)
        self.M = len(self.cols)                      # outputs this view exposes
        self._parent = parent
        shape = tuple(S.shape) if S is not None else tuple(S_dev.shape)
        assert shape == (self.K + self._Mall + 1, self.K + self._Mall + 1)

    @property
    def S(self):
        if self._S is None:
            self._S = self._parent.S if self._parent is not None else self.S_dev.cpu().numpy()
        return self._S

    def _col(self, j):
        return self.K + self.cols[j]

    # --- accessors -------------------------------------------------------------------
    @property
    def G(self):
        return self.S[:self.K, :self.K]

    @property
    def n(self):
        return float(self.S[-1, -1])

    def b(self, j=0):
        return self.S[:self.K, self._col(j)]

    def B(self):
        return self.S[:self.K, [self._col(j) for j in range(self.M)]]

    def yty(self, j=0):
        return float(self.S[self._col(j), self._col(j)])

    def ysum(self, j=0):
        return float(self.S[self._col(j), -1])

    def colsum(self):
        return self.S[:self.K, -1]

    def yvar(self, j=0):
        """Population variance of y_j (np.std(y)**2 in lreg/bcs.py:79)."""
        n = self.n
        mean = self.ysum(j) / n
        return max(self.yty(j) / n - mean * mean, 0.0)

    def rss(self, cf, used=None, j=0):
        """sum_n (y_n - A[n,used] cf)^2 = y^T y - 2 cf^T b_used + cf^T G_used cf."""
        if used is None:
            used = np.arange(self.K)
        Gu = self.G[np.ix_(used, used)]
        bu = self.b(j)[used]
        val = self.yty(j) - 2.0 * float(cf @ bu) + float(cf @ (Gu @ cf))
        if self._residual is not None and val <= self.RESIDUAL_REFINE * self.yty(j):
            full = np.zeros(self.K)
            full[np.asarray(used)] = cf
            val = self._residual(full, self.cols[j])[0]
        return val

    def y_dot_residual(self, cf, j=0):
        """y . (y - A cf) for a full-length cf (lreg/anl.py:61): y^T y - cf^T b on the statistics, or the exact
        pass over the data when that difference has cancelled."""
        val = self.yty(j) - float(self.b(j) @ cf)
        if self._residual is not None and abs(val) <= self.RESIDUAL_REFINE * self.yty(j):
            val = self._residual(np.asarray(cf, dtype=np.float64), self.cols[j])[1]
        return val

    def select(self, j):
        """Statistics of output j alone (a view: the shared G is neither copied nor re-downloaded)."""
        return SuffStats(self._S, self.K, self._Mall, S_dev=self.S_dev, cols=[self.cols[j]], parent=self,
                         residual=self._residual)

    def dev_system(self):
        """(G, B) as device tensors for the exposed outputs, or None when the statistics are host-only."""
        if self.S_dev is None:
            return None
        G = self.S_dev[:self.K, :self.K]
        if self.cols == list(range(self.cols[0], self.cols[0] + self.M)):
            B = self.S_dev[:self.K, self.K + self.cols[0]:self.K + self.cols[0] + self.M]
        else:
            B = self.S_dev[:self.K, [self._col(j) for j in range(self.M)]]
        return G, B

    # --- constructors -------------------------------------------------------------------
    @classmethod
    def from_pc(cls, pcrv, x, y=None, jdim=0):
        """Fused path (K3): statistics of the jdim-th PC basis at germ points x, no A."""
        from ..parallel import suffstats_pc, residual_sums_pc
        st = suffstats_pc(pcrv, x, y, jdim)
        if y is not None:
            st._residual = lambda cf, col: residual_sums_pc(pcrv, x, y, col, jdim, cf)
        return st

    @classmethod
    def from_matrix(cls, Amat, y=None):
        """Materialised path (K4): statistics of a design matrix handed over by the caller
        (lreg.fita(Amat, y), lreg/lreg.py:48-59)."""
        _device.require_device()
        A = np.ascontiguousarray(Amat, dtype=np.float64)
        N, K = A.shape
        if y is None:
            Y, M = None, 0
        else:
            Y = np.ascontiguousarray(y, dtype=np.float64)
            Y = Y.reshape(N, Y.shape[1] if Y.ndim == 2 else 1)
            M = Y.shape[1]
        L = K + M + 1
        S = np.empty((L, L))
        st = _native.lib().pcq_gram(_device.current_device(), _native.ptr(A), N, K, K,
                                    _native.ptr(Y), max(M, 1), M, _native.ptr(S))
        _native.check(st, "pcq_gram")
        out = cls(S, K, M)
        if M:
            from ..parallel import residual_sums_matrix
            out._residual = lambda cf, col: residual_sums_matrix(A, Y, col, cf)
        return out
