"""Linear dimensionality reduction: the part of the reference's ``LinRed`` parent class (linred/linred.py:12-183) that
the KL + PC surrogate needs -- projection onto the modes, truncated reconstruction, explained variance.  Plotting is
left out.

A reduction is (mean (N,), modes (N, K), eigval (K,), xi (M, K)) with  data ~ mean + sum_k sqrt(eigval_k) xi_k mode_k.
The two dense products are plain GEMMs and run on the device through torch.matmul (cuBLAS): glue around the PC path,
not part of it."""
import numpy as np

from .. import device as _device


def _mm(a, b):
    """a @ b for host numpy operands, computed on the current device (float64)."""
    import torch
    _device.require_device()
    where = torch.device("cuda", _device.current_device())
    left, right = (torch.from_numpy(np.ascontiguousarray(m, dtype=np.float64)).to(where) for m in (a, b))
    return torch.matmul(left, right).cpu().numpy()


class LinRed(object):
    r"""Attributes as in the reference: ``eigval (K,)``, ``mean (N,)``, ``modes (N,K)``, ``xi (M,K)``, ``built``."""

    def __init__(self):
        self.mean = self.eigval = self.modes = self.xi = None
        self.built = False

    def build(self, data, plot=True):
        raise NotImplementedError

    def _amplitudes(self, count=None):
        """sqrt(eigval) of the leading ``count`` modes."""
        return np.sqrt(self.eigval[:count])

    def project(self, data, subtract_mean=True):
        r"""data (N, M) -> latent features (M, K): centred rows projected on the modes, scaled to unit variance."""
        assert(self.built)
        rows = data.T - self.mean if subtract_mean else data.T
        return _mm(rows, self.modes / self._amplitudes())

    def eval(self, xi=None, neig=None, add_mean=True):
        r"""Reconstruction from latent features ``xi (M, K)`` (default: the training ones) with the leading ``neig``
        modes -> (N, M)."""
        assert(self.built)
        feats = self.xi if xi is None else xi
        keep = self.modes.shape[0] if neig is None else neig
        weighted_modes = self.modes[:, :keep] * self._amplitudes(keep)
        field = _mm(weighted_modes, feats[:, :keep].T)
        if add_mean:
            field += self.mean[:, np.newaxis]
        return field

    def eigval_clip(self):
        np.maximum(self.eigval, 1.e-14, out=self.eigval)

    def compute_relvar(self):
        """(N, K): fraction of each output's variance captured by the first k modes."""
        assert(self.built)
        share = np.square(self.modes) * self.eigval
        return np.cumsum(share, axis=1) / share.sum(axis=1, keepdims=True)

    def get_neig(self, threshold=0.99):
        """Smallest number of modes whose eigenvalues exceed the given fraction of the total variance."""
        assert(self.built)
        captured = np.cumsum(self.eigval) / np.sum(self.eigval)
        neig = int(np.flatnonzero(captured > threshold)[0]) + 1
        print(f'Number of eigenvalues requested : {neig}')
        return neig

    def modes_clip(self, neig_clip=None):
        keep = self.get_neig(threshold=0.999) if neig_clip is None else neig_clip
        self.modes = self.modes[:, :keep]
