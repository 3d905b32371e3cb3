"""Parent class of linear dimensionality reduction (reference: linred/linred.py:12-183, plotting left out).

The dense products (projection, reconstruction) are plain GEMMs and run on the device through
torch.matmul (cuBLAS): they are glue around the PC path, not part of it."""
import numpy as np

from .. import device as _device


def _dev():
    import torch
    _device.require_device()
    return torch.device("cuda", _device.current_device())


def _mm(a, b):
    """a @ b for host numpy operands, computed on the device."""
    import torch
    d = _dev()
    out = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(d) @ \
        torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).to(d)
    return out.cpu().numpy()


class LinRed(object):
    r"""Attributes as in the reference: ``eigval (K,)``, ``mean (N,)``, ``modes (N,K)``, ``xi (M,K)``, ``built``."""

    def __init__(self):
        self.built = False
        self.mean = None
        self.eigval = None
        self.modes = None
        self.xi = None

    def build(self, data, plot=True):
        raise NotImplementedError

    def project(self, data, subtract_mean=True):
        r"""data (N, M) -> latent features (M, K)  (linred/linred.py:45-59)."""
        assert(self.built)
        return _mm(data.T - int(subtract_mean) * self.mean, self.modes) / np.sqrt(self.eigval)

    def eval(self, xi=None, neig=None, add_mean=True):
        r"""Expansion at latent features xi (M, K) with the first neig modes -> (N, M)  (linred/linred.py:61-84)."""
        assert(self.built)
        nx = self.modes.shape[0]
        if xi is None:
            xi = self.xi
        if neig is None:
            neig = nx
        data_red = int(add_mean) * self.mean + _mm(xi[:, :neig] * np.sqrt(self.eigval[np.newaxis, :neig]),
                                                   self.modes[:, :neig].T)
        return data_red.T

    def eigval_clip(self):
        self.eigval[self.eigval < 1.e-14] = 1.e-14

    def compute_relvar(self):
        assert(self.built)
        tmp = self.modes * np.sqrt(self.eigval)
        return (np.cumsum(tmp * tmp, axis=1) + 0.0) / (np.sum(tmp * tmp, axis=1).reshape(-1, 1) + 0.0)

    def get_neig(self, threshold=0.99):
        assert(self.built)
        explained_variance = np.cumsum(self.eigval) / np.sum(self.eigval)
        neig = np.where(explained_variance > threshold)[0][0] + 1
        print(f'Number of eigenvalues requested : {neig}')
        return neig

    def modes_clip(self, neig_clip=None):
        if neig_clip is None:
            neig_clip = self.get_neig(threshold=0.999)
        self.modes = self.modes[:, :neig_clip]
