"""Karhunen-Loeve expansion of an ensemble of fields (reference: linred/kle.py:9-84).

The field lives on N grid points with trapezoidal quadrature weights w (1/2 at the two ends), M samples.  The modes
are the eigenvectors of the weighted covariance  diag(sqrt w) C diag(sqrt w), mapped back by 1/sqrt(w); the latent
features are the weighted projections on the modes scaled to unit variance.  Covariance GEMM, symmetric
eigendecomposition (cuSOLVER syevd through torch.linalg.eigh) and projection run on the device."""
import numpy as np

from .. import device as _device
from .linred import LinRed, _mm


class KLE(LinRed):
    def __init__(self):
        LinRed.__init__(self)
        self.weights2 = None          # quadrature weights (the reference keeps them under this name)

    def build(self, data, plot=True):
        r"""data (N, M): N conditions (grid points), M samples.  ``plot`` is accepted and ignored."""
        import torch
        _device.require_device()
        where = torch.device("cuda", _device.current_device())
        npts, nsam = data.shape
        quad = np.ones(npts)
        quad[[0, -1]] = 0.5
        self.weights2 = quad

        field = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).to(where)
        centre = field.mean(dim=1)
        dev = field - centre[:, None]
        root_w = torch.from_numpy(np.sqrt(quad)).to(where)
        weighted = dev * root_w[:, None]
        # eigenpairs of the weighted sample covariance (np.cov normalisation: M - 1), largest first
        lam, vec = torch.linalg.eigh((weighted @ weighted.T) / (nsam - 1))
        self.mean = centre.cpu().numpy()
        self.eigval = lam.flip(0).cpu().numpy()
        self.modes = (vec.flip(1) / root_w[:, None]).cpu().numpy()
        self.built = True
        self.eigval_clip()
        self.xi = self.project(data)

    def project(self, data, subtract_mean=True):
        r"""data (N, M) -> (M, K): quadrature-weighted inner products with the modes, scaled by 1 / sqrt(eigval)."""
        assert(self.built)
        kept = self.modes.shape[1]
        rows = data.T - self.mean if subtract_mean else data.T
        basis = (self.modes * self.weights2[:, np.newaxis]) / np.sqrt(self.eigval[:kept])
        return _mm(rows, basis)
