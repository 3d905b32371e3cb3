"""Karhunen-Loeve expansion (reference: linred/kle.py:9-84).  Covariance (one GEMM), symmetric
eigendecomposition (cuSOLVER syevd through torch.linalg.eigh) and projection run on the device."""
import numpy as np

from .linred import LinRed, _dev, _mm


class KLE(LinRed):
    def __init__(self):
        super().__init__()
        self.weights2 = None

    def build(self, data, plot=True):
        r"""data (N, M): N conditions (grid points), M samples.  ``plot`` is accepted and ignored."""
        import torch
        nx, nsam = data.shape
        d = _dev()
        D = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).to(d)
        mean = D.mean(dim=1)
        Dc = D - mean[:, None]
        cov = (Dc @ Dc.T) / (nsam - 1)                     # np.cov(data), kle.py:43

        self.mean = mean.cpu().numpy()
        self.weights2 = np.ones(nx)                        # trapezoidal rule weights, kle.py:46-49
        self.weights2[0] = 0.5
        self.weights2[-1] = 0.5
        weights = np.sqrt(self.weights2)
        w_t = torch.from_numpy(weights).to(d)

        eigval, eigvec = torch.linalg.eigh(torch.outer(w_t, w_t) * cov)
        self.eigval = eigval.flip(0).cpu().numpy()
        self.modes = (eigvec.flip(1) / w_t[:, None]).cpu().numpy()      # nx, neig
        self.built = True
        self.eigval_clip()
        self.xi = self.project(data)

    def project(self, data, subtract_mean=True):
        r"""data (N, M) -> (M, K) with the trapezoidal inner product (kle.py:67-82)."""
        assert(self.built)
        nmodes = self.modes.shape[1]
        return _mm(data.T - int(subtract_mean) * self.mean,
                   self.modes * self.weights2.reshape(-1, 1)) / np.sqrt(self.eigval[:nmodes])
