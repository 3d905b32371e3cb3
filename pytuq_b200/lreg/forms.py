"""Linear and quadratic forms of a design matrix handed over by the caller, on the device.

``lreg.predicta(Amat, msc)``, ``compute_stdev(Amat)`` and ``predicta_samples(Amat)`` (lreg/lreg.py:112-198; the
call pattern of apps/uqpc/uq_pc.py:313-315) form ``Amat @ cf`` and ``|| Amat chol(cf_cov) ||`` with numpy in the
reference.  Here the matrix goes through ``pcq_matrix_forms`` (include/pcq.h): the mean is an HBM-bound pass of
row dot products, the variance a triangular GEMM on the FP64 tensor pipe (K6 on the packed matrix chunk), many
vectors at once (predictive samples, the two GEMMs of msc=2) the linear-forms GEMM.  No host arithmetic of size
N x K.
"""
import numpy as np

from .. import _native, device as _device


def _is_dev(a):
    return type(a).__module__.split(".")[0] == "torch" and bool(getattr(a, "is_cuda", False))


def matrix_forms(Amat, cf=None, C=None):
    """(Y, v):  Y = Amat @ cf.T for cf (M, K)  [None when cf is None],
                v[n] = Amat[n] @ C @ Amat[n] for symmetric C (K, K)  [None when C is None].
    ``Amat``: numpy array (host; chunks are staged through the device) or a CUDA tensor (read in place; the
    results are CUDA tensors then)."""
    _device.require_device()
    lib = _native.lib()
    if _is_dev(Amat):
        import torch
        A = Amat if Amat.dtype == torch.float64 else Amat.to(torch.float64)
        A = A.contiguous()
        N, K = A.shape
        dev = A.device
        cf_t = None if cf is None else torch.as_tensor(np.ascontiguousarray(cf, dtype=np.float64) if isinstance(cf, np.ndarray) else cf,
                                                       dtype=torch.float64, device=dev).contiguous()
        C_t = None if C is None else torch.as_tensor(np.ascontiguousarray(C, dtype=np.float64) if isinstance(C, np.ndarray) else C,
                                                     dtype=torch.float64, device=dev).contiguous()
        M = 0 if cf_t is None else cf_t.shape[0]
        Y = torch.empty((N, M), dtype=torch.float64, device=dev) if M else None
        v = torch.empty(N, dtype=torch.float64, device=dev) if C_t is not None else None
        if N:
            with torch.cuda.device(dev):
                st = lib.pcq_matrix_forms_dev(dev.index, A.data_ptr(), N, K, K,
                                              cf_t.data_ptr() if M else None, K, M, Y.data_ptr() if M else None, max(M, 1),
                                              C_t.data_ptr() if C_t is not None else None, K,
                                              v.data_ptr() if v is not None else None, 1,
                                              torch.cuda.current_stream().cuda_stream)
            _native.check(st, "pcq_matrix_forms_dev")
        return Y, v
    A = np.ascontiguousarray(Amat, dtype=np.float64)
    N, K = A.shape
    cfm = None if cf is None else np.ascontiguousarray(cf, dtype=np.float64).reshape(-1, K)
    Cm = None if C is None else np.ascontiguousarray(C, dtype=np.float64)
    assert Cm is None or Cm.shape == (K, K)
    M = 0 if cfm is None else cfm.shape[0]
    Y = np.zeros((N, M)) if M else None
    v = np.zeros(N) if Cm is not None else None
    if N:
        st = lib.pcq_matrix_forms(_device.current_device(), _native.ptr(A), N, K, K, _native.ptr(cfm), K, M,
                                  _native.ptr(Y), max(M, 1), _native.ptr(Cm), K, _native.ptr(v), 1)
        _native.check(st, "pcq_matrix_forms")
    return Y, v


def psd_part(C):
    """(C', positive_definite): C itself when its Cholesky factorisation exists (checked on the device), else
    u diag(s) u^T from the Hermitian SVD -- the matrix whose quadratic form equals the reference's fallback
    || (Amat u) sqrt(diag(s)) ||^2 (lreg/lreg.py:183-186)."""
    import torch
    Ct = torch.from_numpy(np.ascontiguousarray(C, dtype=np.float64)).to(torch.device("cuda", _device.current_device()))
    _, info = torch.linalg.cholesky_ex(Ct)
    if int(info) == 0:
        return C, True
    u, s, _ = np.linalg.svd(C, hermitian=True)
    return (u * s) @ u.T, False
