"""QR sufficient statistic for least squares.

The Gram form squares the condition number, so ``lsq`` on square (N = K), underdetermined or nearly
collinear systems cannot reproduce the reference's SVD least squares (``scipy.linalg.lstsq(A, y,
cond=1e-13)``, lreg/lreg.py:335).  This module reduces the sample axis with orthogonal transformations
instead: the augmented matrix  Z = [A | Y]  is processed in row chunks, each chunk stacked under the
running triangular factor and re-factorised,

    R  <-  qr([R ; Z_chunk]).R ,        R = [[R_A, Q^T Y], [0, rho]]   ((K+M) x (K+M), upper),

so that ``min ||A c - y||`` becomes the K x K problem ``min ||R_A c - Q^T y||`` with the SAME singular
values as A; the truncated-SVD minimum-norm solution of that small problem is what gelsd returns for A.
The design-matrix chunks are produced on the device by K1 (``pcq_eval_bases_dev`` writing straight into
the left K columns of the chunk buffer); the chunk QR and the K x K SVD are plain library calls
(cuSOLVER geqrf / gesvd through torch.linalg) -- this is the robustness route for ill-conditioned
problems, not the throughput path (that is K3).  Shards combine by stacking their R factors
(``combine_r``): same sharding as the Gram path, the all-reduce replaced by an all-gather of (K+M)^2
doubles and one more QR.
"""
import numpy as np

from .. import _native, device as _device

_CHUNK_BYTES = 1 << 30


def _torch():
    import torch
    return torch


def _qr_r(Z):
    torch = _torch()
    L = Z.shape[1]
    if Z.shape[0] < L:       # fewer rows than columns: zero rows change nothing and keep R square
        Z = torch.cat([Z, torch.zeros((L - Z.shape[0], L), dtype=Z.dtype, device=Z.device)])
    return torch.linalg.qr(Z, mode="r").R


def combine_r(Rs):
    """R factor of the row-stacked matrices whose R factors are given."""
    torch = _torch()
    Rs = list(Rs)
    return Rs[0] if len(Rs) == 1 else _qr_r(torch.cat(Rs))


def r_factor_pc(pcrv, x, y, jdim=0, chunk_bytes=None):
    """R factor of [Psi(x) | y] (device tensor, (K+M) x (K+M)); Psi comes from K1 chunk by chunk."""
    torch = _torch()
    _device.require_device()
    dev = _device.current_device()
    from ..parallel import is_device_tensor
    if is_device_tensor(x):        # resident samples: this route stages its chunks from the host (rare path)
        x = x.cpu().numpy()
        y = None if y is None else y.cpu().numpy()
    x = np.ascontiguousarray(x, dtype=np.float64)
    N, s = x.shape
    Y = None if y is None else np.ascontiguousarray(y, dtype=np.float64).reshape(N, -1)
    M = 0 if Y is None else Y.shape[1]
    mi = pcrv.mindices[jdim]
    K = mi.shape[0]
    plan = pcrv._plan(mi)
    lib = _native.lib()
    L = K + M
    rows = max(L, int((chunk_bytes or _CHUNK_BYTES) // (8 * L)))
    tdev = torch.device("cuda", dev)
    stream = torch.cuda.current_stream(tdev).cuda_stream
    R = None
    for r0 in range(0, max(N, 1), rows):
        r1 = min(N, r0 + rows)
        nr = r1 - r0
        Z = torch.empty((nr, L), dtype=torch.float64, device=tdev)
        if nr:
            xs = torch.from_numpy(x[r0:r1]).to(tdev)
            _native.check(lib.pcq_eval_bases_dev(plan.handle, xs.data_ptr(), nr, s, Z.data_ptr(), L, stream))
            if M:
                Z[:, K:] = torch.from_numpy(Y[r0:r1]).to(tdev)
        R = _qr_r(Z if R is None else torch.cat([R, Z]))
    return R


def r_factor_matrix(Amat, y, chunk_bytes=None):
    """R factor of [Amat | y] for a caller-supplied design matrix (host numpy)."""
    torch = _torch()
    _device.require_device()
    dev = _device.current_device()
    A = np.asarray(Amat, dtype=np.float64)
    N, K = A.shape
    Y = None if y is None else np.asarray(y, dtype=np.float64).reshape(N, -1)
    M = 0 if Y is None else Y.shape[1]
    L = K + M
    rows = max(L, int((chunk_bytes or _CHUNK_BYTES) // (8 * L)))
    tdev = torch.device("cuda", dev)
    R = None
    for r0 in range(0, max(N, 1), rows):
        r1 = min(N, r0 + rows)
        Z = torch.empty((r1 - r0, L), dtype=torch.float64, device=tdev)
        Z[:, :K] = torch.from_numpy(np.ascontiguousarray(A[r0:r1])).to(tdev)
        if M:
            Z[:, K:] = torch.from_numpy(np.ascontiguousarray(Y[r0:r1])).to(tdev)
        R = _qr_r(Z if R is None else torch.cat([R, Z]))
    return R


def allgather_r(R):
    """Combines the R factors of all ranks of the default process group (every rank gets the result)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return R
    torch = _torch()
    L = R.shape[1]
    if R.shape[0] < L:          # a rank with fewer rows than columns: zero rows keep R^T R and make the shapes equal
        R = torch.cat([R, torch.zeros((L - R.shape[0], L), dtype=R.dtype, device=R.device)])
    parts = [torch.empty_like(R) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, R.contiguous())
    return combine_r(parts)


def shared_r(R):
    """R factor of the whole data set when the ranks hold row blocks (parallel.sharded_rows), else R itself."""
    from .. import parallel
    if parallel._sharded and parallel.dist_info()[1] > 1:
        return allgather_r(R)
    return R


def lstsq_from_r(R, K, cond=1.0e-13):
    """Minimum-norm least-squares solution from the R factor of [A | Y]: singular values of R_A (= those of
    A) below ``cond * s_max`` are treated as zero, like gelsd.  Returns (C (K x M) numpy, singular values,
    effective rank)."""
    torch = _torch()
    RA = R[:K, :K]
    Q = R[:K, K:]
    U, S, Vh = torch.linalg.svd(RA)
    smax = float(S[0]) if S.numel() else 0.0
    keep = S > cond * smax
    rank = int(keep.sum())
    W = (U.T @ Q)[keep] / S[keep, None]
    C = Vh[keep].T @ W
    return C.cpu().numpy(), S.cpu().numpy(), rank
