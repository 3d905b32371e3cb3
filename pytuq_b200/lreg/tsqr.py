"""QR sufficient statistic for least squares.

The Gram form squares the condition number, so ``lsq`` on square (N = K), underdetermined or nearly
collinear systems cannot reproduce the reference's SVD least squares (``scipy.linalg.lstsq(A, y,
cond=1e-13)``, lreg/lreg.py:335).  This module reduces the sample axis with orthogonal transformations
instead: the augmented matrix  Z = [A | Y]  is processed in row chunks, each chunk factorised by the
hand-written blocked Householder kernels of ``csrc/pcq_qr.cu`` (``pcq_qr_r_dev``: cooperative panel kernel, DMMA
trailing updates), and the chunk triangles are combined pairwise in a balanced tree by factorising
``[R_a; R_b]`` with the same kernels,

    R = [[R_A, Q^T Y], [0, rho]]   ((K+M) x (K+M), upper),

so that ``min ||A c - y||`` becomes the K x K problem ``min ||R_A c - Q^T y||`` with the SAME singular
values as A; the truncated-SVD minimum-norm solution of that small problem is what gelsd returns for A.
The design-matrix chunks are produced on the device by K1 (``pcq_eval_bases_dev`` writing straight into
the left K columns of the chunk buffer).  Ranks that hold row blocks combine their triangles by a butterfly
(log2(P) pairwise exchanges, both partners factorise the same stacked pair, so every rank ends with the
identical R); a world size that is not a power of two falls back to an all-gather + tree.  The K x K SVD of R_A is
a library call (cuSOLVER through torch.linalg), like the K x K Cholesky of the Gram route: not sample-parallel.
"""
import numpy as np

from .. import _native, device as _device

_CHUNK_BYTES = 1 << 30


def _torch():
    import torch
    return torch


def _qr_r(Z):
    """R factor (L x L, upper) of the device tensor Z (m x L, overwritten): pcq_qr_r_dev."""
    torch = _torch()
    assert Z.is_cuda and Z.dtype == torch.float64 and Z.dim() == 2 and Z.stride(1) == 1
    m, L = Z.shape
    R = torch.empty((L, L), dtype=torch.float64, device=Z.device)
    if m == 0:
        return R.zero_()
    lib = _native.lib()
    dev = Z.device.index
    cap = int(lib.pcq_qr_max_rows(dev))
    if m > cap:                       # taller than one call takes: factor the pieces, then the stacked triangles
        return combine_r([_qr_r(Z[r0:r0 + cap]) for r0 in range(0, m, cap)])
    with torch.cuda.device(Z.device):
        _native.check(lib.pcq_qr_r_dev(dev, Z.data_ptr(), m, Z.stride(0), L, R.data_ptr(), L,
                                       torch.cuda.current_stream().cuda_stream), "pcq_qr_r_dev")
    return R


def combine_r(Rs, qr=None):
    """R factor of the row-stacked matrices whose R factors are given: balanced pairwise tree, every node the
    factorisation of two stacked triangles.  ``qr`` replaces the device kernel (host-logic tests on CPU)."""
    torch = _torch()
    qr = qr or _qr_r
    level = list(Rs)
    while len(level) > 1:
        nxt = [qr(torch.cat([level[i], level[i + 1]])) for i in range(0, len(level) - 1, 2)]
        if len(level) & 1:
            nxt.append(level[-1])
        level = nxt
    return level[0]


class _Tree:
    """Streaming form of the balanced tree: chunk triangles arrive one by one, two triangles of the same height are
    combined at once, so at most log2(#chunks) triangles are alive."""

    def __init__(self):
        self.levels = {}

    def push(self, R):
        h = 0
        while h in self.levels:
            torch = _torch()
            R = _qr_r(torch.cat([self.levels.pop(h), R]))
            h += 1
        self.levels[h] = R

    def result(self):
        R = None
        for h in sorted(self.levels):
            R = self.levels[h] if R is None else _qr_r(_torch().cat([self.levels[h], R]))
        return R


def _chunk_rows(L, dev):
    cap = int(_native.lib().pcq_qr_max_rows(dev))
    return max(1, min(cap, int(_CHUNK_BYTES // (8 * L))))


def r_factor_pc(pcrv, x, y, jdim=0, chunk_rows=None):
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
    rows = chunk_rows or _chunk_rows(L, dev)
    tdev = torch.device("cuda", dev)
    stream = torch.cuda.current_stream(tdev).cuda_stream
    tree = _Tree()
    for r0 in range(0, max(N, 1), rows):
        r1 = min(N, r0 + rows)
        nr = r1 - r0
        Z = torch.empty((nr, L), dtype=torch.float64, device=tdev)
        if nr:
            xs = torch.from_numpy(x[r0:r1]).to(tdev)
            _native.check(lib.pcq_eval_bases_dev(plan.handle, xs.data_ptr(), nr, s, Z.data_ptr(), L, stream))
            if M:
                Z[:, K:] = torch.from_numpy(Y[r0:r1]).to(tdev)
        tree.push(_qr_r(Z))
    return tree.result()


def r_factor_matrix(Amat, y, chunk_rows=None):
    """R factor of [Amat | y] for a caller-supplied design matrix (host numpy)."""
    torch = _torch()
    _device.require_device()
    dev = _device.current_device()
    A = np.asarray(Amat, dtype=np.float64)
    N, K = A.shape
    Y = None if y is None else np.asarray(y, dtype=np.float64).reshape(N, -1)
    M = 0 if Y is None else Y.shape[1]
    L = K + M
    rows = chunk_rows or _chunk_rows(L, dev)
    tdev = torch.device("cuda", dev)
    tree = _Tree()
    for r0 in range(0, max(N, 1), rows):
        r1 = min(N, r0 + rows)
        Z = torch.empty((r1 - r0, L), dtype=torch.float64, device=tdev)
        Z[:, :K] = torch.from_numpy(np.ascontiguousarray(A[r0:r1])).to(tdev)
        if M:
            Z[:, K:] = torch.from_numpy(np.ascontiguousarray(Y[r0:r1])).to(tdev)
        tree.push(_qr_r(Z))
    return tree.result()


def allgather_r(R, qr=None):
    """Combines the R factors of all ranks of the default process group; every rank gets the same result.
    Power-of-two worlds: butterfly -- log2(P) rounds, rank r and r ^ step exchange their triangles and both factorise
    [R_low_rank; R_high_rank] (same input, same kernels: identical output).  Otherwise all-gather + tree."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return R
    torch = _torch()
    qr = qr or _qr_r
    world, rank = dist.get_world_size(), dist.get_rank()
    R = R.contiguous()
    if world & (world - 1) == 0:
        step = 1
        while step < world:
            partner = rank ^ step
            other = torch.empty_like(R)
            ops = [dist.P2POp(dist.isend, R, partner), dist.P2POp(dist.irecv, other, partner)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            R = qr(torch.cat([R, other] if rank < partner else [other, R]))
            step <<= 1
        return R
    parts = [torch.empty_like(R) for _ in range(world)]
    dist.all_gather(parts, R)
    return combine_r(parts, qr=qr)


def shared_r(R):
    """R factor of the whole data set when the ranks hold row blocks (parallel.sharded_rows), else R itself."""
    from .. import parallel
    if parallel._sharded and parallel.dist_info()[1] > 1:
        return allgather_r(R)
    return R


def svd_jacobi(A):
    """(U Sigma)^T, V^T, sigma of the square device tensor A by the one-sided Jacobi kernels (pcq_svd_jacobi_dev);
    sigma is unsorted, A = U diag(sigma) V^T with row j of the first result = sigma_j u_j^T."""
    import ctypes
    torch = _torch()
    n = A.shape[0]
    G = A.t().contiguous()
    Vt = torch.empty((n, n), dtype=torch.float64, device=A.device)
    sig = torch.empty(n, dtype=torch.float64, device=A.device)
    sweeps = ctypes.c_int(0)
    if n:
        with torch.cuda.device(A.device):
            _native.check(_native.lib().pcq_svd_jacobi_dev(A.device.index, G.data_ptr(), n, Vt.data_ptr(), sig.data_ptr(), 60,
                                                           1.0e-15, ctypes.byref(sweeps),
                                                           torch.cuda.current_stream().cuda_stream), "pcq_svd_jacobi_dev")
    return G, Vt, sig, sweeps.value


def lstsq_from_r(R, K, cond=1.0e-13):
    """Minimum-norm least-squares solution from the R factor of [A | Y]: singular values of R_A (= those of
    A) below ``cond * s_max`` are treated as zero, like gelsd.  The SVD of the K x K factor is the hand-written
    one-sided Jacobi (``svd_jacobi``); with  R_A = U S V^T  and  B = U S,  C = V S^-2 B^T (Q^T Y)  on the kept part,
    the two products on the device through the linear-forms kernels.  Returns (C (K x M) numpy, singular values
    (descending), effective rank)."""
    from .forms import matrix_forms
    torch = _torch()
    M = R.shape[1] - K
    if not R.is_cuda:
        # host tensors: only the GPU-less host-logic tests of the collective plumbing come here (tests/test_dist_gloo.py)
        U, S, Vh = torch.linalg.svd(R[:K, :K])
        s = S.numpy()
        keep = s > cond * (float(s[0]) if s.size else 0.0)
        Wh = (U.T @ R[:K, K:])[keep] / S[keep, None]
        return (Vh[keep].T @ Wh).numpy(), s, int(keep.sum())
    Bt, Vt, sig, _ = svd_jacobi(R[:K, :K])
    s = sig.cpu().numpy()
    smax = float(s.max()) if s.size else 0.0
    keep = s > cond * smax
    rank = int(keep.sum())
    if M == 0 or K == 0:
        return np.zeros((K, M)), np.sort(s)[::-1], rank
    Q = R[:K, K:]
    W = matrix_forms(Bt, Q.t().contiguous())[0]                     # (K, M): row j = sigma_j u_j^T (Q^T Y)
    inv2 = np.zeros_like(s)
    inv2[keep] = 1.0 / s[keep] ** 2
    scale = torch.from_numpy(inv2).to(W.device)
    W = W * scale[:, None]
    C = matrix_forms(Vt.t().contiguous(), W.t().contiguous())[0]    # V (S^-2 B^T Q^T Y)
    return C.cpu().numpy(), np.sort(s)[::-1], rank
