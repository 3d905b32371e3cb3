"""Sample sharding of the PC path across the GPUs of one box (one process per GPU).

The path partitions over samples: basis evaluation, surrogate
evaluation and propagation need no communication at all; a fit reduces each rank's row
block to the packed statistics matrix S and performs ONE all-reduce (sum) of S --
``torch.distributed`` over NCCL/NVLink on GPUs, gloo on CPU for the host-logic tests.
torch is used here only as plumbing (process group, device buffers); all arithmetic is in
libpcq.
"""
import numpy as np

from . import _native, device as _device


def dist_info():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_bounds(n, rank, world):
    """Contiguous row block [lo, hi) of rank `rank` out of `world` (sizes differ by <= 1)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_stats(S):
    """Sums the packed statistics over all ranks (the path's only collective).
    numpy in / numpy out; on GPU ranks the reduction runs over NCCL."""
    rank, world = dist_info()
    if world == 1:
        return S
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(S))
    if dist.get_backend() == "nccl":
        t = t.cuda(_device.current_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def suffstats_pc(pcrv, x, y=None, jdim=0, distributed=False):
    """K3 through the host-pointer C-ABI.  With ``distributed=True`` every rank passes its
    own row block (x, y) and receives the statistics of the union."""
    from .lreg.stats import SuffStats
    mindex = pcrv.mindices[jdim]
    pcrv._check_orders(mindex)
    xin = np.ascontiguousarray(x, dtype=np.float64)
    N, s = xin.shape
    assert s == pcrv.sdim
    if y is None:
        Y, M = None, 0
    else:
        Y = np.ascontiguousarray(y, dtype=np.float64)
        Y = Y.reshape(N, Y.shape[1] if Y.ndim == 2 else 1)
        M = Y.shape[1]
    plan = pcrv._plan(mindex)
    K = mindex.shape[0]
    L = K + M + 1
    S = np.empty((L, L))
    st = _native.lib().pcq_suffstats(plan.handle, _native.ptr(xin), N, s, _native.ptr(Y), max(M, 1), M,
                                     _native.ptr(S))
    _native.check(st, "pcq_suffstats")
    if distributed:
        S = allreduce_stats(S)
    return SuffStats(S, K, M)


def propagate(pcrv, nsam, seed=0, exact=False, return_samples=False, first_index=0):
    """K5 on the current device: device germ (Philox) -> evalPC -> power sums.  Under a
    process group each rank draws its own block of the one logical stream and the sums
    are combined with one all-reduce."""
    import torch
    _device.require_device()
    dev = _device.current_device()
    rank, world = dist_info()
    lo, hi = shard_bounds(nsam, rank, world)
    groups = {}
    for j, mindex in enumerate(pcrv.mindices):
        pcrv._check_orders(mindex)
        groups.setdefault(id(mindex), []).append(j)
    kinds = pcrv._germ_kinds()
    tdev = torch.device("cuda", dev)
    sums = np.zeros((pcrv.pdim, 4))
    ysam = np.empty((hi - lo, pcrv.pdim)) if return_samples else None
    lib = _native.lib()
    stream = torch.cuda.current_stream(tdev).cuda_stream
    for outs in groups.values():
        plan = pcrv._plan(pcrv.mindices[outs[0]], dev=dev)
        cf = torch.from_numpy(np.ascontiguousarray(np.stack([pcrv.coefs[j] for j in outs]))).to(tdev)
        dsums = torch.empty((len(outs), 4), dtype=torch.float64, device=tdev)
        dy = torch.empty((hi - lo, len(outs)), dtype=torch.float64, device=tdev) if return_samples else None
        st = lib.pcq_propagate_dev(plan.handle, int(seed), int(first_index + lo), int(hi - lo),
                                   _native.ptr(kinds), cf.data_ptr(), len(outs), dsums.data_ptr(),
                                   dy.data_ptr() if dy is not None else None, len(outs),
                                   0 if exact else 1, stream)
        _native.check(st, "pcq_propagate_dev")
        sums[outs] = dsums.cpu().numpy()
        if return_samples:
            ysam[:, outs] = dy.cpu().numpy()
    if world > 1:
        import torch.distributed as dist
        t = torch.from_numpy(sums[:, :2].copy())
        mn = torch.from_numpy(sums[:, 2].copy())
        mx = torch.from_numpy(sums[:, 3].copy())
        if dist.get_backend() == "nccl":
            t, mn, mx = t.to(tdev), mn.to(tdev), mx.to(tdev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sums = np.column_stack([t.cpu().numpy(), mn.cpu().numpy(), mx.cpu().numpy()])
    n = float(nsam)
    mean = sums[:, 0] / n
    out = dict(n=int(nsam), sum=sums[:, 0], sumsq=sums[:, 1], mean=mean,
               var=np.maximum(sums[:, 1] / n - mean ** 2, 0.0), min=sums[:, 2], max=sums[:, 3])
    if return_samples:
        out["samples"] = ysam
    return out
