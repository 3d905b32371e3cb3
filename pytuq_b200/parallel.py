"""Sample sharding of the PC path across the GPUs of one box (one process per GPU).

The path partitions over samples: basis evaluation, surrogate
evaluation and propagation need no communication at all; a fit reduces each rank's row
block to the packed statistics matrix S and performs ONE all-reduce (sum) of S --
``torch.distributed`` over NCCL/NVLink on GPUs, gloo on CPU for the host-logic tests.
torch is used here only as plumbing (process group, device buffers); all arithmetic is in
libpcq.
"""
import contextlib
import os
import time

import numpy as np

from . import _native, device as _device

# When set (and a process group with more than one rank is initialised) the (x, y) every rank passes to
# lreg.fit / pc_fit / PCE.build are ITS ROW BLOCK of one global data set: the statistics pass all-reduces S and the
# residual pass all-reduces its two sums, so every rank ends with the fit of the whole data set.  Off by default --
# ranks that each hold the full data must not add their statistics up.
_sharded = False


def set_sharded_rows(flag=True):
    global _sharded
    _sharded = bool(flag)


@contextlib.contextmanager
def sharded_rows(flag=True):
    """``with parallel.sharded_rows(): lsq.fit(x_block, y_block)`` -- see above."""
    global _sharded
    old, _sharded = _sharded, bool(flag)
    try:
        yield
    finally:
        _sharded = old


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


_GROUP_STAGES = 8              # stages per call of the statistics kernels once the pipeline has ramped up (1, 2, 4, 8, 8, ..)
_STAGE_BYTES = 64 << 20       # host chunk: large enough that the per-call cost of the statistics kernels (launch tail,
                               # moment reconstruction ~1 ms at K = 1771) stays below a few per cent
_pinned = {}


def _stage_groups(n, rows, max_stages=None):
    """Row ranges (r0, nr) of the host stages of n rows, `rows` per stage, gathered into groups of 1, 2, 4, .. max_stages
    stages: one statistics call per group."""
    max_stages = _GROUP_STAGES if max_stages is None else max_stages
    stages, r0 = [], 0
    while r0 < n:
        stages.append((r0, min(rows, n - r0)))
        r0 += stages[-1][1]
    groups, i, size = [], 0, 1
    while i < len(stages):
        groups.append(stages[i:i + size])
        i += size
        size = min(2 * size, max_stages)
    return groups


def _pinned_pair(nbytes, dev):
    """Two reusable pinned staging buffers per device (double buffering of the H2D pipeline)."""
    import torch
    key = "bufs%d" % dev
    bufs = _pinned.get(key)
    if bufs is None or bufs[0].numel() * 8 < nbytes:
        n = max(int(nbytes) // 8, 1)
        bufs = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
        _pinned[key] = bufs
    return bufs


def _copy_stream(dev, tdev):
    import torch
    key = "stream%d" % dev
    if key not in _pinned:
        _pinned[key] = torch.cuda.Stream(device=tdev)
    return _pinned[key]


def is_device_tensor(a):
    """True for a torch tensor that lives on a CUDA device (samples loaded by utils.io.loadtxt_device, or put
    there by the caller): the fit then reads them where they are."""
    return type(a).__module__.split(".")[0] == "torch" and bool(getattr(a, "is_cuda", False))


def _resident(a, rows=None):
    """float64, contiguous, 2-D view of a device tensor."""
    import torch
    t = a if a.dtype == torch.float64 else a.to(torch.float64)
    if t.dim() == 1:
        t = t.reshape(-1, 1)
    assert t.dim() == 2 and (rows is None or t.shape[0] == rows)
    return t.contiguous()


def _suffstats_pc_resident(pcrv, x, y, jdim, distributed):
    """K3 on samples that are already in HBM (torch CUDA tensors): one pcq_suffstats_dev call, no staging."""
    import torch
    from .lreg.stats import SuffStats
    mindex = pcrv.mindices[jdim]
    pcrv._check_orders(mindex)
    xt = _resident(x)
    N, s = xt.shape
    assert s == pcrv.sdim
    yt = None if y is None else _resident(y, N)
    M = 0 if yt is None else yt.shape[1]
    dev = xt.device.index
    assert yt is None or yt.device == xt.device
    plan = pcrv._plan(mindex, dev=dev)
    K = mindex.shape[0]
    S_t = torch.empty((K + M + 1, K + M + 1), dtype=torch.float64, device=xt.device)
    with torch.cuda.device(xt.device):
        stream = torch.cuda.current_stream()
        _native.check(_native.lib().pcq_suffstats_dev(plan.handle, xt.data_ptr() if N else None, N, s,
                                                      yt.data_ptr() if (M and N) else None, max(M, 1), M,
                                                      S_t.data_ptr(), 0, stream.cuda_stream), "pcq_suffstats_dev")
        if distributed and dist_info()[1] > 1:
            import torch.distributed as dist
            if dist.get_backend() == "nccl":
                dist.all_reduce(S_t, op=dist.ReduceOp.SUM)
            else:
                S_t = torch.from_numpy(allreduce_stats(S_t.cpu().numpy())).to(xt.device)
        stream.synchronize()
    pcrv._distributed_stats = bool(distributed)
    return SuffStats(None, K, M, S_dev=S_t)


def suffstats_pc(pcrv, x, y=None, jdim=0, distributed=None):
    """K3 on device-resident samples (torch CUDA tensors: read in place) or on host (numpy) inputs: the rows are streamed through two pinned staging buffers in L2-sized
    chunks (host memcpy and H2D of chunk i+1 overlap the Gram kernel of chunk i), the statistics are
    accumulated and stay on the device.  With ``distributed=True`` every rank passes its own row
    block (x, y) and the partial statistics are all-reduced on the device (NCCL)."""
    import torch
    from .lreg.stats import SuffStats
    _device.require_device()
    if distributed is None:
        distributed = _sharded and dist_info()[1] > 1
    if is_device_tensor(x):
        return _suffstats_pc_resident(pcrv, x, y, jdim, distributed)
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
    dev = _device.current_device()
    tdev = torch.device("cuda", dev)
    K = mindex.shape[0]
    L = K + M + 1
    lib = _native.lib()
    S_t = torch.empty((L, L), dtype=torch.float64, device=tdev)
    # several GPUs in this process (and no process group sharding the rows already): the rows are cut into one block
    # per device inside the C-ABI (pcq_group_suffstats: pinned pipelines per device, partial S summed on this device)
    nuse = 1 if distributed else _device.fan_out(float(N) * K * (K + 2 * M + 1), 5e11, N)
    if nuse > 1:
        grp = pcrv._group(mindex, nuse)
        assert grp.devices[0] == dev
        torch.cuda.current_stream(tdev).synchronize()
        _native.check(lib.pcq_group_suffstats(grp.handle, _native.ptr(xin), N, s, _native.ptr(Y), max(M, 1), M,
                                              S_t.data_ptr(), None), "pcq_group_suffstats")
        pcrv._distributed_stats = False
        return SuffStats(None, K, M, S_dev=S_t)
    plan = pcrv._plan(mindex, dev=dev)
    ncol = s + M
    rows = max(_STAGE_BYTES // (ncol * 8), 4096)
    pin = _pinned_pair(rows * ncol * 8, dev)
    # Host rows travel stage by stage (two pinned buffers) into one of two device GROUP buffers; the statistics kernels
    # run once per group, not per stage: their per-call cost (launch tails of the task-shape kernels, the K x K
    # reconstruction of the moment form) is paid a handful of times per fit.  Groups ramp up (1, 2, 4, .. _GROUP_STAGES
    # stages) so that the first kernels start after one stage while the copies stay ahead of the kernels.
    groups = _stage_groups(N, rows)
    grows = max([sum(nr for _, nr in g) for g in groups], default=0)
    threaded_copy = os.environ.get("PCQ_PY_HOST_COPY", "threads") != "numpy"
    trace = {"wait_pinned": 0.0, "host_copy": 0.0, "enqueue_h2d": 0.0, "launch_call": 0.0} if os.environ.get("PCQ_PY_TRACE") else None
    with torch.cuda.device(tdev):
        main = torch.cuda.current_stream()
        copy_stream = _copy_stream(dev, tdev)
        gbuf = [torch.empty(max(grows, 1) * ncol, dtype=torch.float64, device=tdev) for _ in range(min(2, max(len(groups), 1)))]
        # the group buffers come from the caching allocator on the main stream: work still queued there on a recycled
        # block must finish before the copy stream overwrites it
        copy_stream.wait_stream(main)
        for t in gbuf:
            t.record_stream(copy_stream)
        copied = [torch.cuda.Event() for _ in range(2)]        # H2D out of pinned buffer b finished
        filled = [torch.cuda.Event() for _ in range(2)]        # every stage of the group in buffer gb has arrived
        consumed = [torch.cuda.Event() for _ in range(2)]      # the kernels of the group in buffer gb are done
        if N == 0:
            _native.check(lib.pcq_suffstats_dev(plan.handle, None, 0, s, None, max(M, 1), M, S_t.data_ptr(), 0,
                                                main.cuda_stream), "pcq_suffstats_dev")
        c = 0
        for gi, grp_stages in enumerate(groups):
            gb = gi & 1
            gn = sum(nr for _, nr in grp_stages)
            gx = gbuf[gb][:gn * s].view(gn, s)
            gy = gbuf[gb][gn * s:gn * ncol].view(gn, M) if M else None
            if gi >= 2:
                copy_stream.wait_event(consumed[gb])       # the kernels of group gi - 2 have read this buffer
            off = 0
            for r0, nr in grp_stages:
                b = c & 1
                _t0 = time.perf_counter() if trace is not None else 0.0
                if c >= 2:
                    copied[b].synchronize()                # pinned buffer b free again
                _t1 = time.perf_counter() if trace is not None else 0.0
                hx = pin[b][:nr * s].view(nr, s)
                hy = pin[b][nr * s:nr * ncol].view(nr, M) if M else None
                if threaded_copy:
                    _native.check(lib.pcq_host_copy_rows(hx.data_ptr(), s, xin[r0:r0 + nr].ctypes.data, s, s, nr), "pcq_host_copy_rows")
                    if M:
                        _native.check(lib.pcq_host_copy_rows(hy.data_ptr(), M, Y[r0:r0 + nr].ctypes.data, M, M, nr), "pcq_host_copy_rows")
                else:
                    np.copyto(hx.numpy(), xin[r0:r0 + nr])
                    if M:
                        np.copyto(hy.numpy(), Y[r0:r0 + nr])
                _t2 = time.perf_counter() if trace is not None else 0.0
                with torch.cuda.stream(copy_stream):
                    gx[off:off + nr].copy_(hx, non_blocking=True)
                    if M:
                        gy[off:off + nr].copy_(hy, non_blocking=True)
                    copied[b].record(copy_stream)
                if trace is not None:
                    _t3 = time.perf_counter()
                    trace["wait_pinned"] += _t1 - _t0
                    trace["host_copy"] += _t2 - _t1
                    trace["enqueue_h2d"] += _t3 - _t2
                off += nr
                c += 1
            filled[gb].record(copy_stream)
            main.wait_event(filled[gb])
            _t4 = time.perf_counter() if trace is not None else 0.0
            _native.check(lib.pcq_suffstats_dev(plan.handle, gx.data_ptr(), gn, s, gy.data_ptr() if M else None, max(M, 1), M,
                                                S_t.data_ptr(), 1 if gi > 0 else 0, main.cuda_stream), "pcq_suffstats_dev")
            consumed[gb].record(main)
            if trace is not None:
                trace["launch_call"] += time.perf_counter() - _t4
        if distributed and dist_info()[1] > 1:
            import torch.distributed as dist
            if dist.get_backend() == "nccl":
                dist.all_reduce(S_t, op=dist.ReduceOp.SUM)
            else:
                S_t = torch.from_numpy(allreduce_stats(S_t.cpu().numpy())).to(tdev)
        _t5 = time.perf_counter()
        main.synchronize()
        if trace is not None:
            trace["final_sync"] = time.perf_counter() - _t5
            trace["groups"] = [sum(nr for _, nr in g) for g in groups]
            print("suffstats_pc host path (seconds):", trace, flush=True)
    pcrv._distributed_stats = bool(distributed)
    return SuffStats(None, K, M, S_dev=S_t)


def residual_sums_pc(pcrv, x, y, col, jdim, cf):
    """(sum_n r_n^2, sum_n y_n r_n) with r = y[:, col] - Psi(x) cf, evaluated over the data: the surrogate
    through K2 (fast mode, no A), the two sums reduced on the device (pcq_residual_ss).  This is the residual
    pass of bcs_fit (lreg/bcs.py:229) and anl (lreg/anl.py:61) for fits whose residual is so small that the
    statistics identity  y^T y - 2 cf^T b + cf^T G cf  has cancelled (noise-free training data)."""
    _device.require_device()
    mindex = pcrv.mindices[jdim]
    if is_device_tensor(x):
        import torch
        xt = _resident(x)
        yt = _resident(y, xt.shape[0])
        plan = pcrv._plan(mindex, dev=xt.device.index)
        cf_t = torch.from_numpy(np.ascontiguousarray(cf, dtype=np.float64)).to(xt.device)
        sums_t = torch.zeros(2, dtype=torch.float64, device=xt.device)
        with torch.cuda.device(xt.device):
            _native.check(_native.lib().pcq_residual_ss_dev(
                plan.handle, xt.data_ptr(), xt.shape[0], xt.shape[1], yt.data_ptr() + 8 * int(col), yt.shape[1],
                cf_t.data_ptr(), sums_t.data_ptr(), 0, torch.cuda.current_stream().cuda_stream), "pcq_residual_ss_dev")
        sums = sums_t.cpu().numpy()
        if dist_info()[1] > 1 and getattr(pcrv, "_distributed_stats", False):
            sums = allreduce_stats(sums)
        return float(sums[0]), float(sums[1])
    xin = np.ascontiguousarray(x, dtype=np.float64)
    N, s = xin.shape
    Y = np.asarray(y, dtype=np.float64).reshape(N, -1)
    ycol = np.ascontiguousarray(Y[:, col])
    sums = np.zeros(2)
    sharded = dist_info()[1] > 1 and getattr(pcrv, "_distributed_stats", False)
    nuse = 1 if sharded else _device.fan_out(float(N) * mindex.shape[0], 4e9, N)
    if nuse > 1:
        _native.check(_native.lib().pcq_group_residual_ss(pcrv._group(mindex, nuse).handle, _native.ptr(xin), N, s,
                                                          _native.ptr(ycol), 1,
                                                          _native.ptr(np.ascontiguousarray(cf, dtype=np.float64)),
                                                          _native.ptr(sums)), "pcq_group_residual_ss")
        return float(sums[0]), float(sums[1])
    plan = pcrv._plan(mindex, dev=_device.current_device())
    _native.check(_native.lib().pcq_residual_ss(plan.handle, _native.ptr(xin), N, s, _native.ptr(ycol), 1,
                                                _native.ptr(np.ascontiguousarray(cf, dtype=np.float64)),
                                                _native.ptr(sums)), "pcq_residual_ss")
    if dist_info()[1] > 1 and getattr(pcrv, "_distributed_stats", False):
        sums = allreduce_stats(sums)
    return float(sums[0]), float(sums[1])


def residual_sums_matrix(A, Y, col, cf):
    """Same two sums for a caller-supplied design matrix (host numpy), chunks contracted on the device."""
    import torch
    _device.require_device()
    tdev = torch.device("cuda", _device.current_device())
    N, K = A.shape
    rows = max(1, (256 << 20) // (8 * K))
    sums = torch.zeros(2, dtype=torch.float64, device=tdev)
    cf_t = torch.from_numpy(np.ascontiguousarray(cf, dtype=np.float64)).to(tdev)
    for r0 in range(0, N, rows):
        r1 = min(N, r0 + rows)
        Ad = torch.from_numpy(np.ascontiguousarray(A[r0:r1])).to(tdev)
        yd = torch.from_numpy(np.ascontiguousarray(Y[r0:r1, col])).to(tdev)
        r = yd - Ad @ cf_t
        sums += torch.stack([torch.dot(r, r), torch.dot(yd, r)])
    out = sums.cpu().numpy()
    return float(out[0]), float(out[1])


def propagate(pcrv, nsam, seed=0, exact=False, return_samples=False, first_index=0, specialize="auto"):
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
    # multi-index-specialised kernels (NVRTC, ~0.2 ms of compile time per term on the GPU box's host, cubins cached on
    # disk afterwards) save ~8e-13 s per sample-term in fast mode: "auto" specialises from 1e12 sample-terms per rank
    # (about a second of generic-kernel time), True / False force it
    if specialize is True or (specialize == "auto" and float(hi - lo) * max(m.shape[0] for m in pcrv.mindices) >= 1.0e12):
        pcrv.specialize(exact=exact, fast=not exact)
    tdev = torch.device("cuda", dev)
    sums = np.zeros((pcrv.pdim, 4))
    ysam = np.empty((hi - lo, pcrv.pdim)) if return_samples else None
    lib = _native.lib()
    stream = torch.cuda.current_stream(tdev).cuda_stream
    nuse = 1
    if world == 1 and not return_samples:
        nuse = _device.fan_out(float(nsam) * max(m.shape[0] for m in pcrv.mindices), 2e10, nsam)
    for outs in groups.values():
        if nuse > 1:
            # one process, several GPUs: disjoint index blocks of the one logical stream per device (pcq_group_propagate)
            mindex = pcrv.mindices[outs[0]]
            grp = pcrv._group(mindex, nuse)
            cfh = np.ascontiguousarray(np.stack([pcrv.coefs[j] for j in outs]), dtype=np.float64)
            part = np.zeros((len(outs), 4))
            _native.check(lib.pcq_group_propagate(grp.handle, int(seed), int(first_index), int(nsam), _native.ptr(kinds),
                                                  _native.ptr(cfh), len(outs), _native.ptr(part), 0 if exact else 1),
                          "pcq_group_propagate")
            sums[outs] = part
            continue
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
