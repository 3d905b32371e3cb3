"""Device selection and the content-keyed plan cache.

PCRV / lreg objects must stay picklable and deep-copyable (the reference pickles them with
dill, utils/xutils.py:36-55, and deep-copies them in surrogates/pce.py:232), so no native
handle is ever stored on them: plans live in this module-level cache, keyed by the content
of (device, multi-index, families, tabulated order).
"""
import collections
import ctypes as C
import os
import threading

import numpy as np

from . import _native

_lock = threading.Lock()
_current_device = None
_PLAN_CACHE_MAX = int(os.environ.get("PCQ_PLAN_CACHE", "32"))
_plans = collections.OrderedDict()


def set_device(index):
    """Selects the CUDA device the hot path runs on (default: LOCAL_RANK, else 0)."""
    global _current_device
    _current_device = int(index)


def current_device():
    global _current_device
    if _current_device is None:
        _current_device = int(os.environ.get("PCQ_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    return _current_device


def require_device():
    n = _native.lib().pcq_device_count()
    if n <= 0:
        raise _native.PcqError("no CUDA device visible: pytuq_b200 has no CPU fallback "
                               f"(pcq_device_count() = {n})")
    return n


class Plan:
    """Owner of one native pcq_plan (device-resident multi-index + recurrence tables)."""

    def __init__(self, handle, sdim, nterms, pmax, device):
        self.handle = handle
        self.sdim, self.nterms, self.pmax, self.device = sdim, nterms, pmax, device

    def info(self, what):
        return int(_native.lib().pcq_plan_info(self.handle, what))

    @property
    def width(self):
        return self.info(3)

    def close(self):
        if self.handle:
            _native.lib().pcq_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _enable_moments(obj, fn_name, ext):
    """Hands the recurrence tables up to twice the total order to a plan / group (moment form of K3, include/pcq.h:
    pcq_plan_enable_moments); once per object."""
    if ext is None or getattr(obj, "_moments", False):
        return
    nord, ra, rb = ext
    ra = np.ascontiguousarray(ra, dtype=np.float64)
    rb = np.ascontiguousarray(rb, dtype=np.float64)
    _native.check(getattr(_native.lib(), fn_name)(obj.handle, int(nord), _native.ptr(ra), _native.ptr(rb)), fn_name)
    obj._moments = True


def get_plan(mi, rec_a, rec_b, p1_scale, pmax, device=None, ext=None):
    """Returns the cached plan for this content, creating it on first use.  ext = (nord, rec_a_ext, rec_b_ext) enables
    the moment form of the sufficient statistics on it.

    Threading: a plan (and a group) is NOT re-entrant -- its host-pointer entry points share per-plan staging
    buffers, streams and events.  Two Python threads may use the library at the same time only on PCRVs with
    different content (different plans); callers keep the returned object referenced for the duration of the native
    call (the LRU below may otherwise destroy a handle another thread is still using)."""
    require_device()
    device = current_device() if device is None else int(device)
    mi32 = np.ascontiguousarray(mi, dtype=np.int32)
    rec_a = np.ascontiguousarray(rec_a, dtype=np.float64)
    rec_b = np.ascontiguousarray(rec_b, dtype=np.float64)
    p1_scale = np.ascontiguousarray(p1_scale, dtype=np.float64)
    key = (device, mi32.shape, mi32.tobytes(), rec_a.tobytes(), rec_b.tobytes(), p1_scale.tobytes(), int(pmax))
    with _lock:
        plan = _plans.get(key)
        if plan is not None:
            _plans.move_to_end(key)
            _enable_moments(plan, "pcq_plan_enable_moments", ext)
            return plan
    K, s = mi32.shape
    handle = C.c_void_p()
    st = _native.lib().pcq_plan_create(C.byref(handle), device, s, K, _native.ptr(mi32), int(pmax),
                                       _native.ptr(rec_a), _native.ptr(rec_b), _native.ptr(p1_scale))
    _native.check(st, "pcq_plan_create")
    plan = Plan(handle, s, K, int(pmax), device)
    _enable_moments(plan, "pcq_plan_enable_moments", ext)
    with _lock:
        _plans[key] = plan
        while len(_plans) > _PLAN_CACHE_MAX:
            _plans.popitem(last=False)
    return plan


# ---------------------------------------------------------------------------------------------------------------
# Several GPUs in ONE process (pcq_group, include/pcq.h): calls that take host (numpy) arrays are cut into row blocks,
# one per device, when the process sees more than one GPU and the call is large enough to pay for the fan-out.
# Policy: PCQ_DEVICES = "all" | "1" | "0,2,3" (explicit list; the first entry is where the K x K statistics land);
# default "all" in a plain process and the rank's own device under torchrun (RANK / LOCAL_RANK in the environment),
# where the sharding is over processes instead (parallel.sharded_rows).  set_devices() overrides the environment.
_devices_override = None
_groups = collections.OrderedDict()


def set_devices(devices):
    """Devices the host-array entry points may use in this process: a list of CUDA indices, or None for the
    environment / default policy.  The first entry holds the reduced statistics and runs the K x K solves."""
    global _devices_override
    _devices_override = None if devices is None else [int(d) for d in devices]


def devices():
    """The device list the policy above yields right now (always at least [current_device()])."""
    cur = current_device()
    if _devices_override is not None:
        devs = list(_devices_override)
    else:
        spec = os.environ.get("PCQ_DEVICES", "").strip().lower()
        in_rank = int(os.environ.get("WORLD_SIZE", "1")) > 1 or "LOCAL_RANK" in os.environ
        if spec in ("", "all"):
            devs = [cur] if (in_rank and spec == "") else list(range(require_device()))
        elif spec == "1":
            devs = [cur]
        else:
            devs = [int(t) for t in spec.split(",") if t.strip() != ""]
    if not devs:
        devs = [cur]
    if cur in devs:                       # the current device leads: statistics and solves stay where plans live
        devs = [cur] + [d for d in devs if d != cur]
    return devs


class Group:
    """Owner of one native pcq_group (one plan per device)."""

    def __init__(self, handle, devs, sdim, nterms, pmax):
        self.handle, self.devices = handle, list(devs)
        self.sdim, self.nterms, self.pmax = sdim, nterms, pmax

    def __len__(self):
        return len(self.devices)

    def close(self):
        if self.handle:
            _native.lib().pcq_group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_group(mi, rec_a, rec_b, p1_scale, pmax, devs, ext=None):
    """Cached pcq_group for this content on this device list (created on first use); ext as in get_plan."""
    require_device()
    mi32 = np.ascontiguousarray(mi, dtype=np.int32)
    rec_a = np.ascontiguousarray(rec_a, dtype=np.float64)
    rec_b = np.ascontiguousarray(rec_b, dtype=np.float64)
    p1_scale = np.ascontiguousarray(p1_scale, dtype=np.float64)
    devs = [int(d) for d in devs]
    key = (tuple(devs), mi32.shape, mi32.tobytes(), rec_a.tobytes(), rec_b.tobytes(), p1_scale.tobytes(), int(pmax))
    with _lock:
        grp = _groups.get(key)
        if grp is not None:
            _groups.move_to_end(key)
            _enable_moments(grp, "pcq_group_enable_moments", ext)
            return grp
    K, s = mi32.shape
    handle = C.c_void_p()
    darr = (C.c_int * len(devs))(*devs)
    st = _native.lib().pcq_group_create(C.byref(handle), len(devs), darr, s, K, _native.ptr(mi32), int(pmax),
                                        _native.ptr(rec_a), _native.ptr(rec_b), _native.ptr(p1_scale))
    _native.check(st, "pcq_group_create")
    grp = Group(handle, devs, s, K, int(pmax))
    _enable_moments(grp, "pcq_group_enable_moments", ext)
    with _lock:
        _groups[key] = grp
        while len(_groups) > max(2, _PLAN_CACHE_MAX // 4):
            _groups.popitem(last=False)
    return grp


def fan_out(work, unit, rows=None):
    """Number of devices to spread a host-array call over: as many of devices() as the call has `unit`s of work (each
    worth a few milliseconds on one GPU), at least 4096 rows per device; 1 means the single-plan path."""
    nd = len(devices())
    if nd == 1:
        return 1
    unit = float(unit) * float(os.environ.get("PCQ_FANOUT_SCALE", "1"))     # (tests shrink the unit to force the fan-out)
    n = int(min(nd, max(1.0, float(work) / unit)))
    if rows is not None:
        n = min(n, max(1, int(rows) // 4096))
    return max(n, 1)


# A plan that keeps being evaluated (MCMC likelihoods, repeated surrogate sweeps) is worth compiling for: once the
# generic evalPC kernel has processed this many sample-terms through one plan (~20 ms of kernel time), its
# multi-index-specialised kernels are built on a BACKGROUND host thread (NVRTC, ~1 s per 2000 terms; cubins are
# cached on disk) and installed when ready -- evaluations never wait for the compiler, they just get 3-4x faster,
# bit-identically in exact mode, from some call on.  PCQ_AUTO_SPECIALIZE=0 disables, any other number replaces
# the threshold.
_AUTO_SPECIALIZE = float(os.environ.get("PCQ_AUTO_SPECIALIZE", "2e10"))
_AUTO_SPECIALIZE_MAX_TERMS = 40000


def note_evalpc_work(plan, nsam, nterms, nout, mode):
    if _AUTO_SPECIALIZE <= 0 or nterms > _AUTO_SPECIALIZE_MAX_TERMS or nout > 8 or plan.sdim * max(plan.pmax, 1) > 256:
        return      # (wide germs: the unrolled kernels would keep the whole 1-D table in registers)
    bit = 2 if (mode & 1) else 1
    if getattr(plan, "_spec_tried", 0) & bit:
        return
    work = getattr(plan, "_evalpc_work", 0.0) + float(nsam) * nterms * nout
    plan._evalpc_work = work
    if work >= _AUTO_SPECIALIZE:
        plan._spec_tried = getattr(plan, "_spec_tried", 0) | bit
        _native.lib().pcq_plan_specialize_async(plan.handle, bit)


def clear_plans():
    """Drops the cached plans; a native handle is destroyed when the last reference to its Plan goes (which also
    waits for a background specialisation build that may still be running)."""
    with _lock:
        _plans.clear()
        _groups.clear()


def _close_all_at_exit():
    # destroy the cached plans (joining background builds) before the interpreter tears the CUDA runtime down
    with _lock:
        plans = list(_plans.values()) + list(_groups.values())
        _plans.clear()
        _groups.clear()
    for plan in plans:
        try:
            plan.close()
        except Exception:
            pass


import atexit  # noqa: E402
atexit.register(_close_all_at_exit)
