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


def get_plan(mi, rec_a, rec_b, p1_scale, pmax, device=None):
    """Returns the cached plan for this content, creating it on first use."""
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
            return plan
    K, s = mi32.shape
    handle = C.c_void_p()
    st = _native.lib().pcq_plan_create(C.byref(handle), device, s, K, _native.ptr(mi32), int(pmax),
                                       _native.ptr(rec_a), _native.ptr(rec_b), _native.ptr(p1_scale))
    _native.check(st, "pcq_plan_create")
    plan = Plan(handle, s, K, int(pmax), device)
    with _lock:
        _plans[key] = plan
        while len(_plans) > _PLAN_CACHE_MAX:
            _plans.popitem(last=False)
    return plan


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


def _close_all_at_exit():
    # destroy the cached plans (joining background builds) before the interpreter tears the CUDA runtime down
    with _lock:
        plans = list(_plans.values())
        _plans.clear()
    for plan in plans:
        try:
            plan.close()
        except Exception:
            pass


import atexit  # noqa: E402
atexit.register(_close_all_at_exit)
