"""Sample files of the PC workflows: ``numpy.savetxt`` / ``numpy.loadtxt`` with their default options, done by
all host cores through ``libpcq_io.so`` (``include/pcq_io.h``).

PyTUQ's applications pass germ samples, parameter samples and model outputs between their steps as text files
written by ``np.savetxt(fname, X)`` and read by ``np.loadtxt(fname)`` (``apps/uqpc/uq_pc.py:202-271``).  The two
functions below keep that format — ``savetxt`` writes the same bytes, ``loadtxt`` returns the same float64 values
and the same shape — so files stay interchangeable with the reference in both directions.  Arguments other than
the ones listed are deliberately not offered: a call that needs them is not this path and should use numpy.

``load`` / ``save`` pick the format from the file name: ``.npy`` goes through numpy's binary format (the one to use
for the 1e7-row configurations: 8 bytes per value instead of 25, no conversion), anything else is text.
"""
import ctypes as C
import os
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("PCQ_IO_LIB", os.path.join(_HERE, "libpcq_io.so"))

# name -> (restype, argtypes); must stay in step with include/pcq_io.h (tests check this)
SIGNATURES = {
    "pcqio_version": (C.c_int, []),
    "pcqio_last_error": (C.c_char_p, []),
    "pcqio_savetxt": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "pcqio_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                             C.c_int]),
    "pcqio_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "pcqio_close": (C.c_int, [C.c_void_p]),
    "pcqio_selftest": (C.c_int64, [C.c_uint64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}

ERR_ARG, ERR_IO, ERR_PARSE, ERR_RAGGED, ERR_ALLOC = -1, -2, -3, -4, -5

_lib = None


def lib():
    """The loaded library; raises when it has not been built (no numpy fallback: a silent one would hide it)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C pytuq_b200/csrc`")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _raise(status, fname):
    msg = (lib().pcqio_last_error() or b"").decode(errors="replace")
    if status == ERR_IO:
        raise OSError(msg or f"I/O error on {fname}")
    if status == ERR_ALLOC:
        raise MemoryError(msg)
    # numpy.loadtxt raises ValueError for unparsable tokens and for ragged rows
    raise ValueError(msg)


def savetxt(fname, X, threads=0):
    """``np.savetxt(fname, X)``: one row per line, values as ``'%.18e'`` separated by one space.

    A 1-D array is written one value per line and integer / float32 input is converted to float64 first, both
    as numpy does (``numpy/lib/_npyio_impl.py``, ``savetxt``)."""
    if type(X).__module__.split(".")[0] == "torch":      # results that live in HBM (or any torch tensor) come to the host first
        X = X.detach().cpu().numpy()
    X = np.asarray(X)
    if X.ndim == 0 or X.ndim > 2:
        raise ValueError("Expected 1D or 2D array, got %dD array instead" % X.ndim)
    if np.iscomplexobj(X) or X.dtype.kind not in "fiub":
        raise TypeError(f"pytuq_b200.utils.io.savetxt writes real numeric arrays only, not {X.dtype}")
    a = np.ascontiguousarray(X if X.ndim == 2 else X[:, None], dtype=np.float64)
    status = lib().pcqio_savetxt(os.fsencode(fname), a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1],
                                 max(a.shape[1], 1), int(threads))
    if status < 0:
        _raise(status, fname)


def loadtxt(fname, ndmin=0, threads=0):
    """``np.loadtxt(fname, ndmin=ndmin)`` for float64 tables: whitespace-separated columns, ``#`` comments,
    blank lines skipped.  Same values bit for bit, same shape rules (squeezed unless ``ndmin=2``)."""
    if ndmin not in (0, 1, 2):
        raise ValueError(f"Illegal value of ndmin keyword: {ndmin}")
    handle = C.c_void_p()
    rows, cols = C.c_int64(), C.c_int64()
    status = lib().pcqio_open(os.fsencode(fname), C.byref(handle), C.byref(rows), C.byref(cols), int(threads))
    if status < 0:
        if status == ERR_IO and not os.path.exists(fname):
            raise FileNotFoundError(f"{fname} not found.")
        _raise(status, fname)
    try:
        out = np.empty((rows.value, cols.value), dtype=np.float64)
        status = lib().pcqio_read(handle, out.ctypes.data_as(C.c_void_p), max(cols.value, 1))
        if status < 0:
            _raise(status, fname)
    finally:
        lib().pcqio_close(handle)
    if out.shape[0] == 0:
        warnings.warn("loadtxt: input contained no data: \"%s\"" % fname, category=UserWarning, stacklevel=2)
        return np.empty((0, 1) if ndmin == 2 else (0,), dtype=np.float64)
    if ndmin == 2:
        return out
    out = np.squeeze(out)
    return np.atleast_1d(out) if ndmin == 1 else out


def loadtxt_device(fname, device=None):
    """The table of ``fname`` as a float64 ``torch`` tensor of shape (rows, cols) in the memory of CUDA device
    ``device`` (default: the path's current device), converted ON the device from the file's text
    (``pcq_text_open`` / ``pcq_text_parse_dev``, ``include/pcq.h``): no host parse and no upload of the binary
    matrix.  Values are bit-identical to ``np.loadtxt(fname, ndmin=2)``.  ``pc_fit`` / ``lreg.fit`` /
    ``PCRV.suffStats`` take such tensors in place of numpy arrays and read them where they are.

    Whitespace-separated tables (what ``savetxt`` writes; inf / nan in numpy's spellings included): a file with
    ``#`` comments raises ``ValueError`` naming the host reader ``loadtxt``; ragged rows and unconvertible tokens
    raise ``ValueError`` as ``np.loadtxt`` does."""
    import torch
    from .. import _native, device as _device
    _device.require_device()
    dev = _device.current_device() if device is None else int(device)
    if not os.path.exists(fname):
        raise FileNotFoundError(f"{fname} not found.")
    lib_ = _native.lib()
    handle, rows, cols = C.c_void_p(), C.c_int64(), C.c_int64()
    status = lib_.pcq_text_open(dev, os.fsencode(fname), C.byref(handle), C.byref(rows), C.byref(cols))
    if status in (-1, -5):          # PCQ_ERR_ARG (ragged / unreadable), PCQ_ERR_UNSUPPORTED (not a plain table)
        raise ValueError((lib_.pcq_last_error() or b"").decode(errors="replace"))
    _native.check(status, "pcq_text_open")
    try:
        out = torch.empty((rows.value, cols.value), dtype=torch.float64, device=torch.device("cuda", dev))
        with torch.cuda.device(out.device):
            status = lib_.pcq_text_parse_dev(handle, out.data_ptr() if out.numel() else None, max(cols.value, 1),
                                             torch.cuda.current_stream().cuda_stream)
        if status in (-1, -5):
            raise ValueError((lib_.pcq_last_error() or b"").decode(errors="replace"))
        _native.check(status, "pcq_text_parse_dev")
    finally:
        lib_.pcq_text_close(handle)
    return out


def save(fname, X, threads=0):
    """``.npy`` -> ``np.save`` (binary, exact); anything else -> :func:`savetxt`."""
    if str(fname).endswith(".npy"):
        np.save(fname, np.asarray(X))
    else:
        savetxt(fname, X, threads=threads)


def load(fname, ndmin=0, threads=0, mmap_mode=None):
    """``.npy`` -> ``np.load`` (optionally memory-mapped, so a 1e7-row sample file is paged in by the chunked
    host->device copies of the fit instead of being read twice); anything else -> :func:`loadtxt`."""
    if str(fname).endswith(".npy"):
        arr = np.load(fname, mmap_mode=mmap_mode)
        if ndmin == 2 and arr.ndim == 1:      # a column, as the text route returns it
            arr = arr.reshape(-1, 1)
        return np.atleast_1d(arr) if ndmin == 1 else arr
    return loadtxt(fname, ndmin=ndmin, threads=threads)
