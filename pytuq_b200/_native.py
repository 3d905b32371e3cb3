"""ctypes binding of libpcq.so (include/pcq.h).  Fails loudly; there is no fallback path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCQ_LIB", os.path.join(_HERE, "libpcq.so"))

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class PcqError(RuntimeError):
    """Raised when a libpcq entry point returns a negative status (kept in ``status``; None when the library or
    the device itself is missing)."""

    def __init__(self, message, status=None):
        super().__init__(message)
        self.status = status


# name -> (restype, argtypes); must stay in step with include/pcq.h (tests check this)
SIGNATURES = {
    "pcq_version": (C.c_int, []),
    "pcq_last_error": (C.c_char_p, []),
    "pcq_device_count": (C.c_int, []),
    "pcq_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pcq_plan_destroy": (C.c_int, [C.c_void_p]),
    "pcq_plan_info": (C.c_int64, [C.c_void_p, C.c_int]),
    "pcq_plan_specialize": (C.c_int, [C.c_void_p, C.c_int]),
    "pcq_plan_specialize_async": (C.c_int, [C.c_void_p, C.c_int]),
    "pcq_debug_horner_source": (C.c_int64, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64]),
    "pcq_plan_enable_moments": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pcq_moment_info": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "pcq_host_copy_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "pcq_group_enable_moments": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pcq_debug_moments_host": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                         C.c_void_p]),
    "pcq_eval_bases_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                     C.c_int64, C.c_void_p]),
    "pcq_eval_bases": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]),
    "pcq_eval_pc_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "pcq_eval_pc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                              C.c_void_p, C.c_int64, C.c_int]),
    "pcq_suffstats_dim": (C.c_int64, [C.c_void_p, C.c_int]),
    "pcq_quadform_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_int64, C.c_void_p]),
    "pcq_quadform": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                               C.c_void_p, C.c_int64]),
    "pcq_matrix_forms_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "pcq_matrix_forms": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "pcq_residual_ss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_void_p]),
    "pcq_residual_ss_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pcq_suffstats_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                    C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "pcq_suffstats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                C.c_int, C.c_void_p]),
    "pcq_gram_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                               C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "pcq_gram": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                           C.c_int, C.c_void_p]),
    "pcq_propagate": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_int]),
    "pcq_propagate_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                    C.c_void_p]),
    "pcq_sample_germ_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p]),
    "pcq_text_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pcq_text_parse_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pcq_text_close": (C.c_int, [C.c_void_p]),
    "pcq_qr_max_rows": (C.c_int64, [C.c_int]),
    "pcq_qr_r_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "pcq_svd_jacobi_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                     C.POINTER(C.c_int), C.c_void_p]),
    "pcq_group_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "pcq_group_destroy": (C.c_int, [C.c_void_p]),
    "pcq_group_size": (C.c_int, [C.c_void_p]),
    "pcq_group_set_active": (C.c_int, [C.c_void_p, C.c_int]),
    "pcq_group_plan": (C.c_void_p, [C.c_void_p, C.c_int]),
    "pcq_group_device": (C.c_int, [C.c_void_p, C.c_int]),
    "pcq_group_specialize": (C.c_int, [C.c_void_p, C.c_int]),
    "pcq_group_suffstats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                      C.c_void_p, C.c_void_p]),
    "pcq_group_eval_bases": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]),
    "pcq_group_eval_pc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_int64, C.c_int]),
    "pcq_group_quadform": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_int64]),
    "pcq_group_residual_ss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "pcq_group_propagate": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_int]),
    "pcq_measure_peak": (C.c_int, [C.c_int, C.c_int, c_double_p]),
    "pcq_launch_count": (C.c_int64, []),
}

_lib = None


def have_library():
    return os.path.exists(LIB_PATH)


def lib():
    """The loaded library.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcqError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C pytuq_b200/csrc`. pytuq_b200 has no CPU fallback.")
        # one process per GPU (torchrun): the ranks of a node share its host cores for the pageable <-> pinned staging
        # copies; without a setting of the user's, each rank takes its share instead of the library's default of 16
        if "PCQ_HOST_THREADS" not in os.environ:
            try:
                local = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
            except ValueError:
                local = 1
            if local > 1:
                os.environ["PCQ_HOST_THREADS"] = str(max(2, (os.cpu_count() or 16) // local))
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what=""):
    if status < 0:
        msg = lib().pcq_last_error()
        raise PcqError(f"{what or 'libpcq'} failed with status {status}: {msg.decode() if msg else ''}", status)
    return status


def ptr(arr):
    """Host pointer of a C-contiguous numpy array (or None)."""
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)
