"""pytuq_b200 -- B200-native (sm_100a) polynomial-chaos hot path behind PyTUQ's API.

Same module layout as the reference for the path it replaces::

    pytuq.rv.pcrv          -> pytuq_b200.rv.pcrv          (PCRV, PC1d, PCRV_iid, PCRV_mvn)
    pytuq.utils.mindex     -> pytuq_b200.utils.mindex     (get_mi, get_npc, ...)
    pytuq.lreg.{lreg,anl,bcs} -> pytuq_b200.lreg.{lreg,anl,bcs}
    pytuq.workflows.{fits,uprop} -> pytuq_b200.workflows.{fits,uprop}
    pytuq.surrogates.pce   -> pytuq_b200.surrogates.pce   (PCE)
    pytuq.gsa.gsa.PCSobol  -> pytuq_b200.gsa.gsa.PCSobol

All arithmetic of the path runs in hand-written CUDA kernels reached through the C-ABI
library ``libpcq.so`` (include/pcq.h) with ctypes.  There is no CPU fallback: importing
this package works without a GPU (so host-side logic can be tested), but the first call
that needs arithmetic raises if the library or a device is missing.
"""
from ._native import lib, PcqError, have_library   # noqa: F401

__version__ = "0.1.0"
