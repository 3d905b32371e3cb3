"""Loader of the unmodified reference copied to oracle/_ref (oracle/Makefile) -- TEST INFRASTRUCTURE ONLY.

PyTUQ imports matplotlib at module import time (func/func.py:7, utils/plotting.py:13-16, lreg/lreg.py:6) and the
image has none; nothing on the arithmetic path touches it, so it is stubbed in sys.modules before the import.
Only tests/, __graft_entry__ and bench.py's CPU legs (cpu_baseline, --impl reference) may import this module.
"""
import os
import sys
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.lines", "matplotlib.patches",
          "matplotlib.ticker", "matplotlib.cm", "matplotlib.gridspec", "matplotlib.collections", "matplotlib.tri",
          "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "pytuq", "rv", "pcrv.py"))


def stub_matplotlib():
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock())


def import_reference():
    """Imports the reference's modules of the path from oracle/_ref and returns them by name."""
    if not available():
        raise ImportError(f"{REF_DIR}/pytuq is missing: run `make -C oracle` where /root/reference exists")
    stub_matplotlib()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from pytuq.rv.pcrv import PCRV, PC1d
    from pytuq.utils.mindex import get_mi, get_npc
    from pytuq.lreg.lreg import lsq
    from pytuq.lreg.anl import anl
    from pytuq.lreg.bcs import bcs, bcs_fit
    from pytuq.workflows.fits import pc_fit
    from pytuq.surrogates.pce import PCE
    return dict(PCRV=PCRV, PC1d=PC1d, get_mi=get_mi, get_npc=get_npc, lsq=lsq, anl=anl, bcs=bcs, bcs_fit=bcs_fit,
                pc_fit=pc_fit, PCE=PCE)
