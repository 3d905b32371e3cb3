"""Host-side utilities of the PC path (multi-index tables, sample-file I/O)."""
from .mindex import get_mi, get_npc, encode_mindex, micf_join, mi_addfront, mi_addfront_cons  # noqa: F401
from . import io  # noqa: F401,E402
