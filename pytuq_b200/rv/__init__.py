from .pcrv import PCRV, PC1d, PCRV_iid, PCRV_mvn  # noqa: F401
