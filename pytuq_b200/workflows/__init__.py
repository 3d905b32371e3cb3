from .fits import pc_fit  # noqa: F401
from .uprop import uprop_proj, uprop_regr  # noqa: F401
