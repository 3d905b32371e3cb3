from .lreg import lreg, lsq, PCBasisEvaluator  # noqa: F401
from .anl import anl  # noqa: F401
from .bcs import bcs, bcs_fit  # noqa: F401
