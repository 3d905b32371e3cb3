from .pce import PCE  # noqa: F401
