from .gsa import PCSobol, SensMethod  # noqa: F401
