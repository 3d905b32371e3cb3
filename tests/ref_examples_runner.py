"""Runs ONE example script of the reference, unmodified, either on the pure reference (CPU) or with the modules of
the PC hot path replaced by the drop-in (GPU) -- helper of tests/test_reference_examples.py, run in a subprocess:

    python tests/ref_examples_runner.py reference|dropin <example.py> <workdir>

Everything that is not on the path (pytuq.utils.plotting / maps / funcbank, pytuq.func.*, ...) comes from the
unmodified reference in oracle/_ref in both modes; matplotlib is stubbed; numpy's global stream is seeded so that
both runs draw the same numbers."""
import importlib
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

# the modules of the hot path: these names resolve to pytuq_b200 in drop-in mode
PATH_MODULES = ("rv.pcrv", "utils.mindex", "lreg.lreg", "lreg.anl", "lreg.bcs", "workflows.fits", "workflows.uprop",
                "surrogates.pce", "gsa.gsa")


def shape_the_matplotlib_stub():
    """The examples unpack what pyplot returns (`fig, ax = plt.subplots()`, `line, = ax.plot(...)`) and feed colours to
    colorsys: give the MagicMock stub return values of the right shape."""
    from unittest.mock import MagicMock
    import matplotlib
    import matplotlib.pyplot as plt
    ax = MagicMock()
    ax.plot.return_value = [MagicMock()]
    ax.errorbar.return_value = (MagicMock(), MagicMock(), MagicMock())
    ax.get_legend_handles_labels.return_value = ([], [])

    def subplots(nrows=1, ncols=1, *a, **kw):
        if nrows * ncols == 1:
            return MagicMock(), ax
        if nrows == 1 or ncols == 1:
            return MagicMock(), tuple(ax for _ in range(nrows * ncols))
        return MagicMock(), tuple(tuple(ax for _ in range(ncols)) for _ in range(nrows))
    plt.subplots.side_effect = subplots
    plt.plot.return_value = [MagicMock()]
    plt.gca.return_value = ax
    plt.errorbar.return_value = (MagicMock(), MagicMock(), MagicMock())
    matplotlib.colors.to_rgb.return_value = (0.3, 0.4, 0.5)
    sys.modules["matplotlib.colors"].to_rgb.return_value = (0.3, 0.4, 0.5)


def main():
    mode, script, workdir = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
    ref_loader.stub_matplotlib()
    shape_the_matplotlib_stub()
    sys.path.insert(0, ref_loader.REF_DIR)
    if mode == "dropin":
        for name in PATH_MODULES:
            parent = "pytuq." + name.rsplit(".", 1)[0]
            pkg = importlib.import_module(parent)                        # the reference's (empty) package
            ours = importlib.import_module("pytuq_b200." + name)
            sys.modules["pytuq." + name] = ours
            setattr(pkg, name.rsplit(".", 1)[1], ours)
    import numpy as np
    np.random.seed(20240229)
    os.chdir(workdir)
    sys.argv = [script]
    runpy.run_path(script, run_name="__main__")
    if mode == "dropin":
        from pytuq_b200 import _native
        print("PCQ_LAUNCHES", _native.lib().pcq_launch_count())


if __name__ == "__main__":
    main()
