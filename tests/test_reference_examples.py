"""The reference's own EXAMPLE scripts, unmodified, as an end-to-end acceptance test of the drop-in.

oracle/Makefile copies the examples that drive the PC path (examples/ex_pcrv*.py, ex_lreg*.py, ex_bcs*.py,
ex_evidence.py, ex_uprop*.py, ex_pcgsa.py, ex_quad.py, surrogates/ex_pce.py, surrogates/ex_genz_bcs.py ...) to
oracle/_ref/examples.  Each one is run twice in a subprocess (tests/ref_examples_runner.py), numpy's stream seeded
identically: once on the pure reference (CPU), once with the path's modules (rv.pcrv, utils.mindex, lreg.*,
workflows.*, surrogates.pce, gsa.gsa) answered by pytuq_b200 (GPU) while everything else -- test functions, maps,
plotting (matplotlib stubbed) -- still comes from the reference.  The numbers the two runs print must agree."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXDIR = os.path.join(ROOT, "oracle", "_ref", "examples")
RUNNER = os.path.join(ROOT, "tests", "ref_examples_runner.py")
EXAMPLES = ["ex_pcrv.py", "ex_pcrv1.py", "ex_pcrv2.py", "ex_pcrv_mvn.py", "ex_lreg.py", "ex_lreg_basiseval.py", "ex_bcs.py",
            "ex_evidence.py", "ex_uprop.py", "ex_uprop2.py", "ex_pcgsa.py", "ex_quad.py", "ex_pcbasis1d.py",
            "ex_bcs_mindex_growth.py", "surrogates/ex_pce.py", "surrogates/ex_genz_bcs.py"]
NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?|[-+]?nan|[-+]?inf")
# relative agreement demanded of the printed numbers (most agree to all printed digits; a cross-validated sparse fit
# or a 10^5-sample Monte Carlo estimate is only as reproducible as its weakest step)
RTOL = {"default": 1e-6}


def _run(mode, script, workdir):
    env = dict(os.environ, PYTHONHASHSEED="0", PCQ_DEVICES="1")
    r = subprocess.run([sys.executable, RUNNER, mode, script, workdir], capture_output=True, text=True, timeout=900, env=env)
    return r.returncode, r.stdout, r.stderr


def _numbers(text):
    lines = [ln for ln in text.splitlines() if not ln.startswith("PCQ_LAUNCHES") and "time" not in ln.lower() and "sec" not in ln.lower()]
    return np.array([float(t) for t in NUM.findall("\n".join(lines))]), lines


@pytest.mark.skipif(not os.path.isdir(EXDIR), reason="oracle/_ref/examples is absent (made by `make -C oracle` where /root/reference exists)")
@pytest.mark.parametrize("example", EXAMPLES)
def test_reference_example_prints_the_same_numbers_on_the_drop_in(example, tmp_path):
    script = os.path.join(EXDIR, example)
    wr, wd = tmp_path / "ref", tmp_path / "dropin"
    wr.mkdir(); wd.mkdir()
    rc_r, out_r, err_r = _run("reference", script, str(wr))
    rc_d, out_d, err_d = _run("dropin", script, str(wd))
    # (a script may end in a plotting call the matplotlib stub cannot serve: both runs then stop at the same place)
    assert rc_d == rc_r, f"reference rc={rc_r}, drop-in rc={rc_d}\n{err_d[-2000:]}"
    launches = [ln for ln in out_d.splitlines() if ln.startswith("PCQ_LAUNCHES")]
    if rc_d == 0:
        assert launches, "the drop-in run did not report its kernel launches"
        if example not in ("ex_pcbasis1d.py", "ex_quad.py"):       # (quadrature rules are host-only work)
            assert int(launches[0].split()[1]) > 0, "no CUDA kernel was launched by the drop-in run"
    a, la = _numbers(out_r)
    b, lb = _numbers(out_d)
    assert len(la) == len(lb), f"different number of printed lines: {len(la)} vs {len(lb)}"
    assert a.shape == b.shape, f"different number of printed values: {a.shape} vs {b.shape}"
    if a.size:
        scale = max(1.0, float(np.nanmax(np.abs(a[np.isfinite(a)]))) if np.isfinite(a).any() else 1.0)
        rtol = RTOL.get(example, RTOL["default"])
        bad = ~np.isclose(b, a, rtol=rtol, atol=1e-9 * scale, equal_nan=True)
        assert not bad.any(), (f"{int(bad.sum())} of {a.size} printed values differ; first: reference {a[bad][:5]} "
                               f"drop-in {b[bad][:5]}")
