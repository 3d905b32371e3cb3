#!/usr/bin/env python
"""FP64 roofline denominators with a clock record (the companion MEASURED_PEAKS.json lacks for FP64).

Runs pcq_measure_peak (csrc/pcq_peaks.cu: register-resident DFMA chain, DMMA m8n8k4 chain, 2 GiB copy) back to back
for a few seconds per kind -- burst = best single measurement, sustained = median over the loop -- while nvidia-smi
samples SM clock, power and throttle reasons.  One JSON object on stdout (committed as profiles/r02_fp64_peaks.json)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]] + sys.argv[1:]
from bench import ClockSampler  # noqa: E402
from pytuq_b200 import _native  # noqa: E402

lib = _native.lib()
out = {"how": "pcq_measure_peak (csrc/pcq_peaks.cu) looped for ~4 s per kind; burst = best, sustained = median; "
              "nvidia-smi -lms 100 during each loop", "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
try:
    import subprocess
    out["gpu_name"] = subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader", "-i", "0"],
                                     capture_output=True, text=True).stdout.strip()
except OSError:
    pass
for kind, name, unit in ((1, "fp64_dmma", "TFLOP/s"), (0, "fp64_dfma", "TFLOP/s"), (2, "hbm_copy", "GB/s")):
    sampler = ClockSampler(0)
    sampler.start()
    vals, t0 = [], time.time()
    while time.time() - t0 < 4.0:
        v = ctypes.c_double()
        _native.check(lib.pcq_measure_peak(0, kind, ctypes.byref(v)))
        vals.append(v.value)
    clocks = sampler.stop()
    out[name] = {"unit": unit, "burst": max(vals), "sustained_median": float(np.median(vals)), "min": min(vals),
                 "measurements": len(vals), "clocks": clocks}
print(json.dumps(out))
