#!/bin/bash
# first GPU contact: parity tests, peak micro-benchmarks
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python - <<'PY' > gpurun_out/peaks.log 2>&1
import ctypes as C, json
from pytuq_b200 import _native
lib = _native.lib()
res = {}
for kind, name in ((0, "fp64_dfma_tflops"), (1, "fp64_dmma_tflops"), (2, "hbm_copy_gbs")):
    vals = []
    for rep in range(3):
        v = C.c_double()
        _native.check(lib.pcq_measure_peak(0, kind, C.byref(v)))
        vals.append(v.value)
    res[name] = vals
print(json.dumps(res))
PY
cat gpurun_out/peaks.log
