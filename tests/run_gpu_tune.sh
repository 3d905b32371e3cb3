#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python tools/tune.py $1 > gpurun_out/tune.log 2>&1; cat gpurun_out/tune.log
for lib in $2; do
  echo "== experiment $lib"
  PCQ_LIB=$PWD/tools/exp/$lib timeout 300 python tools/tune.py k3 2>&1 | grep "C2\|C1 " | tee -a gpurun_out/tune_exp.log
done
