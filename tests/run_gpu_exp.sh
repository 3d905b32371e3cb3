#!/bin/bash
mkdir -p gpurun_out
for lib in $1; do
  echo "== experiment $lib"
  PCQ_GRAM_WS=0 PCQ_LIB=$PWD/tools/exp/$lib timeout 300 python tools/tune.py k3only 2>&1 | grep "C2\|C1 " | tee -a gpurun_out/tune_exp.log
done
