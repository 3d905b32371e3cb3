#!/bin/bash
# usage: run_gpu_ncu.sh <k1|k2|k3> <n> [regex]
set -x
mkdir -p gpurun_out
what=$1; n=${2:-200000}
case $what in k1) PAT=pcq_bases;; k2) PAT=pcq_evalpc;; k3) PAT="pcq_syrk_kernel|pcq_pack";; esac
PAT=${3:-$PAT}
CMD="python tests/prof_driver.py $what $n"
timeout 300 $CMD > gpurun_out/plain_$what.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$PAT -s 2 -c 1 -f -o gpurun_out/prof_$what $CMD > gpurun_out/ncu_$what.log 2>&1
tail -n 3 gpurun_out/ncu_$what.log
