#!/bin/bash
# GPU session: parity tests, smoke, bench, optional ncu passes ($1 = list of: launches k1 k2 k3)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for what in $1; do
  case $what in
    ref)
      timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
    launches)
      CMD="python bench.py --steps 2 --warmup 3 --no-extras --samples-per-gpu 250000 --cpu-samples 500"
      timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1 ;;
    k1|k2|k3)
      case $what in k1) PAT=pcq_bases_kernel;; k2) PAT=pcq_evalpc_kernel;; k3) PAT="pcq_syrk_kernel|pcq_pack";; esac
      CMD="python tests/prof_driver.py $what 200000"
      timeout 300 $CMD > gpurun_out/plain_$what.log 2>&1 &&
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$PAT -s 2 -c 2 -f -o gpurun_out/prof_$what $CMD > gpurun_out/ncu_$what.log 2>&1
      tail -n 3 gpurun_out/ncu_$what.log ;;
  esac
done
