#!/bin/bash
# GPU session: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-extras --samples-per-gpu 250000 --cpu-samples 500"
  timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
  timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pcq_gram_kernel -s 3 -c 1 -o gpurun_out/prof_gram $CMD > gpurun_out/ncu2.log 2>&1
  tail -5 gpurun_out/ncu1.log gpurun_out/ncu2.log
fi
