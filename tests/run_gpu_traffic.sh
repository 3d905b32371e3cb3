#!/bin/bash
# ncu passes on the bench-size workload: (1) DRAM traffic of every K3 launch of one step, (2) launch list of bench.py
set -x
mkdir -p gpurun_out
CMD="python tests/prof_driver.py k3 1250000"
timeout 300 $CMD > gpurun_out/plain_k3.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active --clock-control none -k "regex:pcq_syrk|pcq_pack" -s 28 -c 28 --csv --log-file gpurun_out/k3_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --no-extras --cpu-samples 500"
timeout 600 $CMD2 > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
