#!/bin/bash
# round-2 ncu passes (one gpurun call): (1) every launch of ONE bench step at the metric's configuration with its device
# time and DRAM bytes (launch list + K3 traffic), (2) --set full of the Horner evalPC kernels.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu"
timeout 600 $CMD > gpurun_out/r02_plain_bench.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:pcq_ -s 1020 -c 420 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log
CMD2="python tests/prof_driver.py k2s 2000000"
timeout 300 $CMD2 > gpurun_out/r02_plain_k2s.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pcq_k2 -s 6 -c 6 -o gpurun_out/r02_k2s $CMD2 > gpurun_out/r02_ncu_k2s.log 2>&1
tail -2 gpurun_out/r02_ncu_k2s.log
grep -o '"gpu_launches": [0-9]*' gpurun_out/r02_plain_bench.log
