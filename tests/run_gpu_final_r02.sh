#!/bin/bash
# r02 final-state session: full GPU suite, smoke, bench (+ reference arm), then the ncu launch list with DRAM traffic of
# one bench step ($1 = tag of the output files)
set -x
T=${1:-fin}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; tail -c 400 gpurun_out/${T}_ref.json
CMD="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pcq_ -c 600 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu.log 2>&1
tail -c 300 gpurun_out/${T}_ncu.log
