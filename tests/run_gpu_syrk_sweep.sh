#!/bin/bash
# sweep of the pack+SYRK knobs on the C2 workload (prints ms/step of the resident path)
run() { echo -n "$1: "; env $1 timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), round(d['roofline']['achieved'],2))"; }
for mb in 128 512 1024 4096 8192; do run PCQ_SYRK_CHUNK_MB=$mb; done
for sp in 1 2 7 14; do run PCQ_SYRK_SPLIT=$sp; done
