#!/bin/bash
# development (GPU box): one `ncu --set full` capture of the split kernel(s) inside bench.py (source counters included).
#   gpurun --timeout 900 -- 'bash scripts/ncu_flow.sh flow 2>&1 | tail -5'
set -u
mkdir -p gpurun_out
k=${1:-flow}
pat=$([ "$k" = flow ] && echo k3_flow || echo k3_split)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f -o gpurun_out/prof_$k \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --split-kernel $k > gpurun_out/ncu_$k.log 2>&1
tail -3 gpurun_out/ncu_$k.log
ls -la gpurun_out/prof_$k.ncu-rep
