#!/bin/bash
# development (GPU box): one `ncu --set full` capture of the kernels matching $1 inside bench.py (launch $2 of them, default 3).
#   gpurun --timeout 900 -- 'bash scripts/ncu_kernel.sh k1_scan_tiles 2>&1 | tail -5'
set -u
mkdir -p gpurun_out
pat=$1; skip=${2:-3}; tag=${3:-$1}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o gpurun_out/prof_$tag \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-300
ls -la gpurun_out/prof_$tag.ncu-rep
