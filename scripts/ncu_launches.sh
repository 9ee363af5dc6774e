#!/bin/bash
# development (GPU box): per-kernel durations of two bench steps (ncu launch list, cold-cache and serialised).
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/launches.csv gpurun_out/launch_summary.csv | tail -30
