#!/bin/bash
# development (GPU box): parity tests of the split search and the API layer, then bench.py with both split kernels.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_dev_run.sh 2>&1 | tail -60'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/dev_tests.log
for k in flow level; do
  echo "== split kernel: $k" | tee -a gpurun_out/dev_bench.txt
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --split-kernel $k 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); p = d.get('parity') or {}
        print({k: round(v, 3) for k, v in d['stage_ms'].items()}, 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'parity', p.get('events_bit_exact'), p.get('segments_bit_exact'), d['counts'])
    elif 'rror' in l: print(l.rstrip())
" | tee -a gpurun_out/dev_bench.txt
done
