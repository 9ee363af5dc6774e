#!/bin/bash
# development (GPU box): the whole GPU suite, then bench.py.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_full_run.sh 2>&1 | tail -60'
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/full_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_line.json').read().strip().splitlines()[-1])
    print({k: round(v, 3) for k, v in d['stage_ms'].items()}, 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), d.get('parity'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_err.log').read()[-2000:])
PY
