#!/bin/bash
# development (GPU box with N GPUs): multi-GPU parity tests, then bench.py under torchrun.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_dist_run.sh 2 2>&1 | tail -30'
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/dist_tests_${N}gpu.log 2>&1; tail -6 gpurun_out/dist_tests_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open('gpurun_out/bench_%sgpu.json' % n) if l.startswith('{')][-1])
    print('N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']),
          {k: round(v, 3) for k, v in d['stage_ms'].items()}, d.get('parity'), 'fallbacks', d.get('host_planned_fallback_steps'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_%sgpu.err' % n).read()[-3000:])
PY
