#!/bin/bash
# (GPU box) every BASELINE config through bench.py on N GPUs; lines kept under gpurun_out/bench_<config>_<N>gpu.json
#   gpurun --timeout 1800 -- 'bash scripts/bench_configs.sh 1 "c1 c2 c3 c4 c5" 2>&1 | tail -30'
set -u
N=${1:-1}; CONFIGS=${2:-"c1 c2 c3 c4 c5"}
mkdir -p gpurun_out
for c in $CONFIGS; do
  out=gpurun_out/bench_${c}_${N}gpu.json
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --config $c ${BENCH_ARGS:-} > $out 2> gpurun_out/bench_${c}_${N}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $N --config $c ${BENCH_ARGS:-} > $out 2> gpurun_out/bench_${c}_${N}gpu.err
  fi
  python - $out $c <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    r = d.get('roofline', {})
    print(sys.argv[2], 'N', d['n_gpus'], 'value %.0f' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'],
          'roofline %s %.3f' % (r.get('bound'), r.get('frac', 0)), 'cpu', round((d.get('cpu_baseline') or {}).get('value', 0), 2),
          {k: round(v, 3) for k, v in d.get('stage_ms', {}).items()}, d.get('parity'))
except Exception as e:
    print(sys.argv[2], 'FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
done
