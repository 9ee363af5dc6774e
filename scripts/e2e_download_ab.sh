#!/bin/bash
# development (N-GPU box): end-to-end step time with the result download staged / direct / absent
N=${1:-2}
for v in staged direct none; do
  unset PYPORE_B200_DIRECT_DOWNLOAD PYPORE_B200_BENCH_NO_DOWNLOAD
  [ $v = direct ] && export PYPORE_B200_DIRECT_DOWNLOAD=1
  [ $v = none ] && export PYPORE_B200_BENCH_NO_DOWNLOAD=1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - $v <<'PY'
import json, sys
try:
    d = json.loads([l for l in open('gpurun_out/ab_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'e2e ms/step %.3f' % d['e2e']['ms_per_step'], 'h2d only %.3f' % d['e2e']['h2d_only_ms_per_step'], (d.get('parity') or {}).get('segments_bit_exact'))
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open('gpurun_out/ab_%s.err' % sys.argv[1]).read()[-2000:])
PY
done
