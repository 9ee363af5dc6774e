#!/bin/bash
# development (GPU box): a few named test files, then bench.py stage times.   gpu_quick.sh "tests/a.py tests/b.py" [bench args]
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest $1 -m gpu -x -q 2>&1 | tail -8
shift
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); p = d.get('parity') or {}
        print({k: round(v, 3) for k, v in d['stage_ms'].items()}, 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'parity', p.get('events_bit_exact'), p.get('segments_bit_exact'))
    elif 'rror' in l: print(l.rstrip())
"
