#!/bin/bash
# development: stage times of bench.py for every library variant under build/
for lib in build/lib_*.so; do
  echo "== $lib"
  PYPORE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: round(v, 3) for k, v in d['stage_ms'].items()}, 'ms/step', round(d['ms_per_step'], 3))
"
done
