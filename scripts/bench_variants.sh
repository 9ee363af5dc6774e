#!/bin/bash
# development: stage times of bench.py for every library variant under build/ (EVENTS="5000 500" event counts)
for lib in build/lib_*.so; do
  for n in ${EVENTS:-5000}; do
  echo "== $lib events=$n"
  PYPORE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --events-per-gpu $n 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: round(v, 3) for k, v in d['stage_ms'].items()}, 'ms/step', round(d['ms_per_step'], 3))
"
  done
done
