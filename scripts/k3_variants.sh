#!/bin/bash
# development (GPU box): build k3_split variants (name:-Dflag:-Dflag ...) next to the product library and time each
# through bench.py (stage times + parity of the end-to-end tables).
#   gpurun --timeout 900 -- 'bash scripts/k3_variants.sh out.txt base: t64c14:-DK3_CFG_THREADS=64:-DK3_CFG_CTAS=14 2>&1 | tail -40'
set -u
mkdir -p build gpurun_out
out=gpurun_out/$1; shift
: > $out
for spec in "$@"; do
  name=${spec%%:*}; flags=$(echo "${spec#*:}" | tr ':' ' ')
  ( python -m pypore_b200.build $flags --out=build/lib_$name.so >/dev/null 2>build/$name.err || echo "build $name failed" ) &
done
wait
for spec in "$@"; do
  name=${spec%%:*}
  for kern in ${KERNELS:-flow}; do
  echo "== $spec [$kern]" | tee -a $out
  PYPORE_B200_LIB=$PWD/build/lib_$name.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --split-kernel $kern 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); p = d.get('parity') or {}
        print({k: round(v, 3) for k, v in d['stage_ms'].items()}, 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'parity', p.get('events_bit_exact'), p.get('segments_bit_exact'))
" | tee -a $out
  done
done
