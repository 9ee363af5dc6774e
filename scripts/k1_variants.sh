#!/bin/bash
# development (GPU box): threshold-scan variants -- tile size and residency -- timed
# through bench.py's stage times (the parity tests of the scan run against every variant first).
#   gpurun --timeout 600 -- 'bash scripts/k1_variants.sh 2>&1 | tail -40'
set -u
mkdir -p build gpurun_out
python -m pypore_b200.build -DK1_CFG_MINBLOCKS=8 --out=build/lib_k1_mb8.so || exit 1
python -m pypore_b200.build -DK1_CFG_ROWS=8 --out=build/lib_k1_rows8.so || exit 1
python -m pypore_b200.build -DK1_CFG_ROWS=2 -DK1_CFG_MINBLOCKS=8 --out=build/lib_k1_rows2_mb8.so || exit 1
for lib in pypore_b200/libpypore_b200.so build/lib_k1_mb8.so build/lib_k1_rows8.so build/lib_k1_rows2_mb8.so; do
  echo "== $lib"
  PYPORE_B200_LIB=$PWD/$lib timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "threshold or pipeline" 2>&1 | tail -1
  PYPORE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: round(v, 3) for k, v in d['stage_ms'].items()}, 'ms/step', round(d['ms_per_step'], 3))
" | tee -a gpurun_out/k1_variants.txt
done
