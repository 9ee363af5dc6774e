#!/bin/bash
# development (GPU box): the K3_CFG_HALFKEY variant of the split search ("planned next";
# csrc/split.cuh) -- built next to the product library, run through the split / pipeline parity tests, then timed
# against the product build.  Written on a box without a GPU: the variant has NOT run yet.
#   gpurun --timeout 600 -- 'bash scripts/halfkey_variant.sh 2>&1 | tail -40'
set -u
mkdir -p build gpurun_out
python -m pypore_b200.build -DK3_CFG_HALFKEY=1 --out=build/lib_halfkey.so || exit 1
echo "== parity (variant)"
PYPORE_B200_LIB=$PWD/build/lib_halfkey.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -5
for lib in pypore_b200/libpypore_b200.so build/lib_halfkey.so; do
  echo "== $lib"
  PYPORE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: round(v, 3) for k, v in d['stage_ms'].items()}, 'ms/step', round(d['ms_per_step'], 3))
" | tee -a gpurun_out/halfkey_variant.txt
done
