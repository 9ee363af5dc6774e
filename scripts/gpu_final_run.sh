#!/bin/bash
# (GPU box, one GPU) the evidence of one state of the tree in one call: whole GPU suite, default bench line, ncu launch
# list + one `--set full` capture of the main kernels, compute-sanitizer memcheck / racecheck of the smoke pass, then
# the other BASELINE configs.   gpurun --timeout 1300 -- 'bash scripts/gpu_final_run.sh 2>&1 | tail -60'
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/final_gpu_tests.log
echo "== bench c2"; bash scripts/bench_configs.sh 1 "c2" 2>&1 | tail -3
echo "== ncu launches"; bash scripts/ncu_launches.sh 2>&1 | tail -24
echo "== ncu full"
timeout 500 ncu --set full --clock-control none --import-source on \
  -k regex:'k1_scan_tiles|k1_stitch|k1_select_events|k2_event_scan|k3_split|k4_segment_stats' -s 18 -c 6 -f \
  -o gpurun_out/prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_main.log 2>&1
tail -2 gpurun_out/ncu_main.log | cut -c1-200; ls -la gpurun_out/prof_main.ncu-rep
echo "== sanitizer memcheck"
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1
tail -6 gpurun_out/sanitizer_memcheck.log
echo "== sanitizer racecheck"
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1
tail -6 gpurun_out/sanitizer_racecheck.log
echo "== other configs"; bash scripts/bench_configs.sh 1 "c1 c3 c4 c5" 2>&1 | tail -8
