#!/bin/bash
# Suggested FIRST GPU call of round 2 (1 GPU, ~3 min): what round 1's GPU budget did not cover.
#   1. the whole parity suite, the opt-in experimental tile configurations included
#   2. compute-sanitizer (memcheck + racecheck) on every kernel family, the round-1 late kernels included
#   3. A/B sweeps of the tile configurations that were compiled but never timed
out=gpurun_out; mkdir -p $out
FDB_TEST_EXPERIMENTAL=1 timeout -s KILL 600 python -m pytest tests -m gpu -q > $out/r02_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 $out/r02_tests.log
for tool in memcheck racecheck; do
  timeout -s KILL 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > $out/r02_sanitizer_$tool.log 2>&1
  tail -3 $out/r02_sanitizer_$tool.log
done
# fused upwind: shipped default (3:0) against the two-rows-per-thread tiles (3:8, 3:9)
SWEEP_FUSED=3:0,3:8,3:9 SWEEP_CIS=0 timeout -s KILL 200 python tools/sweep_fused.py 512 > $out/r02_fused_512.txt 2>&1; cat $out/r02_fused_512.txt
SWEEP_FUSED=3:0,3:8,3:9 SWEEP_CIS=0 timeout -s KILL 200 python tools/sweep_fused.py 1024 > $out/r02_fused_1024.txt 2>&1; cat $out/r02_fused_1024.txt
# fused 7-point: shipped default (7) against the shuffled-neighbour variants (12, 13)
SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_lapfused.py 1024 7 12 13 > $out/r02_lapf_1024.txt 2>&1; cat $out/r02_lapf_1024.txt
SWEEP_CIS=0 timeout -s KILL 200 python tools/sweep_lapfused.py 512 7 12 13 > $out/r02_lapf_512.txt 2>&1; cat $out/r02_lapf_512.txt
