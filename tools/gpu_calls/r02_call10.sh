#!/bin/bash
# round 2, call 10 (1 GPU): lean + split formulation of the fused 7-point kernel -- parity, then A/B
out=gpurun_out; mkdir -p $out
FDB_LAPF_IMPL=2 timeout -s KILL 900 python -m pytest tests/test_stencil_gpu.py tests/test_random_gpu.py tests/test_persistent_gpu.py -m gpu -q -x > $out/r02k_lapf_tests.log 2>&1; echo "tests rc=$?"; tail -3 $out/r02k_lapf_tests.log
for impl in 1 2; do
  echo "== FDB_LAPF_IMPL=$impl"
  FDB_LAPF_IMPL=$impl SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_lapfused.py 1024 7 0 2 3
  FDB_LAPF_IMPL=$impl SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_lapfused.py 512 7 0 3
done 2>&1 | tee $out/r02k_lapf_ab.txt
FDB_LAPF_IMPL=2 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:lap7_fused2 -s 2 -c 1 -o $out/r02k_prof_lapf_lean -f python tools/prof_lapfused.py 1024 > $out/r02k_ncu.log 2>&1; tail -2 $out/r02k_ncu.log
