#!/bin/bash
# round 2, call 5 (1 GPU): lean consumer formulation of the fused upwind kernel -- parity, then A/B against the first one
out=gpurun_out; mkdir -p $out
FDB_FUSED_IMPL=2 timeout -s KILL 900 python -m pytest tests/test_upwind_gpu.py tests/test_random_gpu.py tests/test_persistent_gpu.py -m gpu -q -x > $out/r02f_lean_tests.log 2>&1; echo "lean tests rc=$?"; tail -5 $out/r02f_lean_tests.log
for impl in 1 2; do
  echo "== FDB_FUSED_IMPL=$impl"
  FDB_FUSED_IMPL=$impl SWEEP_FUSED=3:0,3:8,2:0,4:0 SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_fused.py 512
  FDB_FUSED_IMPL=$impl SWEEP_FUSED=3:0,3:8 SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_fused.py 1024
done 2>&1 | tee $out/r02f_lean_ab.txt
FDB_FUSED_IMPL=2 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:upwind3d_fused -s 4 -c 1 -o $out/r02f_prof_lean -f python tools/prof_upwind.py 512 > $out/r02f_ncu.log 2>&1; tail -2 $out/r02f_ncu.log
