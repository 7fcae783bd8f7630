#!/bin/bash
# round 2, call 7 (1 GPU): lean + split exchange-tile layout -- parity, then A/B against lean
out=gpurun_out; mkdir -p $out
FDB_FUSED_IMPL=4 timeout -s KILL 900 python -m pytest tests/test_upwind_gpu.py tests/test_random_gpu.py tests/test_persistent_gpu.py -m gpu -q -x > $out/r02h_split_tests.log 2>&1; echo "split tests rc=$?"; tail -5 $out/r02h_split_tests.log
for impl in 2 4; do
  echo "== FDB_FUSED_IMPL=$impl"
  FDB_FUSED_IMPL=$impl SWEEP_FUSED=3:0,3:8,2:0,4:0 SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_fused.py 512
  FDB_FUSED_IMPL=$impl SWEEP_FUSED=3:0,3:8 SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_fused.py 1024
done 2>&1 | tee $out/r02h_split_ab.txt
FDB_FUSED_IMPL=4 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:upwind3d_fused -s 4 -c 1 -o $out/r02h_prof_split -f python tools/prof_upwind.py 512 > $out/r02h_ncu.log 2>&1; tail -2 $out/r02h_ncu.log
