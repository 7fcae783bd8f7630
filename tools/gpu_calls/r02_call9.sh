#!/bin/bash
# round 2, call 9 (1 GPU): early stage wait (impl 5) vs split (impl 4); pruned tile tables incl. the 17-warp tile
out=gpurun_out; mkdir -p $out
FDB_FUSED_IMPL=5 timeout -s KILL 900 python -m pytest tests/test_upwind_gpu.py tests/test_random_gpu.py tests/test_persistent_gpu.py -m gpu -q -x > $out/r02j_tests.log 2>&1; echo "tests rc=$?"; tail -3 $out/r02j_tests.log
for impl in 4 5; do
  echo "== FDB_FUSED_IMPL=$impl"
  FDB_FUSED_IMPL=$impl SWEEP_FUSED=3:0,3:1,3:2,3:3,4:0,4:1,2:0,2:1 SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_fused.py 512
  FDB_FUSED_IMPL=$impl SWEEP_FUSED=3:0,3:1,3:2,4:0 SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_fused.py 1024
done 2>&1 | tee $out/r02j_sweep.txt
