#!/bin/bash
# round 2, call 8 (1 GPU): split layout + single TMA wait for non-first k-tiles; four-rows-per-thread tiles
out=gpurun_out; mkdir -p $out
FDB_FUSED_IMPL=4 timeout -s KILL 900 python -m pytest tests/test_upwind_gpu.py tests/test_random_gpu.py tests/test_persistent_gpu.py -m gpu -q -x > $out/r02i_tests.log 2>&1; echo "tests rc=$?"; tail -3 $out/r02i_tests.log
{ FDB_FUSED_IMPL=4 SWEEP_FUSED=3:0,3:8,3:10,3:11,3:12,3:6,3:1,4:0,2:0 SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_fused.py 512
  FDB_FUSED_IMPL=4 SWEEP_FUSED=3:0,3:8,3:10,3:12 SWEEP_CIS=0 timeout -s KILL 300 python tools/sweep_fused.py 1024; } 2>&1 | tee $out/r02i_sweep.txt
