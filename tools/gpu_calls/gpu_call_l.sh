#!/bin/bash
# round-1 call l: second layout of the fused upwind kernel -- parity, A/B sweep against the first
mkdir -p gpurun_out
FDB_FUSED_IMPL=2 timeout -s KILL 400 python -m pytest tests/test_upwind_gpu.py tests/test_random_gpu.py -m gpu -q -x > gpurun_out/t16_fused2.log 2>&1
echo "upwind tests with the second layout rc=$?"; tail -6 gpurun_out/t16_fused2.log
: > gpurun_out/fused2_512.txt
FDB_FUSED_IMPL=1 SWEEP_FUSED=3:0 SWEEP_CIS=0 timeout -s KILL 200 python tools/sweep_fused.py 512 >> gpurun_out/fused2_512.txt 2>&1
FDB_FUSED_IMPL=2 SWEEP_FUSED=3:0,3:1,3:2,3:3,3:4,3:5,3:6,3:7,2:1,2:3,4:0,4:1,4:2,4:3 SWEEP_CIS=0,64 timeout -s KILL 400 python tools/sweep_fused.py 512 >> gpurun_out/fused2_512.txt 2>&1
cat gpurun_out/fused2_512.txt
: > gpurun_out/fused2_1024.txt
FDB_FUSED_IMPL=1 SWEEP_FUSED=3:0 SWEEP_CIS=0 timeout -s KILL 200 python tools/sweep_fused.py 1024 >> gpurun_out/fused2_1024.txt 2>&1
FDB_FUSED_IMPL=2 SWEEP_FUSED=3:0,3:1,3:2,3:3,3:4,3:7,4:2 SWEEP_CIS=0,128 timeout -s KILL 400 python tools/sweep_fused.py 1024 >> gpurun_out/fused2_1024.txt 2>&1
cat gpurun_out/fused2_1024.txt
