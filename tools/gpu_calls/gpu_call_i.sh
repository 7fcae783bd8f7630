#!/bin/bash
# round-1 call i: fused two-apply 7-point kernel -- parity, sweep, ncu
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_stencil_gpu.py -m gpu -q -k "fused or fuse" > gpurun_out/t14_fused.log 2>&1
echo "fused tests rc=$?"; tail -15 gpurun_out/t14_fused.log
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/t14.log 2>&1
echo "all gpu tests rc=$?"; tail -8 gpurun_out/t14.log
timeout -s KILL 200 python tools/sweep_lapfused.py 512 > gpurun_out/lapf512_v2.txt 2>&1; tail -30 gpurun_out/lapf512_v2.txt
timeout -s KILL 300 python tools/sweep_lapfused.py 1024 > gpurun_out/lapf1024_v2.txt 2>&1; tail -30 gpurun_out/lapf1024_v2.txt
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:lap7_fused2 --launch-skip 1 --launch-count 1 \
   -f -o gpurun_out/prof_lapf_r01j python tools/prof_lapfused.py 1024 > gpurun_out/ncu_lapf_v2.log 2>&1; tail -3 gpurun_out/ncu_lapf_v2.log
