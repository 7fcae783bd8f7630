#!/bin/bash
# round 2, call 20 (4 GPUs): reversed ring (negative slab-axis velocity) on more than two slabs, in one process and
# one process per GPU; the other multi-GPU tests of the final code
out=gpurun_out; mkdir -p $out
timeout -s KILL 400 python -m pytest tests/test_upwind_gpu.py tests/test_stencil_gpu.py tests/test_dist_gpu.py tests/test_persistent_gpu.py -m gpu -q -x --timeout 90 -k "slab or in_process or halo or nccl" > $out/r02u_tests_n4.log 2>&1; echo "rc=$?"; tail -5 $out/r02u_tests_n4.log
