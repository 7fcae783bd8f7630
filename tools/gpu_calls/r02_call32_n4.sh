#!/bin/bash
# round 2, call 32 (4 GPUs): T = 4 sweeps (the default) on four slabs, in one process and one process per GPU
timeout -s KILL 150 python -m pytest tests/test_upwind_gpu.py tests/test_dist_gpu.py -m gpu -q --timeout 80 -k "fused_in_process or nccl or negative_velocities_on_slab or transports" > gpurun_out/r02ff_tests_n4.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02ff_tests_n4.log
