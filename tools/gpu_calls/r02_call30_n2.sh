#!/bin/bash
# round 2, call 30 (2 GPUs): ragged 7-point planes on slabs
timeout -s KILL 200 python -m pytest tests/test_stencil_gpu.py -m gpu -q --timeout 90 -k "ragged or in_process" > gpurun_out/r02dd_tests_n2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02dd_tests_n2.log
