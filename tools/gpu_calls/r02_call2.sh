#!/bin/bash
# round 2, call 2 (1 GPU): new persistent-kernel tests, the full default bench line (parity + also + e2e_process),
# an ncu --set full capture of the shipped fused upwind kernel (baseline for the kernel work), host topology
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_persistent_gpu.py -m gpu -q -x > $out/r02b_persistent_tests.log 2>&1; echo "persistent tests rc=$?"; tail -5 $out/r02b_persistent_tests.log
timeout -s KILL 900 python bench.py > $out/r02b_bench_default.json 2> $out/r02b_bench_default.err; echo "bench rc=$?"; tail -c 1500 $out/r02b_bench_default.err; head -c 6000 $out/r02b_bench_default.json
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:upwind3d_fused -s 4 -c 1 -o $out/r02b_prof_fused -f python tools/prof_upwind.py 512 > $out/r02b_ncu.log 2>&1; tail -3 $out/r02b_ncu.log
{ nvidia-smi topo -m; lscpu | head -30; numactl -H 2>/dev/null; free -g; } > $out/r02b_host_topology.txt 2>&1
