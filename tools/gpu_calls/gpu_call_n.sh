#!/bin/bash
# round-1 call n (final): whole GPU suite, smoke, bench lines, launch list + ncu of the shipped upwind kernel
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests -m gpu -q > $out/t18.log 2>&1; echo "gpu tests rc=$?"; tail -12 $out/t18.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke18.log 2>&1; tail -2 $out/smoke18.log
timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > $out/bench18.log 2>&1; tail -1 $out/bench18.log | cut -c1-700
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r01n.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $out/ncu_launch_n.log 2>&1; tail -1 $out/ncu_launch_n.log | cut -c1-200
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:upwind3d_fused --launch-skip 4 --launch-count 2 \
   -f -o $out/prof_fused_r01n python tools/prof_upwind.py 512 > $out/ncu_full_n.log 2>&1; tail -2 $out/ncu_full_n.log
SWEEP_FUSED=2:0,2:1,2:2,2:3,4:0,4:1,4:2,4:3 SWEEP_CIS=0 timeout -s KILL 200 python tools/sweep_fused.py 512 > $out/fused_t24_512.txt 2>&1; cat $out/fused_t24_512.txt
timeout -s KILL 200 python bench.py --workload upwind1024 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $out/bench18_1024.log 2>&1; tail -1 $out/bench18_1024.log | cut -c1-300
timeout -s KILL 100 python tools/prof_upwind.py 512 -1 1 1 2>&1 | tail -1
timeout -s KILL 100 python tools/prof_upwind.py 512 1 -1 -1 2>&1 | tail -1
FDB_NO_FLIP=1 timeout -s KILL 100 python tools/prof_upwind.py 512 -1 1 1 2>&1 | tail -1
