#!/bin/bash
# round-1 call k: fused 7-point kernel -- default config, L2 promotion A/B, DRAM bytes, bench lines
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_stencil_gpu.py -m gpu -q -k "fuse" > gpurun_out/t15_fused.log 2>&1
echo "fused tests rc=$?"; tail -4 gpurun_out/t15_fused.log
SWEEP_PROMOS=256,128,0 SWEEP_CIS=0,128 timeout -s KILL 300 python tools/sweep_lapfused.py 1024 7 8 2 0 > gpurun_out/lapf1024_promo.txt 2>&1; cat gpurun_out/lapf1024_promo.txt
SWEEP_PROMOS=256,0 SWEEP_CIS=0 timeout -s KILL 200 python tools/sweep_lapfused.py 512 7 2 8 > gpurun_out/lapf512_promo.txt 2>&1; cat gpurun_out/lapf512_promo.txt
for promo in 256 0; do
  FDB_TMA_L2PROMO=$promo timeout -s KILL 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
     -k regex:lap7_fused2 --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/dram_lapf_promo$promo.csv python tools/prof_lapfused.py 1024 > /dev/null 2>&1
  grep -E "dram__|gpu__time" gpurun_out/dram_lapf_promo$promo.csv | awk -F'","' '{print "promo='$promo'", $(NF-2), $(NF-1), $NF}'
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:lap7_fused2 --launch-skip 1 --launch-count 1 \
   -f -o gpurun_out/prof_lapf_r01k python tools/prof_lapfused.py 1024 > gpurun_out/ncu_lapf_k.log 2>&1; tail -2 gpurun_out/ncu_lapf_k.log
timeout -s KILL 400 python bench.py --workload lap1024 --steps 5 --warmup 3 > gpurun_out/bench_lap1024.log 2>&1; tail -1 gpurun_out/bench_lap1024.log | cut -c1-1500
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench15.log 2>&1; tail -1 gpurun_out/bench15.log | cut -c1-400
