#!/bin/bash
# 8-GPU weak-scaling bench with the fused halo push (default) and with copy-engine transfers
mkdir -p gpurun_out; : > gpurun_out/n8_push_vs_copy.txt; port=29960
for h in direct copy; do
  port=$((port+1))
  FDB_HALO=$h timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus 8 --steps 6 --warmup 3 --no-e2e 2>&1 | grep '^{' | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('halo=$h', 'GCUPS=%.1f' % j['value'], 'ms/step=%.2f' % j['ms_per_step'], j['clocks'])" >> gpurun_out/n8_push_vs_copy.txt 2>&1
done
cat gpurun_out/n8_push_vs_copy.txt
