#!/bin/bash
# 2-GPU probe of the halo transports: bash tools/n2_probe.sh <out> [nproc]
out=${1:-gpurun_out/n2probe.txt}; np=${2:-2}; : > $out; port=29950
run() { port=$((port+1)); echo "== $*" >> $out; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 4 --warmup 3 --no-e2e 2>&1 | grep -E '^\{|rror' | python -c "
import sys,json
for l in sys.stdin:
    try:
        j=json.loads(l); print('GCUPS=%.1f ms/step=%.2f %s'%(j['value'], j['ms_per_step'], j['config']['workload']))
    except Exception: print(l.strip()[:300])" >> $out 2>&1; }
run FDB_HALO=direct WL=upwind512
run FDB_HALO=nccl WL=upwind512
run FDB_HALO=direct FDB_COMM_SMS=0
run FDB_HALO=nccl FDB_COMM_SMS=0
cat $out
