#!/bin/bash
# round-1 call m (2 GPUs): slab ring with the new kernels -- dist parity, in-process slabs, 2-GPU bench lines
out=gpurun_out; mkdir -p $out
(timeout -s KILL 500 python -m pytest tests/test_dist_gpu.py tests/test_upwind_gpu.py tests/test_stencil_gpu.py -q -m gpu --timeout=400 --timeout-method=thread \
   -k "nccl or in_process or slabs" 2>&1 | tail -15) > $out/t17_n2.log; cat $out/t17_n2.log
port=29940
for wl in upwind512 lap1024; do
  port=$((port+1))
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl 2>&1 | grep '^{' > $out/bench17_${wl}_n2.json
  python - <<PY
import json
try:
    j=json.load(open("$out/bench17_${wl}_n2.json")); print("$wl n=2 GCUPS=%.1f e2e=%s kernel=%s halo=%s %s"%(j["value"], j["e2e"] and round(j["e2e"]["value"],1), j["config"]["kernel"], j.get("halo_bytes_per_gpu"), j["clocks"]))
except Exception as e: print("$wl n=2 FAILED", e)
PY
done
for h in direct copy; do
  port=$((port+1))
  FDB_HALO=$h timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
     tools/dist_heavy.py upwind 512 100 2>&1 | grep -E "upwind|Error|error" | tail -2
done
port=$((port+1))
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
     tools/dist_heavy.py lap 512 10 2>&1 | grep -E "laplacian|Error|error" | tail -2
