#!/bin/bash
# short 8-GPU check: dist parity at 8 ranks + weak/strong bench at N=4,8
out=gpurun_out; mkdir -p $out; port=29970
(timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu --timeout=500 --timeout-method=thread -k "8" 2>&1 | tail -5) > $out/t_dist8b.log
for spec in "8 upwind512" "4 upwind512" "8 upwind1024"; do
  set -- $spec; n=$1; wl=$2; port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus $n --steps 5 --warmup 3 --workload $wl $([ "$wl" = upwind1024 ] && echo --no-e2e) 2>&1 | grep '^{' > $out/scale2_${wl}_n$n.json
  python - <<PY
import json
try:
    j=json.load(open("$out/scale2_${wl}_n$n.json")); print("$wl n=$n GCUPS=%.1f e2e=%s %s"%(j["value"], j["e2e"] and round(j["e2e"]["value"],1), j["clocks"]))
except Exception as e: print("$wl n=$n FAILED", e)
PY
done
(timeout 200 ./drivers/bin/upwindCuda -numCells 1024 -numSteps 100 -ngpus 8 -timing 2>&1 | tail -3)
tail -2 $out/t_dist8b.log
