#!/bin/bash
# Scaling series on one box: bash tools/scale_run.sh <outdir> [workload] [extra bench args]
out=${1:-gpurun_out}; wl=${2:-upwind512}; shift 2
mkdir -p $out
port=29800
for n in 1 2 4 8; do
  port=$((port+1))
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $wl --no-cpu-baseline "$@" 2>&1 | grep '^{' > $out/scale_${wl}_n$n.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 5 --warmup 3 --workload $wl "$@" 2>&1 | grep '^{' > $out/scale_${wl}_n$n.json
  fi
  python - <<PY
import json
try:
    j=json.load(open("$out/scale_${wl}_n$n.json")); print("$wl", "n=$n", "GCUPS=%.1f"%j["value"], "e2e=%s"%(j["e2e"] and round(j["e2e"]["value"],1)), j["config"]["workload"], j["clocks"])
except Exception as e: print("$wl n=$n FAILED", e)
PY
done
