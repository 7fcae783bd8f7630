#!/bin/bash
# round 2, call 3 (1 GPU): whole GPU suite after the ADVICE fixes / drivers / graph capture; CUDA-graph A/B on config 1
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > $out/r02c_tests.log 2>&1; echo "gpu tests rc=$?"; tail -8 $out/r02c_tests.log
for g in 0 1; do
  FDB_GRAPH=$g timeout -s KILL 300 python bench.py --workload upwind128 --tsteps 10 --steps 200 --warmup 20 --no-cpu-baseline --no-also --no-parity 2>&1 | grep '^{' > $out/r02c_upwind128_graph$g.json
  python - <<PY
import json
j=json.load(open("$out/r02c_upwind128_graph$g.json")); print("FDB_GRAPH=$g upwind128 x10: GCUPS=%.1f ms/advect=%.4f e2e=%.1f launches=%d"%(j["value"],j["ms_per_step"],j["e2e"]["value"],j["gpu_launches"]))
PY
done
