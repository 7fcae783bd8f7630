#!/bin/bash
# round 2, call 28 (2 GPUs): multi-GPU tests and the N=2 bench line with the T = 4 default
out=gpurun_out; mkdir -p $out
timeout -s KILL 500 python -m pytest tests/test_upwind_gpu.py tests/test_stencil_gpu.py tests/test_dist_gpu.py tests/test_persistent_gpu.py -m gpu -q --timeout 120 -k "slab or in_process or halo or nccl" > $out/r02bb_tests_n2.log 2>&1; echo "rc=$?"; tail -3 $out/r02bb_tests_n2.log
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29861 bench.py --gpus 2 --steps 8 --warmup 3 > $out/r02bb_bench_n2.json 2> $out/r02bb_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r02bb_bench_n2.json") if l.startswith("{")][-1]); p=j["parity"]
print("N=2 GCUPS=%.1f kernel=%s traffic=%s e2e=%.1f parity=%s/%s"%(j["value"],j["config"]["kernel"],j["roofline"]["traffic"],j["e2e"]["value"],p["random_bitexact"],p["corner_bitexact"]))
for k,v in (j.get("also") or {}).items(): print("   also",k,v.get("value"),(v.get("parity") or {}).get("ok"),v.get("error"))
PY
