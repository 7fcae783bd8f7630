#!/bin/bash
# round 2, call 11 (2 GPUs): single-launch ring sweeps -- multi-GPU parity tests, then bench N=2 in both forms
out=gpurun_out; mkdir -p $out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x > $out/r02l_tests_n2.log 2>&1; echo "gpu tests rc=$?"; tail -6 $out/r02l_tests_n2.log
for halo in single push2; do
  FDB_HALO=$halo timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 2 --steps 8 --warmup 3 --no-also --no-e2e > $out/r02l_bench_n2_$halo.json 2> $out/r02l_bench_n2_$halo.err; echo "bench $halo rc=$?"
  python - <<PY
import json
try:
    j=json.loads([l for l in open("$out/r02l_bench_n2_$halo.json") if l.startswith("{")][-1]); print("FDB_HALO=$halo N=2 GCUPS=%.1f avg_launch_ms=%.4f launches=%d parity=%s"%(j["value"],j["roofline"]["avg_launch_ms"],j["gpu_launches"],j["parity"]["random_bitexact"] and j["parity"]["corner_bitexact"]))
except Exception as e: print("FAILED", e); print(open("$out/r02l_bench_n2_$halo.err").read()[-1500:])
PY
done
timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-also --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('N=1 GCUPS=%.1f avg_launch_ms=%.4f'%(j['value'],j['roofline']['avg_launch_ms']))"
