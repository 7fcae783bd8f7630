#!/bin/bash
# round 2, call 12 (2 GPUs): single-launch ring sweeps with the ghost wait on the device and the cheaper PUSH variant
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_upwind_gpu.py tests/test_dist_gpu.py tests/test_persistent_gpu.py -m gpu -q -x > $out/r02m_tests_n2.log 2>&1; echo "gpu tests rc=$?"; tail -4 $out/r02m_tests_n2.log
for halo in single push2 single push2; do
  FDB_HALO=$halo timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 2 --steps 8 --warmup 3 --no-also --no-e2e > $out/r02m_bench_n2_$halo.json 2> $out/r02m_bench_n2_$halo.err; echo "bench $halo rc=$?"
  python - <<PY
import json
try:
    j=json.loads([l for l in open("$out/r02m_bench_n2_$halo.json") if l.startswith("{")][-1]); print("FDB_HALO=$halo N=2 GCUPS=%.1f avg_launch_ms=%.4f launches=%d parity=%s clocks=%s"%(j["value"],j["roofline"]["avg_launch_ms"],j["gpu_launches"],j["parity"]["random_bitexact"] and j["parity"]["corner_bitexact"], j["clocks"]["sm_mhz"]))
except Exception as e: print("FAILED", e); print(open("$out/r02m_bench_n2_$halo.err").read()[-1500:])
PY
done
