#!/bin/bash
# round 2, call 13 (8 GPUs): the whole GPU suite with nothing skipped, the bench line at N=8 and N=4 (parity + also),
# single-launch vs two-launch ring sweeps, T=4 at the 8-GPU slab shape, host->device ceiling
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{ nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|Model name|NUMA|Socket"; free -g | head -2; } > $out/r02n_topo_n8.txt 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q > $out/r02n_tests_n8.log 2>&1; echo "gpu tests rc=$?"; tail -6 $out/r02n_tests_n8.log
timeout -s KILL 900 $TR --nproc-per-node 8 --master-port 29831 bench.py --gpus 8 --steps 8 --warmup 3 > $out/r02n_bench_n8.json 2> $out/r02n_bench_n8.err; echo "bench n8 rc=$?"; tail -c 600 $out/r02n_bench_n8.err
FDB_HALO=push2 timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29832 bench.py --gpus 8 --steps 8 --warmup 3 --no-also --no-e2e > $out/r02n_bench_n8_push2.json 2> $out/r02n_bench_n8_push2.err; echo "bench n8 push2 rc=$?"
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29833 bench.py --gpus 8 --steps 8 --warmup 3 --no-also --no-e2e --fuse 4 > $out/r02n_bench_n8_T4.json 2> $out/r02n_bench_n8_T4.err; echo "bench n8 T4 rc=$?"
timeout -s KILL 900 $TR --nproc-per-node 4 --master-port 29834 bench.py --gpus 4 --steps 8 --warmup 3 > $out/r02n_bench_n4.json 2> $out/r02n_bench_n4.err; echo "bench n4 rc=$?"
timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $out/r02n_bench_n1.json 2> $out/r02n_bench_n1.err; echo "bench n1 rc=$?"
timeout -s KILL 300 $TR --nproc-per-node 8 --master-port 29835 tools/h2d_ceiling.py > $out/r02n_h2d_ceiling_n8.txt 2>&1; cat $out/r02n_h2d_ceiling_n8.txt | tail -5
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02n_bench_*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        p=j.get("parity") or {}
        print(f.split("/")[-1], "N=%d GCUPS=%.1f avg_launch_ms=%.4f e2e=%s parity=%s/%s"%(j["n_gpus"],j["value"],j["roofline"]["avg_launch_ms"], j["e2e"] and round(j["e2e"]["value"],1), p.get("random_bitexact"), p.get("corner_bitexact")))
        for k,v in (j.get("also") or {}).items(): print("   also", k, v.get("value"), (v.get("parity") or {}).get("ok"), v.get("error"))
    except Exception as e: print(f, "FAILED", e)
PY
