#!/bin/bash
# round 2, call 26 (2 GPUs): final code -- whole GPU suite on 2 GPUs, bench lines at N=2 and (lap1024) N=1, T=3 vs T=4 sustained
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 170 > $out/r02zz_tests_n2.log 2>&1; echo "gpu tests rc=$?"; tail -3 $out/r02zz_tests_n2.log
timeout -s KILL 600 $TR --nproc-per-node 2 --master-port 29841 bench.py --gpus 2 --steps 8 --warmup 3 > $out/r02zz_bench_n2.json 2> $out/r02zz_bench_n2.err; echo "bench n2 rc=$?"
timeout -s KILL 600 python bench.py --workload lap1024 --steps 3 --warmup 3 --no-also > $out/r02zz_bench_lap1024.json 2> $out/r02zz_bench_lap1024.err; echo "bench lap1024 rc=$?"; tail -c 300 $out/r02zz_bench_lap1024.err
for fuse in 3 4 3 4; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-also --no-parity --no-cpu-baseline --no-e2e --fuse $fuse 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('N=1 sustained, fuse $fuse: GCUPS=%.1f kernel=%s clocks=%s'%(j['value'],j['config']['kernel'],j['clocks']['sm_mhz']))"
done
for fuse in 3 4; do
  timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 2985$fuse bench.py --gpus 2 --steps 10 --warmup 3 --no-also --no-parity --no-e2e --fuse $fuse 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('N=2 sustained, fuse $fuse: GCUPS=%.1f'%(j['value']))"
done
python - <<'PY'
import json
for f in ("r02zz_bench_n2.json","r02zz_bench_lap1024.json"):
    try:
        j=json.loads([l for l in open("gpurun_out/"+f) if l.startswith("{")][-1]); p=j.get("parity") or {}
        print(f, "N=%d GCUPS=%.1f kernel=%s traffic=%s e2e=%s parity=%s/%s"%(j["n_gpus"],j["value"],j["config"]["kernel"],j["roofline"]["traffic"],j["e2e"] and round(j["e2e"]["value"],1),p.get("random_bitexact"),p.get("corner_bitexact")))
        for k,v in (j.get("also") or {}).items(): print("   also",k,v.get("value"),v.get("kernel"),(v.get("parity") or {}).get("ok"),v.get("error"))
    except Exception as e: print(f,"FAILED",e)
PY
