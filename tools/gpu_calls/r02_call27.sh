#!/bin/bash
# round 2, call 27 (1 GPU): default T = 4 -- whole GPU suite, DRAM traffic per launch for the slab shapes bench.py emits, default bench line
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 170 > $out/r02aa_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $out/r02aa_tests.log
probe() {
  timeout -s KILL 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$2 -s 2 -c 1 --csv --log-file $out/r02aa_traffic_$1_$3x$4x$5.csv python tools/traffic_probe.py $1 $3 $4 $5 > $out/r02aa_probe_$3x$4x$5.log 2>&1
  echo "$1 $3x$4x$5 $(grep -E 'dram__bytes|gpu__time' $out/r02aa_traffic_$1_$3x$4x$5.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ') $(grep -o 'upwind3d[a-z_]*<T=[0-9]>' $out/r02aa_probe_$3x$4x$5.log | head -1)"
}
probe upwind upwind3d_fused 512 512 512
probe upwind upwind3d_fused 256 1024 512
probe upwind upwind3d_fused 128 1024 1024
probe upwind upwind3d_fused 512 1024 1024
probe upwind upwind3d_fused 1024 1024 1024
probe upwind upwind3d_fused 256 2048 2048
timeout -s KILL 600 python bench.py > $out/r02aa_bench_default.json 2> $out/r02aa_bench_default.err; echo "bench rc=$?"; tail -c 300 $out/r02aa_bench_default.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r02aa_bench_default.json") if l.startswith("{")][-1])
print("N=1 GCUPS=%.1f kernel=%s e2e=%.1f parity=%s/%s launches=%d"%(j["value"],j["config"]["kernel"],j["e2e"]["value"],j["parity"]["random_bitexact"],j["parity"]["corner_bitexact"],j["gpu_launches"]))
for k,v in j["also"].items(): print("  also",k,v.get("value"),v.get("kernel"),(v.get("parity") or {}).get("ok"),v.get("error"))
PY
