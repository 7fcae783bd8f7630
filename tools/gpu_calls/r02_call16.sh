#!/bin/bash
# round 2, call 16 (1 GPU): sanitizers on the new default kernels, DRAM traffic per launch for every slab shape bench.py
# emits, the full default bench line (with the 2-D default laplacian case), the fresh-line contract test
out=gpurun_out; mkdir -p $out
for tool in memcheck synccheck; do
  timeout -s KILL 400 compute-sanitizer --tool $tool python tools/sanitize_small.py > $out/r02q_sanitizer_$tool.log 2>&1
  tail -2 $out/r02q_sanitizer_$tool.log
done
probe() {  # kind regex n0 n1 n2
  timeout -s KILL 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$2 -s 2 -c 1 --csv --log-file $out/r02q_traffic_$1_$3x$4x$5.csv python tools/traffic_probe.py $1 $3 $4 $5 > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time" $out/r02q_traffic_$1_$3x$4x$5.csv | awk -F'","' '{print "'$1' '$3'x'$4'x'$5'", $(NF-3), $(NF-2), $(NF-1), $NF}'
}
probe upwind upwind3d_fused 512 512 512
probe upwind upwind3d_fused 256 1024 512
probe upwind upwind3d_fused 128 1024 1024
probe upwind upwind3d_fused 512 1024 1024
probe upwind upwind3d_fused 1024 1024 1024
probe upwind upwind3d_fused 256 2048 2048
probe upwind upwind3d_fused 128 128 128
probe lap lap7_fused2 1024 1024 1024
probe lap lap7_fused2 128 1024 1024
timeout -s KILL 900 python bench.py > $out/r02q_bench_default.json 2> $out/r02q_bench_default.err; echo "bench rc=$?"; tail -c 400 $out/r02q_bench_default.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r02q_bench_default.json") if l.startswith("{")][-1])
print("N=1 GCUPS=%.1f e2e=%.1f parity=%s/%s clocks=%s"%(j["value"],j["e2e"]["value"],j["parity"]["random_bitexact"],j["parity"]["corner_bitexact"],j["clocks"]))
for k,v in j["also"].items(): print("  also",k,v.get("value"),v.get("kernel"),(v.get("parity") or {}).get("ok"),v.get("single_apply",{}).get("value"),v.get("error"))
print("  e2e_process", j["e2e_process"])
PY
timeout -s KILL 400 python -m pytest tests/test_bench_contract.py -m gpu -q > $out/r02q_contract_test.log 2>&1; tail -3 $out/r02q_contract_test.log
