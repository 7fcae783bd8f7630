#!/bin/bash
# round 2, call 25 (1 GPU): final code -- whole GPU suite, the three sanitizers, smoke(), launch list of the bench command
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 170 > $out/r02z_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $out/r02z_tests.log
for tool in synccheck memcheck racecheck; do
  timeout -s KILL 400 compute-sanitizer --tool $tool --print-limit 3 python tools/sanitize_small.py > $out/r02z_sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small ok' $out/r02z_sanitizer_$tool.log | tr '\n' ' ')"
done
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r02z_launches.csv python bench.py --steps 2 --warmup 3 --no-also --no-parity --no-cpu-baseline > $out/r02z_bench_under_ncu.log 2>&1; echo "launch list rc=$? lines=$(wc -l < $out/r02z_launches.csv)"
for fuse in 3 4; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-also --no-parity --no-cpu-baseline --no-e2e --fuse $fuse 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('sustained bench, fuse $fuse: GCUPS=%.1f avg_launch_ms=%.4f clocks=%s %s'%(j['value'],j['roofline']['avg_launch_ms'],j['clocks']['sm_mhz'],j['clocks']['reasons']))"
done
