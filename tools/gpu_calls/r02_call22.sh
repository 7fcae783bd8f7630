#!/bin/bash
out=gpurun_out; mkdir -p $out
SAN_FUSES=1 SAN_SKIP_MIRROR=1 timeout -s KILL 200 compute-sanitizer --tool synccheck --print-limit 1 python tools/sanitize_small.py > $out/r02w_sync_lap.log 2>&1
echo "single-step upwind + generic + 7-point kernels: $(grep -E 'ERROR SUMMARY: [0-9]+ errors$|sanitize_small ok' $out/r02w_sync_lap.log | tr '\n' ' ')"; grep -m1 -A3 -E "Barrier error" $out/r02w_sync_lap.log | cut -c1-260; grep -E "Host Frame: <module>" $out/r02w_sync_lap.log | sort | uniq -c
