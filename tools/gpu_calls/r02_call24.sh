#!/bin/bash
out=gpurun_out; mkdir -p $out
SAN_ONLY_FUSED=1 SAN_FUSES=3 timeout -s KILL 200 compute-sanitizer --tool synccheck --print-limit 1 python tools/sanitize_small.py > $out/r02y_sync_no_landed_wait.log 2>&1
echo "consumers wait on full only (debug build): $(grep -E 'ERROR SUMMARY: [0-9]+ errors$|sanitize_small ok' $out/r02y_sync_no_landed_wait.log | tr '\n' ' ')"; grep -m1 -A4 "Barrier error" $out/r02y_sync_no_landed_wait.log | cut -c1-200
