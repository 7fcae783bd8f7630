#!/bin/bash
out=gpurun_out; mkdir -p $out
run() { # label, env...
  label=$1; shift
  env "$@" SAN_ONLY_FUSED=1 SAN_FUSES=3 timeout -s KILL 200 compute-sanitizer --tool synccheck --print-limit 1 python tools/sanitize_small.py > $out/r02t_sync_$label.log 2>&1
  echo "$label: $(grep -E 'ERROR SUMMARY: [0-9]+ errors$|sanitize_small ok' $out/r02t_sync_$label.log | tr '\n' ' ') $(grep -m1 'located' $out/r02t_sync_$label.log)"
}
run exact_tiles SAN_SHAPE=7,38,128
run exact_tiles_1ktile SAN_SHAPE=7,19,128
run cfg1 FDB_FUSED_CFG=1
run cfg3 FDB_FUSED_CFG=3
run nograph FDB_GRAPH=0 SAN_SHAPE=7,38,128
run one_cta FDB_MAX_CTAS=1
