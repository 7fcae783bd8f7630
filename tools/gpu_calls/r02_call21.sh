#!/bin/bash
out=gpurun_out; mkdir -p $out
for cfg in 0 3 2; do
  FDB_LAPF_CFG=$cfg SAN_FUSES= SAN_SKIP_MIRROR=1 timeout -s KILL 200 compute-sanitizer --tool synccheck --print-limit 1 python tools/sanitize_small.py > $out/r02v_sync_lap_cfg$cfg.log 2>&1
  echo "lap cfg $cfg: $(grep -E 'ERROR SUMMARY: [0-9]+ errors$|sanitize_small ok' $out/r02v_sync_lap_cfg$cfg.log | tr '\n' ' ') $(grep -m1 -E '     at ' $out/r02v_sync_lap_cfg$cfg.log | cut -c1-200)"
done
