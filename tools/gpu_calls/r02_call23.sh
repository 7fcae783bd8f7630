#!/bin/bash
out=gpurun_out; mkdir -p $out
for cfg in 1 0 5 6; do
  FDB_LAP_CFG=$cfg timeout -s KILL 120 compute-sanitizer --tool synccheck --print-limit 1 python tools/synccheck_probe.py > $out/r02x_sync_lap7_cfg$cfg.log 2>&1
  echo "lap7_tma cfg $cfg: $(grep -E 'ERROR SUMMARY: [0-9]+ errors$|probe ok' $out/r02x_sync_lap7_cfg$cfg.log | cut -c1-120 | tr '\n' ' ') $(grep -m1 'located' $out/r02x_sync_lap7_cfg$cfg.log)"
done
