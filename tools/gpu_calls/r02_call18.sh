#!/bin/bash
out=gpurun_out; mkdir -p $out
for f in 3 4 ""; do
  SAN_FUSES=$f timeout -s KILL 300 compute-sanitizer --tool synccheck --print-limit 2 python tools/sanitize_small.py > $out/r02s_synccheck_f$f.log 2>&1; echo "SAN_FUSES='$f':"; grep -E "ERROR SUMMARY|sanitize_small ok|     at |Device Frame.*kernel|located" $out/r02s_synccheck_f$f.log | sort | uniq -c | head -6
done
