#!/bin/bash
# round 2, call 17 (1 GPU): what synccheck's "Missing init" on the first-formulation T=2 kernel depends on
out=gpurun_out; mkdir -p $out
FDB_GRAPH=0 timeout -s KILL 300 compute-sanitizer --tool synccheck --print-limit 3 python tools/sanitize_small.py > $out/r02r_synccheck_nograph.log 2>&1; echo "graph=0 impl default:"; grep -E "ERROR SUMMARY|sanitize_small ok|     at " $out/r02r_synccheck_nograph.log | sort | uniq -c | head -5
FDB_FUSED_IMPL=4 timeout -s KILL 300 compute-sanitizer --tool synccheck --print-limit 3 python tools/sanitize_small.py > $out/r02r_synccheck_impl4.log 2>&1; echo "graph default, impl 4 for every T:"; grep -E "ERROR SUMMARY|sanitize_small ok|     at " $out/r02r_synccheck_impl4.log | sort | uniq -c | head -5
