#!/bin/bash
# round 2, call 31 (1 GPU): last check of the final tree -- smoke(), the whole GPU suite, a short bench line
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 500 python -m pytest tests -m gpu -q --timeout 170 > gpurun_out/r02ee_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02ee_tests.log
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-also 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('bench: GCUPS=%.1f kernel=%s e2e=%.1f parity=%s/%s cpu=%.3f e2e_process_512=%s'%(j['value'],j['config']['kernel'],j['e2e']['value'],j['parity']['random_bitexact'],j['parity']['corner_bitexact'],j['cpu_baseline']['value'],j['e2e_process']['upwind -numCells 512 -numSteps 10'].get('speedup')))"
