#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
run() { echo -n "$* : "; timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 256 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('value', round(d['value'],1), 'host submit ms', round(d['host_submit_ms_per_step'],1), 'dev ms', round(d['device_ms_per_step'],1))"; }
run --streams 4
run --streams 6
run --streams 8
run --streams 12
run --streams 16
run --config C4 --streams 4 --frames-per-step 48
run --config C4 --streams 8 --frames-per-step 48
run --config C4 --streams 12 --frames-per-step 48
