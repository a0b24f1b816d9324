#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
run() { echo -n "$* : "; env "$@" timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 192 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('value', round(d['value'],1), 'tile_ms', round(r['kernel_ms_per_launch'],4), 'pe/frame %.3g'%r['pair_evals_per_frame'])"; }
run CMX_ROWDIV=2
run CMX_ROWDIV=3 CMX_QSIDE=5.0
run CMX_ROWDIV=3 CMX_QSIDE=4.5
run CMX_ROWDIV=3 CMX_QSIDE=5.5
run CMX_ROWDIV=2 CMX_QSIDE=5.0
run CMX_ROWDIV=3 CMX_QSIDE=5.0 CMX_CULLDIV=7
run CMX_ROWDIV=3 CMX_QSIDE=5.0 CMX_CULLDIV=4
run CMX_ROWDIV=3 CMX_QSIDE=5.0 CMX_XSIDE=1.875
