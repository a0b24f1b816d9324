#!/bin/bash
# experiment: smaller grids for the latency-bound kernels, so that the searches of the other batches can share the SMs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
Q="--steps 6 --no-cpu-baseline --no-e2e --no-job --no-secondary --no-hbm-kernel --no-roofline"
run() { cfg=$1; shift; extra=""; while [ "$1" = "--arg" ]; do extra="$extra $2"; shift; shift; done
  v=$(env "$@" timeout 120 python bench.py --config $cfg $Q $extra 2>/dev/null | python -c "
import json,sys
try: d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1))
except Exception as e: print('failed', e)")
  echo "$cfg $extra $* | value $v"; }
run C4
run C4 CMX_GRID_SCALE=0.5
run C4 CMX_GRID_SCALE=0.25
run C4 CMX_GRID_SCALE=0.125
run C4 --arg "--streams 4" CMX_GRID_SCALE=0.25
run C4 --arg "--streams 6" CMX_GRID_SCALE=0.25
run C2 
run C2 CMX_GRID_SCALE=0.25
