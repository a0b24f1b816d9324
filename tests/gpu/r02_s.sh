#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for cfg in C4 C2; do
echo "== $cfg"; CMX_TRACE=30:2 timeout 300 python bench.py --config $cfg --steps 4 --streams 1 --no-cpu-baseline --no-e2e --no-hbm-kernel --no-job --no-secondary 2>&1 >/dev/null | grep "cmx trace\]  " | awk '{printf "%s %s | ", $3, $6} END{print ""}'
timeout 300 python bench.py --config $cfg --steps 5 --no-cpu-baseline --no-e2e --no-hbm-kernel --no-job --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2))"
done
