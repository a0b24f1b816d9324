#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "namd or synthetic or custom or weights or toy or c2 or edge" 2>&1 | tail -5
for fb in 74 148 296 592 1184; do
echo "== fin blocks $fb: $(CMX_FIN_BLOCKS=$fb CMX_TRACE=16:2 timeout 300 python bench.py --config C4 --steps 2 --streams 1 --no-cpu-baseline --no-e2e --no-hbm-kernel 2>&1 >/dev/null | grep -E "finalise<rand>|finalise<real>" | tr '\n' ' ')"
done
for cfg in C4 C2; do
timeout 300 python bench.py --config $cfg --steps 5 --no-cpu-baseline --no-e2e --no-hbm-kernel 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2))"
done
