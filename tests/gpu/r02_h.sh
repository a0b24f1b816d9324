#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for d in 0 1 2 4 8 16 32 64 7 31; do
echo "== dbg $d: $(CMX_DBG=$d CMX_TRACE=16:2 timeout 300 python bench.py --config C4 --steps 2 --streams 1 --no-cpu-baseline --no-e2e --no-hbm-kernel 2>&1 >/dev/null | grep -E "finalise<rand>|finalise<real>" | tr '\n' ' ')"
done
