#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in C2 C4; do
for st in 1 3; do
echo "== $cfg streams $st"
CMX_TRACE=60:2 timeout 300 python bench.py --config $cfg --steps 3 --streams $st --no-cpu-baseline --no-e2e --no-hbm-kernel 2>&1 >/dev/null | grep "cmx trace" | head -40
done; done
