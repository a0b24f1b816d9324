#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="--steps 1 --warmup 3 --frames-per-step 8 --streams 1 --no-cpu-baseline --no-e2e --no-hbm-kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile_search" --launch-skip 20 --launch-count 2 -f -o gpurun_out/r02k_search_C4 python bench.py --config C4 $B > gpurun_out/r02k_ncu_C4.log 2>&1
tail -2 gpurun_out/r02k_ncu_C4.log
