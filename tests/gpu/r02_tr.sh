#!/bin/bash
# `ncu --set full` of the random-phase search for the two configurations whose bench line still said "traffic": null
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LL="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-job --no-secondary --no-hbm-kernel --streams 1"
timeout 200 ncu --set full --clock-control none -k regex:k_tile_search -s 10 -c 2 -f -o gpurun_out/r02tr_prof_search_C2urea python bench.py --config C2urea $LL --frames-per-step 16 > gpurun_out/r02tr_C2urea.log 2>&1; tail -1 gpurun_out/r02tr_C2urea.log | cut -c1-100
T0=$(date +%s)
timeout 240 ncu --set full --clock-control none -k regex:k_tile_search -s 8 -c 2 -f -o gpurun_out/r02tr_prof_search_C5 python bench.py --config C5 $LL --frames-per-step 2 > gpurun_out/r02tr_C5.log 2>&1; tail -1 gpurun_out/r02tr_C5.log | cut -c1-100
echo "C5 capture took $(( $(date +%s) - T0 )) s"
