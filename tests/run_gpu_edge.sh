#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs -s 3 -c 1 -f -o gpurun_out/prof_pairs \
   python bench.py --config C3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 8 --streams 1 > gpurun_out/ncu_pairs.log 2>&1
tail -1 gpurun_out/ncu_pairs.log | cut -c1-200
