#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 120 --csv --log-file gpurun_out/launches_c5.csv \
   python bench.py --config C5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 4 --streams 1 > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log | cut -c1-300
