#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
G=${1:-8}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 16 --streams 1 > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 8 -c 2 -f -o gpurun_out/prof_search \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 8 --streams 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
