#!/bin/bash
# round 2, first call: baseline launch lists of C4 / C3 / C5 with the round-1 code + the C4 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
for cfg in C4 C3 C5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02a_launches_$cfg.csv \
    python bench.py --config $cfg --steps 1 --warmup 3 --frames-per-step 4 --streams 1 --no-cpu-baseline --no-e2e --no-hbm-kernel > gpurun_out/r02a_ncu_$cfg.log 2>&1
  python profiles/ncu_summary.py launches gpurun_out/r02a_launches_$cfg.csv 32 > gpurun_out/r02a_launch_summary_$cfg.txt 2>&1
  head -30 gpurun_out/r02a_launch_summary_$cfg.txt
done
timeout 600 python bench.py --config C4 --steps 5 --no-hbm-kernel > gpurun_out/r02a_bench_C4.json 2> gpurun_out/r02a_bench_C4.err; tail -2 gpurun_out/r02a_bench_C4.err
cat gpurun_out/r02a_bench_C4.json | cut -c1-600
