#!/bin/bash
# GPU round: parity tests, smoke, bench (N=1), ncu launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 32 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
