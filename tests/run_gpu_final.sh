#!/bin/bash
# round-end style validation: parity tests, smoke, bench (N=1) + reference arm, ncu launch list + full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 900 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 16 --streams 1 > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 8 -c 2 -f -o gpurun_out/prof_search \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 8 --streams 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
