#!/bin/bash
# final sweep of the round (second session): parity tests, smoke, bench of the named configurations, reference arm,
# ncu launch list + full captures (tile search, group reduction), extras.  Steps are skipped once the time budget is spent.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s); LIMIT=${1:-540}
left() { [ $(( $(date +%s) - T0 )) -lt $LIMIT ]; }
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { C=$1; shift; left || { echo "skip $C (time)"; return; }; timeout 400 python bench.py --config $C "$@" > gpurun_out/final_$C.json 2> gpurun_out/final_$C.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/final_$C.json").read().strip().splitlines()[-1])
    cb=d.get("cpu_baseline") or {}
    print("$C value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", round(cb.get("value",0),3), "cores", cb.get("cores"), "equal", cb.get("counts_equal_device"), "| kernel", d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"],4), "share", round(d["roofline"]["kernel_share_of_frame"],3), "frac", round(d["roofline"]["frac"],4), "pair-evals/s", "%.3g"%d["roofline"]["pair_evals_per_s"], "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"]["samples"])
except Exception as e:
    print("$C failed", e); print(open("gpurun_out/final_$C.err").read()[-800:])
PY
}
run C2 --steps 20 --warmup 3
left && { timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference_C2.json 2>/dev/null; tail -c 300 gpurun_out/final_reference_C2.json; }
left && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 260 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 16 --streams 1 > gpurun_out/ncu_bench.log 2>&1
left && { timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 8 -c 2 -f -o gpurun_out/prof_search \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 8 --streams 1 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-100; }
left && { timeout 300 python bench_extras.py reduce --repeat 2 > gpurun_out/extras_reduce.json 2> gpurun_out/extras_reduce.err; cat gpurun_out/extras_reduce.json | cut -c1-700; }
run C2urea --steps 10 --warmup 3 --cpu-frames 16
run C4 --steps 5 --warmup 3 --cpu-frames 2
run C3 --steps 4 --warmup 3 --cpu-frames 4
left && { timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_reduce_rows -c 1 -f -o gpurun_out/prof_reduce \
   python bench_extras.py reduce --repeat 1 --rows 500000 > gpurun_out/ncu_reduce.log 2>&1; tail -1 gpurun_out/ncu_reduce.log | cut -c1-100; }
echo "elapsed $(( $(date +%s) - T0 )) s"
