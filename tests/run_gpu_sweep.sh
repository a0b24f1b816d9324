#!/bin/bash
# round-end sweep: all named configurations (bench JSON per config) + profiles of the default config
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { C=$1; shift; timeout 600 python bench.py --config $C "$@" > gpurun_out/sweep_$C.json 2> gpurun_out/sweep_$C.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_$C.json").read().strip().splitlines()[-1])
    cb=d.get("cpu_baseline") or {}
    print("$C value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", round(cb.get("value",0),3), "cores", cb.get("cores"), "equal", cb.get("counts_equal_device"), "| kernel", d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"],4), "share", round(d["roofline"]["kernel_share_of_frame"],3), "frac", round(d["roofline"]["frac"],4), "pair-evals/s", "%.3g"%d["roofline"]["pair_evals_per_s"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"]["samples"])
except Exception as e:
    print("$C failed", e); print(open("gpurun_out/sweep_$C.err").read()[-800:])
PY
}
run C2 --steps 20 --warmup 3
run C2urea --steps 10 --warmup 3 --cpu-frames 16
run C3 --steps 5 --warmup 3 --cpu-frames 4
run C4 --steps 5 --warmup 3 --cpu-frames 2
run C5 --steps 2 --warmup 3 --no-cpu-baseline --frames-per-step 16
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/sweep_reference_C2.json 2>/dev/null; tail -c 300 gpurun_out/sweep_reference_C2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 16 --streams 1 > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 8 -c 2 -f -o gpurun_out/prof_search \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-step 8 --streams 1 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-100
