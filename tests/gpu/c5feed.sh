#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 200 python bench_extras.py feed --frames 256 > gpurun_out/extras_feed.json 2> gpurun_out/extras_feed.err; tail -2 gpurun_out/extras_feed.err; cat gpurun_out/extras_feed.json
timeout 230 python bench.py --config C5 --steps 2 --warmup 3 --no-cpu-baseline --frames-per-step 16 > gpurun_out/final_C5.json 2> gpurun_out/final_C5.err; tail -2 gpurun_out/final_C5.err; python - <<PY
import json
d=json.loads(open("gpurun_out/final_C5.json").read().strip().splitlines()[-1])
print("C5 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"],4), "share", round(d["roofline"]["kernel_share_of_frame"],3), "pair-evals/s", "%.3g"%d["roofline"]["pair_evals_per_s"])
PY
