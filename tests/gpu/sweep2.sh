#!/bin/bash
# sweep: search blocks per SM x frames in flight (C2), then the feed bench with the fixed-cost breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
show() { python - "$1" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "submit_ms", round(d["host_submit_ms_per_step"],1), "dev_ms", round(d["device_ms_per_step"],1), "kernel_ms", round(r["kernel_ms_per_launch"],4), "clk", d["clocks"]["sm_mhz"])
PY
}
for CFG in "3 8" "4 8" "3 16" "4 16" "5 16" "5 12" "2 16"; do
  set -- $CFG
  export CMX_SEARCH_BLOCKS_PER_SM=$1
  timeout 300 python bench.py --steps 4 --warmup 3 --frames-per-step 256 --no-cpu-baseline --streams $2 > gpurun_out/sw_b$1_s$2.json 2> gpurun_out/sw.err; show gpurun_out/sw_b$1_s$2.json; tail -2 gpurun_out/sw.err
done
unset CMX_SEARCH_BLOCKS_PER_SM
timeout 600 python bench_extras.py feed --frames 256 > gpurun_out/extras_feed.json 2> gpurun_out/extras_feed.err; tail -2 gpurun_out/extras_feed.err; cat gpurun_out/extras_feed.json
