#!/bin/bash
# tests + A/B of the search grid (one resident wave with the tile queue vs 8 blocks/SM) + feed bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
show() { python - "$1" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "launches/frame", d["gpu_launches"]/(d["steps"]*d["config"]["frames_per_step"]), "submit_ms", round(d["host_submit_ms_per_step"],1), "dev_ms", round(d["device_ms_per_step"],1), "kernel_ms", round(r["kernel_ms_per_launch"],4), "share", round(r["kernel_share_of_frame"],3), "clk", d["clocks"]["sm_mhz"])
PY
}
for B in 0 8 6; do
  if [ $B -gt 0 ]; then export CMX_SEARCH_BLOCKS_PER_SM=$B; else unset CMX_SEARCH_BLOCKS_PER_SM; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_C2_b$B.json 2> gpurun_out/ab.err; show gpurun_out/ab_C2_b$B.json; tail -2 gpurun_out/ab.err
done
unset CMX_SEARCH_BLOCKS_PER_SM
timeout 300 python bench.py --config C2urea --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_C2urea.json 2> gpurun_out/ab.err; show gpurun_out/ab_C2urea.json; tail -2 gpurun_out/ab.err
timeout 300 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_C4.json 2> gpurun_out/ab.err; show gpurun_out/ab_C4.json; tail -2 gpurun_out/ab.err
timeout 600 python bench_extras.py feed --frames 256 > gpurun_out/extras_feed.json 2> gpurun_out/extras_feed.err; tail -2 gpurun_out/extras_feed.err; cat gpurun_out/extras_feed.json
