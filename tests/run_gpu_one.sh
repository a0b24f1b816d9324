#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
SECONDS=0; timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench wall seconds: $SECONDS"; tail -2 gpurun_out/bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fps/step", d["config"]["frames_per_step"], "ms/step", round(d["ms_per_step"],2), "host submit", round(d["host_submit_ms_per_step"],2), "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"], "traffic", d["roofline"]["traffic"], "l2", d["config"]["l2"])
PY
