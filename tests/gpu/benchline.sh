#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 280 python bench.py > gpurun_out/final_C2.json 2> gpurun_out/final_C2.err; tail -3 gpurun_out/final_C2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/final_C2.json").read().strip().splitlines()[-1])
print("C2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"]["value"], "hbm kernel", d["roofline_hbm_kernel"])
PY
