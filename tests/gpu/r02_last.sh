#!/bin/bash
# last check of the committed tree: GPU suite, smoke, the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r02last_bench_default.json 2> gpurun_out/r02last_bench_default.err; tail -2 gpurun_out/r02last_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02last_bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "clocks", d["clocks"], "job", round(d["job"]["frames_per_s"],1), "cpu", round(d["cpu_baseline"]["value"],3), d["cpu_baseline"]["counts_equal_device"], "launches", d["gpu_launches"], "traffic", d["roofline"]["traffic"], "frac", round(d["roofline"]["frac"],3))
s=d["secondary"]; print("secondary C2", round(s["value"],1), round(s["e2e"]["value"],1))
PY
timeout 120 python bench.py --impl reference --steps 1 --warmup 1 | cut -c1-260
