#!/bin/bash
# the default bench line on one GPU (clock sampler via NVML), then compute-sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02b1_bench_default.json 2> gpurun_out/r02b1_bench_default.err; tail -3 gpurun_out/r02b1_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b1_bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ceiling", round(d["e2e"]["h2d_ceiling"]["frames_per_s_ceiling"]), "clocks", d["clocks"], "job fps", round(d["job"]["frames_per_s"],1), "traffic", d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["counts_equal_device"])
PY
bash tests/gpu/r02_san.sh
