#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py --steps 5 > gpurun_out/r02o_bench_default.json 2> gpurun_out/r02o_bench_default.err; tail -3 gpurun_out/r02o_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02o_bench_default.json").read().strip().splitlines()[-1])
print("workload", d["config"]["workload"][:40], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "job fps", round(d["job"]["frames_per_s"],1), "job", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["job"].items() if k!="what"})
print("launches/frame", d["launches_per_frame"], "host submit/wait/cpu ms", round(d["host_submit_ms_per_step"],2), round(d["host_wait_ms_per_step"],2), round(d["host_cpu_ms_per_step"],2), "of", round(d["ms_per_step"],2))
print("cpu", d["cpu_baseline"])
print("roofline", {k:v for k,v in d["roofline"].items() if k in ("kernel","achieved","frac","kernel_ms_per_frame","kernel_share_of_frame","pair_evals_per_frame","frames_per_launch")})
print("secondary", d["secondary"]["value"], d["secondary"]["e2e"]["value"], d["secondary"]["launches_per_frame"], d["secondary"]["host_submit_ms_per_step"], d["secondary"]["host_wait_ms_per_step"], d["secondary"]["ms_per_step"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02o_bench_ref.json 2> gpurun_out/r02o_bench_ref.err; cut -c1-400 gpurun_out/r02o_bench_ref.json
