#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for S in 8; do
timeout 420 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $S > gpurun_out/bench_one.json 2> gpurun_out/bench_one.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_one.json").read().strip().splitlines()[-1])
print("streams $S value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step dev", round(d["device_ms_per_step"],2), "wall", round(d["wall_ms_per_step"],2), "host submit", round(d["host_submit_ms_per_step"],2), "launches/frame", d["gpu_launches"]/d["config"]["frames_per_step"]/d["steps"])
PY
done
