#!/bin/bash
# CPU arm (the fp64 port of the reference path on the GPU box's host cores) for every configuration + the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/final_C2.json 2> gpurun_out/final_C2.err; tail -2 gpurun_out/final_C2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/final_C2.json").read().strip().splitlines()[-1])
print("C2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"])
PY
for CF in "C2 16" "C2urea 32" "C3 16" "C4 16"; do set -- $CF
timeout 200 python bench.py --impl reference --config $1 --steps 1 --warmup 1 --cpu-frames $2 > gpurun_out/cpuarm_$1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/cpuarm_$1.json").read().strip().splitlines()[-1])
print("$1 cpu arm", round(d["value"],3), "frames/s", d["cpu_baseline"]["cores"], "threads", d["config"]["frames_per_step"], "frames/step")
PY
done
