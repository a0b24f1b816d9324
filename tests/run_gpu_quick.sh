#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for G in 1 8; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --group-lanes 8 --streams $G > gpurun_out/bench_G$G.json 2> gpurun_out/bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_G$G.json").read().strip().splitlines()[-1])
print("streams=$G value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "roof", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("kernel","kernel_ms_per_launch","kernel_share_of_frame","pair_evals_per_frame","frac")}, "deferred/frame", d["config"]["deferred_to_exact_per_frame"])
PY
tail -3 gpurun_out/bench.err
done
timeout 900 python bench.py --config C3 --steps 2 --warmup 3 --no-cpu-baseline --group-lanes 8 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_C3.json").read().strip().splitlines()[-1])
print("C3 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "deferred/frame", d["config"]["deferred_to_exact_per_frame"], d["roofline"]["kernel"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["kernel_share_of_frame"])
PY
