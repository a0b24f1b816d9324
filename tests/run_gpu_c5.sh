#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for C in ${CONFIGS:-C2 C3}; do
EXTRA=""; [ $C = C5 ] && EXTRA="--frames-per-step 16"
timeout 420 python bench.py --config $C --steps 3 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/bench_$C.json 2> gpurun_out/bench_$C.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$C.json").read().strip().splitlines()[-1])
    print("$C value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fps/step", d["config"]["frames_per_step"], "host submit ms", round(d["host_submit_ms_per_step"],2), "dev ms", round(d["device_ms_per_step"],2), "roof", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("kernel","kernel_ms_per_launch","kernel_share_of_frame","pair_evals_per_frame")}, "deferred", d["config"]["deferred_to_exact_per_frame"])
except Exception as e:
    print("$C failed", e); print(open("gpurun_out/bench_$C.err").read()[-1500:])
PY
done
