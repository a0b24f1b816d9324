#!/bin/bash
# N-GPU bench only (C4 = the 8-GPU target configuration of BASELINE.json, then C2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
for C in C4 C2; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $C --steps 8 --warmup 3 > gpurun_out/bench_${C}_N$N.json 2> gpurun_out/bench_N$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${C}_N$N.json").read().strip().splitlines()[-1])
    print("$C N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2), "host submit", round(d["host_submit_ms_per_step"],2), "clocks", d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_N$N.err").read()[-2000:])
PY
done
