#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_N$N.json 2> gpurun_out/bench_N$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_N$N.json").read().strip().splitlines()[-1])
    print("N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2), "host submit", round(d["host_submit_ms_per_step"],2), "clocks", d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_N$N.err").read()[-2000:])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 0 --cpu-frames 8 2>/dev/null | tail -c 400
