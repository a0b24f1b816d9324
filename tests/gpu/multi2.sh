#!/bin/bash
# N-GPU: public mddf (native + host feed, NCCL all-reduce) == single GPU; bench at N GPUs (C2 and C4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_mddf.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -5
for C in C2 C4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $C --steps 8 --warmup 3 > gpurun_out/bench_${C}_N$N.json 2> gpurun_out/bench_N$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${C}_N$N.json").read().strip().splitlines()[-1])
    print("$C N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2), "host submit", round(d["host_submit_ms_per_step"],2), "clocks", d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_N$N.err").read()[-2000:])
PY
done
