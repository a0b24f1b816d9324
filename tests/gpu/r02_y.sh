#!/bin/bash
# N GPUs (gpurun --gpus N): multi-GPU tests (group handle, driver under torchrun), bench.py under torchrun (C4 default + C2 secondary,
# e2e, job leg, guard), group-handle leg of bench_extras
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -8; nvidia-smi topo -m 2>/dev/null | head -12; nproc; numactl -H 2>/dev/null | head -4
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_final.py -q -m gpu 2>&1 | tail -6; fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02y_n$N.json 2> gpurun_out/r02y_n$N.err; tail -3 gpurun_out/r02y_n$N.err
python - $N <<'PY'
import json, sys
N=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02y_n{N}.json").read().strip().splitlines()[-1])
print("N=%s value"%N, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "guard", d["multi_gpu_sum_equals_single_gpu"], "clocks", d["clocks"])
print("job", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["job"].items() if k!="what"})
s=d["secondary"]; print("secondary C2 value", round(s["value"],1), "e2e", round(s["e2e"]["value"],1))
PY
timeout 600 python bench_extras.py group --devices $(seq -s, 0 $((N-1))) 2>&1 | tail -3
