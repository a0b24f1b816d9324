#!/bin/bash
# N GPUs: bench.py under torchrun exactly as the driver launches it (default workload), nothing else
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout ${2:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02y_n$N.json 2> gpurun_out/r02y_n$N.err; tail -3 gpurun_out/r02y_n$N.err | cut -c1-300
python - $N <<'PY'
import json, sys
N=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02y_n{N}.json").read().strip().splitlines()[-1])
print("N=%s value"%N, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "h2d ceiling", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"]["h2d_ceiling"].items() if k!="what"}, "guard", d["multi_gpu_sum_equals_single_gpu"], "clocks", d["clocks"])
print("job", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["job"].items() if k!="what"})
s=d["secondary"]; print("secondary C2 value", round(s["value"],1), "e2e", round(s["e2e"]["value"],1), "")
PY
echo "elapsed $(( $(date +%s) - T0 )) s"
