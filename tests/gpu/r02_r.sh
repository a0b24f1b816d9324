#!/bin/bash
# 2 GPUs: group handle on two devices, driver under torchrun (mixed weights), bench at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02r_n2.json 2> gpurun_out/r02r_n2.err; tail -3 gpurun_out/r02r_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02r_n2.json").read().strip().splitlines()[-1])
print("N=2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "guard", d["multi_gpu_sum_equals_single_gpu"], "job", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["job"].items() if k!="what"})
print("secondary", round(d["secondary"]["value"],1), round(d["secondary"]["e2e"]["value"],1))
PY
timeout 600 python bench_extras.py group --devices 0,1 2>&1 | tail -3
