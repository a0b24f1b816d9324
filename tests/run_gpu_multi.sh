#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --group-lanes 8 > gpurun_out/bench_N$N.json 2> gpurun_out/bench_N$N.err
tail -c 2500 gpurun_out/bench_N$N.json; tail -5 gpurun_out/bench_N$N.err
# sharded public driver: mddf() under torchrun equals the single-process result
cat > /tmp/shard_check.py <<'PY'
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cmx_b200 as cm
from common import namd
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
d = namd()
sel_p = cm.AtomSelection(np.arange(1, 1464), nmols=1); sel_t = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
opt = cm.Options(bulk_range=(8.0, 10.0), seed=321, silent=True, n_random_samples=5)
fr = np.concatenate([d["protein"], d["tmao"]], axis=1)
R = cm.mddf(cm.ArrayTrajectory(fr, d["cells"], sel_p, sel_t), opt)
if dist.get_rank() == 0:
    np.save("gpurun_out/shard_mddf.npy", np.stack([R.md_count, R.md_count_random, R.mddf, R.kb]))
    print("sharded mddf ok: sum md_count", R.md_count.sum(), "volume", R.volume.total)
dist.destroy_process_group()
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 /tmp/shard_check.py 2>&1 | tail -3
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cmx_b200 as cm
from common import namd
d = namd()
sel_p = cm.AtomSelection(np.arange(1, 1464), nmols=1); sel_t = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
opt = cm.Options(bulk_range=(8.0, 10.0), seed=321, silent=True, n_random_samples=5)
fr = np.concatenate([d["protein"], d["tmao"]], axis=1)
R = cm.mddf(cm.ArrayTrajectory(fr, d["cells"], sel_p, sel_t), opt)
ref = np.stack([R.md_count, R.md_count_random, R.mddf, R.kb])
got = np.load("gpurun_out/shard_mddf.npy")
print("single == sharded: raw md_count exact", np.array_equal(ref[0], got[0]), "| all within 1e-12", np.allclose(ref, got, rtol=1e-12, atol=0), "| max rel", np.max(np.abs(ref - got) / np.maximum(np.abs(ref), 1e-300)))
PY
