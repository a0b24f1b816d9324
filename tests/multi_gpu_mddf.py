"""Multi-GPU check of the public driver (run under torchrun, one rank per GPU; NOT collected by pytest):
mddf() on a DCD file with the native feed shards the frames over the ranks, sums the integer counters with ONE
NCCL all-reduce and must reproduce the single-GPU result bit for bit (rank 0 recomputes it alone and compares).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_mddf.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import cmx_b200 as cm
from cmx_b200 import synthetic as syn
from common import write_dcd


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    s = syn.config_c2(0.25)
    sol, wat = s.selections["solute"], s.selections["water"]
    nf = 23                                    # not a multiple of the world size
    path = os.path.join(tempfile.gettempdir(), "cmx_multi_c2.dcd")
    if rank == 0:
        frames = np.stack([s.frame(k + 1)[0] for k in range(nf)]).astype(np.float32)
        write_dcd(path, frames, np.asarray(s.cell, dtype=np.float64))
    dist.barrier()
    opt = cm.Options(bulk_range=(10.0, 15.0), n_random_samples=5, seed=321, silent=True, irefatom=1)
    ok = True
    # all weights equal (integer all-reduce); equal within a rank but different between ranks; different everywhere
    # (the last two take the f64 exchange: every rank applies its own weights before the sum)
    for case, w in (("uniform", [1.0] * nf), ("per-rank", [1.0 if k % world == 0 else 2.0 for k in range(nf)]),
                    ("mixed", [(0.5, 1.0, 2.0)[k % 3] for k in range(nf)])):
        out = {}
        for feed in ("native", "host"):
            out[feed] = cm.mddf(path, sol, wat, opt, frame_weights=w, feed=feed, device=local)
        if rank == 0:
            dist.barrier()
            # single-process reference on this rank's GPU (the process group is ignored)
            R1 = cm.mddf(path, sol, wat, opt, frame_weights=w, feed="native", device=local, distributed=False)
            for feed, R in out.items():
                for key in ("md_count", "md_count_random", "rdf_count", "rdf_count_random", "solute_group_count",
                            "solvent_group_count", "solute_group_count_random", "mddf", "kb"):
                    same = np.array_equal(getattr(R, key), getattr(R1, key))
                    ok &= bool(same)
                    if not same:
                        print("MISMATCH", case, feed, key)
                ok &= R.volume.total == R1.volume.total
            print(f"MULTI_GPU_MDDF case={case} world={world} frames={nf} hits={R1.md_count.sum():.1f} identical so far={ok}")
        else:
            dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_MDDF world={world} identical={ok}")
        os.remove(path)
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
