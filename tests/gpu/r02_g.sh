#!/bin/bash
# ncu --set full of the heavy kernels of one C4 batch (batched code), for source-level attribution
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="--steps 1 --warmup 3 --frames-per-step 8 --streams 1 --no-cpu-baseline --no-e2e --no-hbm-kernel"
# per batch: 23 launches; prefill 2 batches + 3 warmup x 2 + 1 timed x 2 = 10 batches -> skip 230, capture the 23 launches of one batch of the profile pass
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 230 --launch-count 23 -f -o gpurun_out/r02g_full_C4 python bench.py --config C4 $B > gpurun_out/r02g_ncu_C4.log 2>&1
tail -2 gpurun_out/r02g_ncu_C4.log
ls -la gpurun_out/r02g_full_C4.ncu-rep
