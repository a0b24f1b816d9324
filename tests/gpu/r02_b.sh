#!/bin/bash
# round 2, second call: ncu --set full of one whole frame's heavy kernels for C4, C3 and C5 (round-1 code)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="--steps 1 --warmup 3 --frames-per-step 4 --streams 1 --no-cpu-baseline --no-e2e --no-hbm-kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile_search|k_finalise|k_gen_rand|k_filter_rand|k_filter_real|k_edt_xy|k_qscatter|k_resolve" \
  --launch-skip 240 --launch-count 12 -f -o gpurun_out/r02b_full_C4 python bench.py --config C4 $B > gpurun_out/r02b_ncu_C4.log 2>&1
tail -2 gpurun_out/r02b_ncu_C4.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pairs|k_pair_resolve|k_ref_lists|k_pair_random|k_mol_prep|k_bulk_compact" \
  --launch-skip 140 --launch-count 7 -f -o gpurun_out/r02b_full_C3 python bench.py --config C3 $B > gpurun_out/r02b_ncu_C3.log 2>&1
tail -2 gpurun_out/r02b_ncu_C3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile_search|k_finalise|k_gen_rand" \
  --launch-skip 120 --launch-count 6 -f -o gpurun_out/r02b_full_C5 python bench.py --config C5 $B > gpurun_out/r02b_ncu_C5.log 2>&1
tail -2 gpurun_out/r02b_ncu_C5.log
ls -la gpurun_out/*.ncu-rep
