#!/bin/bash
# ncu launch list of the default bench command (C4, the library's own batching: 3 batches of 4 frames in flight), timed region only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02ll_launches_default.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-job --no-secondary --no-hbm-kernel --no-cpu-baseline --no-roofline > gpurun_out/r02ll.log 2>&1
tail -2 gpurun_out/r02ll.log | cut -c1-200; wc -l gpurun_out/r02ll_launches_default.csv
