#!/bin/bash
# round 2: launch lists of the batched pipeline (C2 batch 16, C4 batch 4) + stream / search-grid sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in C2 C4; do
  fps=32; [ "$cfg" = C4 ] && fps=8
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02d_launches_$cfg.csv \
    python bench.py --config $cfg --steps 1 --warmup 3 --frames-per-step $fps --streams 1 --no-cpu-baseline --no-e2e --no-hbm-kernel > gpurun_out/r02d_ncu_$cfg.log 2>&1
  python profiles/ncu_summary.py launches gpurun_out/r02d_launches_$cfg.csv 1 > gpurun_out/r02d_launch_summary_$cfg.txt 2>&1
  head -30 gpurun_out/r02d_launch_summary_$cfg.txt
done
run() { # cfg streams searchblocks
  CMX_SEARCH_BLOCKS_PER_SM=$3 timeout 300 python bench.py --config $1 --steps 5 --streams $2 --no-cpu-baseline --no-e2e --no-hbm-kernel > gpurun_out/r02d_tmp.json 2> gpurun_out/r02d_tmp.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02d_tmp.json").read().strip().splitlines()[-1])
    print("$1 streams $2 searchblocks $3 value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2))
except Exception as e:
    print("$1 $2 $3 failed", e); print(open("gpurun_out/r02d_tmp.err").read()[-500:])
PY
}
for st in 1 2 4; do run C2 $st 0; done
for sb in 1 3 5; do run C2 3 $sb; done
for st in 1 2 4; do run C4 $st 0; done
for sb in 1 3 5; do run C4 3 $sb; done
