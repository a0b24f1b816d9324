#!/bin/bash
# round 2: parity suite after the batched-launch refactor + quick bench sweep over batch sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for cfg in C2 C4; do
  for b in 0 1 4 16; do
    [ "$cfg" = C4 ] && [ "$b" = 16 ] && continue
    CMX_BATCH=$b timeout 300 python bench.py --config $cfg --steps 5 --no-cpu-baseline --no-e2e --no-hbm-kernel > gpurun_out/r02c_${cfg}_b$b.json 2> gpurun_out/r02c_${cfg}_b$b.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_${cfg}_b$b.json").read().strip().splitlines()[-1])
    print("$cfg batch $b value", round(d["value"],1), "launches", d["gpu_launches"], "host_submit_ms", round(d["host_submit_ms_per_step"],2), "ms/step", round(d["ms_per_step"],2), "search ms", round(d["roofline"]["kernel_ms_per_launch"],4))
except Exception as e:
    print("$cfg batch $b failed", e); print(open("gpurun_out/r02c_${cfg}_b$b.err").read()[-800:])
PY
  done
done
