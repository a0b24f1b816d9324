#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --config C3 --steps 5 --cpu-frames 4 --no-secondary --no-hbm-kernel > gpurun_out/r02v2_bench_C3.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02v2_bench_C3.json').read().strip().splitlines()[-1]); print('C3 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['counts_equal_device'], 'launches/frame', d['launches_per_frame'])"
