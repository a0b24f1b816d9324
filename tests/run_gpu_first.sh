#!/bin/bash
# first GPU bring-up: build, tests with compute-sanitizer on one small case, then the parity suite
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nvidia-smi --query-gpu=name,memory.total --format=csv
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -40
