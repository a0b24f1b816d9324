#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_gpu_feed.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -12
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench_extras.py feed --format xtc --frames 128 2>&1 | tail -2
timeout 600 python bench_extras.py feed --format dcd --frames 128 2>&1 | tail -2
