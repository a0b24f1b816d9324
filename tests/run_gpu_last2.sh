#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 80 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
