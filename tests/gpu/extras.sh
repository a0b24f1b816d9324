#!/bin/bash
# GPU validation of the SURVEY 8(f) rows: native DCD feed, device-side group reduction, merge; then the extras bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench_extras.py feed --frames 256 > gpurun_out/extras_feed.json 2> gpurun_out/extras_feed.err; tail -2 gpurun_out/extras_feed.err; cat gpurun_out/extras_feed.json
timeout 600 python bench_extras.py reduce > gpurun_out/extras_reduce.json 2> gpurun_out/extras_reduce.err; tail -2 gpurun_out/extras_reduce.err; cat gpurun_out/extras_reduce.json
