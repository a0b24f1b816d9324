#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
show() { python - "$1" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "submit_ms", round(d["host_submit_ms_per_step"],1), "dev_ms", round(d["device_ms_per_step"],1), "kernel_ms", round(r["kernel_ms_per_launch"],4), "share", round(r["kernel_share_of_frame"],3), "clk", d["clocks"]["sm_mhz"])
PY
}
run() { F=gpurun_out/ck4_$1_s$2.json; timeout 300 python bench.py --config $1 --steps $3 --warmup 3 --no-cpu-baseline --streams $2 > $F 2> gpurun_out/ck4.err; show $F; tail -2 gpurun_out/ck4.err; }
run C2 0 8
run C2 16 8
run C2 4 8
run C2urea 0 8
run C4 0 4
run C4 8 4
run C3 0 3
