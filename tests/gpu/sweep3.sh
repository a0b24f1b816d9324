#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
show() { python - "$1" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "submit_ms", round(d["host_submit_ms_per_step"],1), "dev_ms", round(d["device_ms_per_step"],1), "kernel_ms", round(r["kernel_ms_per_launch"],4), "clk", d["clocks"]["sm_mhz"])
PY
}
run() { # blocks streams maxconn config
  export CMX_SEARCH_BLOCKS_PER_SM=$1
  if [ "$3" != "0" ]; then export CUDA_DEVICE_MAX_CONNECTIONS=$3; else unset CUDA_DEVICE_MAX_CONNECTIONS; fi
  F=gpurun_out/sw3_$4_b$1_s$2_c$3.json
  timeout 300 python bench.py --config $4 --steps 4 --warmup 3 $5 --no-cpu-baseline --streams $2 > $F 2> gpurun_out/sw.err; show $F; tail -2 gpurun_out/sw.err
}
run 1 16 0 C2 "--frames-per-step 256"
run 2 8 0 C2 "--frames-per-step 256"
run 2 16 32 C2 "--frames-per-step 256"
run 3 16 32 C2 "--frames-per-step 256"
run 2 12 0 C2 "--frames-per-step 256"
run 2 16 0 C2urea "--frames-per-step 256"
run 2 16 32 C2urea "--frames-per-step 256"
run 2 16 0 C4 ""
run 2 8 0 C4 ""
run 3 8 0 C4 ""
