#!/bin/bash
# round 2, third session: (1) GPU tests incl. the device finalresults/contributions, with the x-limited rings of the search;
# (2) A/B of the search variants (x-limited rings on/off, ring width) and of k_gen_rand variants (sincospi, occupancy)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s); LIMIT=${1:-1200}
left() { [ $(( $(date +%s) - T0 )) -lt $LIMIT ]; }
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
Q="--steps 5 --no-cpu-baseline --no-e2e --no-job --no-secondary --no-hbm-kernel"
run() { # cfg, label, env...
  cfg=$1; label=$2; shift; shift
  left || { echo "skip $cfg $label (time)"; return; }
  env "$@" CMX_TRACE=30:2 timeout 300 python bench.py --config $cfg $Q > gpurun_out/r02w_tmp.json 2> gpurun_out/r02w_tmp.err
  python - "$cfg" "$label $*" <<'PY'
import json, sys, re
try:
    d=json.loads(open("gpurun_out/r02w_tmp.json").read().strip().splitlines()[-1])
    tr={m.group(2):float(m.group(1)) for m in re.finditer(r"\[cmx trace\]\s+([\d.]+) us\s+[\d.]+%\s+(\S+)", open("gpurun_out/r02w_tmp.err").read())}
    r=d["roofline"]
    print(sys.argv[1], sys.argv[2], "| value", round(d["value"],1), "| search rand/real ms per frame", round(r["kernel_ms_per_frame"],4), "| pair evals/frame %.4g" % r["pair_evals_per_frame"], "| gen_rand/fin_rand us per batch", tr.get("gen_rand",0)/2, tr.get("finalise<rand>",0)/2, "| deferred", round(d["config"]["deferred_to_exact_per_frame"],1), flush=True)
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e); print(open("gpurun_out/r02w_tmp.err").read()[-600:])
PY
}
rebuild() { CMX_NVCC_EXTRA="$1 -Xptxas -v" python -c "
import sys; sys.path.insert(0,'.')
from cmx_b200 import engine; engine.build(force=True)" > gpurun_out/r02w_build.log 2>&1 || tail -5 gpurun_out/r02w_build.log; }
run C4 xring1
run C2 xring1
run C4 xring1 CMX_RING=1.5
run C4 xring1 CMX_RING=3.5
run C2 xring1 CMX_RING=3.5
run C4 xring1 CMX_QSIDE=6
run C5 xring1
rebuild "-DCMX_XRING=0"
run C4 xring0
run C2 xring0
run C5 xring0
rebuild "-DCMX_SINCOSPI"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
run C4 sincospi
run C2 sincospi
rebuild "-DCMX_SINCOSPI -DCMX_GEN_MINBLOCKS=6"
run C4 sincospi_mb6
run C2 sincospi_mb6
rebuild "-DCMX_SINCOSPI -DCMX_GEN_MINBLOCKS=8"
run C4 sincospi_mb8
rebuild ""
echo "elapsed $(( $(date +%s) - T0 )) s"
