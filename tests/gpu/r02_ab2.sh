#!/bin/bash
# A/B: x-limited rings in the random phase only (default) vs both phases; ring width 3.5 vs 4.5.  Parity suite first.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s); LIMIT=${1:-600}
left() { [ $(( $(date +%s) - T0 )) -lt $LIMIT ]; }
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3
Q="--steps 8 --no-cpu-baseline --no-e2e --no-job --no-secondary --no-hbm-kernel"
run() { cfg=$1; label=$2; shift; shift
  left || { echo "skip $cfg $label (time)"; return; }
  env "$@" timeout 300 python bench.py --config $cfg $Q > gpurun_out/r02ab_tmp.json 2> gpurun_out/r02ab_tmp.err
  python - "$cfg" "$label $*" <<'PY'
import json, sys
try:
    d=json.loads(open("gpurun_out/r02ab_tmp.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], sys.argv[2], "| value", round(d["value"],1), "| rand search ms/frame", round(r["kernel_ms_per_frame"],4), "share", round(r["kernel_share_of_frame"],3), "| pair evals/frame %.4g" % r["pair_evals_per_frame"], flush=True)
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e); print(open("gpurun_out/r02ab_tmp.err").read()[-600:])
PY
}
rebuild() { CMX_NVCC_EXTRA="$1" python -c "
import sys; sys.path.insert(0,'.')
from cmx_b200 import engine; engine.build(force=True)" > gpurun_out/r02ab_build.log 2>&1 || tail -5 gpurun_out/r02ab_build.log; }
run C4 randonly
run C2 randonly
run C4 randonly CMX_RING=4.5
run C2 randonly CMX_RING=4.5
run C4 randonly
run C2 randonly
run C4 randonly CMX_RING=4.5
run C2 randonly CMX_RING=4.5
rebuild "-DCMX_XRING_REAL=1"
run C4 both
run C2 both
run C4 both CMX_RING=4.5
run C2 both CMX_RING=4.5
rebuild ""
echo "elapsed $(( $(date +%s) - T0 )) s"
