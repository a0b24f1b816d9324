#!/bin/bash
# round 2, third session: full GPU suite with the x-limited rings + bulk-copy tile prefetch; A/B of the prefetch, ring width,
# batches in flight / frames per batch, search blocks per SM; ncu of the new search kernel (C4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s); LIMIT=${1:-1200}
left() { [ $(( $(date +%s) - T0 )) -lt $LIMIT ]; }
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02x_pytest.log 2>&1; tail -4 gpurun_out/r02x_pytest.log; grep -E "^(FAILED|ERROR)|Max (abs|rel)|Mismatch" gpurun_out/r02x_pytest.log | head -20
Q="--steps 5 --no-cpu-baseline --no-e2e --no-job --no-secondary --no-hbm-kernel"
run() { # cfg, label, [bench args --] env...
  cfg=$1; label=$2; shift; shift
  extra=""; while [ "$1" = "--arg" ]; do extra="$extra $2"; shift; shift; done
  left || { echo "skip $cfg $label (time)"; return; }
  env "$@" timeout 300 python bench.py --config $cfg $Q $extra > gpurun_out/r02x_tmp.json 2> gpurun_out/r02x_tmp.err
  python - "$cfg" "$label $extra $*" <<'PY'
import json, sys
try:
    d=json.loads(open("gpurun_out/r02x_tmp.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[1], sys.argv[2], "| value", round(d["value"],1), "| search ms/frame", round(r["kernel_ms_per_frame"],4), "share", round(r["kernel_share_of_frame"],3), "| pair evals/frame %.4g" % r["pair_evals_per_frame"], "| launches/frame", round(d["launches_per_frame"],2), "| submit/wait", round(d["host_submit_ms_per_step"],1), round(d["host_wait_ms_per_step"],1), "of", round(d["ms_per_step"],1), flush=True)
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e); print(open("gpurun_out/r02x_tmp.err").read()[-600:])
PY
}
rebuild() { CMX_NVCC_EXTRA="$1" python -c "
import sys; sys.path.insert(0,'.')
from cmx_b200 import engine; engine.build(force=True)" > gpurun_out/r02x_build.log 2>&1 || tail -5 gpurun_out/r02x_build.log; }
run C4 tma1
run C2 tma1
run C5 tma1
run C4 tma1 CMX_RING=4.5
run C4 tma1 CMX_RING=5.5
run C2 tma1 CMX_RING=4.5
run C4 tma1 --arg "--streams 4"
run C4 tma1 --arg "--streams 6"
run C4 tma1 --arg "--batch 8"
run C4 tma1 --arg "--batch 2" --arg "--streams 6"
run C4 tma1 CMX_SEARCH_BLOCKS_PER_SM=4
run C4 tma1 CMX_SEARCH_BLOCKS_PER_SM=1
run C2 tma1 CMX_SEARCH_BLOCKS_PER_SM=4
LL="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-job --no-secondary --no-hbm-kernel --streams 1"
left && { timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 10 -c 2 -f -o gpurun_out/r02x_prof_search_C4 \
   python bench.py --config C4 $LL --frames-per-step 16 > gpurun_out/r02x_ncu_full_C4.log 2>&1; tail -1 gpurun_out/r02x_ncu_full_C4.log | cut -c1-120; }
rebuild "-DCMX_TILE_TMA=0"
run C4 tma0
run C2 tma0
rebuild ""
echo "elapsed $(( $(date +%s) - T0 )) s"
