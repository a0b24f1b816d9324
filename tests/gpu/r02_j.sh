#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "namd or synthetic or custom or weights or toy or c2 or edge or cutoffs" 2>&1 | tail -5
run() { # cfg, env...
  cfg=$1; shift
  env "$@" CMX_TRACE=30:2 timeout 300 python bench.py --config $cfg --steps 3 --no-cpu-baseline --no-e2e --no-hbm-kernel > gpurun_out/r02j_tmp.json 2> gpurun_out/r02j_tmp.err
  python - "$cfg" "$*" <<'PY'
import json, sys, re
try:
    d=json.loads(open("gpurun_out/r02j_tmp.json").read().strip().splitlines()[-1])
    tr={m.group(2):float(m.group(1)) for m in re.finditer(r"\[cmx trace\]\s+([\d.]+) us\s+[\d.]+%\s+(\S+)", open("gpurun_out/r02j_tmp.err").read())}
    print(sys.argv[1], sys.argv[2], "| value", round(d["value"],1), "| search rand/real us per batch", tr.get("tile_search<rand>",0)/2, tr.get("tile_search<real>",0)/2, "| pair evals/frame %.3g" % d["roofline"]["pair_evals_per_frame"], "| deferred", round(d["config"]["deferred_to_exact_per_frame"],1))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e); print(open("gpurun_out/r02j_tmp.err").read()[-400:])
PY
}
run C4 CMX_RING=2.5
run C4 CMX_RING=1.5
run C4 CMX_RING=4
run C4 CMX_RING=2.5 CMX_ROWDIV=4
run C4 CMX_RING=2.5 CMX_ROWDIV=5
run C4 CMX_RING=2.5 CMX_ROWDIV=6
run C4 CMX_RING=2.5 CMX_ROWDIV=4 CMX_QSIDE=4
run C4 CMX_RING=2.5 CMX_ROWDIV=4 CMX_QSIDE=6.5
run C4 CMX_RING=2.5 CMX_ROWDIV=4 CMX_XSIDE=1.9
run C2 CMX_RING=2.5
run C2 CMX_RING=2.5 CMX_ROWDIV=4
run C2 CMX_RING=2.5 CMX_ROWDIV=5
