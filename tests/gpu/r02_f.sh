#!/bin/bash
# sweep of the search geometry on C4 (and C2): query-cell size, row division, x cell size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { # cfg, env...
  cfg=$1; shift
  env "$@" CMX_TRACE=30:2 timeout 300 python bench.py --config $cfg --steps 3 --no-cpu-baseline --no-e2e --no-hbm-kernel > gpurun_out/r02f_tmp.json 2> gpurun_out/r02f_tmp.err
  python - "$cfg" "$*" <<'PY'
import json, sys, re
try:
    d=json.loads(open("gpurun_out/r02f_tmp.json").read().strip().splitlines()[-1])
    tr={m.group(2):float(m.group(1)) for m in re.finditer(r"\[cmx trace\]\s+([\d.]+) us\s+[\d.]+%\s+(\S+)", open("gpurun_out/r02f_tmp.err").read())}
    print(sys.argv[1], sys.argv[2], "| value", round(d["value"],1), "| search rand/real us per batch", tr.get("tile_search<rand>",0)/2, tr.get("tile_search<real>",0)/2, "| pair evals/frame %.3g" % d["roofline"]["pair_evals_per_frame"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e); print(open("gpurun_out/r02f_tmp.err").read()[-400:])
PY
}
for q in 3.2 4 5 6.5; do run C4 CMX_QSIDE=$q; done
for r in 2 4 5; do run C4 CMX_ROWDIV=$r; done
for x in 1.9 3.75 5; do run C4 CMX_XSIDE=$x; done
run C4 CMX_QSIDE=4 CMX_ROWDIV=4
run C4 CMX_QSIDE=4 CMX_ROWDIV=4 CMX_XSIDE=3.75
for q in 3.2 4 6.5; do run C2 CMX_QSIDE=$q; done
