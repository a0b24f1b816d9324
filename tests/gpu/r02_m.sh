#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
run() { # cfg
  CMX_TRACE=30:2 timeout 300 python bench.py --config $1 --steps 3 --no-cpu-baseline --no-e2e --no-hbm-kernel > gpurun_out/r02m_tmp.json 2> gpurun_out/r02m_tmp.err
  python - "$1" "$2" <<'PY'
import json, sys, re
try:
    d=json.loads(open("gpurun_out/r02m_tmp.json").read().strip().splitlines()[-1])
    tr={m.group(2):float(m.group(1)) for m in re.finditer(r"\[cmx trace\]\s+([\d.]+) us\s+[\d.]+%\s+(\S+)", open("gpurun_out/r02m_tmp.err").read())}
    print(sys.argv[1], sys.argv[2], "| value", round(d["value"],1), "| search rand/real us per batch", tr.get("tile_search<rand>",0)/2, tr.get("tile_search<real>",0)/2)
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e); print(open("gpurun_out/r02m_tmp.err").read()[-400:])
PY
}
for v in "-DCMX_SEARCH_MINBLOCKS=4 -DCMX_STAGE=256 -DCMX_SWEEP_UNROLL=2" "-DCMX_SEARCH_MINBLOCKS=4 -DCMX_STAGE=256 -DCMX_SWEEP_UNROLL=4" "-DCMX_SEARCH_MINBLOCKS=3 -DCMX_STAGE=384 -DCMX_SWEEP_UNROLL=2" "-DCMX_SEARCH_MINBLOCKS=2 -DCMX_STAGE=512 -DCMX_SWEEP_UNROLL=4"; do
  CMX_NVCC_EXTRA="$v -Xptxas -v" python -c "
import sys; sys.path.insert(0,'.')
from cmx_b200 import engine; engine.build(force=True)" 2>&1 | grep -A2 "tile_searchILb0ELb1" | grep -E "Used|spill" | head -2
  run C4 "$v"; run C2 "$v"
done
