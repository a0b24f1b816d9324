#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
CMX_TRACE=100000:0 timeout 600 python bench.py --steps 5 --no-hbm-kernel > gpurun_out/r02q.json 2> gpurun_out/r02q.err; grep "create:" gpurun_out/r02q.err | tail -4
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02q.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "job", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["job"].items() if k!="what"})
print("secondary", round(d["secondary"]["value"],1), round(d["secondary"]["e2e"]["value"],1))
PY
