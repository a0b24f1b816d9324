#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
CMX_TRACE=100000:0 timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-hbm-kernel --no-secondary --no-e2e > gpurun_out/r02p.json 2> gpurun_out/r02p.err; grep "create:" gpurun_out/r02p.err | tail -12
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02p.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "job", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["job"].items() if k!="what"})
PY
