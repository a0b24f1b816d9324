#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
cat > /tmp/san.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cmx_b200 as cm
from common import Problem, namd, assert_counters_equal
import __graft_entry__ as g
g.smoke()
d = namd()
TMAO = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)
PROT = cm.AtomSelection(np.arange(1, 1464), nmols=1)
o = cm.Options(bulk_range=(8.0, 10.0), n_random_samples=3, silent=True)
for path in (2, 1):
    p = Problem(TMAO, TMAO, o, d["tmao"][:2], None, d["cells"][:2], autocorrelation=True)
    eng = p.engine(path=path, n_streams=2); dev = p.run_engine(eng); eng.close()
    orc_, _ = p.oracle(); assert_counters_equal(dev, orc_)
p = Problem(PROT, TMAO, o, d["protein"], d["tmao"], d["cells"], weights=[1.0, 2.0, 1.0])
eng = p.engine(n_streams=3); dev = p.run_engine(eng); eng.close()
orc_, _ = p.oracle(); assert_counters_equal(dev, orc_)
print("sanitizer workload ok")
PY
for TOOL in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $TOOL --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_$TOOL.log 2>&1
  echo "== $TOOL: $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_$TOOL.log) summaries"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload ok|Error|hazard" gpurun_out/sanitizer_$TOOL.log | sort | uniq -c | head -8
done
