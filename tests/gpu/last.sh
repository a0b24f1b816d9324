#!/bin/bash
# last validation of the round: parity suite after the k_scan_small barrier fix, racecheck of the smoke frame
# (scan, tile queue, fused transform), memcheck of the native feed + group reduction
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 100 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
cat > /tmp/san2.py <<'PY'
import sys, os, numpy as np, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cmx_b200 as cm
from cmx_b200.engine import DcdFile, Engine
from common import namd, write_dcd
import __graft_entry__ as g
g.smoke()
if len(sys.argv) > 1:
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    path = os.path.join(tempfile.gettempdir(), "san.dcd"); write_dcd(path, frames, d["cells"])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1); tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    eng = Engine(solute=sol, solvent=tm, options=cm.Options(bulk_range=(8.0, 10.0), n_random_samples=2, silent=True), irefatom=1, autocorrelation=False, n_streams=2)
    f = DcdFile(path); eng.run_dcd(f, sol.indices, tm.indices, [0, 1, 2, 1], n_reader_threads=2)
    r = eng.reduce_groups("solute_group_count", [np.arange(0, 1463, 2), np.arange(1, 1463, 2)])
    c = eng.finish(); assert np.array_equal(r.sum(axis=0), c["md_count"]); f.close(); eng.close()
print("sanitizer workload ok")
PY
timeout 70 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san2.py > gpurun_out/sanitizer2_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|sanitizer workload ok|hazard" gpurun_out/sanitizer2_racecheck.log | sort | uniq -c | head -5
timeout 50 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san2.py feed > gpurun_out/sanitizer2_memcheck.log 2>&1
grep -E "ERROR SUMMARY|sanitizer workload ok|Invalid" gpurun_out/sanitizer2_memcheck.log | sort | uniq -c | head -5
