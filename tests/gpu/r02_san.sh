#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the round-2 kernels: batched grid path with the x-limited-ring search, chained
# scans, molecule-pair path, device finalresults / contributions, native DCD feed, device XTC decode
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/san3.py <<'PY'
import sys, os, numpy as np, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cmx_b200 as cm
from cmx_b200.engine import DcdFile, Engine, XtcFile
from common import Problem, namd, write_dcd, assert_counters_equal
import __graft_entry__ as g
g.smoke()
d = namd()
TMAO = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)
PROT = cm.AtomSelection(np.arange(1, 1464), nmols=1)
o = cm.Options(bulk_range=(8.0, 10.0), n_random_samples=3, silent=True)
p = Problem(PROT, TMAO, o, d["protein"], d["tmao"], d["cells"], weights=[1.0, 2.0, 1.0])
eng = p.engine(n_streams=2, batch_frames=2); dev = p.run_engine(eng)
orc_, _ = p.oracle(); assert_counters_equal(dev, orc_)
fin = eng.final_results()
groups = [np.arange(k, 1463, 7) for k in range(7)]
for t in ("mddf", "coordination_number", "md_count", "kbi"):
    eng.contributions("solute", groups, t)
assert np.allclose(eng.contributions("solute", [np.arange(1463)], "mddf")[0], fin["mddf"], rtol=1e-12)
eng.close()
if len(sys.argv) > 1:
    pa = Problem(TMAO, TMAO, o, d["tmao"][:2], None, d["cells"][:2], autocorrelation=True)
    eng = pa.engine(path=2, n_streams=2); dev = pa.run_engine(eng); eng.close()
    orc_, _ = pa.oracle(); assert_counters_equal(dev, orc_)
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    path = os.path.join(tempfile.gettempdir(), "san.dcd"); write_dcd(path, frames, d["cells"])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1); tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    eng = Engine(solute=sol, solvent=tm, options=cm.Options(bulk_range=(8.0, 10.0), n_random_samples=2, silent=True), irefatom=1, autocorrelation=False, n_streams=2)
    f = DcdFile(path); eng.run_dcd(f, sol.indices, tm.indices, [0, 1, 2, 1], n_reader_threads=2); eng.finish(); f.close(); eng.close()
    x = XtcFile(os.path.join(os.getcwd(), "tests", "golden", "nucleic_frame0.xtc"))
    a = x.read_frame(0)[0]; b = x.read_frame_device(0)[0]; x.close()
    assert np.array_equal(a, b)
print("sanitizer workload ok")
PY
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san3.py all > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "== memcheck"; grep -E "ERROR SUMMARY|sanitizer workload ok|Invalid|Traceback|Error" gpurun_out/r02_sanitizer_memcheck.log | sort | uniq -c | head -8
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san3.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "== racecheck"; grep -E "RACECHECK SUMMARY|sanitizer workload ok|hazard|Traceback|Error" gpurun_out/r02_sanitizer_racecheck.log | sort | uniq -c | head -8
