import os, sys, time, tempfile, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cmx_b200 as cm
from cmx_b200 import synthetic as syn
from cmx_b200.engine import DcdFile, Engine
from common import write_dcd
s = syn.config_c2(1.0)
sol, wat = s.selections["solute"], s.selections["water"]
opt = cm.Options(bulk_range=(10.0, 15.0), n_random_samples=10, seed=321, silent=True, irefatom=1)
nf = 256
path = os.path.join(tempfile.mkdtemp(), "c2.dcd")
frames = np.stack([s.frame(k + 1)[0] for k in range(nf)]).astype(np.float32)
write_dcd(path, frames, np.asarray(s.cell, dtype=np.float64))
def T(label, t0): print("%-40s %8.1f ms" % (label, 1e3 * (time.perf_counter() - t0)), flush=True)
for rep in range(3):
    print("---- rep", rep)
    t0 = time.perf_counter(); eng = Engine(solute=sol, solvent=wat, options=opt, irefatom=1, autocorrelation=False); T("create", t0)
    f = DcdFile(path)
    t0 = time.perf_counter(); eng.run_dcd(f, sol.indices, wat.indices, list(range(64)), n_reader_threads=2); eng.sync(); T("run 64 frames (first)", t0)
    t0 = time.perf_counter(); eng.run_dcd(f, sol.indices, wat.indices, list(range(nf)), n_reader_threads=2); eng.sync(); T("run 256 frames (second)", t0)
    t0 = time.perf_counter(); eng.run_dcd(f, sol.indices, wat.indices, list(range(nf)), n_reader_threads=4); eng.sync(); T("run 256 frames (third, 4 threads)", t0)
    t0 = time.perf_counter(); c = eng.finish(copy=False); T("finish", t0)
    t0 = time.perf_counter(); c = eng.finish(copy=False); T("finish again", t0)
    t0 = time.perf_counter(); f.close(); eng.close(); T("close", t0)
t0 = time.perf_counter(); R = cm.mddf(path, sol, wat, cm.Options(bulk_range=(10.0, 15.0), n_random_samples=10, seed=321, silent=True, irefatom=1, lastframe=64)); T("public mddf 64 frames", t0)
t0 = time.perf_counter(); R = cm.mddf(path, sol, wat, cm.Options(bulk_range=(10.0, 15.0), n_random_samples=10, seed=321, silent=True, irefatom=1, lastframe=64)); T("public mddf 64 frames again", t0)
