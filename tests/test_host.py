"""CPU tests of the host-side mirrors and of the C-ABI library (symbols only: no GPU calls)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import cmx_b200 as cm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_options_defaults_and_errors():
    """src/Options.jl:194-220 (testitem "Options")."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = cm.Options()
    assert (o.dbulk, o.cutoff, o.usecutoff) == (10.0, 10.0, False)
    o = cm.Options(bulk_range=(10.0, 14.0))
    assert (o.dbulk, o.cutoff, o.usecutoff) == (10.0, 14.0, True)
    o = cm.Options(dbulk=8.0, usecutoff=True, silent=True)
    assert o.cutoff == 12.0
    for bad in (dict(stride=0), dict(firstframe=3, lastframe=2), dict(n_random_samples=0), dict(bulk_range=(8.0, 12.0), dbulk=8.0),
                dict(bulk_range=(12.0, 8.0)), dict(dbulk=8.0, cutoff=12.0), dict(bulk_range=(8.0, 12.01)), dict(bulk_range=(1, 2, 3))):
        with pytest.raises(ValueError):
            cm.Options(silent=True, **bad)


def test_atom_selection_and_group_csr():
    s = cm.AtomSelection([1, 2, 3, 4, 5, 6], natomspermol=3)
    assert (s.nmols, s.natomspermol, s.n_groups, s.custom_groups) == (2, 3, 3, False)
    assert s.group_csr() == (None, None)
    with pytest.raises(ValueError):
        cm.AtomSelection([1, 2, 3, 4], natomspermol=3)
    with pytest.raises(ValueError):
        cm.AtomSelection([1, 2, 3], nmols=1, group_atom_indices=[[1, 1]], group_names=["a"])
    with pytest.raises(ValueError):
        cm.AtomSelection([1, 2, 3], nmols=1, group_atom_indices=[[7]], group_names=["a"])
    g = cm.AtomSelection([10, 11, 12, 13], nmols=1, group_atom_indices=[[10, 12], [12, 13], [11]], group_names=["a", "b", "c"])
    off, ids = g.group_csr()
    assert off.tolist() == [0, 1, 2, 4, 5] and ids.tolist() == [0, 2, 0, 1, 1]
    assert g.n_groups == 3


def test_frame_selection_and_sharding():
    from cmx_b200.driver import frames_to_compute, shard
    o = cm.Options(firstframe=2, lastframe=9, stride=3, silent=True)
    fr = frames_to_compute(o, 9, [1, 1, 1, 1, 0, 1, 1, 2.5, 1])
    assert fr == [(2, 1.0), (8, 2.5)]                   # frame 5 has zero weight -> skipped (src/mddf.jl:102)
    todo = [(k, 1.0) for k in range(1, 12)]
    parts = [shard(todo, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == todo and max(map(len, parts)) - min(map(len, parts)) <= 1


def test_cell_conventions():
    m = cm.cell_from_lengths_angles(10, 11, 12, 90, 90, 90)
    assert np.array_equal(m, np.diag([10.0, 11.0, 12.0]))
    m = cm.cell_from_lengths_angles(10, 11, 12, 0, 0, 0)       # DCD without angles (NamdDCD.jl:182-186)
    assert np.array_equal(m, np.diag([10.0, 11.0, 12.0]))
    m = cm.cell_from_lengths_angles(10, 11, 12, 80, 85, 70)
    a, b, c = m[:, 0], m[:, 1], m[:, 2]
    ang = lambda u, v: np.degrees(np.arccos(u @ v / np.linalg.norm(u) / np.linalg.norm(v)))
    assert np.allclose([np.linalg.norm(a), np.linalg.norm(b), np.linalg.norm(c)], [10, 11, 12])
    assert np.allclose([ang(b, c), ang(a, c), ang(a, b)], [80, 85, 70])
    from cmx_b200.engine import cell_to_c
    assert cell_to_c(m).reshape(3, 3)[0].tolist() == a.tolist()   # column-major: first 3 doubles = first lattice vector


def test_result_json_roundtrip_and_finalresults(tmp_path):
    """save/load keep the reference's JSON schema; finalresults == the oracle's independent restatement."""
    from oracle import cmx_oracle as orc
    from common import Problem, namd
    d = namd()
    protein = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tmao = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)
    opt = cm.Options(bulk_range=(8.0, 10.0), seed=321, silent=True, n_random_samples=5)
    p = Problem(protein, tmao, opt, d["protein"], d["tmao"], d["cells"])
    o, _ = p.oracle()
    tr = cm.ArrayTrajectory(np.concatenate([d["protein"], d["tmao"]], axis=1), d["cells"],
                            cm.AtomSelection(np.arange(1, 1464), nmols=1), cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14))
    meta = cm.trajectory_metadata(tr, opt)
    assert meta.irefatom == 1 and meta.nframes_read == 3
    from cmx_b200.results import new_result
    R = new_result(tr, opt, meta)
    c = o.counters()
    for k, v in c.items():
        if k != "volume_total":
            setattr(R, k, v.copy())
    R.volume.total = c["volume_total"]
    R = cm.finalresults(R, opt)
    f = orc.finalresults(c, nmols_solute=1, nmols_solvent=181, autocorrelation=False, n_random_samples=5, binstep=0.02,
                         dbulk=8.0, cutoff=10.0, usecutoff=True, Q=3.0)
    for a, b in ((R.mddf, f.mddf), (R.kb, f.kb), (R.rdf, f.rdf), (R.kb_rdf, f.kb_rdf), (R.coordination_number, f.coordination_number),
                 (R.volume.shell, f.volume_shell), (R.md_count_random, f.md_count_random)):
        assert np.allclose(a, b, rtol=1e-13, atol=0)
    assert np.isclose(R.volume.bulk, f.volume_bulk) and np.isclose(R.density.solvent_bulk, f.density_solvent_bulk)
    # contributions: solute rows and solvent rows both sum to the total (src/tools/contributions.jl:320-348)
    # (atom-index selections of a multi-molecule selection are per molecule: contributions.jl:186-193)
    tot = sum(cm.contributions(R, cm.SolventGroup([int(i)])) for i in R.solvent.indices[:14]) * R.solvent.nmols
    assert np.allclose(tot, R.mddf, rtol=1e-12)
    tot = sum(cm.contributions(R, cm.SoluteGroup([int(i)]), type="md_count") for i in R.solute.indices[:50])
    assert np.allclose(tot, R.solute_group_count[:50].sum(axis=0), rtol=1e-12)
    assert np.allclose(np.cumsum(R.solute_group_count.sum(axis=0)), R.coordination_number)
    fn = cm.save(R, str(tmp_path / "r.json"))
    R2 = cm.load(fn)
    assert np.array_equal(R2.mddf, R.mddf) and R2.options.cutoff == 10.0 and R2.solvent.natomspermol == 14
    assert np.array_equal(R2.solute_group_count, R.solute_group_count)


def test_loads_reference_json_schema():
    """a JSON with exactly the reference's fields (as in test/data/NAMD/tmao_tmao.json) loads."""
    from common import kat
    g = kat()["golden_json_sums"]["tmao_tmao"]
    assert g["nbins"] == 500 and g["solute_nmols"] == 181 and g["solute_first_index"] == 1479


def test_cabi_exports_every_declared_symbol():
    """libcmx_b200.so loads and exports every function include/cmx_b200.h declares (no compute call)."""
    from cmx_b200 import engine
    hdr = open(os.path.join(ROOT, "include", "cmx_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cmx_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    engine.build()
    lib = ctypes.CDLL(engine.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(engine.EXPORTS) == declared
    lib.cmx_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.cmx_version()
    # struct layouts agree with the header (compiled with the host compiler)
    src = '#include "cmx_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu", sizeof(cmx_config), sizeof(cmx_counters), sizeof(cmx_md), sizeof(cmx_stats), sizeof(cmx_dcd_info), sizeof(cmx_xtc_info), sizeof(cmx_final));}'
    exe = os.path.join("/tmp", f"cmx_sz_{os.getpid()}")
    subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    sizes = list(map(int, subprocess.run([exe], capture_output=True, check=True).stdout.split()))
    os.remove(exe)
    assert sizes == [ctypes.sizeof(engine.CmxConfig), ctypes.sizeof(engine.CmxCounters), engine.MD_DTYPE.itemsize, ctypes.sizeof(engine.CmxStats),
                     ctypes.sizeof(engine.CmxDcdInfo), ctypes.sizeof(engine.CmxXtcInfo), ctypes.sizeof(engine.CmxFinal)]


def test_julia_shim_struct_layouts_match_the_abi():
    """julia/CMXB200.jl cannot be executed here (no Julia): its `struct` mirrors are parsed instead and their C layout
    (field order, names, sizes, natural alignment) compared with the ctypes mirrors, which are checked against the header
    by the test above."""
    from cmx_b200 import engine
    jl = open(os.path.join(ROOT, "julia", "CMXB200.jl")).read()
    size_of = {"Int32": 4, "UInt32": 4, "Int64": 8, "UInt64": 8, "Float64": 8, "Float32": 4}

    def layout(name):
        body = re.search(r"struct " + name + r"\n(.*?)\nend", jl, re.S).group(1)
        fields = []
        for line in body.splitlines():
            line = line.split("#")[0]
            for m in re.finditer(r"(\w+)::([\w{}]+)", line):
                fields.append((m.group(1), 8 if m.group(2).startswith("Ptr") else size_of[m.group(2)]))
        off, out = 0, []
        for fname, sz in fields:
            off = (off + sz - 1) // sz * sz
            out.append((fname, off, sz))
            off += sz
        return out, (off + 7) // 8 * 8

    for jname, cstruct in (("CmxConfig", engine.CmxConfig), ("CmxCounters", engine.CmxCounters), ("CmxFinal", engine.CmxFinal)):
        fields, total = layout(jname)
        assert total == ctypes.sizeof(cstruct), jname
        assert [f[0] for f in fields] == [f[0] for f in cstruct._fields_], jname
        for (fname, off, sz), (cname, ctype) in zip(fields, cstruct._fields_):
            assert off == getattr(cstruct, cname).offset and sz == ctypes.sizeof(ctype), (jname, fname)


def test_product_has_no_cpu_fallback_and_never_imports_oracle():
    """the product package must not reference oracle/ and must fail loudly without its CUDA library."""
    pkg = os.path.join(ROOT, "complexmixtures.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "cmx_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
    from cmx_b200 import engine
    with pytest.raises(RuntimeError):
        saved, engine._lib = engine._lib, None
        try:
            engine.load_library("/nonexistent/libcmx_b200.so")
        finally:
            engine._lib = saved


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import cmx_b200 as cm
from cmx_b200.driver import frames_to_compute, shard, weights_agree
from common import namd
from oracle import cmx_oracle as orc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = namd()
protein = cm.AtomSelection(np.arange(1, 1464), nmols=1); tmao = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)
opt = cm.Options(bulk_range=(8.0, 10.0), seed=321, silent=True, n_random_samples=3)

def allreduce_min(values):
    t = torch.tensor(values, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MIN); return t.tolist()

# frames 1, 2, 3, 2 (4 computed frames), three weight patterns: all equal; equal within each rank but different between
# the ranks (the case that silently gave w_rank * (cnt0 + cnt1)); different within a rank
for case, weights in (("uniform", [1.0, 1.0, 1.0, 1.0]), ("per-rank", [1.0, 2.0, 1.0, 2.0]), ("mixed", [0.5, 2.0, 1.0, 2.0])):
    frames = [1, 2, 3, 2]
    todo = list(zip(frames, weights))
    mine = shard(todo, rank, world)
    # the driver's decision, taken collectively BEFORE the exchange: every rank must reach the same branch
    agree = weights_agree([w for _, w in mine], allreduce_min)
    flags = torch.tensor([1.0 if agree else 0.0]); lo = flags.clone(); hi = flags.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(lo) == float(hi), "ranks disagree on the exchange branch"
    assert agree == (case == "uniform"), (case, agree)
    # the per-rank engine of this CPU test is the oracle; integers (uniform) or f64 with the weights applied (otherwise)
    o = orc.Oracle.from_problem(protein, tmao, opt, 1, False)
    for f, w in mine:
        o.frame(d["protein"][f - 1], d["tmao"][f - 1], d["cells"][f - 1], weight=(1.0 if agree else w), frame_index=f)
    c = o.counters()
    keys = [k for k in sorted(c) if k != "volume_total"]
    flat = torch.from_numpy(np.concatenate([c[k].ravel() for k in keys]).astype(np.int64 if agree else np.float64))
    dist.all_reduce(flat)                                              # the single exchange step
    if rank == 0:
        ref = orc.Oracle.from_problem(protein, tmao, opt, 1, False)
        for f, w in todo:
            ref.frame(d["protein"][f - 1], d["tmao"][f - 1], d["cells"][f - 1], weight=w, frame_index=f)
        r = ref.counters()
        want = np.concatenate([r[k].ravel() for k in keys])
        got = flat.numpy().astype(np.float64) * (weights[0] if agree else 1.0)
        assert np.array_equal(got, want), case + ": sharded + all-reduced counters differ from the single-rank run"
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_frame_sharding_allreduce_gloo_world2(tmp_path):
    """N>1 host logic on CPU (gloo, world size 2): frames dealt round-robin, the driver's COLLECTIVE decision between
    the integer exchange (all weights one number) and the f64 exchange (weights differ between or within ranks --
    the per-rank case used to return w_rank * (cnt0 + cnt1)), and the sum equal to the single-rank run; the Philox
    stream is keyed by the global frame index so the random phase does not depend on the sharding.  (The engine itself
    is run sharded against single-device in the -m gpu suite: tests/test_gpu_multi.py.)"""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_dcd_reader_roundtrip(tmp_path):
    """NamdDCD reader (src/trajectory_formats/NamdDCD.jl:141-188): records, selection gather, unit cell."""
    from common import namd, write_dcd
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    path = str(tmp_path / "t.dcd")
    tri = np.array([[40.0, 8.0, 5.0], [0.0, 38.0, 7.0], [0.0, 0.0, 36.0]])
    write_dcd(path, frames, [d["cells"][0], tri, d["cells"][2]])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    t = cm.make_trajectory(path, sol, tm)
    assert isinstance(t, cm.NamdDCD) and t.nframes == 3 and t.natoms_file == 3997
    t.open()
    for k in range(3):
        xs, xv = t.nextframe()
        assert np.array_equal(xs, d["protein"][k]) and np.array_equal(xv, d["tmao"][k])
        want = tri if k == 1 else d["cells"][k]
        assert np.allclose(t.getunitcell(), want, rtol=1e-12, atol=1e-9)
    t.close()
    meta = cm.trajectory_metadata(t, cm.Options(silent=True, lastframe=2))
    assert meta.irefatom == 1 and meta.lastframe_read == 2 and meta.n_groups_solute == 1463 and meta.n_groups_solvent == 14
    with pytest.raises(ValueError):
        cm.trajectory_metadata(t, cm.Options(silent=True, lastframe=7))
    with pytest.raises(ValueError):
        cm.trajectory_metadata(t, cm.Options(silent=True, irefatom=15))


def test_native_dcd_reader_matches_host_reader(tmp_path):
    """cmx_dcd_* (pure host code of the library, no GPU): header parsing, frame count from the file size,
    X/Y/Z records and the unit cell agree with the NamdDCD mirror of src/trajectory_formats/NamdDCD.jl:141-230."""
    from common import namd, write_dcd
    from cmx_b200.engine import CmxError, DcdFile
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    path = str(tmp_path / "t.dcd")
    tri = np.array([[40.0, 8.0, 5.0], [0.0, 38.0, 7.0], [0.0, 0.0, 36.0]])
    cells = [d["cells"][0], tri, d["cells"][2]]
    write_dcd(path, frames, cells)
    f = DcdFile(path)
    assert (f.natoms, f.nframes) == (3997, 3) and f.frame_bytes == 56 + 3 * (8 + 4 * 3997)
    sel = cm.AtomSelection(np.arange(1, 3998), nmols=1)
    t = cm.make_trajectory(path, sel, sel)
    t.open()
    for k in (0, 1, 2):
        x, cell = f.read_frame(k)
        xs, _ = t.nextframe()
        assert np.array_equal(x, frames[k]) and np.array_equal(x, xs)
        assert np.allclose(cell, t.getunitcell(), rtol=1e-15, atol=1e-12)
        if k != 1:
            assert np.array_equal(cell, np.asarray(cells[k]))          # orthorhombic: exact
    t.close()
    with pytest.raises(CmxError):
        f.read_frame(3)
    f.close()
    # trailing partial frame is ignored (frame count from the size), garbage and missing files are errors
    with open(path, "ab") as fh:
        fh.write(b"\0" * 100)
    g = DcdFile(path); assert g.nframes == 3; g.close()
    bad = tmp_path / "bad.dcd"; bad.write_bytes(b"\x54\0\0\0XXXX" + b"\0" * 200)
    with pytest.raises(CmxError):
        DcdFile(str(bad))
    with pytest.raises(CmxError):
        DcdFile(str(tmp_path / "missing.dcd"))


def test_merge_weights_and_errors():
    """merge (src/tools/merge.jl:10-148) on hand-made Results: frame-weighted averages, file weights, error paths."""
    from cmx_b200.results import Result, TrajectoryFileOptions, merge, sum_frame_weights
    sol = cm.AtomSelection([1, 2, 3], nmols=1); solv = cm.AtomSelection(np.arange(4, 10), natomspermol=3)

    def mk(nframes, fw, fill, **okw):
        o = cm.Options(silent=True, **okw)
        nb = cm.setbin(o.cutoff, o.binstep)
        R = Result(nbins=nb, dbulk=o.dbulk, cutoff=o.cutoff, autocorrelation=False, solute=sol, solvent=solv,
                   files=[TrajectoryFileOptions("f.dcd", o, 1, nframes, nframes, np.asarray(fw, dtype=float))])
        R.md_count[:] = fill; R.mddf[:] = 2 * fill; R.solute_group_count[:] = fill; R.volume.total = 100.0 * fill
        R.density.solvent_bulk = fill
        return R
    A, B = mk(2, [1.0, 1.0], 1.0), mk(6, [1.0] * 6, 3.0)
    M = merge([A, B])
    assert M.weights == [0.25, 0.75] and len(M.files) == 2
    assert np.allclose(M.md_count, 0.25 * 1 + 0.75 * 3) and np.allclose(M.mddf, 2 * 2.5) and np.isclose(M.volume.total, 250.0)
    assert np.allclose(M.solute_group_count, 2.5) and np.isclose(M.density.solvent_bulk, 2.5)
    # custom frame weights change the data weights but not the file weights (merge.jl:112-121)
    Aw = mk(2, [3.0, 3.0], 1.0)
    Mw = merge([Aw, B])
    assert Mw.weights == [0.25, 0.75] and np.allclose(Mw.md_count, 0.5 * 1 + 0.5 * 3)
    assert sum_frame_weights(Mw) == 12.0
    with pytest.raises(ValueError, match="number of bins"):
        merge([A, mk(2, [1, 1], 1.0, binstep=0.05)])
    with pytest.raises(ValueError, match="cutoff distance"):
        merge([A, mk(2, [1, 1], 1.0, bulk_range=(4.0, 5.0), binstep=0.01)])
    other = mk(2, [1, 1], 1.0); other.solute = cm.AtomSelection([1, 2, 4], nmols=1)
    with pytest.raises(ValueError, match="selections"):
        merge([A, other])


def test_native_xtc_reader_small_uncompressed(tmp_path):
    """cmx_xtc_* on frames of <= 9 atoms (stored as plain floats by the format): frame index, units (nm -> A), box
    rows -> lattice vectors as matrix columns, selection gather of XTCTraj, truncated tail, garbage."""
    from common import write_xtc_small
    from cmx_b200.engine import CmxError, XtcFile
    rng = np.random.default_rng(2)
    fr = rng.uniform(0, 3, size=(4, 7, 3)).astype(np.float32)
    boxes = np.array([np.diag([3.0, 3.1, 3.2]), [[3.0, 0, 0], [0.4, 2.9, 0], [0.3, 0.5, 2.8]], np.diag([3.0, 3.0, 3.0]), np.diag([2.5, 3.0, 3.5])])
    path = str(tmp_path / "s.xtc")
    write_xtc_small(path, fr, boxes)
    x = XtcFile(path)
    assert (x.natoms, x.nframes) == (7, 4)
    for k in (2, 0, 3, 1):                                        # any order: frames are indexed
        xyz, cell, step, time = x.read_frame(k)
        assert np.array_equal(xyz, (fr[k].astype(np.float64) * 10.0).astype(np.float32)) and step == 10 * k and time == 2.0 * k
        assert np.allclose(cell, 10.0 * boxes[k].T, rtol=1e-7)    # columns = lattice vectors
    with pytest.raises(CmxError):
        x.read_frame(4)
    x.close()
    sel_a, sel_b = cm.AtomSelection([2, 5], nmols=1), cm.AtomSelection([1, 3, 4, 7], natomspermol=2)
    t = cm.make_trajectory(path, sel_a, sel_b)
    assert isinstance(t, cm.XTCTraj) and t.nframes == 4
    t.open()
    xs, xv = t.nextframe()
    assert np.allclose(xs, 10 * fr[0][[1, 4]], rtol=1e-6) and np.allclose(xv, 10 * fr[0][[0, 2, 3, 6]], rtol=1e-6)
    assert np.allclose(t.getunitcell(), 10.0 * boxes[0].T, rtol=1e-7)
    t.close()
    with open(path, "ab") as fh:
        fh.write(b"\x00\x00\x07\xcb" + b"\x00" * 20)              # a truncated extra frame is ignored
    y = XtcFile(path); assert y.nframes == 4; y.close()
    bad = tmp_path / "bad.xtc"; bad.write_bytes(b"\x00" * 300)
    with pytest.raises(CmxError):
        XtcFile(str(bad))
    with pytest.raises(ValueError):
        cm.make_trajectory(path, cm.AtomSelection([8], nmols=1), sel_b)


def test_native_xtc_reader_reference_fixture():
    """the compressed coordinate block, on the reference's own fixture test/data/nucleic/trajectory.xtc (present in
    the build container only): the decoder is pinned by physics -- the last 1000 molecules are TIP3P water and must
    come out rigid (O-H 0.9572 A, H-H 1.5139 A within the 0.01 A grid of the format) in every frame -- and against
    the committed summary tests/golden/xtc_nucleic.json."""
    from cmx_b200.engine import XtcFile
    src = "/root/reference/test/data/nucleic/trajectory.xtc"
    if not os.path.exists(src):
        pytest.skip("reference fixture not present on this machine")
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "xtc_nucleic.json")))
    x = XtcFile(src)
    assert (x.natoms, x.nframes) == (g["natoms"], g["nframes"]) == (95988, 6)
    for k, fr in enumerate(g["frames"]):
        xyz, cell, step, time = x.read_frame(k)
        w = xyz[-3000:].astype(np.float64).reshape(-1, 3, 3)
        oh1, oh2 = np.linalg.norm(w[:, 1] - w[:, 0], axis=1), np.linalg.norm(w[:, 2] - w[:, 0], axis=1)
        hh = np.linalg.norm(w[:, 2] - w[:, 1], axis=1)
        assert abs(oh1.mean() - 0.9572) < 2e-3 and abs(oh2.mean() - 0.9572) < 2e-3 and abs(hh.mean() - 1.5139) < 2e-3
        assert max(np.abs(oh1 - 0.9572).max(), np.abs(oh2 - 0.9572).max(), np.abs(hh - 1.5139).max()) < 0.03
        L = np.diag(cell)
        assert np.all(xyz.min(axis=0) > -0.05 * L) and np.all(xyz.max(axis=0) < 1.05 * L)
        assert step == fr["step"] and time == fr["time"] and np.array_equal(cell, np.array(fr["cell"]))
        assert np.array_equal(xyz[:4].astype(float), np.array(fr["first_atoms"])) and np.array_equal(xyz[-3:].astype(float), np.array(fr["last_atoms"]))
        assert np.allclose(np.sum(xyz.astype(np.float64), axis=0), fr["sum"], rtol=1e-12)
    x.close()


def test_native_xtc_reader_golden_frame():
    """the first frame of the reference's XTC fixture travels with the repository (tests/golden/nucleic_frame0.xtc, made by
    tests/golden/make_golden_xtc.py): the host decoder against the committed summary, on every machine."""
    from cmx_b200.engine import XtcFile
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "xtc_nucleic.json")))
    x = XtcFile(os.path.join(ROOT, "tests", "golden", "nucleic_frame0.xtc"))
    assert (x.natoms, x.nframes) == (95988, 1)
    xyz, cell, step, time = x.read_frame(0)
    x.close()
    fr = g["frames"][0]
    assert step == fr["step"] and time == fr["time"] and np.array_equal(cell, np.array(fr["cell"]))
    assert np.array_equal(xyz[:4].astype(float), np.array(fr["first_atoms"])) and np.array_equal(xyz[-3:].astype(float), np.array(fr["last_atoms"]))
    assert np.allclose(np.sum(xyz.astype(np.float64), axis=0), fr["sum"], rtol=1e-12)
    assert np.isclose(float(np.sum(xyz.astype(np.float64) ** 2)), fr["sum_sq"], rtol=1e-12)


def test_native_dcd_reader_reference_fixture():
    """the native DCD reader on the reference's own file test/data/NAMD/traj_duplicated_first_frame.dcd (build container
    only): frame count from the file size, unit cells, and the coordinates of the selections equal to the golden arrays
    tests/golden/make_golden.py extracted from it with an independent numpy reader."""
    from cmx_b200.engine import DcdFile
    from common import namd
    src = "/root/reference/test/data/NAMD/traj_duplicated_first_frame.dcd"
    if not os.path.exists(src):
        pytest.skip("reference fixture not present on this machine")
    d = namd()
    f = DcdFile(src)
    assert (f.natoms, f.nframes) == (62026, 3)
    for k in range(3):
        xyz, cell = f.read_frame(k)
        assert np.array_equal(xyz[0:1463], d["protein"][k]) and np.array_equal(xyz[1478:4012], d["tmao"][k])
        assert np.allclose(cell, d["cells"][k], rtol=0, atol=1e-12)
    assert np.array_equal(f.read_frame(0)[0], f.read_frame(1)[0])          # "duplicated first frame"
    f.close()


def test_xtc_compressed_block_roundtrip(tmp_path):
    """decoder of the compressed coordinate block against the test-suite's independent writer (tests/common.py:
    xtc_compress): water-like runs (swapped first pair, adaptive small range), unordered atoms (no runs), a mixture,
    coordinates beyond the 24-bit range (components stored separately), negative coordinates, several frames."""
    from common import write_xtc
    from cmx_b200.engine import XtcFile
    rng = np.random.default_rng(0)

    def water_box(nmol, L):
        o = rng.uniform(0, L, size=(nmol, 1, 3))
        return np.concatenate([o, o + rng.normal(0, 0.06, size=(nmol, 2, 3))], axis=1).reshape(-1, 3)
    cases = {"water": [water_box(400, 3.0), water_box(400, 3.0)], "random": [rng.uniform(-2, 9, size=(500, 3))],
             "mixed": [np.concatenate([rng.uniform(0, 5, size=(37, 3)), water_box(100, 5.0), rng.uniform(0, 5, size=(11, 3))])],
             "big": [rng.uniform(-9000, 9000, size=(50, 3))], "ten": [rng.uniform(0, 1, size=(10, 3))]}
    for name, frames in cases.items():
        path = str(tmp_path / f"{name}.xtc")
        box = np.array([[5.0, 0, 0], [0.5, 5.0, 0], [0.3, 0.2, 5.0]])
        quant = write_xtc(path, np.stack(frames), np.stack([box] * len(frames)))
        f = XtcFile(path)
        assert (f.natoms, f.nframes) == (len(frames[0]), len(frames))
        for k in range(f.nframes):
            xyz, cell, step, time = f.read_frame(k)
            want = ((quant[k].astype(np.float32) * np.float32(1.0 / np.float32(1000.0))).astype(np.float64) * 10.0).astype(np.float32)
            assert np.array_equal(xyz, want), name
            assert np.allclose(cell, 10.0 * box.T, rtol=1e-7) and step == 10 * k
        f.close()


def test_xtc_writer_and_reader_conform_to_the_reference_fixture():
    """format conformance of BOTH directions on a file GROMACS wrote: decoding test/data/nucleic/trajectory.xtc and
    re-encoding the decoded integers reproduces every frame of the file byte for byte (build container only)."""
    from common import xtc_compress
    from cmx_b200.engine import XtcFile
    src = "/root/reference/test/data/nucleic/trajectory.xtc"
    if not os.path.exists(src):
        pytest.skip("reference fixture not present on this machine")
    raw = open(src, "rb").read()
    f = XtcFile(src)
    off = 0
    for k in range(f.nframes):
        xyz, _, _, _ = f.read_frame(k)
        ints = np.rint(xyz.astype(np.float64) * 100.0).astype(np.int64)     # the 0.01 A grid of precision 1000 / nm
        blk = xtc_compress(ints, 1000.0)
        assert raw[off + 56: off + 56 + len(blk)] == blk, f"frame {k}"
        off += 56 + len(blk)
    assert off == len(raw)
    f.close()


def test_c_example_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """examples/mddf_dcd.c: the C ABI from plain C (header is valid C, every symbol links); without a CUDA device the
    program must stop at cmx_create with the library's error -- there is no CPU fallback behind the ABI."""
    from common import namd, write_dcd
    from cmx_b200 import engine
    engine.build()
    exe = str(tmp_path / "mddf_dcd")
    subprocess.run(["/usr/bin/gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "mddf_dcd.c"),
                    "-L", os.path.dirname(engine.LIB_PATH), "-lcmx_b200", "-Wl,-rpath," + os.path.dirname(engine.LIB_PATH),
                    "-Wl,--allow-shlib-undefined", "-o", exe], check=True)
    d = namd()
    path = str(tmp_path / "t.dcd")
    write_dcd(path, np.concatenate([d["protein"], d["tmao"]], axis=1), d["cells"])
    r = subprocess.run([exe, path, "1", "1463", "1464", "2534", "14"], capture_output=True, text=True)
    assert "3997 atoms, 3 frames" in r.stdout
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert r.returncode == 0 and "solvent molecules within" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 1 and "cmx_create" in r.stderr, r.stdout + r.stderr


def test_bench_contract_pieces_on_cpu():
    """bench.py without a GPU: the clock summary (samples inside the timed region, nearest ones for a region shorter than
    the sampling period, 'unavailable' without samples), the Julia probe's answer in this image, and the reference arm's
    JSON line (contract keys; config of the GPU arm, CPU sample stated separately)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.nvml, s.proc = None, None
    row = lambda clk, power: [str(clk), "1965.0", "Not Active", "Not Active", "Not Active", power]
    s.samples = [(1.0, row(1200, "Not Active")), (2.0, row(1965, "Not Active")), (2.1, row(1950, "Active")), (3.0, row(300, "Not Active"))]
    s.t0, s.t1 = 1.9, 2.2
    out = s.summary()
    assert out["samples"] == 2 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    s.t0, s.t1 = 2.04, 2.05                       # shorter than the sampling period: the nearest samples are used
    assert s.summary()["samples"] == 3
    s.samples = []
    assert s.summary()["reasons"] == ["unavailable"]
    ok, why = bench.julia_probe()
    assert isinstance(ok, bool) and isinstance(why, str) and (ok or why)
    import argparse
    args = argparse.Namespace(gpus=1, steps=1, warmup=0, impl="reference", config="C2", frames_per_step=0, scale=0.02, cpu_frames=2,
                              n_random_samples=2)
    line = bench.reference_arm(args, 2)
    for k in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("port", "julia") and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"].startswith("C2 ") and line["config"]["cpu_sample_frames_per_step"] == 2
