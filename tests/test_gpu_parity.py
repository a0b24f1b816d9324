"""GPU parity tests: the CUDA path (through the C ABI) against the fp64 oracle on the same inputs.

Bars (BASELINE.json north_star): minimum distances within 1e-4 A (here: the counted distances
are finalised in fp64 with the oracle's arithmetic, so the bar used is 1e-9 A / bit equality),
integer histogram and contribution counts bit-exact, final mddf/KB within 1e-3 relative.
"""
import numpy as np
import pytest

import cmx_b200 as cm
from common import Problem, assert_counters_equal, assert_lists_equal, namd, toy
from oracle import cmx_oracle as orc

pytestmark = pytest.mark.gpu

PROTEIN = cm.AtomSelection(np.arange(1, 1464), nmols=1)
TMAO = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)
WATER = cm.AtomSelection(np.arange(4013, 62027), natomspermol=3)


def opts(**kw):
    kw.setdefault("silent", True)
    kw.setdefault("seed", 321)
    return cm.Options(**kw)


def check(problem, *, lists=True, dtol=1e-9, engine_kw=None, nsolute_lists=1):
    o, olists = problem.oracle(want_lists=lists)
    eng = problem.engine(keep_lists=lists, **(engine_kw or {}))
    nf = len(problem.xv)
    if lists:
        # frame by frame so that the per-frame lists can be read back
        for k in range(nf):
            problem.run_engine(eng, frames=[k])
            real, rnd = olists[k]
            for isol in range(min(nsolute_lists, problem.solute.nmols)):
                assert_lists_equal(eng.minimum_distances(isol), real[isol], dtol=dtol, what=f"frame {k} real list solute {isol}")
            if not problem.cn_only:
                for s in range(problem.options.n_random_samples):
                    assert_lists_equal(eng.random_minimum_distances(s), rnd[s], dtol=dtol, what=f"frame {k} random sample {s}")
        dev = eng.finish()
    else:
        dev = problem.run_engine(eng)
    stats = eng.stats()
    eng.close()
    assert_counters_equal(dev, o)
    return dev, o, stats


def test_namd_protein_tmao():
    """C1: protein (1 molecule, 1463 atoms) x TMAO, cubic cell, the reference's test/namd.jl:16-22 setup."""
    d = namd()
    p = Problem(PROTEIN, TMAO, opts(bulk_range=(8.0, 10.0), n_random_samples=10), d["protein"], d["tmao"], d["cells"])
    assert p.irefatom == 1
    dev, o, stats = check(p)
    assert dev["md_count"].sum() > 0 and dev["md_count_random"].sum() > 0


def test_namd_kat_coordination_numbers():
    """src/tools/coordination_number.jl:126-134 on frame 1: CN(d>3)=7, CN(d>5)=14, sum CN(O1)=1171."""
    d = namd()
    p = Problem(PROTEIN, TMAO, opts(n_random_samples=1), d["protein"][:1], d["tmao"][:1], d["cells"][:1], coordination_number_only=True)
    eng = p.engine()
    c = p.run_engine(eng)
    eng.close()
    dd = orc.shellradius(np.arange(1, 501), 0.02)
    cn = np.cumsum(c["md_count"])
    assert cn[np.argmax(dd > 3)] == 7.0 and cn[np.argmax(dd > 5)] == 14.0
    assert np.cumsum(c["solvent_group_count"][4]).sum() == 1171.0


@pytest.mark.parametrize("path", [2, 1])
def test_namd_tmao_water_cross(path):
    """solute = 181 TMAO molecules, solvent = 19338 waters (water_tmao.json setup)."""
    d = namd()
    p = Problem(TMAO, WATER, opts(bulk_range=(8.0, 10.0), n_random_samples=3), d["tmao"][:1], [d["water_frame1"]], d["cells"][:1])
    assert p.irefatom == 1
    check(p, lists=(path == 2), engine_kw=dict(path=path), nsolute_lists=3)


@pytest.mark.parametrize("path", [2, 1])
def test_namd_tmao_self(path):
    """autocorrelation of TMAO (tmao_tmao.json setup), both device paths."""
    d = namd()
    p = Problem(TMAO, TMAO, opts(bulk_range=(8.0, 10.0), n_random_samples=10), d["tmao"], None, d["cells"], autocorrelation=True)
    check(p, engine_kw=dict(path=path), nsolute_lists=4)


def test_namd_water_self():
    """19338-molecule water autocorrelation, frame 1 (water_water.json setup)."""
    d = namd()
    p = Problem(WATER, WATER, opts(bulk_range=(8.0, 10.0), n_random_samples=2), [d["water_frame1"]], None, d["cells"][:1], autocorrelation=True)
    check(p, lists=False)


def test_default_options_no_cutoff():
    """usecutoff=false: cutoff = dbulk, bulk = molecules outside the cutoff (src/mddf.jl:55-57)."""
    d = namd()
    p = Problem(PROTEIN, TMAO, opts(dbulk=8.0, n_random_samples=4), d["protein"][:2], d["tmao"][:2], d["cells"][:2])
    check(p)


@pytest.mark.parametrize("n_streams", [1, 3, 8])
def test_frames_in_flight_on_several_streams(n_streams):
    """frames overlap on n_streams compute streams (private scratch, shared integer accumulators):
    the counters do not depend on how many frames are in flight."""
    d = namd()
    p = Problem(PROTEIN, TMAO, opts(bulk_range=(8.0, 10.0), n_random_samples=4), [d["protein"][k % 3] for k in range(9)],
                [d["tmao"][k % 3] for k in range(9)], d["cells"][0])
    check(p, lists=False, engine_kw=dict(n_streams=n_streams))


def _synthetic(triclinic, seed=7, nprot=400, nwat=600, nco=60):
    from cmx_b200 import synthetic as syn
    cell = np.array([[46.0, 9.0, 6.0], [0.0, 44.0, 8.0], [0.0, 0.0, 43.0]]) if triclinic else [44.0, 46.0, 45.0]
    return syn.make_system("t", cell=cell, solute_atoms=nprot, solvents=[("co", "tmao", nco), ("water", "water", nwat)], seed=seed)


@pytest.mark.parametrize("triclinic", [False, True])
def test_synthetic_protein_water(triclinic):
    s = _synthetic(triclinic)
    fr = [s.frame(k)[0] for k in range(3)]
    sol, wat = s.selections["solute"], s.selections["water"]
    p = Problem(sol, wat, opts(bulk_range=(6.0, 9.0), n_random_samples=5), [f[sol.indices - 1] for f in fr],
                [f[wat.indices - 1] for f in fr], s.cell)
    check(p)


@pytest.mark.parametrize("triclinic", [False, True])
@pytest.mark.parametrize("path", [2, 1])
def test_synthetic_self_and_cross_small_molecules(triclinic, path):
    s = _synthetic(triclinic, seed=11)
    fr = [s.frame(k)[0] for k in range(2)]
    co, wat = s.selections["co"], s.selections["water"]
    p = Problem(co, co, opts(bulk_range=(7.0, 10.0), n_random_samples=6), [f[co.indices - 1] for f in fr], None, s.cell, autocorrelation=True)
    check(p, engine_kw=dict(path=path), nsolute_lists=3)
    p = Problem(co, wat, opts(bulk_range=(7.0, 10.0), n_random_samples=6), [f[co.indices - 1] for f in fr],
                [f[wat.indices - 1] for f in fr], s.cell)
    check(p, engine_kw=dict(path=path), nsolute_lists=3)


def test_custom_groups_overlapping():
    """custom groups on both sides, overlapping (src/update_counters.jl:27-33, contributions.jl:375-386)."""
    d = namd()
    g1 = [np.arange(1, 500), np.arange(400, 900), np.arange(1000, 1464)]
    prot = cm.AtomSelection(np.arange(1, 1464), nmols=1, group_atom_indices=g1, group_names=["a", "b", "c"])
    tm_idx = np.arange(1479, 4013)
    g2 = [tm_idx[tm_idx % 14 == 5], tm_idx[:700]]
    tm = cm.AtomSelection(tm_idx, natomspermol=14, group_atom_indices=g2, group_names=["x", "y"])
    p = Problem(prot, tm, opts(bulk_range=(8.0, 10.0), n_random_samples=3), d["protein"][:2], d["tmao"][:2], d["cells"][:2])
    check(p, lists=False)
    # autocorrelation with custom groups (md.i resolves into the first molecule: update_counters.jl:27)
    p = Problem(tm, tm, opts(bulk_range=(8.0, 10.0), n_random_samples=3), d["tmao"][:1], None, d["cells"][:1], autocorrelation=True)
    check(p, lists=False)


def test_frame_weights():
    d = namd()
    for w in ([2.0, 1.0, 0.5], [0.3, 0.0, 0.7]):
        p = Problem(PROTEIN, TMAO, opts(bulk_range=(8.0, 10.0), n_random_samples=2), d["protein"], d["tmao"], d["cells"], weights=w)
        o, _ = p.oracle()
        eng = p.engine()
        dev = p.run_engine(eng)
        eng.close()
        dyadic = all(float(x) in (0.0, 0.5, 1.0, 2.0) for x in w)
        assert_counters_equal(dev, o, exact=dyadic, rtol=1e-12)


def test_toy_mddf_cross_and_self():
    """src/mddf.jl:587-624 through the public mddf()/coordination_number() drivers."""
    t = toy()
    allat = np.arange(1, 11)
    fr = t["cross"]
    protein = cm.AtomSelection([10], nmols=1)
    water = cm.AtomSelection(np.arange(1, 10), natomspermol=3)
    for lastframe in (1, 2):
        tr = cm.ArrayTrajectory(fr, t["cross_cells"], protein, water)
        o = cm.Options(seed=321, silent=True, n_random_samples=10 ** 4, lastframe=lastframe)
        R = cm.mddf(tr, o)
        assert R.volume.total == 27000.0
        assert np.isclose(R.volume.domain, R.volume.total - R.volume.bulk)
        assert np.isclose(R.volume.domain, 4 * np.pi / 3 * R.dbulk ** 3, rtol=0.03)
        assert np.isclose(R.density.solvent_bulk, 2 / R.volume.bulk)
        assert np.isclose(R.md_count.sum(), 1.0) and np.isclose(R.coordination_number.sum(), 51.0)
        tr = cm.ArrayTrajectory(fr, t["cross_cells"], protein, water)
        Cn = cm.coordination_number(tr, o)
        assert np.array_equal(Cn.md_count, R.md_count) and Cn.volume.total == R.volume.total
    atom = cm.AtomSelection([1, 2], natomspermol=1)
    tr = cm.ArrayTrajectory(t["self_monoatomic"], t["self_monoatomic_cells"], atom, atom)
    R = cm.mddf(tr, cm.Options(seed=321, silent=True, n_random_samples=10 ** 4, lastframe=1))
    assert R.volume.total == 27000.0 and np.isclose(R.md_count.sum(), 1.0)
    assert np.isclose(R.density.solute, 2 / R.volume.total)


# ---------------------------------------------------------------------------------------------
# size-independent properties at the full size of the bench configuration (C2: 100k atoms)
# ---------------------------------------------------------------------------------------------
def _c2_problem(nframes=4, nrand=10):
    from cmx_b200 import synthetic as syn
    s = syn.config_c2()
    sol, wat = s.selections["solute"], s.selections["water"]
    fr = [s.frame(k + 1)[0] for k in range(nframes)]
    p = Problem(sol, wat, opts(bulk_range=(10.0, 15.0), n_random_samples=nrand), [f[sol.indices - 1] for f in fr],
                [f[wat.indices - 1] for f in fr], s.cell)
    return s, p


def test_full_size_properties_c2():
    s, p = _c2_problem()
    eng = p.engine()
    full = p.run_engine(eng)
    # (1) every hit credits exactly one solute atom and one solvent atom type (src/tools/contributions.jl:320-348)
    assert np.array_equal(full["solute_group_count"].sum(axis=0), full["md_count"])
    assert np.array_equal(full["solvent_group_count"].sum(axis=0), full["md_count"])
    assert np.array_equal(full["solute_group_count_random"].sum(axis=0), full["md_count_random"])
    assert np.array_equal(full["solvent_group_count_random"].sum(axis=0), full["md_count_random"])
    # (2) a molecule's reference atom is within the cutoff at most as often as the molecule itself
    assert full["rdf_count"].sum() <= full["md_count"].sum() and np.all(np.cumsum(full["rdf_count"]) <= np.cumsum(full["md_count"]))
    # (3) additivity over frames / order independence / sharding: two engines with disjoint frame sets sum to the full run
    eng.reset()
    a = p.run_engine(eng, frames=[0, 2]); eng.reset()
    b = p.run_engine(eng, frames=[3, 1]); eng.reset()
    for k in ("md_count", "md_count_random", "rdf_count", "rdf_count_random", "solute_group_count", "solvent_group_count_random"):
        assert np.array_equal(a[k] + b[k], full[k]), k
    # (4) linearity in the frame weight: weight 2 on every frame doubles every counter
    p2 = Problem(p.solute, p.solvent, p.options, p.xs, p.xv, s.cell, weights=[2.0] * 4)
    d2 = p2.run_engine(eng)
    for k in ("md_count", "md_count_random", "solute_group_count"):
        assert np.array_equal(d2[k], 2 * full[k]), k
    assert np.isclose(d2["volume_total"], 2 * full["volume_total"])
    eng.close()
    # (5) the oracle agrees on one of the frames at this size (bit-exact counters)
    p1 = Problem(p.solute, p.solvent, p.options, p.xs[:1], p.xv[:1], s.cell)
    check(p1, lists=False)


def test_full_size_residue_groups_c2():
    """C2's 375-residue custom-group run: group rows are sums of the per-atom rows."""
    from cmx_b200 import synthetic as syn
    s, p = _c2_problem(nframes=2, nrand=2)
    eng = p.engine()
    per_atom = p.run_engine(eng); eng.close()
    res = syn.residue_groups(p.solute, 16)
    pr = Problem(res, p.solvent, p.options, p.xs, p.xv, s.cell)
    eng = pr.engine()
    grouped = pr.run_engine(eng); eng.close()
    assert grouped["solute_group_count"].shape[0] == 375
    want = per_atom["solute_group_count"].reshape(375, 16, -1).sum(axis=1)
    assert np.array_equal(grouped["solute_group_count"], want)
    assert np.array_equal(grouped["md_count"], per_atom["md_count"])


def test_final_mddf_and_kb_within_tolerance():
    """final mddf / KB / random normalisation within 1e-3 relative of the oracle (BASELINE.json north_star)."""
    d = namd()
    opt = opts(bulk_range=(8.0, 10.0), n_random_samples=10)
    p = Problem(PROTEIN, TMAO, opt, d["protein"], d["tmao"], d["cells"])
    o, _ = p.oracle()
    ref = orc.finalresults(o.counters(), nmols_solute=1, nmols_solvent=181, autocorrelation=False, n_random_samples=10,
                           binstep=0.02, dbulk=8.0, cutoff=10.0, usecutoff=True, Q=3.0)
    tr = cm.ArrayTrajectory(np.concatenate([d["protein"], d["tmao"]], axis=1), d["cells"],
                            cm.AtomSelection(np.arange(1, 1464), nmols=1), cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14))
    R = cm.mddf(tr, opt)
    for a, b in ((R.mddf, ref.mddf), (R.kb, ref.kb), (R.rdf, ref.rdf), (R.kb_rdf, ref.kb_rdf), (R.md_count_random, ref.md_count_random)):
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-12)
    assert np.isclose(R.volume.total, ref.volume_total) and np.isclose(R.density.solvent_bulk, ref.density_solvent_bulk, rtol=1e-3)


def test_error_paths():
    from cmx_b200.engine import CmxError
    d = namd()
    with pytest.raises(CmxError):     # irefatom larger than the molecule (src/Trajectory.jl:191-193)
        Problem(PROTEIN, TMAO, opts(n_random_samples=1), d["protein"][:1], d["tmao"][:1], d["cells"][:1], irefatom=15).engine()
    p = Problem(PROTEIN, TMAO, opts(bulk_range=(8.0, 10.0), n_random_samples=1), d["protein"][:1], d["tmao"][:1], d["cells"][:1])
    eng = p.engine()
    with pytest.raises(CmxError):     # cell narrower than 2*cutoff (CellListMap's requirement)
        eng.submit_arrays(p.xs[0], p.xv[0], np.diag([19.0, 84.0, 84.0]), frame_index=1)
    with pytest.raises(CmxError):     # zero-weight frames are skipped by the driver, never submitted
        eng.submit_arrays(p.xs[0], p.xv[0], p.cells[0], frame_index=1, weight=0.0)
    eng.submit_arrays(p.xs[0], p.xv[0], p.cells[0], frame_index=1)   # the handle is still usable
    assert eng.finish()["md_count"].sum() == 24.0                    # 24 of 181 TMAO within 10 A
    eng.close()


def test_slab_solute_spanning_the_cell():
    """C5-like geometry at small scale: the solute slab spans the cell in x and y, so the search relies on
    the periodic images of the solute atoms on every face."""
    from cmx_b200 import synthetic as syn
    s = syn.config_c5(0.004)       # ~4000 slab atoms, ~5300 waters, 63 x 63 x 50 A
    sol, wat = s.selections["solute"], s.selections["water"]
    fr = [s.frame(k + 1)[0] for k in range(2)]
    p = Problem(sol, wat, opts(bulk_range=(10.0, 15.0), n_random_samples=4), [f[sol.indices - 1] for f in fr],
                [f[wat.indices - 1] for f in fr], s.cell)
    dev, o, stats = check(p)
    assert dev["md_count"].sum() > 1000


# ---------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------
def _tiny(nsol_atoms, solvent_kind, nsolvent, cell, seed=3):
    from cmx_b200 import synthetic as syn
    return syn.make_system("tiny", cell=cell, solute_atoms=nsol_atoms, solvents=[("s", solvent_kind, nsolvent)], seed=seed)


@pytest.mark.parametrize("path", [1, 2])
def test_edge_single_atom_solute_and_single_solvent_molecule(path):
    s = _tiny(1, "water", 1, [40.0, 41.0, 42.0])
    sol, sv = s.selections["solute"], s.selections["s"]
    x = s.frame(0)[0]
    # put the water next to the solute atom so that there is exactly one hit
    x[sv.indices - 1] += x[sol.indices - 1][0] - x[sv.indices - 1][0] + np.float32(3.0)
    p = Problem(sol, sv, opts(bulk_range=(6.0, 9.0), n_random_samples=50), [x[sol.indices - 1]], [x[sv.indices - 1]], s.cell)
    dev, o, _ = check(p, engine_kw=dict(path=path))
    assert dev["md_count"].sum() == 1.0


def test_edge_nothing_within_cutoff():
    s = _tiny(50, "water", 40, [80.0, 80.0, 80.0])
    sol, sv = s.selections["solute"], s.selections["s"]
    x = s.frame(0)[0]
    one = x[sv.indices - 1][:3].copy()
    one = one - one[0] + x[sol.indices - 1].mean(axis=0) + np.float32(35.0)   # every molecule stacked 35 A away
    far = np.tile(one, (40, 1)).astype(np.float32)
    p = Problem(sol, sv, opts(bulk_range=(5.0, 8.0), n_random_samples=3), [x[sol.indices - 1]], [far], s.cell)
    dev, o, _ = check(p)
    assert dev["md_count"].sum() == 0.0 and dev["md_count_random"].sum() > 0   # no bulk molecule -> any molecule is drawn (src/mddf.jl:77-81)


def test_edge_monoatomic_solvent_and_nondefault_irefatom():
    from cmx_b200 import synthetic as syn
    s = syn.make_system("m", cell=[36.0, 36.0, 36.0], solute_atoms=120, solvents=[("ion", "water", 90)], seed=9)
    sol, wat = s.selections["solute"], s.selections["ion"]
    x = s.frame(1)[0]
    # monoatomic: every atom of the 3-site selection as its own molecule
    ions = cm.AtomSelection(wat.indices, natomspermol=1)
    p = Problem(sol, ions, opts(bulk_range=(6.0, 9.0), n_random_samples=5), [x[sol.indices - 1]], [x[ions.indices - 1]], s.cell)
    check(p)
    # reference atom = last atom of the molecule (Options(irefatom=3))
    p = Problem(sol, wat, opts(bulk_range=(6.0, 9.0), n_random_samples=5, irefatom=3), [x[sol.indices - 1]], [x[wat.indices - 1]], s.cell)
    assert p.irefatom == 3
    check(p)
    p = Problem(wat, wat, opts(bulk_range=(6.0, 9.0), n_random_samples=5, irefatom=2), [x[wat.indices - 1]], None, s.cell, autocorrelation=True)
    check(p, nsolute_lists=2)


def test_edge_large_solvent_molecules_and_unwrapped_split_molecules():
    """solvent molecules of 14 atoms whose atoms sit in different periodic images (a molecule 'broken' by
    wrapping, as in trajectories written with per-atom wrapping)."""
    from cmx_b200 import synthetic as syn
    s = syn.make_system("b", cell=[38.0, 40.0, 39.0], solute_atoms=150, solvents=[("co", "glycerol", 60)], seed=21)
    sol, co = s.selections["solute"], s.selections["co"]
    x = s.frame(2)[0].astype(np.float64)
    rng = np.random.default_rng(5)
    xv = x[co.indices - 1]
    xv += rng.integers(-1, 2, size=xv.shape) * np.array([38.0, 40.0, 39.0])   # every ATOM shifted by its own lattice vector
    p = Problem(sol, co, opts(bulk_range=(6.0, 9.0), n_random_samples=6), [x[sol.indices - 1].astype(np.float32)], [xv.astype(np.float32)], s.cell)
    check(p)
    p = Problem(co, co, opts(bulk_range=(6.0, 9.0), n_random_samples=6), [xv.astype(np.float32)], None, s.cell, autocorrelation=True)
    check(p, nsolute_lists=2)


def test_edge_exact_ties_lattice():
    """atoms on an exact lattice: many exactly equal distances -> every tie must be resolved by (d, j, i)."""
    g = np.arange(0, 24, 3.0)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    sol = cm.AtomSelection(np.arange(1, 65), nmols=1)
    sv = cm.AtomSelection(np.arange(65, 65 + 448), natomspermol=2)
    p = Problem(sol, sv, opts(bulk_range=(4.5, 6.0), n_random_samples=3), [pts[:64]], [pts[64:]], np.diag([24.0, 24.0, 24.0]))
    dev, o, stats = check(p)
    assert stats["deferred"] > 0    # the ties went through the exact kernel
    sv1 = cm.AtomSelection(np.arange(1, 513), natomspermol=2)
    p = Problem(sv1, sv1, opts(bulk_range=(4.5, 6.0), n_random_samples=3), [pts], None, np.diag([24.0, 24.0, 24.0]), autocorrelation=True)
    check(p, nsolute_lists=3)


def test_counters_device_block_is_the_allreduce_payload():
    """cmx_counters_device exposes ONE contiguous uint64 block (the payload of the single all-reduce that
    replaces sum!, src/results.jl:629-649); summing two engines' blocks on the device and reading the result
    back through cmx_finish equals one engine that processed all frames."""
    import torch
    d = namd()
    o = opts(bulk_range=(8.0, 10.0), n_random_samples=3)
    p = Problem(PROTEIN, TMAO, o, d["protein"], d["tmao"], d["cells"])
    full = p.engine(); ref = p.run_engine(full); full.close()

    def wrap(eng):
        ptr, n = eng.counters_device()

        class W:
            pass
        w = W()
        w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}
        return torch.as_tensor(w, device="cuda:0"), n
    a, b = p.engine(), p.engine()
    for k, (xs, xv, cell, fid) in enumerate(zip(p.xs, p.xv, p.cells, p.frame_ids)):
        (a if k % 2 == 0 else b).submit_arrays(xs, xv, cell, frame_index=fid)
    a.sync(); b.sync()
    ta, n = wrap(a); tb, _ = wrap(b)
    nb = 500
    assert n == nb * (4 + 2 * 1463 + 2 * 14)
    ta += tb                              # what ncclAllReduce(sum) does across ranks
    torch.cuda.synchronize()
    got = a.finish()
    a.close(); b.close()
    for k in ("md_count", "md_count_random", "rdf_count", "rdf_count_random", "solute_group_count", "solvent_group_count_random"):
        assert np.array_equal(got[k], ref[k]), k


def test_public_mddf_from_dcd_file(tmp_path):
    """mddf(trajectory_file, solute, solvent, options) through the DCD reader feeding the pinned ring, with
    firstframe / stride / frame_weights handled like goto_nextframe! (src/mddf.jl:95-111)."""
    from common import write_dcd
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    path = str(tmp_path / "t.dcd")
    write_dcd(path, frames, d["cells"])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    opt = opts(bulk_range=(8.0, 10.0), n_random_samples=4)
    R = cm.mddf(path, sol, tm, opt, frame_weights=[1.0, 0.0, 2.0])
    p = Problem(sol, tm, opt, d["protein"], d["tmao"], d["cells"], weights=[1.0, 0.0, 2.0])
    o, _ = p.oracle()
    ref = orc.finalresults(o.counters(), nmols_solute=1, nmols_solvent=181, autocorrelation=False, n_random_samples=4,
                           binstep=0.02, dbulk=8.0, cutoff=10.0, usecutoff=True, Q=3.0)
    assert np.allclose(R.md_count, ref.md_count, rtol=1e-13) and np.allclose(R.mddf, ref.mddf, rtol=1e-12)
    assert np.isclose(R.volume.total, ref.volume_total) and R.files[0].irefatom == 1
    Rs = cm.mddf(path, tm, opts(bulk_range=(8.0, 10.0), n_random_samples=2, firstframe=2))   # mddf(file, solute_and_solvent, options)
    assert Rs.autocorrelation and Rs.md_count.sum() > 0
    C = cm.coordination_number(path, sol, tm, opts(lastframe=1))
    assert C.coordination_number[np.argmax(C.d > 3)] == 7.0 and np.all(C.md_count_random == 0)


@pytest.mark.parametrize("path", [1, 2])
def test_random_phase_in_chunks_of_samples(path):
    """the random phase runs over chunks of samples (bounded scratch for any n_random_samples): same counters
    for any chunk size."""
    d = namd()
    o = opts(bulk_range=(8.0, 10.0), n_random_samples=7)
    if path == 1:
        p = Problem(PROTEIN, TMAO, o, d["protein"][:2], d["tmao"][:2], d["cells"][:2])
    else:
        p = Problem(TMAO, TMAO, o, d["tmao"][:2], None, d["cells"][:2], autocorrelation=True)
    ref, _ = p.oracle()
    for chunk in (1, 3):
        eng = p.engine(path=path)
        eng.set_option("sample_chunk", chunk)
        dev = p.run_engine(eng); eng.close()
        assert_counters_equal(dev, ref)


def test_toy_with_the_reference_sample_count():
    """n_random_samples = 10^5 as in the reference's toy tests (src/mddf.jl:601,641): more samples than a grid
    dimension holds, on both device paths."""
    t = toy()
    protein = cm.AtomSelection([10], nmols=1)
    water = cm.AtomSelection(np.arange(1, 10), natomspermol=3)
    o = cm.Options(seed=321, silent=True, n_random_samples=10 ** 5, lastframe=1)
    R = cm.mddf(cm.ArrayTrajectory(t["cross"], t["cross_cells"], protein, water), o)
    assert R.volume.total == 27000.0 and np.isclose(R.md_count.sum(), 1.0) and np.isclose(R.coordination_number.sum(), 51.0)
    assert np.isclose(R.volume.domain, 4 * np.pi / 3 * R.dbulk ** 3, rtol=0.01)      # the reference's tolerance (:607)
    assert np.isclose(R.density.solvent_bulk, 2 / R.volume.bulk)
    atom = cm.AtomSelection([1, 2], natomspermol=1)
    R = cm.mddf(cm.ArrayTrajectory(t["self_monoatomic"], t["self_monoatomic_cells"], atom, atom), o)
    assert R.volume.total == 27000.0 and np.isclose(R.md_count.sum(), 1.0)
    assert np.isclose(R.volume.domain, 4 * np.pi / 3 * R.dbulk ** 3, rtol=0.1)       # (:639)


@pytest.mark.parametrize("bulk_range", [(3.0, 5.0), (2.0, 3.5), (14.0, 20.0)])
def test_small_and_large_cutoffs(bulk_range):
    """cutoffs far from the usual 10-15 A: a small cutoff makes the search grid fine relative to a tile (several
    row chunks per tile), a large one (with a triclinic cell) many cells per row."""
    from cmx_b200 import synthetic as syn
    big = bulk_range[1] > 10
    cell = np.array([[52.0, 9.0, 6.0], [0.0, 50.0, 8.0], [0.0, 0.0, 49.0]]) if big else [30.0, 31.0, 32.0]
    s = syn.make_system("c", cell=cell, solute_atoms=500 if big else 250, solvents=[("water", "water", 900 if big else 700)], seed=13)
    sol, wat = s.selections["solute"], s.selections["water"]
    fr = [s.frame(k)[0] for k in range(2)]
    p = Problem(sol, wat, opts(bulk_range=bulk_range, n_random_samples=4, binstep=0.05 if not big else 0.02), [f[sol.indices - 1] for f in fr],
                [f[wat.indices - 1] for f in fr], s.cell)
    dev, o, _ = check(p)
    assert dev["md_count"].sum() > 0


# ---------------------------------------------------------------------------------------------
# native DCD feed, device-side group reduction, merge
# ---------------------------------------------------------------------------------------------
def _dcd_with_padding(tmp_path, triclinic_second=False):
    """a DCD whose selected atoms are a strict, shuffled-order subset of the file (so that the device gather matters)"""
    from common import write_dcd
    d = namd()
    nf = d["protein"].shape[0]
    rng = np.random.default_rng(11)
    pad0 = rng.uniform(0, 80, size=(nf, 7, 3)).astype(np.float32)
    pad1 = rng.uniform(0, 80, size=(nf, 13, 3)).astype(np.float32)
    frames = np.concatenate([pad0, d["protein"], pad1, d["tmao"]], axis=1)       # protein at 8..1470, TMAO at 1484..4017
    cells = [np.asarray(c, dtype=np.float64) for c in d["cells"]]
    if triclinic_second:
        cells[1] = np.array([[84.4, 6.0, 4.0], [0.0, 84.0, 5.0], [0.0, 0.0, 83.5]])
    path = str(tmp_path / "padded.dcd")
    write_dcd(path, frames, cells)
    sol = cm.AtomSelection(np.arange(8, 8 + 1463), nmols=1)
    tm = cm.AtomSelection(np.arange(1484, 1484 + 2534), natomspermol=14)
    return path, sol, tm, d


@pytest.mark.parametrize("threads", [1, 3])
def test_native_dcd_feed_equals_the_staging_slot_path(tmp_path, threads):
    """cmx_run_dcd (reader threads -> pinned ring -> raw H2D -> device gather) == acquire/submit with the host
    reader == oracle, bit for bit; frames by number (any order/subset), weights, triclinic frame included."""
    from cmx_b200.engine import DcdFile, Engine
    path, sol, tm, d = _dcd_with_padding(tmp_path, triclinic_second=True)
    opt = opts(bulk_range=(8.0, 10.0), n_random_samples=3)
    f = DcdFile(path)
    cells = [f.read_frame(k)[1] for k in range(3)]
    # many passes over the 3 frames so that the ring (4+ slots) wraps several times
    order = [0, 1, 2, 2, 1, 0, 1, 1, 2, 0, 0, 2, 1]
    weights = [1.0, 2.0, 1.0, 1.0, 2.0, 1.0, 2.0, 2.0, 1.0, 1.0, 1.0, 1.0, 2.0]
    # one stream -> a ring of 4 slots, so the 13 frames wrap it three times; three threads -> the default ring (18 slots)
    eng = Engine(solute=sol, solvent=tm, options=opt, irefatom=1, autocorrelation=False, n_streams=1 if threads == 1 else 0)
    eng.run_dcd(f, sol.indices, tm.indices, order, weights, n_reader_threads=threads)
    dev = eng.finish()
    st = eng.stats()
    assert st["frames"] == len(order) and st["h2d_bytes"] == len(order) * f.frame_bytes
    # same frames through the staging-slot path (frame key = frame number + 1)
    eng2 = Engine(solute=sol, solvent=tm, options=opt, irefatom=1, autocorrelation=False)
    for k, w in zip(order, weights):
        eng2.submit_arrays(d["protein"][k], d["tmao"][k], cells[k], frame_index=k + 1, weight=w)
    dev2 = eng2.finish(); eng2.close()
    for key in dev:
        assert np.array_equal(dev[key], dev2[key]), key
    p = Problem(sol, tm, opt, [d["protein"][k] for k in order], [d["tmao"][k] for k in order], [cells[k] for k in order],
                weights=weights, frame_ids=[k + 1 for k in order], irefatom=1)
    o, _ = p.oracle()
    assert_counters_equal(dev, o, exact=True)
    # a second run on the same handle re-uses the ring; an autocorrelation ignores the solute indices
    eng.reset()
    eng.run_dcd(f, sol.indices, tm.indices, [2])
    one = eng.finish(); eng.close()
    p1 = Problem(sol, tm, opt, [d["protein"][2]], [d["tmao"][2]], [cells[2]], frame_ids=[3], irefatom=1)
    assert_counters_equal(one, p1.oracle()[0])
    enga = Engine(solute=tm, solvent=tm, options=opt, irefatom=1, autocorrelation=True)
    enga.run_dcd(f, None, tm.indices, [0, 2])
    da = enga.finish(); enga.close()
    pa = Problem(tm, tm, opt, [d["tmao"][0], d["tmao"][2]], None, [cells[0], cells[2]], autocorrelation=True, frame_ids=[1, 3], irefatom=1)
    assert_counters_equal(da, pa.oracle()[0])
    f.close()


def test_native_feed_errors_and_public_driver(tmp_path):
    from cmx_b200.engine import CmxError, DcdFile, Engine
    path, sol, tm, d = _dcd_with_padding(tmp_path)
    opt = opts(bulk_range=(8.0, 10.0), n_random_samples=2)
    f = DcdFile(path)
    eng = Engine(solute=sol, solvent=tm, options=opt, irefatom=1, autocorrelation=False)
    with pytest.raises(CmxError):
        eng.run_dcd(f, sol.indices, tm.indices, [3])                       # frame outside the file
    with pytest.raises(CmxError):
        eng.run_dcd(f, sol.indices, tm.indices + 100000, [0])              # index outside the file
    with pytest.raises(CmxError):
        eng.run_dcd(f, sol.indices, tm.indices, [0], [0.0])                # zero weight
    eng.run_dcd(f, sol.indices, tm.indices, [])                            # nothing to do
    assert eng.stats()["frames"] == 0
    eng.close(); f.close()
    # public driver: native feed (default for DCD) == host feed, incl. firstframe/stride/weights
    for kw in (dict(), dict(firstframe=2), dict(stride=2)):
        o = opts(bulk_range=(8.0, 10.0), n_random_samples=2, **kw)
        Rn = cm.mddf(path, sol, tm, o, frame_weights=[1.0, 3.0, 0.5], feed="native", reader_threads=2)
        Rh = cm.mddf(path, sol, tm, o, frame_weights=[1.0, 3.0, 0.5], feed="host")
        for key in ("md_count", "md_count_random", "rdf_count", "solute_group_count", "solvent_group_count_random", "mddf", "kb"):
            assert np.array_equal(getattr(Rn, key), getattr(Rh, key)), (kw, key)
        assert Rn.volume.total == Rh.volume.total
    with pytest.raises(ValueError):
        cm.mddf(cm.ArrayTrajectory(d["protein"], d["cells"], PROTEIN, PROTEIN), opts(bulk_range=(8.0, 10.0)), feed="native")


@pytest.mark.parametrize("auto", [False, True])
def test_reduce_groups_on_device(auto):
    """cmx_reduce_groups == summing the rows of cmx_finish's arrays (contributions / ResidueContributions count stage)."""
    d = namd()
    opt = opts(bulk_range=(8.0, 10.0), n_random_samples=3)
    if auto:
        p = Problem(TMAO, TMAO, opt, d["tmao"][:2], None, d["cells"][:2], autocorrelation=True, irefatom=1)
    else:
        p = Problem(PROTEIN, TMAO, opt, d["protein"][:2], d["tmao"][:2], d["cells"][:2], irefatom=1)
    eng = p.engine()
    eng.set_option("profile", 1)
    dev = p.run_engine(eng)
    nrows = dev["solute_group_count"].shape[0]
    rng = np.random.default_rng(5)
    groups = [np.arange(0, nrows)[k::7] for k in range(7)]                     # disjoint cover
    groups += [rng.choice(nrows, size=min(nrows, 300), replace=False), np.array([0]), np.zeros(0, dtype=int), np.arange(nrows)]
    for which in ("solute_group_count", "solute_group_count_random", "solvent_group_count", "solvent_group_count_random"):
        arr = dev[which]
        gs = groups if arr.shape[0] == nrows else [np.arange(arr.shape[0]), np.array([1, 3]), np.zeros(0, dtype=int)]
        got = eng.reduce_groups(which, gs)
        want = np.stack([arr[np.asarray(g, dtype=int)].sum(axis=0) if len(g) else np.zeros(arr.shape[1]) for g in gs])
        assert np.array_equal(got, want), which
    assert eng.stats()["gpu_ms_reduce"] > 0
    # the seven disjoint groups add up to md_count (tools/contributions.jl:320-348); half weights in an autocorrelation
    tot = eng.reduce_groups("solute_group_count", groups[:7]).sum(axis=0)
    assert np.array_equal(tot, dev["md_count"])
    from cmx_b200.engine import CmxError
    with pytest.raises(CmxError):
        eng.reduce_groups("solute_group_count", [np.array([nrows])])
    eng.close()
    # varying frame weights: fp64 accumulators are part of the sum
    pw = Problem(PROTEIN, TMAO, opt, d["protein"], d["tmao"], d["cells"], weights=[1.0, 0.3, 2.5], irefatom=1)
    eng = pw.engine()
    dev = pw.run_engine(eng)
    gs = [np.arange(0, 1463)[k::3] for k in range(3)]
    got = eng.reduce_groups("solute_group_count", gs); eng.close()
    want = np.stack([dev["solute_group_count"][g].sum(axis=0) for g in gs])
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)


def test_merge_of_two_halves_equals_reference_semantics():
    """merge (src/tools/merge.jl:150-222): toy two-atom system, halves of the trajectory, weighted frames."""
    t = toy()
    at1, at2 = cm.AtomSelection([1], nmols=1), cm.AtomSelection([2], nmols=1)
    tr = lambda: cm.ArrayTrajectory(t["self_monoatomic"], t["self_monoatomic_cells"], at1, at2)
    R1 = cm.mddf(tr(), opts(lastframe=1, n_random_samples=100))
    R2 = cm.mddf(tr(), opts(firstframe=2, n_random_samples=100))
    assert R1.md_count.sum() == 1 and R2.md_count.sum() == 0
    R = cm.merge([R1, R2])
    assert R.weights == [0.5, 0.5] and R.md_count.sum() == 0.5 and len(R.files) == 2
    R1w = cm.mddf(tr(), opts(lastframe=1, n_random_samples=100), frame_weights=[2.0])
    assert R1w.md_count.sum() == 1
    R = cm.merge([R1w, R2])
    assert np.isclose(R.md_count.sum(), 2 / 3) and np.isclose(R.solute_group_count.sum(), 2 / 3) and np.isclose(R.solvent_group_count.sum(), 2 / 3)
    assert R.volume.total == 27000.0


def test_mddf_many_reuses_one_engine(tmp_path):
    """mddf_many: several DCD files through ONE engine (cmx_reset between them) == separate mddf calls, and the
    merged Result equals merge() of the parts (frame-weighted, src/tools/merge.jl)."""
    from common import write_dcd
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    pa, pb = str(tmp_path / "a.dcd"), str(tmp_path / "b.dcd")
    write_dcd(pa, frames[:2], d["cells"][:2]); write_dcd(pb, frames[2:], d["cells"][2:])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    o = opts(bulk_range=(8.0, 10.0), n_random_samples=2)
    parts, merged = cm.mddf_many([pa, pb], sol, tm, o)
    Ra, Rb = cm.mddf(pa, sol, tm, o), cm.mddf(pb, sol, tm, o)
    for got, want in zip(parts, (Ra, Rb)):
        for key in ("md_count", "md_count_random", "solute_group_count", "rdf_count", "mddf", "kb"):
            assert np.array_equal(getattr(got, key), getattr(want, key)), key
    M = cm.merge([Ra, Rb])
    assert merged.weights == [2 / 3, 1 / 3] and np.allclose(merged.md_count, M.md_count, rtol=0, atol=0)
    assert np.allclose(merged.md_count, (2 * Ra.md_count + Rb.md_count) / 3, rtol=1e-15)


def test_xtc_native_and_host_feed_small(tmp_path):
    """GROMACS XTC through cmx_run_xtc (reader threads decode into the pinned ring, device gather of xyz triplets) and
    through the host reader XTCTraj: same Result as an in-memory trajectory of the decoded frames.  Frames of <= 9
    atoms (plain floats in the format; the compressed block is covered on the CPU by tests/test_host.py)."""
    from common import write_xtc_small
    from cmx_b200.engine import XtcFile
    t = toy()
    fr, cells = t["self"], t["self_cells"]                                  # 6 atoms = 2 molecules x 3, 2 frames
    nrep = 7                                                                # more frames than ring slots of a 1-stream run
    frames_nm = np.concatenate([fr] * nrep).astype(np.float64) / 10.0
    boxes_nm = np.concatenate([np.transpose(cells, (0, 2, 1))] * nrep) / 10.0   # XTC rows = box vectors
    path = str(tmp_path / "toy.xtc")
    write_xtc_small(path, frames_nm, boxes_nm)
    x = XtcFile(path)
    dec = [x.read_frame(k) for k in range(x.nframes)]
    x.close()
    sel = cm.AtomSelection(np.arange(1, 7), natomspermol=3)
    o = opts(n_random_samples=50)
    w = [1.0, 2.0] * nrep
    Rn = cm.mddf(path, sel, o, frame_weights=w, feed="native", reader_threads=3, _engine_kw={"n_streams": 1})   # ring of 5 slots: wraps
    Rh = cm.mddf(path, sel, o, frame_weights=w, feed="host")
    Ra = cm.mddf(cm.ArrayTrajectory(np.stack([d[0] for d in dec]), np.stack([d[1] for d in dec]), sel, sel), o, frame_weights=w)
    assert Rn.md_count.sum() > 0 and Rn.volume.total == Ra.volume.total
    for key in ("md_count", "md_count_random", "rdf_count", "rdf_count_random", "solute_group_count", "solvent_group_count", "mddf", "kb"):
        assert np.array_equal(getattr(Rn, key), getattr(Ra, key)), key
        assert np.array_equal(getattr(Rh, key), getattr(Ra, key)), key
    # a cross-correlation with a strict subset of the file's atoms (the gather matters), unit weights
    a, b = cm.AtomSelection([1, 2, 3], nmols=1), cm.AtomSelection([4, 5, 6], natomspermol=3)
    Rn = cm.mddf(path, a, b, o, feed="native")
    Ra = cm.mddf(cm.ArrayTrajectory(np.stack([d[0] for d in dec]), np.stack([d[1] for d in dec]), a, b), o)
    assert np.array_equal(Rn.md_count, Ra.md_count) and np.array_equal(Rn.md_count_random, Ra.md_count_random)


def test_xtc_compressed_feed(tmp_path):
    """a compressed XTC (protein + TMAO of the NAMD fixture, written by the test-suite's own XTC writer, which
    reproduces GROMACS' bytes on the reference's fixture) through cmx_run_xtc with decoding reader threads: same
    counters as the decoded frames through the staging-slot path, and as the oracle on those frames."""
    from common import write_xtc
    from cmx_b200.engine import Engine, XtcFile
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)                 # 3997 atoms, Angstrom
    boxes = np.stack([np.asarray(c, dtype=np.float64).T / 10.0 for c in d["cells"]])
    path = str(tmp_path / "c.xtc")
    write_xtc(path, frames.astype(np.float64) / 10.0, boxes)
    x = XtcFile(path)
    assert (x.natoms, x.nframes) == (3997, 3)
    dec = [x.read_frame(k) for k in range(3)]
    assert max(np.abs(dec[k][0] - frames[k]).max() for k in range(3)) < 0.0051   # the format's 0.01 A grid
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    opt = opts(bulk_range=(8.0, 10.0), n_random_samples=3)
    order = [0, 1, 2, 1, 0, 2, 2]
    eng = Engine(solute=sol, solvent=tm, options=opt, irefatom=1, autocorrelation=False, n_streams=2)
    eng.run_xtc(x, sol.indices, tm.indices, order, n_reader_threads=3)
    dev = eng.finish(); eng.close(); x.close()
    p = Problem(sol, tm, opt, [dec[k][0][:1463] for k in order], [dec[k][0][1463:] for k in order], [dec[k][1] for k in order],
                frame_ids=[k + 1 for k in order], irefatom=1)
    o, _ = p.oracle()
    assert_counters_equal(dev, o)
    eng2 = p.engine()
    dev2 = p.run_engine(eng2); eng2.close()
    for key in dev:
        assert np.array_equal(dev[key], dev2[key]), key


@pytest.mark.parametrize("seed", range(10))
def test_random_geometries_against_the_oracle(seed):
    """Randomised systems (the search takes every branch of its ring / span bookkeeping somewhere): random orthorhombic or
    triclinic cell, solute blob anywhere in it -- across faces, shifted by whole cell vectors --, solute density from
    sparse to packed, solvent of 1 / 3 / 5 atoms per molecule with unwrapped coordinates, random cutoffs, several frames
    per batch.  Counters bit-exact against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    L = rng.uniform(38.0, 64.0, size=3)
    cell = np.diag(L)
    if seed % 2:
        cell[0, 1], cell[0, 2], cell[1, 2] = rng.uniform(-0.2, 0.2, size=3) * L[[0, 0, 1]]      # columns = lattice vectors
    ib, ic = int(rng.integers(8, 16)), int(rng.integers(2, 8))
    dbulk, cutoff = 0.5 * ib, 0.5 * (ib + ic)               # multiples of 0.5 A (cutoff must be a multiple of binstep = 0.02)
    ns = int(rng.integers(1, 900))
    napm = int(rng.choice([1, 3, 5]))
    nmol = int(rng.integers(50, 1500))
    spread = float(rng.uniform(2.0, 12.0))
    tmpl = rng.normal(scale=0.9, size=(napm, 3))
    frames_s, frames_v = [], []
    for f in range(3):
        centre = cell @ rng.uniform(0.0, 1.0, size=3)
        xs = centre + rng.normal(scale=spread, size=(ns, 3)) + cell @ rng.integers(-2, 3, size=3)
        com = (cell @ rng.uniform(0.0, 1.0, size=(3, nmol))).T + (cell @ rng.integers(-1, 2, size=(3, nmol))).T
        q = rng.normal(size=(nmol, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
        w, x, y, z = q.T
        R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                      np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                      np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
        xv = (com[:, None, :] + np.einsum("mij,aj->mai", R, tmpl)).reshape(-1, 3)
        frames_s.append(xs.astype(np.float32)); frames_v.append(xv.astype(np.float32))
    sol = cm.AtomSelection(np.arange(1, ns + 1), nmols=1)
    solv = cm.AtomSelection(np.arange(ns + 1, ns + 1 + nmol * napm), natomspermol=napm)
    p = Problem(sol, solv, opts(bulk_range=(dbulk, cutoff), n_random_samples=int(rng.integers(1, 5))), frames_s, frames_v, cell)
    dev, o, stats = check(p, lists=False, engine_kw=dict(batch_frames=int(rng.choice([1, 2, 3])), n_streams=int(rng.choice([1, 2]))))
    assert stats["frames"] == 3
