"""CPU tests: the oracle against the reference's own known-answer tests and golden values.

These pin the CPU restatement (oracle/) before it is trusted as the checker of the CUDA path.
Every expected value is quoted from the reference's tests (file:line in each docstring).
"""
import numpy as np
import pytest

import cmx_b200 as cm
from common import Problem, kat, namd, toy
from oracle import cmx_oracle as orc

PROTEIN = cm.AtomSelection(np.arange(1, 1464), nmols=1)
TMAO = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)
WATER = cm.AtomSelection(np.arange(4013, 62027), natomspermol=3)


def test_philox_random123_vectors():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_update_md():
    """src/minimum_distances.jl:41-46."""
    out = orc.update_md((1, 1, 2, 1, 1.0, 1.0), (1, 1, 2, 1, 0.5, 0.5))
    assert out == (1, 1, 2, 1, 0.5, 0.5)


def test_setbin_shellradius_volume():
    """src/results.jl:28, :259-264, :277-281, :486-497."""
    assert orc.setbin(0.0, 0.02) == 1 and orc.setbin(0.02, 0.02) == 1 and orc.setbin(0.0201, 0.02) == 2
    assert orc.setbin(10.0, 0.02) == 500 == cm.setbin(10.0, 0.02)
    k = kat()["results.jl:277-281"]
    assert np.isclose(orc.shellradius(1, 0.1), k["shellradius(1,0.1)"], rtol=1e-14)
    assert np.isclose(orc.shellradius(5, 0.3), k["shellradius(5,0.3)"], rtol=1e-14)
    assert np.isclose(cm.shellradius(5, 0.3), k["shellradius(5,0.3)"], rtol=1e-14)
    assert np.isclose(orc.sphericalshellvolume(1, 1.0), 4 * np.pi / 3)
    assert np.isclose(orc.sphericalshellvolume(3, 1.0), 4 * np.pi / 3 * (27 - 8))
    assert np.isclose(cm.sphericalshellvolume(2, 1.0), 4 * np.pi / 3 * 7)


def test_eulermat_and_move():
    """src/rigid_body.jl:59-65 and :82-96."""
    I = np.eye(3)
    assert np.allclose(orc.eulermat(0, 0, 0), I)
    assert np.allclose(orc.eulermat(np.pi, 0, 0), np.diag([1, -1, -1]))
    assert np.allclose(orc.eulermat(0, np.pi, 0), np.diag([-1, 1, -1]))
    assert np.allclose(orc.eulermat(0, 0, np.pi), np.diag([-1, -1, 1]))
    x = np.array([[1.0, 0, 0], [0, 0, 0]])
    assert np.allclose(orc.move(x, [0, 0, 0], 0, 0, 0), [[0.5, 0, 0], [-0.5, 0, 0]])
    assert np.allclose(orc.move(x, [1, 1, 1], 0, 0, 0), [[1.5, 1, 1], [0.5, 1, 1]])
    assert np.allclose(orc.move(x, [0, 0, 0], np.pi, 0, 0), [[0.5, 0, 0], [-0.5, 0, 0]])
    assert np.allclose(orc.move(x, [0, 0, 0], 0, np.pi, 0), [[-0.5, 0, 0], [0.5, 0, 0]])
    assert np.allclose(orc.move(x, [0, 0, 0], 0, 0, np.pi), [[-0.5, 0, 0], [0.5, 0, 0]])


@pytest.mark.parametrize("cell", [np.diag([10.0, 10.0, 10.0]), np.array([[10.0, 5.0, 0.0], [0, 10.0, 0], [0, 0, 10.0]])])
@pytest.mark.parametrize("shift", [0.0, -9.0, 4.0])
def test_random_move_is_rigid(cell, shift):
    """src/rigid_body.jl:139-190: internal distances are preserved, orthorhombic and triclinic cells."""
    rng = np.random.default_rng(3)
    x = -1.0 + 2 * rng.uniform(size=(5, 3)) + shift
    for k in range(5):
        y = orc.random_move(cell, x, 0, seed=321, slot=k, sample=1, frame=2)
        dx = np.linalg.norm(x[:, None] - x[None], axis=-1)
        dy = np.linalg.norm(y[:, None] - y[None], axis=-1)
        assert np.allclose(dx, dy, atol=1e-12)
        assert not np.allclose(x, y)


def test_random_move_splits_are_healed():
    """a molecule broken across the periodic boundary is re-assembled about its reference atom (:129-131)."""
    cell = np.diag([10.0, 10.0, 10.0])
    x = np.array([[9.8, 5.0, 5.0], [0.3, 5.0, 5.0], [9.9, 5.6, 5.0]])   # atom 1 is the image of 10.3
    y = orc.random_move(cell, x, 0, seed=1)
    assert np.isclose(np.linalg.norm(y[0] - y[1]), 0.5) and np.isclose(np.linalg.norm(y[0] - y[2]), np.hypot(0.1, 0.6))


def test_default_irefatom_matches_golden_json():
    """TrajectoryMetaData default irefatom (src/Trajectory.jl:206-213) vs the values stored in the
    reference's golden JSON files test/data/NAMD/{tmao_tmao,water_tmao,water_water}.json."""
    from common import default_irefatom
    d, k = namd(), kat()["irefatom"]
    assert default_irefatom(d["tmao"][0], 14) == k["tmao_tmao"]
    assert default_irefatom(d["water_frame1"], 3) == k["water_water"]
    assert default_irefatom(d["water_frame1"], 3) == k["water_tmao"]


def test_namd_frame1_coordination_numbers():
    """src/tools/coordination_number.jl:126-134 ("checked with VMD"): frame 1, protein x TMAO."""
    d, k = namd(), kat()["coordination_number.jl:126-134"]
    p = Problem(PROTEIN, TMAO, cm.Options(lastframe=1, seed=321, silent=True, n_random_samples=1), d["protein"][:1], d["tmao"][:1], d["cells"][:1])
    assert p.irefatom == 1
    for use_clist in (False, True):
        o, _ = p.oracle(use_clist=use_clist)
        f = orc.finalresults(o.counters(), nmols_solute=1, nmols_solvent=181, autocorrelation=False, n_random_samples=1,
                             binstep=0.02, dbulk=10.0, cutoff=10.0, usecutoff=False, Q=1.0)
        assert f.coordination_number[np.argmax(f.d > 3)] == k["cn_first_d_gt_3"]
        assert f.coordination_number[np.argmax(f.d > 5)] == k["cn_first_d_gt_5"]
        # O1 is the 5th atom of TMAO in structure.pdb
        assert np.cumsum(f.solvent_group_count[4]).sum() == k["sum_cn_O1"]
        # self-consistency (:120): sum of solute group counts == sum of solvent group counts == md_count
        assert np.isclose(f.solute_group_count.sum(), f.solvent_group_count.sum())
        assert np.allclose(f.solute_group_count.sum(axis=0), f.md_count)
    assert np.isclose(d["cells"][0][0, 0], kat()["minimum_distances.jl:204"]["unitcell"])


def test_namd_golden_json_consistency():
    """The golden JSONs were produced from frames 1,6,11,16 of trajectory.dcd (stride 5); only frame 1
    exists here.  Necessary condition: every bin of 4*nmols*md_count(JSON) holds at least our frame-1 hits."""
    d, k = namd(), kat()["golden_json_sums"]
    o5 = cm.Options(stride=5, seed=321, silent=True, bulk_range=(8.0, 10.0), n_random_samples=1)
    p = Problem(TMAO, TMAO, o5, d["tmao"][:1], None, d["cells"][:1], autocorrelation=True)
    o, _ = p.oracle()
    g = k["tmao_tmao"]
    total = np.array(g["md_count"]) * 4 * g["solute_nmols"]
    assert np.all(total + 1e-6 >= o.md_count)
    assert np.allclose(total, np.round(total), atol=1e-6)     # integer counts, as our counters
    assert 0.15 < o.md_count.sum() / total.sum() < 0.40       # one frame out of four


def test_toy_cross(  ):
    """src/mddf.jl:587-624: 1 C atom + 3 waters in a 30 A box."""
    t = toy()
    protein = cm.AtomSelection([10], nmols=1)
    water = cm.AtomSelection(np.arange(1, 10), natomspermol=3)
    fr = t["cross"]
    for lastframe in (1, 2):
        opt = cm.Options(seed=321, silent=True, n_random_samples=20000, lastframe=lastframe)
        p = Problem(protein, water, opt, fr[:lastframe, 9:10], fr[:lastframe, 0:9], t["cross_cells"][:lastframe], irefatom=1)
        o, _ = p.oracle()
        f = orc.finalresults(o.counters(), nmols_solute=1, nmols_solvent=3, autocorrelation=False, n_random_samples=20000,
                             binstep=0.02, dbulk=10.0, cutoff=10.0, usecutoff=False, Q=float(lastframe))
        assert f.volume_total == 27000.0
        assert np.isclose(f.volume_domain, f.volume_total - f.volume_bulk)
        assert np.isclose(f.volume_domain, 4 * np.pi / 3 * 10.0 ** 3, rtol=0.02)
        assert np.isclose(f.density_solute, 1 / 27000.0) and np.isclose(f.density_solvent, 3 / 27000.0)
        assert np.isclose(f.density_solvent_bulk, 2 / f.volume_bulk)
        assert np.isclose(f.md_count.sum(), 1.0) and np.isclose(f.coordination_number.sum(), 51.0)


def test_toy_self_monoatomic_weights():
    """src/mddf.jl:626-722: frame-weight identities on the two-atom system."""
    t = toy()
    atom = cm.AtomSelection([1, 2], natomspermol=1)
    fr, cells = t["self_monoatomic"], t["self_monoatomic_cells"]
    opt = cm.Options(seed=321, silent=True, n_random_samples=50)

    def run(frames, cells_, weights=None):
        p = Problem(atom, atom, opt, frames, None, cells_, autocorrelation=True, weights=weights, irefatom=1)
        o, _ = p.oracle()
        Q = float(np.sum(weights)) if weights is not None else float(len(frames))
        return o, orc.finalresults(o.counters(), nmols_solute=2, nmols_solvent=2, autocorrelation=True, n_random_samples=50,
                                   binstep=0.02, dbulk=10.0, cutoff=10.0, usecutoff=False, Q=Q)
    _, f1 = run(fr[:1], cells[:1])
    assert np.isclose(f1.md_count.sum(), 1.0)                     # :641  (d = 5 A, both ordered pairs / nmols)
    _, f2 = run(fr[1:2], cells[1:2])
    _, f2w = run(fr, cells, weights=[0.0, 1.0])
    assert np.array_equal(f2.md_count, f2w.md_count) and np.array_equal(f2.rdf_count, f2w.rdf_count)   # :667-670
    _, fa = run(fr, cells)
    _, fb = run(fr, cells, weights=[0.3, 0.3])
    assert np.allclose(fa.md_count, fb.md_count, rtol=1e-15)      # :689-692
    dup, dupc = t["self_monoatomic_duplicated_first_frame"], t["self_monoatomic_duplicated_first_frame_cells"]
    _, fd = run(dup, dupc)
    _, fw = run(fr, cells, weights=[2.0, 1.0])
    assert np.array_equal(fd.md_count, fw.md_count) and np.array_equal(fd.rdf_count, fw.rdf_count)     # :693-700
    assert np.isclose(fw.md_count.sum(), 2 / 3)                   # :703
    _, fw2 = run(fr, cells, weights=[1.0, 2.0])
    assert np.isclose(fw2.md_count.sum(), 1 / 3)                  # :705


def test_frame_weight_identities_real_data():
    """src/mddf.jl:863-879 on the 3-frame DCD whose first two frames are identical."""
    d = namd()
    opt = cm.Options(seed=321, silent=True, n_random_samples=1)
    a = Problem(TMAO, TMAO, opt, d["tmao"][1:3], None, d["cells"][1:3], autocorrelation=True, frame_ids=[2, 3])
    b = Problem(TMAO, TMAO, opt, d["tmao"], None, d["cells"], autocorrelation=True, weights=[1.0, 0.0, 1.0], frame_ids=[1, 2, 3])
    oa, _ = a.oracle(); ob, _ = b.oracle()
    # frames 1 and 2 are identical, so {2,3} == {1,3} for the deterministic counters
    assert np.array_equal(oa.md_count, ob.md_count) and np.array_equal(oa.solute_group_count, ob.solute_group_count)
    assert np.isclose(oa.volume_total, ob.volume_total)
    c = Problem(TMAO, TMAO, opt, d["tmao"], None, d["cells"], autocorrelation=True, weights=[2.0, 0.0, 1.0])
    e = Problem(TMAO, TMAO, opt, d["tmao"], None, d["cells"], autocorrelation=True)
    oc, _ = c.oracle(); oe, _ = e.oracle()
    assert np.array_equal(oc.md_count, oe.md_count)


def test_brute_force_equals_cell_list_triclinic():
    from cmx_b200 import synthetic as syn
    s = syn.make_system("t", cell=np.array([[46.0, 9.0, 6.0], [0.0, 44.0, 8.0], [0.0, 0.0, 43.0]]), solute_atoms=200,
                        solvents=[("water", "water", 300)], seed=3)
    x, cell = s.frame(1)
    sol, wat = s.selections["solute"], s.selections["water"]
    p = Problem(sol, wat, cm.Options(bulk_range=(6.0, 9.0), n_random_samples=3, silent=True), [x[sol.indices - 1]], [x[wat.indices - 1]], cell)
    a, la = p.oracle(use_clist=False, want_lists=True)
    b, lb = p.oracle(use_clist=True, want_lists=True)
    assert np.array_equal(la[0][0], lb[0][0]) and np.array_equal(la[0][1], lb[0][1])
    for k in ("md_count", "md_count_random", "rdf_count", "solute_group_count", "solvent_group_count_random"):
        assert np.array_equal(a.counters()[k], b.counters()[k])
    # distances agree with an independent numpy minimum-image computation
    real = la[0][0][0]
    inv = np.linalg.inv(cell)
    xs, xv = x[sol.indices - 1].astype(np.float64), x[wat.indices - 1].astype(np.float64)
    for m in np.flatnonzero(real["within_cutoff"])[:25]:
        dr = xv[real["j"][m]] - xs[real["i"][m]]
        f = inv @ dr; f -= np.round(f)
        assert np.isclose(np.linalg.norm(cell @ f), real["d"][m], rtol=1e-12)
        allf = (inv @ (xv[3 * m:3 * m + 3, None, :] - xs[None, :, :]).reshape(-1, 3).T); allf -= np.round(allf)
        assert np.isclose(np.linalg.norm(cell @ allf, axis=0).min(), real["d"][m], rtol=1e-12)


def test_multithreaded_run_frames_equals_serial():
    """the frame-parallel chunk loop (src/mddf.jl:285-338) sums to the same counters for any thread count."""
    d = namd()
    opt = cm.Options(bulk_range=(8.0, 10.0), seed=321, silent=True, n_random_samples=3)
    res = []
    for nt in (1, 3):
        o = orc.Oracle.from_problem(PROTEIN, TMAO, opt, 1, False)
        o.run_frames(d["protein"], d["tmao"], d["cells"][0], frame_ids=[1, 2, 3], nthreads=nt)
        res.append(o.counters())
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]) if k != "volume_total" else np.isclose(res[0][k], res[1][k])


@pytest.mark.parametrize("triclinic", [False, True])
@pytest.mark.parametrize("bulk_range", [(5.0, 7.0), (12.0, 16.0), (15.0, 20.0)])
def test_cell_list_paths_equal_brute_force(triclinic, bulk_range):
    """the CPU baseline's cell list (cells of cut/2, wrapped-coordinate pre-test with per-cell periodic shifts; plain
    27..125-cell scan when an axis has <= 5 cells) visits exactly the pairs of the brute-force search: identical
    lists and counters, unwrapped coordinates, several molecules of solute (random phase included)."""
    from cmx_b200 import synthetic as syn
    cell = np.array([[46.0, 9.0, 6.0], [0.0, 44.0, 8.0], [0.0, 0.0, 43.0]]) if triclinic else np.diag([44.0, 47.0, 42.0])
    s = syn.make_system("t", cell=cell, solute_atoms=150, solvents=[("water", "water", 250), ("co", "urea", 30)], seed=9)
    x, cell = s.frame(2)                              # frame(k) leaves the molecules unwrapped by up to +-2 cells
    sol, co, wat = s.selections["solute"], s.selections["co"], s.selections["water"]
    opt = cm.Options(bulk_range=bulk_range, n_random_samples=2, silent=True, seed=7)
    for solute, solvent, auto in ((sol, wat, False), (co, wat, False), (co, co, True)):
        p = Problem(solute, solvent, opt, [x[solute.indices - 1]], None if auto else [x[solvent.indices - 1]], cell, autocorrelation=auto)
        a, la = p.oracle(use_clist=False, want_lists=True)
        b, lb = p.oracle(use_clist=True, want_lists=True)
        for k in range(len(la[0][0])):
            assert np.array_equal(la[0][0][k], lb[0][0][k])
        for k in range(len(la[0][1])):
            assert np.array_equal(la[0][1][k], lb[0][1][k])
        ca, cb = a.counters(), b.counters()
        assert all(np.array_equal(ca[key], cb[key]) for key in ca if key != "volume_total")
        assert ca["md_count"].sum() > 0
