"""GPU tests of the step right after the hot path, done on the device: cmx_final_results
(finalresults!, src/results.jl:311-469) and cmx_contributions (src/tools/contributions.jl:70-248, the matrix of
ResidueContributions) against the oracle's numpy finalresults on the oracle's counters.

Tolerance: fp64 with the reference's operation order; the only differences are pow() of CUDA vs libm in shellradius
(<= 2 ulp), the summation order of the four scalar sums (pairwise in numpy, tree here) and warp-scan vs serial cumulative
sums in the contributions.  rtol 1e-12 (north_star: final mddf / KB integrals within 1e-3 relative)."""
import numpy as np
import pytest

import cmx_b200 as cm
from common import Problem, namd
from oracle import cmx_oracle as orc

pytestmark = pytest.mark.gpu

PROTEIN = cm.AtomSelection(np.arange(1, 1464), nmols=1)
TMAO = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)
RTOL = 1e-12
VECS = ("d", "md_count", "md_count_random", "coordination_number", "coordination_number_random", "mddf", "kb", "rdf_count",
        "rdf_count_random", "sum_rdf_count", "sum_rdf_count_random", "rdf", "kb_rdf", "volume_shell")
SCALARS = ("volume_total", "volume_bulk", "volume_domain", "density_solute", "density_solvent", "density_solvent_bulk")


def opts(**kw):
    kw.setdefault("silent", True)
    kw.setdefault("seed", 321)
    return cm.Options(**kw)


def close(a, b, what, atol=1e-300):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=atol, err_msg=what)


def kbi_atol(ref, side, groups):
    """:kbi is a difference of two cumulative sums: where they cancel, the rounding of the sums (1e-16 of THEIR size)
    is all that is left, so the bar is absolute, 1e-12 of the size of the terms"""
    gc, gcr = getattr(ref, side + "_group_count"), getattr(ref, side + "_group_count_random")
    big = max(max(float(gc[g].sum()), float(gcr[g].sum())) if len(g) else 0.0 for g in groups)
    return 1e-12 * orc.ANGS3_TO_CM3_PER_MOL / ref.density_solvent_bulk * max(big, 1e-300)


def oracle_final(p, o, Q):
    c = o.counters()
    opt = p.options
    return c, orc.finalresults(c, nmols_solute=p.solute.nmols, nmols_solvent=p.solvent.nmols, autocorrelation=p.auto,
                               n_random_samples=opt.n_random_samples, binstep=opt.binstep, dbulk=opt.dbulk, cutoff=opt.cutoff,
                               usecutoff=opt.usecutoff, Q=Q, coordination_number_only=p.cn_only)


def check_final(p, fin, ref):
    for k in VECS:
        if hasattr(ref, k):
            atol = 1e-300
            if k in ("kb", "kb_rdf"):      # differences of cumulative sums: absolute bar where they cancel (see kbi_atol)
                cs = ref.coordination_number if k == "kb" else ref.sum_rdf_count
                atol = 1e-12 * orc.ANGS3_TO_CM3_PER_MOL / ref.density_solvent_bulk * float(np.abs(cs).max())
            close(fin[k], getattr(ref, k), k, atol=atol)
    for k in SCALARS:
        if hasattr(ref, k):
            close(fin[k], getattr(ref, k), k)


def contrib_ref(ref, side, rows, type):
    """contributions() of the reference on the oracle's final arrays, for a group given as rows of the group array"""
    gc = getattr(ref, side + "_group_count")
    sel = gc[rows].sum(axis=0)
    if type == "md_count":
        return sel
    if type == "coordination_number":
        return np.cumsum(sel)
    if type == "mddf":
        return np.where(ref.md_count_random != 0.0, sel / np.where(ref.md_count_random != 0.0, ref.md_count_random, 1.0), 0.0)
    selr = getattr(ref, side + "_group_count_random")[rows].sum(axis=0)
    return orc.ANGS3_TO_CM3_PER_MOL * (1 / ref.density_solvent_bulk) * (np.cumsum(sel) - np.cumsum(selr))


@pytest.mark.parametrize("weights", [None, [1.0, 0.3, 2.5]])
def test_final_results_and_contributions_protein_tmao(weights):
    """C1 protein x TMAO, unit and varying frame weights (the fp64 twin of the accumulators takes part)."""
    d = namd()
    p = Problem(PROTEIN, TMAO, opts(bulk_range=(8.0, 10.0), n_random_samples=5), d["protein"], d["tmao"], d["cells"], weights=weights, irefatom=1)
    o, _ = p.oracle()
    eng = p.engine()
    p.run_engine(eng)
    Q = float(sum(p.weights))
    c, ref = oracle_final(p, o, Q)
    fin = eng.final_results()
    assert fin["sum_weights"] == Q
    check_final(p, fin, ref)
    assert np.all(fin["mddf"][ref.md_count_random == 0] == 0.0)
    # residue-like groups of the protein's per-atom rows (disjoint, ragged, one empty) and single TMAO atom types
    rng = np.random.default_rng(3)
    cuts = np.sort(rng.choice(np.arange(1, 1463), size=90, replace=False))
    groups = [np.arange(a, b) for a, b in zip(np.r_[0, cuts], np.r_[cuts, 1463])] + [np.zeros(0, dtype=np.int32)]
    for type in ("mddf", "coordination_number", "md_count", "kbi"):
        got = eng.contributions("solute", groups, type)
        want = np.stack([contrib_ref(ref, "solute", g, type) for g in groups])
        close(got, want, f"solute {type}", atol=kbi_atol(ref, "solute", groups) if type == "kbi" else 1e-300)
        vgroups = [[k] for k in range(14)] + [list(range(14))]
        gotv = eng.contributions("solvent", vgroups, type)
        wantv = np.stack([contrib_ref(ref, "solvent", g, type) for g in vgroups])
        close(gotv, wantv, f"solvent {type}", atol=kbi_atol(ref, "solvent", vgroups) if type == "kbi" else 1e-300)
    # all atoms of a side together give the total distribution (src/tools/contributions.jl:320-348)
    close(eng.contributions("solute", [np.arange(1463)], "mddf")[0], fin["mddf"], "sum of solute contributions")
    close(eng.contributions("solvent", [np.arange(14)], "coordination_number")[0], fin["coordination_number"], "sum of solvent contributions")
    # a caller-supplied Q / volume sum (what a multi-process driver passes after its all-reduce)
    fin2 = eng.final_results(sum_weights=2 * Q, volume_sum=2 * c["volume_total"])
    close(fin2["volume_total"], fin["volume_total"], "volume with explicit sums")
    close(fin2["md_count"], fin["md_count"] / 2, "md_count with explicit Q")
    from cmx_b200.engine import CmxError
    with pytest.raises(CmxError):
        eng.contributions("solute", [np.array([1463])], "mddf")
    eng.close()


def test_final_results_autocorrelation_and_usecutoff():
    """TMAO self-correlation (group counts carry w/2, solvent groups == solute groups, samples nmols-1) with
    usecutoff = true (bulk = the shell between dbulk and cutoff)."""
    d = namd()
    sel = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)
    p = Problem(sel, sel, opts(bulk_range=(6.0, 9.0), n_random_samples=3), d["tmao"], d["tmao"], d["cells"], autocorrelation=True, irefatom=1)
    assert p.options.usecutoff
    o, _ = p.oracle()
    eng = p.engine()
    p.run_engine(eng)
    c, ref = oracle_final(p, o, 3.0)
    fin = eng.final_results()
    check_final(p, fin, ref)
    for type in ("mddf", "coordination_number", "md_count", "kbi"):
        for side in ("solute", "solvent"):
            gs = [[0], [1, 5, 13], list(range(14))]
            got = eng.contributions(side, gs, type)
            want = np.stack([contrib_ref(ref, "solute", g, type) for g in gs])
            close(got, want, f"{side} {type}", atol=kbi_atol(ref, "solute", gs) if type == "kbi" else 1e-300)
    eng.close()


def test_final_results_coordination_number_only():
    d = namd()
    p = Problem(PROTEIN, TMAO, opts(bulk_range=(8.0, 10.0)), d["protein"], d["tmao"], d["cells"], coordination_number_only=True, irefatom=1)
    o, _ = p.oracle()
    eng = p.engine()
    p.run_engine(eng)
    c, ref = oracle_final(p, o, 3.0)
    fin = eng.final_results()
    check_final(p, fin, ref)
    assert not fin["mddf"].any() and not fin["kb"].any() and not fin["md_count_random"].any()
    close(eng.contributions("solvent", [[0, 1]], "coordination_number")[0], contrib_ref(ref, "solvent", [0, 1], "coordination_number"), "cn")
    from cmx_b200.engine import CmxError
    with pytest.raises(CmxError):
        eng.contributions("solvent", [[0]], "mddf")
    eng.close()


def test_device_final_results_match_the_public_result():
    """the Result of the public mddf() (host finalresults) and the device evaluation of the same run"""
    d = namd()
    tr = cm.ArrayTrajectory(np.concatenate([d["protein"], d["tmao"]], axis=1), d["cells"],
                            cm.AtomSelection(np.arange(1, 1464), nmols=1), cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14))
    o = opts(bulk_range=(8.0, 10.0), n_random_samples=4)
    cache = {}
    R = cm.mddf(tr, o, _engine_cache=cache)
    (eng,) = cache.values()
    fin = eng.final_results()
    for k in ("d", "md_count", "md_count_random", "coordination_number", "coordination_number_random", "mddf", "kb", "rdf", "kb_rdf"):
        close(fin[k], getattr(R, k), k, atol=1e-9 if k in ("kb", "kb_rdf") else 1e-300)
    close(fin["volume_domain"], R.volume.domain, "volume.domain"); close(fin["density_solvent_bulk"], R.density.solvent_bulk, "density.solvent_bulk")
    got = eng.contributions("solvent", [[0], [3]], "mddf")
    close(got[0], cm.contributions(R, cm.SolventGroup([int(tr.solvent.indices[0])]), type="mddf") * tr.solvent.nmols, "SolventGroup first atom type")
    eng.close()


def test_group_arrays_left_on_the_device():
    """mddf(..., group_arrays=False): the O(nbins) Result is the same, the group arrays never travel, and R.device
    evaluates contributions() on the device == the host contributions() of the full Result."""
    d = namd()
    mk = lambda: cm.ArrayTrajectory(np.concatenate([d["protein"], d["tmao"]], axis=1), d["cells"],
                                    cm.AtomSelection(np.arange(1, 1464), nmols=1), cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14))
    o = opts(bulk_range=(8.0, 10.0), n_random_samples=4)
    Rfull = cm.mddf(mk(), o, frame_weights=[1.0, 2.0, 0.5])
    R = cm.mddf(mk(), o, frame_weights=[1.0, 2.0, 0.5], group_arrays=False)
    assert R.solute_group_count.shape == (0, R.nbins) and R.solvent_group_count_random.shape == (0, R.nbins)
    for k in ("md_count", "md_count_random", "coordination_number", "mddf", "kb", "rdf", "kb_rdf"):
        assert np.array_equal(getattr(R, k), getattr(Rfull, k)), k
    residues = [list(range(a, min(a + 16, 1463))) for a in range(0, 1463, 16)]
    for type in ("mddf", "coordination_number", "md_count"):
        got = R.device.contributions("solute", residues, type)
        want = np.stack([cm.contributions(Rfull, cm.SoluteGroup([r + 1 for r in res]), type=type) for res in residues])
        close(got, want, f"residue {type}")
    got = R.device.contributions("solute", residues[:5], "kbi")
    want = np.stack([cm.contributions(Rfull, cm.SoluteGroup([r + 1 for r in res]), type="kbi") for res in residues[:5]])
    close(got, want, "residue kbi", atol=1e-9)
    fin = R.device.final_results()
    close(fin["mddf"], Rfull.mddf, "mddf")
    R.device.close()


def test_c_example_runs_on_the_gpu_and_agrees_with_the_python_host(tmp_path):
    """examples/mddf_dcd.c (plain C over the ABI: cmx_run_dcd -> cmx_finish -> cmx_final_results -> cmx_contributions)
    prints the same numbers as the Python host obtains for the same run."""
    import os, re, subprocess
    from common import write_dcd
    from cmx_b200 import engine
    from cmx_b200.engine import DcdFile, Engine
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "mddf_dcd")
    subprocess.run(["/usr/bin/gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "mddf_dcd.c"),
                    "-L", os.path.dirname(engine.LIB_PATH), "-lcmx_b200", "-Wl,-rpath," + os.path.dirname(engine.LIB_PATH),
                    "-Wl,--allow-shlib-undefined", "-o", exe], check=True)
    d = namd()
    path = str(tmp_path / "t.dcd")
    write_dcd(path, np.concatenate([d["protein"], d["tmao"]], axis=1), d["cells"])
    r = subprocess.run([exe, path, "1", "1463", "1464", "2534", "14"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    peak, at, kbv = map(float, re.search(r"mddf peak ([\d.]+) at ([\d.]+) A, KB integral (-?[\d.]+)", r.stdout).groups())
    cn = float(re.search(r"coordination number of the first 731 solute atoms at the cutoff: ([\d.]+)", r.stdout).group(1))
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    eng = Engine(solute=sol, solvent=tm, options=opts(bulk_range=(8.0, 10.0), n_random_samples=10), irefatom=1, autocorrelation=False)
    f = DcdFile(path)
    eng.run_dcd(f, sol.indices, tm.indices, [0, 1, 2], n_reader_threads=2)
    fin = eng.final_results()
    b = int(np.argmax(fin["mddf"]))
    assert abs(peak - fin["mddf"][b]) < 1e-3 and abs(at - fin["d"][b]) < 1e-2 and abs(kbv - fin["kb"][-1]) < 0.1
    assert abs(cn - eng.contributions("solute", [np.arange(731)], "coordination_number")[0, -1]) < 1e-3
    f.close(); eng.close()
