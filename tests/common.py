"""Shared helpers of the test-suite: fixtures, oracle/engine drivers and comparisons."""
import json
import os

import numpy as np

import cmx_b200 as cm
from oracle import cmx_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COUNTER_KEYS = ("md_count", "md_count_random", "rdf_count", "rdf_count_random", "solute_group_count",
                "solute_group_count_random", "solvent_group_count", "solvent_group_count_random")


def namd():
    return np.load(os.path.join(GOLDEN, "namd_fixture.npz"))


def toy():
    return np.load(os.path.join(GOLDEN, "toy.npz"))


def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


def default_irefatom(xsolvent, napm):
    """TrajectoryMetaData default (src/Trajectory.jl:206-213), 1-based."""
    first = np.asarray(xsolvent[:napm], dtype=np.float64)
    return int(np.argmin(np.linalg.norm(first - first.mean(axis=0), axis=1))) + 1


class Problem:
    """Selections + options + frames in memory; runs the oracle and (on a GPU) the engine."""

    def __init__(self, solute, solvent, options, frames_solute, frames_solvent, cells, *, autocorrelation=False,
                 irefatom=None, weights=None, frame_ids=None, coordination_number_only=False):
        self.solute, self.solvent, self.options = solute, solvent, options
        self.auto = autocorrelation
        self.xs = [np.ascontiguousarray(f, dtype=np.float32) for f in frames_solute]
        self.xv = self.xs if autocorrelation else [np.ascontiguousarray(f, dtype=np.float32) for f in frames_solvent]
        nf = len(self.xv)
        self.cells = [np.asarray(cells, dtype=np.float64)] * nf if np.ndim(cells) <= 2 else [np.asarray(c, dtype=np.float64) for c in cells]
        self.weights = [1.0] * nf if weights is None else list(weights)
        self.frame_ids = list(range(1, nf + 1)) if frame_ids is None else list(frame_ids)
        self.irefatom = irefatom if irefatom is not None else (
            options.irefatom if options.irefatom > 0 else default_irefatom(self.xv[0], solvent.natomspermol))
        self.cn_only = coordination_number_only

    def oracle(self, use_clist=True, want_lists=False):
        o = orc.Oracle.from_problem(self.solute, self.solvent, self.options, self.irefatom, self.auto, self.cn_only)
        lists = []
        for xs, xv, cell, w, fid in zip(self.xs, self.xv, self.cells, self.weights, self.frame_ids):
            if w == 0:
                continue
            lists.append(o.frame(xs, xv, cell, weight=w, frame_index=fid, use_clist=use_clist, want_lists=want_lists))
        return o, lists

    def engine(self, **kw):
        from cmx_b200.engine import Engine
        return Engine(solute=self.solute, solvent=self.solvent, options=self.options, irefatom=self.irefatom,
                      autocorrelation=self.auto, coordination_number_only=self.cn_only, **kw)

    def run_engine(self, eng, frames=None):
        for k, (xs, xv, cell, w, fid) in enumerate(zip(self.xs, self.xv, self.cells, self.weights, self.frame_ids)):
            if w == 0 or (frames is not None and k not in frames):
                continue
            eng.submit_arrays(xs, xv, cell, frame_index=fid, weight=w)
        return eng.finish()


def assert_counters_equal(dev: dict, oracle, *, exact=True, rtol=0.0):
    ref = oracle.counters()
    for k in COUNTER_KEYS:
        a, b = np.asarray(dev[k]), np.asarray(ref[k])
        assert a.shape == b.shape, k
        if exact:
            bad = np.argwhere(a != b)
            assert len(bad) == 0, f"{k}: {len(bad)} cells differ, first {bad[:5].tolist()} dev={a[tuple(bad[0])]} ref={b[tuple(bad[0])]}"
        else:
            np.testing.assert_allclose(a, b, rtol=rtol, atol=0, err_msg=k)
    np.testing.assert_allclose(dev["volume_total"], ref["volume_total"], rtol=1e-13)


def assert_lists_equal(dev, ref, *, dtol=0.0, what="list"):
    """dev: engine MD array (1-based i/j, 0 = empty); ref: oracle MD array (0-based, -1 = empty)."""
    assert len(dev) == len(ref)
    w_dev, w_ref = dev["within_cutoff"] != 0, ref["within_cutoff"] != 0
    bad = np.flatnonzero(w_dev != w_ref)
    assert len(bad) == 0, f"{what}: within_cutoff differs for molecules {bad[:10].tolist()}"
    m = w_ref
    assert np.array_equal(dev["i"][m], ref["i"][m] + 1), f"{what}: i differs at {np.flatnonzero(m)[dev['i'][m] != ref['i'][m] + 1][:10]}"
    assert np.array_equal(dev["j"][m], ref["j"][m] + 1), f"{what}: j differs"
    dd = np.abs(dev["d"][m] - ref["d"][m])
    assert dd.size == 0 or dd.max() <= dtol, f"{what}: max |d - d_ref| = {dd.max()}"
    r_dev, r_ref = dev["ref_atom_within_cutoff"] != 0, ref["ref_atom_within_cutoff"] != 0
    assert np.array_equal(r_dev[m], r_ref[m]), f"{what}: ref_atom_within_cutoff differs"
    mr = m & r_ref
    dr = np.abs(dev["d_ref_atom"][mr] - ref["d_ref_atom"][mr])
    assert dr.size == 0 or dr.max() <= dtol, f"{what}: max |d_ref - d_ref_ref| = {dr.max()}"


def write_dcd(path, frames, cells):
    """Minimal NAMD/CHARMM DCD writer with unit-cell records (layout read by src/trajectory_formats/NamdDCD.jl:
    header, title, natoms; per frame [A, gamma, B, beta, alpha, C] + x, y, z Float32 records)."""
    import struct
    frames = np.asarray(frames, dtype=np.float32)
    nf, n = frames.shape[0], frames.shape[1]

    def rec(f, payload):
        f.write(struct.pack("<i", len(payload))); f.write(payload); f.write(struct.pack("<i", len(payload)))
    with open(path, "wb") as f:
        icntrl = [0] * 20
        icntrl[0], icntrl[1], icntrl[2], icntrl[3], icntrl[10], icntrl[19] = nf, 0, 1, nf, 1, 24
        hdr = b"CORD" + struct.pack("<9i", *icntrl[:9]) + struct.pack("<f", 1.0) + struct.pack("<10i", *icntrl[10:])
        rec(f, hdr)
        rec(f, struct.pack("<i", 1) + b"written by the cmx-b200 test-suite".ljust(80))
        rec(f, struct.pack("<i", n))
        for k in range(nf):
            c = np.asarray(cells[k] if np.ndim(cells) == 3 else cells, dtype=np.float64)
            a, b, cc = c[:, 0], c[:, 1], c[:, 2]
            A, B, C_ = np.linalg.norm(a), np.linalg.norm(b), np.linalg.norm(cc)
            ang = lambda u, v: np.degrees(np.arccos(np.clip(u @ v / np.linalg.norm(u) / np.linalg.norm(v), -1, 1)))
            rec(f, struct.pack("<6d", A, ang(a, b), B, ang(a, cc), ang(b, cc), C_))
            for d in range(3):
                rec(f, frames[k, :, d].astype("<f4").tobytes())


def write_xtc_small(path, frames_nm, boxes_nm, steps=None, dt=2.0):
    """XTC writer for AT MOST 9 atoms per frame: such frames are stored by the format as plain big-endian floats
    (no compression), which makes a writer a few lines.  frames_nm [nframes, natoms, 3] and boxes_nm [nframes, 3, 3]
    (rows = box vectors) in nm."""
    import struct
    frames_nm = np.asarray(frames_nm, dtype=np.float32)
    nf, n = frames_nm.shape[0], frames_nm.shape[1]
    assert n <= 9
    with open(path, "wb") as f:
        for k in range(nf):
            f.write(struct.pack(">iiif", 1995, n, k * 10 if steps is None else steps[k], k * dt))
            f.write(struct.pack(">9f", *np.asarray(boxes_nm[k], dtype=np.float32).reshape(9)))
            f.write(struct.pack(">i", n))
            f.write(frames_nm[k].astype(">f4").tobytes())
