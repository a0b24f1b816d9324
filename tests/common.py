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


# ---- XTC writer with the compressed coordinate block (test infrastructure; pure Python, serial) -----------------------
_XTC_MAGICINTS = [0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
                  645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
                  32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561,
                  832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607,
                  10568983, 13316085, 16777216]
_XTC_FIRSTIDX, _XTC_LASTIDX = 9, len(_XTC_MAGICINTS) - 1


class _BitWriter:
    def __init__(self):
        self.out, self.acc, self.nacc = bytearray(), 0, 0

    def bits(self, nbits, value):
        self.acc = (self.acc << nbits) | (value & ((1 << nbits) - 1))
        self.nacc += nbits
        while self.nacc >= 8:
            self.nacc -= 8
            self.out.append((self.acc >> self.nacc) & 0xff)
        self.acc &= (1 << self.nacc) - 1

    def ints3(self, nbits, sizes, nums):
        """three integers as one mixed-radix number (radices `sizes`), sent least-significant byte first"""
        v = (nums[0] * sizes[1] + nums[1]) * sizes[2] + nums[2]
        nbytes = []
        while v:
            nbytes.append(v & 0xff); v >>= 8
        # the reader takes whole bytes while more than 8 bits remain, then the rest in one piece
        k, left = 0, nbits
        while left > 8:
            self.bits(8, nbytes[k] if k < len(nbytes) else 0); k += 1; left -= 8
        if left > 0:
            self.bits(left, nbytes[k] if k < len(nbytes) else 0)

    def bytes(self):
        if self.nacc:
            return bytes(self.out) + bytes([(self.acc << (8 - self.nacc)) & 0xff])
        return bytes(self.out)


def _xtc_sizeofint(size):
    num, nbits = 1, 0
    while size >= num and nbits < 32:
        nbits += 1; num <<= 1
    return nbits


def _xtc_sizeofints(sizes):
    return max(1, (sizes[0] * sizes[1] * sizes[2]).bit_length())


def xtc_compress(ints, precision=1000.0):
    """the compressed coordinate block of one frame from quantised coordinates ints[natoms,3] (more than 9 atoms)."""
    import struct
    M = _XTC_MAGICINTS
    c = [list(map(int, r)) for r in np.asarray(ints)]
    n = len(c)
    mn = [min(r[k] for r in c) for k in range(3)]; mx = [max(r[k] for r in c) for k in range(3)]
    sizeint = [mx[k] - mn[k] + 1 for k in range(3)]
    if max(sizeint) > 0xffffff:
        bitsizeint, bitsize = [_xtc_sizeofint(s) for s in sizeint], 0
    else:
        bitsizeint, bitsize = None, _xtc_sizeofints(sizeint)
    mindiff = min((sum(abs(c[i][k] - c[i - 1][k]) for k in range(3)) for i in range(1, n)), default=0x7fffffff)
    smallidx = _XTC_FIRSTIDX
    while smallidx < _XTC_LASTIDX and M[smallidx] < mindiff:
        smallidx += 1
    smallidx0 = smallidx
    maxidx = min(_XTC_LASTIDX, smallidx + 8); minidx = maxidx - 8
    smaller, smallnum, larger = M[max(_XTC_FIRSTIDX, smallidx - 1)] // 2, M[smallidx] // 2, M[maxidx] // 2
    w = _BitWriter()
    i, prevrun, prev = 0, -1, [0, 0, 0]
    close = lambda a, b, lim: all(abs(a[k] - b[k]) < lim for k in range(3))
    while i < n:
        is_small = 0
        if smallidx < maxidx and i >= 1 and close(c[i], prev, larger):
            is_smaller = 1
        elif smallidx > minidx:
            is_smaller = -1
        else:
            is_smaller = 0
        if i + 1 < n and close(c[i], c[i + 1], smallnum):
            c[i], c[i + 1] = c[i + 1], c[i]          # first and second atom of a run are stored swapped
            is_small = 1
        tmp = [c[i][k] - mn[k] for k in range(3)]
        if bitsize == 0:
            for k in range(3):
                w.bits(bitsizeint[k], tmp[k])
        else:
            w.ints3(bitsize, sizeint, tmp)
        prev = c[i]; i += 1
        run, small = 0, []
        if is_small == 0 and is_smaller == -1:
            is_smaller = 0
        while is_small and run < 8 * 3:
            if is_smaller == -1 and sum((c[i][k] - prev[k]) ** 2 for k in range(3)) >= smaller * smaller:
                is_smaller = 0
            small.append([c[i][k] - prev[k] + smallnum for k in range(3)]); run += 3
            prev = c[i]; i += 1
            is_small = 1 if (i < n and close(c[i], prev, smallnum)) else 0
        if run != prevrun or is_smaller != 0:
            prevrun = run
            w.bits(1, 1); w.bits(5, run + is_smaller + 1)
        else:
            w.bits(1, 0)
        sz = [M[smallidx]] * 3
        for t in small:
            w.ints3(smallidx, sz, t)
        if is_smaller != 0:
            smallidx += is_smaller
            if is_smaller < 0:
                smallnum = smaller; smaller = M[smallidx - 1] // 2
            else:
                smaller = smallnum; smallnum = M[smallidx] // 2
    body = w.bytes()
    pad = (-len(body)) % 4
    return (struct.pack(">f3i3ii", precision, *mn, *mx, smallidx0) + struct.pack(">i", len(body)) + body + b"\0" * pad)


def write_xtc(path, frames_nm, boxes_nm, precision=1000.0, steps=None, dt=2.0):
    """XTC writer (any number of atoms): coordinates are quantised to round(x * precision) as the format does."""
    import struct
    frames_nm = np.asarray(frames_nm, dtype=np.float32)
    nf, n = frames_nm.shape[0], frames_nm.shape[1]
    if n <= 9:
        return write_xtc_small(path, frames_nm, boxes_nm, steps=steps, dt=dt)
    quant = []
    with open(path, "wb") as f:
        for k in range(nf):
            f.write(struct.pack(">iiif", 1995, n, k * 10 if steps is None else steps[k], k * dt))
            f.write(struct.pack(">9f", *np.asarray(boxes_nm[k], dtype=np.float32).reshape(9)))
            f.write(struct.pack(">i", n))
            x = frames_nm[k].astype(np.float64) * precision
            ints = np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)
            quant.append(ints)
            f.write(xtc_compress(ints, precision))
    return quant
