"""The two-phase restructuring of the XTC decoder planned for the device (oracle/xtc_two_phase.c: serial skeleton walk ->
one byte per group; prefix sums; independent per-group decode) against the product's serial host decoder
(complexmixtures.jl_b200/csrc/cmx_xtc.inl) -- CPU only."""
import os
import struct

import numpy as np
import pytest

from common import write_xtc
from cmx_b200.engine import XtcFile
from oracle import cmx_oracle as orc


def _blocks(path):
    """(natoms, [coordinate block bytes of every frame]) of an XTC file with more than 9 atoms per frame"""
    raw = open(path, "rb").read()
    out, off, natoms = [], 0, None
    while off + 56 + 36 <= len(raw):
        magic, natoms = struct.unpack(">ii", raw[off:off + 8])
        assert magic == 1995
        nbytes = struct.unpack(">i", raw[off + 56 + 32:off + 56 + 36])[0]
        ln = 36 + ((nbytes + 3) // 4) * 4
        out.append(raw[off + 56:off + 56 + ln])
        off += 56 + ln
    return natoms, out


def _check(path):
    natoms, blocks = _blocks(path)
    f = XtcFile(path)
    assert f.nframes == len(blocks) and f.natoms == natoms
    ngroups = []
    for k, blk in enumerate(blocks):
        want, _, _, _ = f.read_frame(k)
        got, codes = orc.xtc_two_phase_decode(blk, natoms)
        assert np.array_equal(got, want), (path, k)
        ngroups.append(len(codes))
    f.close()
    return natoms, ngroups


def test_two_phase_equals_serial_decoder_synthetic(tmp_path):
    rng = np.random.default_rng(4)

    def water_box(nmol, L):
        o = rng.uniform(0, L, size=(nmol, 1, 3))
        return np.concatenate([o, o + rng.normal(0, 0.06, size=(nmol, 2, 3))], axis=1).reshape(-1, 3)
    cases = {"water": [water_box(700, 4.0), water_box(700, 4.0)], "random": [rng.uniform(-3, 12, size=(800, 3))],
             "mixed": [np.concatenate([rng.uniform(0, 6, size=(53, 3)), water_box(200, 6.0), rng.uniform(0, 6, size=(17, 3))])],
             "big": [rng.uniform(-9000, 9000, size=(64, 3))], "chain": [np.cumsum(rng.normal(0, 0.05, size=(600, 3)), axis=0) + 3.0]}
    box = np.diag([6.0, 6.0, 6.0])
    for name, frames in cases.items():
        path = str(tmp_path / f"{name}.xtc")
        write_xtc(path, np.stack(frames), np.stack([box] * len(frames)))
        natoms, ngroups = _check(path)
        assert all(0 < g <= natoms for g in ngroups)


def test_two_phase_equals_serial_decoder_reference_fixture():
    src = "/root/reference/test/data/nucleic/trajectory.xtc"
    if not os.path.exists(src):
        pytest.skip("reference fixture not present on this machine")
    natoms, ngroups = _check(src)
    # one skeleton byte per group: the extra H2D next to the compressed block
    assert natoms == 95988 and max(ngroups) < 0.4 * natoms
