"""GPU tests of the frame feed: the DEVICE decoder of GROMACS XTC frames against the host decoder, on a
real GROMACS-written frame (tests/golden/nucleic_frame0.xtc = the first frame of the reference's
test/data/nucleic/trajectory.xtc, extracted by tests/golden/make_golden_xtc.py) and on frames written by the
test-suite's independent XTC writer; the native feeds with device decoding and with partial DCD record reads."""
import json
import os

import numpy as np
import pytest

import cmx_b200 as cm
from common import COUNTER_KEYS, Problem, assert_counters_equal, namd, write_dcd, write_xtc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_xtc_device_decoder_on_a_gromacs_frame():
    """95 988 atoms, 29 813 groups: device == host decoder bit for bit, == the committed summary of the fixture."""
    from cmx_b200.engine import XtcFile
    x = XtcFile(os.path.join(GOLDEN, "nucleic_frame0.xtc"))
    assert (x.natoms, x.nframes) == (95988, 1)
    host, cell, step, time = x.read_frame(0)
    dev, cell_d = x.read_frame_device(0)
    x.close()
    assert np.array_equal(host.view(np.uint32), dev.view(np.uint32))
    assert np.array_equal(cell, cell_d)
    fr = json.load(open(os.path.join(GOLDEN, "xtc_nucleic.json")))["frames"][0]
    assert np.array_equal(dev[:4].astype(float), np.array(fr["first_atoms"])) and np.array_equal(dev[-3:].astype(float), np.array(fr["last_atoms"]))
    assert np.allclose(np.sum(dev.astype(np.float64), axis=0), fr["sum"], rtol=1e-12)
    # TIP3P water of the fixture comes out rigid (the physics pin of the host decoder, src of the numbers: test_host.py)
    w = dev[-3000:].astype(np.float64).reshape(-1, 3, 3)
    assert abs(np.linalg.norm(w[:, 1] - w[:, 0], axis=1).mean() - 0.9572) < 2e-3


def test_xtc_device_decoder_on_written_frames(tmp_path):
    """water-like runs (swapped first pair, adaptive small range), unordered atoms (no runs), a mixture, coordinates
    beyond the 24-bit range (components stored separately: high precision), negative coordinates, several frames."""
    from cmx_b200.engine import XtcFile
    rng = np.random.default_rng(5)

    def water_box(nmol, L):
        o = rng.uniform(0, L, size=(nmol, 1, 3))
        return np.concatenate([o, o + rng.normal(0, 0.06, size=(nmol, 2, 3))], axis=1).reshape(-1, 3)
    cases = {"water": ([water_box(3000, 6.0), water_box(3000, 6.0)], 1000.0),
             "random": ([rng.uniform(-2, 9, size=(5000, 3))], 1000.0),
             "mixed": ([np.concatenate([rng.uniform(0, 5, size=(37, 3)), water_box(500, 5.0), rng.uniform(0, 5, size=(11, 3))])], 1000.0),
             "wide": ([rng.uniform(-3, 14, size=(700, 3))], 1.0e7)}
    for name, (frames, precision) in cases.items():
        path = str(tmp_path / f"{name}.xtc")
        boxes = np.stack([np.eye(3) * 6.0] * len(frames))
        write_xtc(path, np.stack(frames), boxes, precision=precision)
        x = XtcFile(path)
        for k in range(len(frames)):
            host = x.read_frame(k)[0]
            dev = x.read_frame_device(k)[0]
            assert np.array_equal(host.view(np.uint32), dev.view(np.uint32)), (name, k)
            assert np.abs(host - 10.0 * frames[k]).max() < 10.0 * (0.51 / precision) + 1e-5 * 140, name
        x.close()


def test_xtc_feed_device_decode_equals_host_decode(tmp_path):
    """cmx_run_xtc with the device decoder (default) == the same with decoding reader threads == the oracle"""
    from cmx_b200.engine import Engine, XtcFile
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)                 # 3997 atoms, Angstrom
    boxes = np.stack([np.asarray(c, dtype=np.float64).T / 10.0 for c in d["cells"]])
    path = str(tmp_path / "c.xtc")
    write_xtc(path, frames.astype(np.float64) / 10.0, boxes)
    x = XtcFile(path)
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    opt = cm.Options(bulk_range=(8.0, 10.0), n_random_samples=3, seed=321, silent=True)
    order = [0, 1, 2, 1, 0, 2, 2, 1]
    out = {}
    for mode in ("device", "host"):
        eng = Engine(solute=sol, solvent=tm, options=opt, irefatom=1, autocorrelation=False)
        eng.set_option("xtc_host_decode", 1.0 if mode == "host" else 0.0)
        eng.run_xtc(x, sol.indices, tm.indices, order, n_reader_threads=3)
        out[mode] = (eng.finish(), eng.stats()["h2d_bytes"])
        eng.close()
    for k in COUNTER_KEYS:
        assert np.array_equal(out["device"][0][k], out["host"][0][k]), k
    assert out["device"][1] < 0.8 * out["host"][1]          # the compressed frames (+ group records) travel, not 12 B/atom
    dec = [x.read_frame(k) for k in range(3)]
    x.close()
    p = Problem(sol, tm, opt, [dec[k][0][:1463] for k in order], [dec[k][0][1463:] for k in order], [dec[k][1] for k in order],
                frame_ids=[k + 1 for k in order], irefatom=1)
    assert_counters_equal(out["device"][0], p.oracle()[0])


def test_dcd_feed_reads_records_up_to_the_last_selected_atom(tmp_path):
    """the selections sit in the first sixth of the file (as protein + cosolvent do in a solvated system): the reader
    threads pread -- and the copy engine moves -- only the head of the X, Y, Z records (lastatom,
    src/trajectory_formats/NamdDCD.jl:86-118); same counters as the staging-slot path."""
    from cmx_b200.engine import DcdFile, Engine
    d = namd()
    nf = d["protein"].shape[0]
    rng = np.random.default_rng(3)
    tail = rng.uniform(0, 80, size=(nf, 20000, 3)).astype(np.float32)          # "water" that is not selected
    frames = np.concatenate([d["protein"], d["tmao"], tail], axis=1)
    path = str(tmp_path / "head.dcd")
    write_dcd(path, frames, d["cells"])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    opt = cm.Options(bulk_range=(8.0, 10.0), n_random_samples=2, seed=321, silent=True)
    f = DcdFile(path)
    eng = Engine(solute=sol, solvent=tm, options=opt, irefatom=1, autocorrelation=False)
    order = [0, 1, 2, 2, 0, 1, 1]
    eng.run_dcd(f, sol.indices, tm.indices, order, n_reader_threads=2)
    dev = eng.finish()
    h2d = eng.stats()["h2d_bytes"]
    eng.close()
    assert h2d == len(order) * 3 * (4 + 4 * 3997) and h2d < 0.2 * len(order) * f.frame_bytes
    cells = [f.read_frame(k)[1] for k in range(3)]
    f.close()
    p = Problem(sol, tm, opt, [d["protein"][k] for k in order], [d["tmao"][k] for k in order], [cells[k] for k in order],
                frame_ids=[k + 1 for k in order], irefatom=1)
    assert_counters_equal(dev, p.oracle()[0])
