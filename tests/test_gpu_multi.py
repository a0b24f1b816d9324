"""GPU tests of the multi-device paths: the ENGINE sharded over several device contexts against the single-device run.

* one handle, several device contexts (cmx_config.n_devices / device_ids -- the in-library counterpart of the
  reference's chunk tasks + sum!, src/parallel_setup.jl:7-57, src/results.jl:629-649): with one visible GPU the two
  contexts live on the same device (device_ids = [0, 0]), with two or more on different GPUs (peer-access merge);
* one process per GPU under torchrun with the NCCL all-reduce of the public driver (tests/multi_gpu_mddf.py), collected
  here when at least two GPUs are visible.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import cmx_b200 as cm
from common import COUNTER_KEYS, Problem, assert_counters_equal, namd, write_dcd

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROTEIN = cm.AtomSelection(np.arange(1, 1464), nmols=1)
TMAO = cm.AtomSelection(np.arange(1479, 4013), natomspermol=14)


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _devices():
    return [0, 1] if _ngpus() >= 2 else [0, 0]


def _opts(**kw):
    kw.setdefault("silent", True); kw.setdefault("seed", 321)
    return cm.Options(**kw)


def _problem(order, weights=None, nrand=4, solvent=TMAO, auto=False):
    d = namd()
    xs = [d["protein"][k] for k in order]
    xv = [d["tmao"][k] for k in order]
    cells = [d["cells"][k] for k in order]
    if auto:
        return Problem(solvent, solvent, _opts(bulk_range=(8.0, 10.0), n_random_samples=nrand), xv, xv, cells, autocorrelation=True,
                       weights=weights, frame_ids=[k + 1 for k in range(len(order))])
    return Problem(PROTEIN, solvent, _opts(bulk_range=(8.0, 10.0), n_random_samples=nrand), xs, xv, cells, weights=weights,
                   frame_ids=[k + 1 for k in range(len(order))], irefatom=1)


@pytest.mark.parametrize("auto", [False, True])
def test_group_handle_equals_single_device_and_oracle(auto):
    """7 frames dealt to two device contexts of ONE handle == one context == the oracle, bit for bit (integer sums;
    the Philox stream is keyed by the frame index, not by the device)."""
    p = _problem([0, 1, 2, 1, 0, 2, 1], auto=auto)
    single = p.engine(); ref = p.run_engine(single); single.close()
    eng = p.engine(devices=_devices())
    dev = p.run_engine(eng)
    st = eng.stats()
    assert st["frames"] == 7
    # the merged integer block is what a multi-process driver would all-reduce
    eng.reset()
    again = p.run_engine(eng)
    eng.close()
    for k in COUNTER_KEYS:
        assert np.array_equal(dev[k], ref[k]), k
        assert np.array_equal(again[k], ref[k]), k
    assert np.isclose(dev["volume_total"], ref["volume_total"], rtol=1e-14) and dev["sum_weights"] == 7.0
    o, _ = p.oracle()
    assert_counters_equal(dev, o)


def test_group_handle_with_frame_weights():
    """weights that differ between the devices' frames and within them: every context counts in units of the FIRST
    weight of the run (integers) and adds the other weights in f64; the merge moves both blocks to the first device."""
    for weights in ([1.0, 2.0, 1.0, 2.0, 1.0, 2.0], [0.5, 2.0, 1.0, 2.0, 0.5, 1.0]):
        p = _problem([0, 1, 2, 2, 1, 0], weights=weights, nrand=2)
        eng = p.engine(devices=_devices())
        dev = p.run_engine(eng)
        eng.close()
        o, _ = p.oracle()
        assert_counters_equal(dev, o)          # dyadic weights: the f64 sums are exact
        assert dev["sum_weights"] == sum(weights)


def test_group_handle_native_dcd_feed_and_group_reduction(tmp_path):
    """cmx_run_dcd on a group handle (one reader/consumer team per device) and cmx_reduce_groups on the merged counters."""
    from cmx_b200.engine import DcdFile, Engine
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    path = str(tmp_path / "g.dcd")
    write_dcd(path, frames, d["cells"])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    opt = _opts(bulk_range=(8.0, 10.0), n_random_samples=3)
    order = [0, 1, 2, 2, 1, 0, 1, 2, 0]
    out = {}
    for name, devices in (("single", None), ("group", _devices())):
        eng = Engine(solute=sol, solvent=tm, options=opt, irefatom=1, autocorrelation=False, devices=devices)
        f = DcdFile(path)
        eng.run_dcd(f, sol.indices, tm.indices, order, n_reader_threads=2)
        red = eng.reduce_groups("solute_group_count", [np.arange(0, 1463, 2), np.arange(1, 1463, 2)])
        out[name] = (eng.finish(), red)
        f.close(); eng.close()
    for k in COUNTER_KEYS:
        assert np.array_equal(out["single"][0][k], out["group"][0][k]), k
    assert np.array_equal(out["single"][1], out["group"][1])
    assert np.array_equal(out["group"][1].sum(axis=0), out["group"][0]["md_count"])


def test_group_handle_public_driver(tmp_path):
    """mddf(file, ...; devices=[...]) == mddf(file, ...) (Result arrays and the final mddf / kb)."""
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    path = str(tmp_path / "p.dcd")
    write_dcd(path, frames, d["cells"])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    opt = _opts(bulk_range=(8.0, 10.0), n_random_samples=3, irefatom=1)
    R1 = cm.mddf(path, sol, tm, opt, frame_weights=[1.0, 3.0, 0.5])
    for feed in ("native", "host"):
        R2 = cm.mddf(path, sol, tm, opt, frame_weights=[1.0, 3.0, 0.5], devices=_devices(), feed=feed)
        for key in ("md_count", "md_count_random", "rdf_count", "solute_group_count", "solvent_group_count_random", "mddf", "kb"):
            assert np.array_equal(getattr(R1, key), getattr(R2, key)), (feed, key)
        assert R1.volume.total == R2.volume.total


def test_group_handle_errors():
    from cmx_b200.engine import CmxError
    p = _problem([0])
    with pytest.raises(CmxError):
        p.engine(devices=[0, 99])                      # no such device
    with pytest.raises(CmxError):
        p.engine(devices=[0, 0], keep_lists=True)      # the parity hooks read ONE context's scratch
    eng = p.engine(devices=[0, 0])
    with pytest.raises(CmxError):
        eng.minimum_distances(0)
    eng.close()


def test_stop_file_ends_the_native_feed(tmp_path, monkeypatch):
    """the reference's cooperative interrupt (src/mddf.jl:301-304) in the native feed: a file named
    stop_complexmixtures in the working directory ends the frame loop; the frames enqueued so far are finished."""
    from cmx_b200.engine import DcdFile, Engine
    d = namd()
    frames = np.concatenate([d["protein"], d["tmao"]], axis=1)
    path = str(tmp_path / "s.dcd")
    write_dcd(path, frames, d["cells"])
    sol = cm.AtomSelection(np.arange(1, 1464), nmols=1)
    tm = cm.AtomSelection(np.arange(1464, 1464 + 2534), natomspermol=14)
    monkeypatch.chdir(tmp_path)
    eng = Engine(solute=sol, solvent=tm, options=_opts(bulk_range=(8.0, 10.0), n_random_samples=1), irefatom=1, autocorrelation=False)
    f = DcdFile(path)
    (tmp_path / "stop_complexmixtures").write_text("")
    eng.run_dcd(f, sol.indices, tm.indices, [0, 1, 2] * 20, n_reader_threads=2)
    assert eng.stats()["frames"] == 0                  # polled before the first frame
    os.remove(tmp_path / "stop_complexmixtures")
    eng.run_dcd(f, sol.indices, tm.indices, [0, 1, 2], n_reader_threads=2)
    assert eng.stats()["frames"] == 3
    f.close(); eng.close()


@pytest.mark.skipif("_ngpus() < 2")
def test_driver_under_torchrun_two_gpus():
    """one process per GPU + NCCL all-reduce (tests/multi_gpu_mddf.py): sharded == single, incl. weights that differ
    between the ranks."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", os.path.join(ROOT, "tests", "multi_gpu_mddf.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "identical=True" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
