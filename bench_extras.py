#!/usr/bin/env python
"""bench_extras.py -- measurements of the SURVEY 8(f) rows next to the hot path (not the headline bench.py):

  feed    frames/s of a whole mddf run FROM A DCD FILE on disk: the library's native feed (cmx_run_dcd: reader
          threads -> pinned ring -> raw-frame H2D -> device gather) with 1/2/4 reader threads, against the host
          reader of this package writing into the staging slot (the structure of the reference's frame loop,
          src/mddf.jl:296-334 + NamdDCD.jl:141-169); counters of both feeds must be identical.
  reduce  the device-side group reduction (cmx_reduce_groups) over a per-atom contribution array of C5's shape
          (1e6 rows x 750 bins x 8 B = 6 GB): kernel time (CUDA events) -> achieved GB/s against the measured
          HBM peak (this kernel IS HBM-bound: algorithmic bytes = rows x nbins x 8).

Prints one JSON line per measurement.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def write_xtc_replicated(path, system, nf, distinct=4):
    """an XTC of `nf` frames for feed measurements: `distinct` frames are compressed by the test-suite's (pure Python,
    ~0.7 s per 100k-atom frame) writer and their bytes repeated with increasing step numbers."""
    import struct
    from common import xtc_compress
    box = (np.asarray(system.cell, dtype=np.float64).T / 10.0).astype(np.float32)
    blocks = []
    for k in range(distinct):
        x = system.frame(k + 1)[0].astype(np.float64) / 10.0 * 1000.0
        ints = np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)
        blocks.append(xtc_compress(ints, 1000.0))
    n = system.natoms
    with open(path, "wb") as f:
        for k in range(nf):
            f.write(struct.pack(">iiif", 1995, n, 10 * k, 2.0 * k) + struct.pack(">9f", *box.reshape(9)) + struct.pack(">i", n))
            f.write(blocks[k % distinct])


def feed(args):
    import cmx_b200 as cm
    from cmx_b200 import synthetic as syn
    from cmx_b200.engine import DcdFile, Engine, XtcFile
    from common import write_dcd
    s = syn.config_c2(args.scale)
    sol, wat = s.selections["solute"], s.selections["water"]
    opt = cm.Options(bulk_range=(10.0, 15.0), n_random_samples=10, seed=321, silent=True)
    nf = args.frames
    tmp = tempfile.mkdtemp(prefix="cmx_feed_", dir=args.tmpdir)
    xtc = args.format == "xtc"
    path = os.path.join(tmp, "c2.xtc" if xtc else "c2.dcd")
    if xtc:
        write_xtc_replicated(path, s, nf)
    else:
        frames = np.stack([s.frame(k + 1)[0] for k in range(nf)]).astype(np.float32)
        write_dcd(path, frames, np.asarray(s.cell, dtype=np.float64))
        del frames
    size = os.path.getsize(path)
    first = s.frame(1)[0][wat.indices - 1][:3].astype(np.float64)
    iref = int(np.argmin(np.linalg.norm(first - first.mean(axis=0), axis=1))) + 1
    out = {"what": "feed", "format": args.format,
           "workload": f"C2 (scale {args.scale}) mddf(protein, water) from a {args.format.upper()} file of {nf} frames, {size / 1e6:.0f} MB (page cache warm)",
           "frames": nf, "file_bytes": size}
    ref = None
    Reader = XtcFile if xtc else DcdFile
    for threads in ((2, 4, 8, 16) if xtc else (1, 2, 4)):
        eng = Engine(solute=sol, solvent=wat, options=opt, irefatom=iref, autocorrelation=False)
        f = Reader(path)
        run = eng.run_xtc if xtc else eng.run_dcd
        run(f, sol.indices, wat.indices, list(range(min(nf, 16))), n_reader_threads=threads); eng.sync(); eng.reset()   # warm-up
        t0 = time.perf_counter()
        run(f, sol.indices, wat.indices, list(range(nf)), n_reader_threads=threads)
        c = eng.finish(copy=False)
        dt = time.perf_counter() - t0
        out[f"native_{threads}_threads_frames_per_s"] = nf / dt
        out[f"native_{threads}_threads_file_GBps"] = size / dt / 1e9
        if ref is None:
            ref = {k: np.array(v) for k, v in c.items() if isinstance(v, np.ndarray)}
        else:
            assert all(np.array_equal(ref[k], c[k]) for k in ref), "native feed: counters depend on the reader thread count"
        f.close(); eng.close()
    # where the fixed cost of a short run goes: create / first run (ring allocation) / finish / destroy
    t0 = time.perf_counter(); eng = Engine(solute=sol, solvent=wat, options=opt, irefatom=iref, autocorrelation=False); t1 = time.perf_counter()
    f = Reader(path); (eng.run_xtc if xtc else eng.run_dcd)(f, sol.indices, wat.indices, list(range(min(nf, 64))), n_reader_threads=2); t2 = time.perf_counter()
    eng.finish(copy=False); t3 = time.perf_counter(); f.close(); eng.close(); t4 = time.perf_counter()
    out["fixed_costs_ms"] = {"create": 1e3 * (t1 - t0), "run_64_frames_incl_ring_alloc": 1e3 * (t2 - t1), "finish": 1e3 * (t3 - t2), "destroy": 1e3 * (t4 - t3)}
    # host reader of this package -> staging slot (first `host_frames` frames only: it is the slow side)
    nh = min(nf, args.host_frames)
    o2 = cm.Options(bulk_range=(10.0, 15.0), n_random_samples=10, seed=321, silent=True, lastframe=nh, irefatom=iref)
    t0 = time.perf_counter()
    Rh = cm.mddf(path, sol, wat, o2, feed="host")
    dth = time.perf_counter() - t0
    t0 = time.perf_counter()
    Rn = cm.mddf(path, sol, wat, o2, feed="native", reader_threads=2)
    dtn = time.perf_counter() - t0
    out["public_mddf_host_feed_frames_per_s"] = nh / dth
    out["public_mddf_native_feed_frames_per_s"] = nh / dtn
    out["public_mddf_frames"] = nh
    out["feeds_identical"] = bool(np.array_equal(Rh.md_count, Rn.md_count) and np.array_equal(Rh.solute_group_count, Rn.solute_group_count)
                                  and np.array_equal(Rh.md_count_random, Rn.md_count_random))
    os.remove(path); os.rmdir(tmp)
    print(json.dumps(out))


def reduce_measure(nrows=1000000, repeat=3, check=True, device=0):
    """cmx_reduce_groups over a per-atom solute_group_count array of `nrows` rows x 750 bins (C5's shape for 1e6 rows):
    kernel time from CUDA events (option "profile") -> achieved GB/s against the measured HBM peak."""
    import cmx_b200 as cm
    from cmx_b200.engine import Engine
    nsolv = 100000
    sol = cm.AtomSelection(np.arange(1, nrows + 1), nmols=1)
    wat = cm.AtomSelection(np.arange(nrows + 1, nrows + 1 + 3 * nsolv), natomspermol=3)
    opt = cm.Options(bulk_range=(10.0, 15.0), n_random_samples=2, seed=321, silent=True)
    rng = np.random.default_rng(1)
    cell = np.diag([400.0, 400.0, 312.5])
    xs = (rng.uniform(0, 1, size=(nrows, 3)) * np.array([400.0, 400.0, 50.0]) + np.array([0, 0, 130.0])).astype(np.float32)
    ctr = rng.uniform(0, 1, size=(nsolv, 1, 3)) * np.array([400.0, 400.0, 312.5])
    xv = (ctr + rng.normal(0, 0.5, size=(nsolv, 3, 3))).reshape(-1, 3).astype(np.float32)
    eng = Engine(solute=sol, solvent=wat, options=opt, irefatom=1, autocorrelation=False, n_streams=2, device=device)
    try:
        eng.set_option("profile", 1)
        for k in range(2):
            eng.submit_arrays(xs, xv, cell, frame_index=k + 1)
        eng.sync()
        nb = eng.nbins
        per = 16
        groups = [np.arange(g * per, min((g + 1) * per, nrows)) for g in range((nrows + per - 1) // per)]
        peak, src = hbm_peak()
        res = {"what": "reduce", "kernel": "k_reduce_rows (cmx_reduce_groups)", "rows": nrows, "nbins": nb,
               "algorithmic_bytes": nrows * nb * 8, "peak_GBps": peak, "peak_source": src}
        for name, gs in (("residues_of_16_rows", groups), ("one_group_of_all_rows", [np.arange(nrows)])):
            best = 1e30
            for _ in range(repeat):
                got = eng.reduce_groups("solute_group_count", gs)
                best = min(best, eng.stats()["gpu_ms_reduce"])
            res[name] = {"n_groups": len(gs), "kernel_ms": best, "achieved_GBps": nrows * nb * 8 / (best * 1e-3) / 1e9,
                         "frac_of_hbm_peak": nrows * nb * 8 / (best * 1e-3) / 1e9 / peak, "d2h_bytes": int(got.nbytes)}
        if check:
            full = eng.finish(copy=False)
            want = full["solute_group_count"].reshape(len(groups), per, nb).sum(axis=1) if nrows % per == 0 else None
            got = eng.reduce_groups("solute_group_count", groups)
            res["equals_host_sum_of_rows"] = bool(want is None or np.array_equal(got, want))
            res["hits"] = float(full["md_count"].sum())
    finally:
        eng.close()
    return res


def reduce(args):
    print(json.dumps(reduce_measure(args.rows, args.repeat)))


def group(args):
    """ONE process, ONE handle, several GPUs (cmx_config.n_devices / device_ids): the whole C4 (or --config) trajectory
    through create -> acquire/submit from pinned host frames -> finish (peer-access merge onto the first device) ->
    destroy, wall clock -- the in-library counterpart of bench.py's `job` leg under torchrun."""
    import ctypes as C
    import time
    import bench
    from cmx_b200.engine import Engine, cell_to_c
    devices = [int(x) for x in args.devices.split(",")]
    ns = argparse.Namespace(scale=args.scale, n_random_samples=10, frames_per_step=0)
    w = bench.build_workload(ns, args.config, 0, 1)
    fps = min(w["fps"], 32)
    xs, xv = bench.frames_of(w, [1 + k for k in range(fps)])
    iref = bench.irefatom_of(w, xv[0])
    total = args.frames or w["traj_frames"]
    out = {"config": args.config, "frames": total, "runs": []}
    warm = Engine(solute=w["solute"], solvent=w["solvent"], options=w["opt"], irefatom=iref, autocorrelation=w["auto"], device=devices[0])
    warm.submit_arrays(xs[0], xv[0], w["system"].cell, frame_index=1); warm.finish(); warm.close()     # CUDA context, module load
    for devs in ([devices[0]], devices):
        t0 = time.perf_counter()
        eng = Engine(solute=w["solute"], solvent=w["solvent"], options=w["opt"], irefatom=iref, autocorrelation=w["auto"],
                     device=devs[0], devices=devs if len(devs) > 1 else None)
        t_create = time.perf_counter() - t0
        lib, h = eng.lib, eng.h
        cellp = cell_to_c(w["system"].cell).ctypes.data_as(C.POINTER(C.c_double))
        ps, pv = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        nfill = 32 * len(devs)     # every ring slot of every device is filled once -- with the SAME coordinates, so that the
        for k in range(total):     # counters of the two runs can be compared (the frame ids, i.e. the random phases, differ)
            if k < nfill:
                a_s, a_v = eng.acquire()
                a_v[...] = xv[0]
                if not w["auto"]:
                    a_s[...] = xs[0]
            elif lib.cmx_acquire_frame_buffer(h, C.byref(ps), C.byref(pv)):
                raise RuntimeError(lib.cmx_last_error(h).decode())
            if lib.cmx_submit_frame(h, 1 + k, 1.0, cellp):
                raise RuntimeError(lib.cmx_last_error(h).decode())
        eng.sync()
        t_feed = time.perf_counter() - t0
        res = eng.finish(copy=False)
        t_fin = time.perf_counter() - t0
        hits = float(res["md_count"].sum()), float(res["md_count_random"].sum())
        st = eng.stats()
        eng.close()
        wall = time.perf_counter() - t0
        out["runs"].append({"devices": devs, "wall_s": wall, "frames_per_s": total / wall, "create_s": t_create, "feed_s": t_feed - t_create,
                            "merge_finish_s": t_fin - t_feed, "hits": hits, "host_submit_ms": st["host_submit_ms"], "host_wait_ms": st["host_wait_ms"]})
    r = out["runs"]
    out["identical_counters"] = r[0]["hits"] == r[-1]["hits"]
    out["speedup"] = r[0]["wall_s"] / r[-1]["wall_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["feed", "reduce", "group"])
    ap.add_argument("--devices", default="0,0", help="group: CUDA ordinals behind the one handle")
    ap.add_argument("--config", default="C4")
    ap.add_argument("--frames", type=int, default=0, help="feed: frames of the generated file (default 256); group: frames of the run (default: the configuration's trajectory)")
    ap.add_argument("--host-frames", type=int, default=64)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--tmpdir", default=None)
    ap.add_argument("--format", default="dcd", choices=["dcd", "xtc"], help="trajectory format of the feed measurement")
    a = ap.parse_args()
    if a.what == "feed":
        a.frames = a.frames or 256
        feed(a)
    elif a.what == "group":
        group(a)
    else:
        reduce(a)
