#!/usr/bin/env python
"""bench.py -- frames/s of the full mddf hot path (real + random phases + counters) on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (CPU arm: the fp64 port of the reference path)

A "step" is one pass of the hot path over one batch of `--frames-per-step` synthetic frames of
the workload (default: BASELINE.json configs[1] = C2, synthetic 100k-atom protein in water/urea,
mddf(protein, water) with per-atom contributions).  Prints ONE JSON line (rank 0).

  value     frames/s with the frames already resident in HBM (cmx_submit_frame_device), device-timed
            with CUDA events on the library's compute stream, max over ranks
  e2e       frames/s through the public C-ABI feed: frames in PINNED HOST memory (the staging ring),
            H2D copy of every frame + kernels + D2H read of the counters inside the timed region
  roofline  dominant kernel (random-phase search) vs the measured HBM peak, algorithmic bytes per
            SURVEY.md section 8(d)
  cpu_baseline  the oracle port of the reference's CPU path on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (builder, solute selection, solvent selection (None = self), bulk_range, description)
    "C2": ("config_c2", "solute", "water", (10.0, 15.0),
           "C2 synthetic 100k-atom protein(6000)+urea(800x8)+water(29200x3), cubic 100 A: mddf(protein, water), per-atom contributions"),
    "C2urea": ("config_c2", "solute", "urea", (10.0, 15.0),
               "C2 synthetic 100k-atom system: mddf(protein, urea), per-atom contributions"),
    "C3": ("config_c3", "glycerol", None, (20.0, 25.0),
           "C3 synthetic 200k-atom glycerol(5000x14)+water triclinic: glycerol self-MDDF"),
    "C4": ("config_c4", "solute", "water", (10.0, 15.0),
           "C4 synthetic 1M-atom protein(20000)+cosolvent(5000x14)+water(303333x3), cubic 216 A: mddf(protein, water)"),
    "C5": ("config_c5", "solute", "water", (10.0, 15.0),
           "C5 synthetic 5M-atom slab(1e6)+water(1333333x3): mddf(slab, water), per-atom contributions"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=list(CONFIGS))
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the system (testing only; reported in config)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--n-random-samples", type=int, default=10)
    ap.add_argument("--group-lanes", type=int, default=0)
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--no-hbm-kernel", action="store_true",
                    help="skip the extra measurement of the HBM-bound kernel of the path's tail (cmx_reduce_groups, 6 GB)")
    return ap.parse_args()


def ncu_traffic(config, phase):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t[config][phase]
    except Exception:
        return None


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region: one long-running `nvidia-smi -lms 50`
    whose lines are time-stamped on arrival; only samples inside [start, stop] are summarised."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.samples, self.t0, self.t1, self.proc = index, [], None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _run(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [t.strip() for t in line.strip().split(",")]))

    def __enter__(self):
        self.t0 = time.perf_counter(); return self

    def __exit__(self, *a):
        self.t1 = time.perf_counter()

    def close(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        inside = [s for t, s in self.samples if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) and len(s) >= 6]
        if not inside:   # region shorter than the sampling period: fall back to the nearest samples
            inside = [s for _, s in self.samples[-3:] if len(s) >= 6]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        sm = sorted(float(s[0]) for s in inside)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in inside)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(inside[0][1]), "reasons": reasons, "samples": len(inside)}


def build_workload(args, rank, world):
    import cmx_b200 as cm
    from cmx_b200 import synthetic as syn
    builder, sol_name, solv_name, bulk_range, desc = CONFIGS[args.config]
    system = getattr(syn, builder)(args.scale)
    solute = system.selections[sol_name]
    solvent = system.selections[solv_name] if solv_name else solute
    auto = solv_name is None
    opt = cm.Options(bulk_range=bulk_range, n_random_samples=args.n_random_samples, seed=321, silent=True)
    fps = args.frames_per_step
    if fps <= 0:
        in_bytes = 12 * (solvent.natoms if auto else solute.natoms + solvent.natoms)
        # > 126 MB of distinct input per step (larger than L2); small systems take half a trajectory (500 of
        # C2's 1000 frames) per step so that the per-step cmx_finish weighs as it does in a real run
        fps = max(int(np.ceil(150e6 / in_bytes)), min(512, int(600e6 / in_bytes)), 16)
    # weak scaling: every rank gets its own `fps` frames per step (frame ids interleaved as in the sharded driver)
    frame_ids = [1 + rank + world * k for k in range(fps)]
    return dict(cm=cm, system=system, solute=solute, solvent=solvent, auto=auto, opt=opt, fps=fps, frame_ids=frame_ids, desc=desc)


def gather_frames(w):
    s = w["system"]
    xs, xv = [], []
    for fid in w["frame_ids"]:
        x, _ = s.frame(fid)
        xv.append(x[w["solvent"].indices - 1])
        if not w["auto"]:
            xs.append(x[w["solute"].indices - 1])
    xv = np.ascontiguousarray(np.stack(xv), dtype=np.float32)
    xs = xv if w["auto"] else np.ascontiguousarray(np.stack(xs), dtype=np.float32)
    return xs, xv


def algorithmic_bytes(w):
    ns, nv = w["solute"].natoms, w["solvent"].natoms
    n_in = nv if w["auto"] else ns + nv
    nr = w["opt"].n_random_samples
    return 12 * n_in, 12 * nr * nv     # (input coordinates read once, source-molecule gather of every random molecule)


def irefatom_of(w, xv0):
    first = xv0[: w["solvent"].natomspermol].astype(np.float64)
    return int(np.argmin(np.linalg.norm(first - first.mean(axis=0), axis=1))) + 1


def cpu_arm(w, xs, xv, nframes, nthreads):
    from oracle import cmx_oracle as orc
    o = orc.Oracle.from_problem(w["solute"], w["solvent"], w["opt"], irefatom_of(w, xv[0]), w["auto"])
    t0 = time.perf_counter()
    o.run_frames(xs[:nframes], xv[:nframes], w["system"].cell, frame_ids=w["frame_ids"][:nframes], use_clist=True, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return nframes / dt, dt, o


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    if args.impl == "reference":
        # the reference's own CPU implementation of the path = the fp64 oracle port (Julia is not available)
        if rank != 0:
            return
        w = build_workload(args, 0, 1)
        nf = args.cpu_frames or max(ncores, 8)
        w["frame_ids"] = w["frame_ids"][:nf] if len(w["frame_ids"]) >= nf else [1 + k for k in range(nf)]
        w["fps"] = nf
        xs, xv = gather_frames(w)
        for _ in range(min(args.warmup, 1)):
            cpu_arm(w, xs, xv, min(nf, ncores), ncores)
        times = []
        for _ in range(args.steps):
            fps_, dt, _ = cpu_arm(w, xs, xv, nf, ncores)
            times.append(dt)
        val = nf * len(times) / sum(times)
        sample = f"{nf} frames/step of the same workload, {ncores} OpenMP threads, frame-parallel as src/mddf.jl:285-338"
        print(json.dumps({"metric": "frames/sec of full mddf (real + random phases + counters)", "value": val, "unit": "frames/s",
                          "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": w["desc"], "frames_per_step": nf, "scale": args.scale, "n_random_samples": args.n_random_samples},
                          "cpu_baseline": {"value": val, "unit": "frames/s", "cores": ncores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"), timeout=datetime.timedelta(seconds=180))
    from cmx_b200.engine import Engine
    import ctypes as C

    w = build_workload(args, rank, world)
    cm, opt, fps = w["cm"], w["opt"], w["fps"]
    xs, xv = gather_frames(w)
    iref = irefatom_of(w, xv[0])
    cell = w["system"].cell
    eng = Engine(solute=w["solute"], solvent=w["solvent"], options=opt, irefatom=iref, autocorrelation=w["auto"],
                 device=local_rank, ring_slots=fps, group_lanes=args.group_lanes, n_streams=args.streams)
    lib, h = eng.lib, eng.h
    cellc = cm.engine.cell_to_c(cell)
    cellp = cellc.ctypes.data_as(C.POINTER(C.c_double))
    # ---- device-resident copies (value) and the pinned staging ring pre-filled (e2e) ----
    d_xv = torch.from_numpy(xv).cuda()
    d_xs = d_xv if w["auto"] else torch.from_numpy(xs).cuda()
    sv_stride, ss_stride = d_xv[0].numel() * 4, d_xs[0].numel() * 4
    pv, ps = d_xv.data_ptr(), d_xs.data_ptr()
    for k in range(fps):   # fill every pinned slot once (untimed): slot k holds frame k
        a_s, a_v = eng.acquire()
        a_v[...] = xv[k]
        if not w["auto"]:
            a_s[...] = xs[k]
        eng.submit(w["frame_ids"][k], 1.0, cell)
    eng.sync(); eng.reset()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allreduce_counts():
        if world > 1:
            ptr, n = eng.counters_device()

            class W_:
                pass
            wobj = W_()
            wobj.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}
            t = torch.as_tensor(wobj, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    submit_s = [0.0]

    def step_device(collective=True):
        t_sub = time.perf_counter()
        for k in range(fps):
            rc = lib.cmx_submit_frame_device(h, C.c_void_p(ps + k * ss_stride), C.c_void_p(pv + k * sv_stride), w["frame_ids"][k], 1.0, cellp)
            if rc:
                raise RuntimeError(lib.cmx_last_error(h).decode())
        submit_s[0] += time.perf_counter() - t_sub
        eng.sync()
        if collective:
            allreduce_counts()

    null_s, null_v = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()

    def step_e2e():
        for k in range(fps):
            rc = lib.cmx_acquire_frame_buffer(h, C.byref(null_s), C.byref(null_v))   # slot k already holds frame k (pinned)
            rc = rc or lib.cmx_submit_frame(h, w["frame_ids"][k], 1.0, cellp)
            if rc:
                raise RuntimeError(lib.cmx_last_error(h).decode())
        eng.sync()
        allreduce_counts()
        return eng.finish(copy=False)          # D2H read of the step's result (into the engine's pinned result arrays)

    # ---- value: frames resident in HBM, device-timed ----
    for _ in range(max(args.warmup, 3)):
        step_device()
    eng.reset()
    barrier()
    st0 = eng.stats()
    submit_s[0] = 0.0
    sampler = ClockSampler(local_rank)
    time.sleep(0.15)
    with sampler as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_device()
        barrier()
        wall = time.perf_counter() - t0
    sampler.close()
    st1 = eng.stats()
    host_submit_ms = 1e3 * submit_s[0] / args.steps
    dev_ms = st1["gpu_ms_total"] - st0["gpu_ms_total"]
    t_dev = max(dev_ms * 1e-3, 1e-9)
    tt = torch.tensor([t_dev, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, wall = float(tt[0]), float(tt[1])
    # device events cover first-kernel..last-kernel of each step; the wall bracket (barrier+sync both sides)
    # is what the job takes: report the slower of the two views
    t_used = max(t_dev, wall)
    total_frames = fps * args.steps * world
    value = total_frames / t_used
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    deferred = (st1["deferred"] - st0["deferred"]) / max(1, fps * args.steps)
    counters_check = eng.finish()
    hits = float(counters_check["md_count"].sum())

    # ---- e2e ----
    e2e = None
    if not args.no_e2e:
        eng.reset()
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        eng.reset()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step_e2e()
        barrier()
        te = time.perf_counter() - t0
        tt = torch.tensor([te], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt[0])
        in_bytes = (xv[0].nbytes if w["auto"] else xs[0].nbytes + xv[0].nbytes)
        out_bytes = sum(v.nbytes for v in res.values() if isinstance(v, np.ndarray))
        e2e = {"value": total_frames / te, "unit": "frames/s", "h2d_bytes_per_step": int(in_bytes * fps), "d2h_bytes_per_step": int(out_bytes)}

    # ---- roofline of the dominant kernel (separate profiled pass: CUDA events around the search kernels) ----
    roof = None
    if rank == 0:
        eng.reset(); eng.set_option("active_streams", 1); eng.set_option("profile", 1)   # one frame at a time: clean per-kernel times
        s0 = eng.stats()
        step_device(collective=False)
        s1 = eng.stats()
        eng.set_option("profile", 0)
        ms_rand = (s1["gpu_ms_search_random"] - s0["gpu_ms_search_random"]) / fps
        ms_real = (s1["gpu_ms_search_real"] - s0["gpu_ms_search_real"]) / fps
        ms_frame = (s1["gpu_ms_total"] - s0["gpu_ms_total"]) / fps
        b_in, b_rand = algorithmic_bytes(w)
        peak, peak_src = peaks()
        pair_path = w["solute"].nmols > 1 and w["solute"].natomspermol <= 64     # the library's auto rule (cmx_config.path = 0)
        kname = ("k_pair_random", "k_pairs") if pair_path else ("k_tile_search", "k_tile_search")
        dom_ms, dom_bytes, dom = (ms_rand, b_rand, "random-phase search") if ms_rand >= ms_real else (ms_real, b_in, "real-phase search")
        kernel_name = kname[0] if ms_rand >= ms_real else kname[1]
        ach = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        eng.reset(); eng.set_option("count_pairs", 1)
        step_device(collective=False)
        pe = eng.stats()["pair_evals"] / fps
        eng.set_option("count_pairs", 0); eng.set_option("active_streams", 0)
        roof = {"bound": "hbm", "kernel": f"{kernel_name} ({dom})", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ncu_traffic(args.config, "random" if dom.startswith("random") else "real"), "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "kernel_ms_per_launch": dom_ms, "kernel_share_of_frame": dom_ms / ms_frame if ms_frame > 0 else None,
                "frame_algorithmic_bytes": b_in + b_rand, "frame_achieved_GBps": (b_in + b_rand) * value / world / 1e9,
                "pair_evals_per_frame": pe, "pair_evals_per_s": pe * value,
                # second view (the kernel is instruction-bound, not HBM-bound): pair evaluations per second inside the
                # dominant kernel against lanes x clock / 12.5 instructions (the SASS inner loop of k_tile_search)
                "alu_view": {"kernel_pair_evals_per_s": (pe * (ms_rand / max(ms_rand + ms_real, 1e-12)) / (dom_ms * 1e-3)) if dom_ms > 0 else None,
                             "inner_loop_peak_pair_evals_per_s": 148 * 128 * 1.965e9 / 12.5}}

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nfc = args.cpu_frames or min(fps, max(ncores, 8))
        cpu_arm(w, xs, xv, min(nfc, 2), ncores)   # warm the library
        val_cpu, dt, o = cpu_arm(w, xs, xv, nfc, ncores)
        # parity guard on the same frames: the device counters of these frames must equal the oracle's
        eng.reset()
        for k in range(nfc):
            eng.submit_device(ps + k * ss_stride, pv + k * sv_stride, cell, frame_index=w["frame_ids"][k])
        devc = eng.finish()
        ok = all(np.array_equal(devc[k], getattr(o, k)) for k in ("md_count", "md_count_random", "rdf_count", "rdf_count_random"))
        cpu = {"value": val_cpu, "unit": "frames/s", "cores": ncores, "kind": "port",
               "sample": f"{nfc} frames of the same workload, oracle/cmx_oracle.c cell-list path, {ncores} OpenMP threads, frame-parallel",
               "counts_equal_device": bool(ok)}

    # ---- the one HBM-bound kernel next to the path (rank 0, N=1): per-residue sums of a per-atom contribution array of
    # C5's shape (1e6 rows x 750 bins x 8 B = 6 GB), cmx_reduce_groups; reported beside the dominant kernel's roofline
    hbm_kernel = None
    nbins = eng.nbins
    if rank == 0 and world == 1 and not args.no_hbm_kernel and args.scale == 1.0:
        eng.close()
        eng = None
        try:
            import bench_extras
            r = bench_extras.reduce_measure(1000000, repeat=2, check=False, device=local_rank)
            k = r["residues_of_16_rows"]
            hbm_kernel = {"kernel": r["kernel"], "bound": "hbm", "achieved": k["achieved_GBps"], "peak": r["peak_GBps"], "unit": "GB/s",
                          "frac": k["frac_of_hbm_peak"], "kernel_ms_per_launch": k["kernel_ms"], "algorithmic_bytes_per_launch": r["algorithmic_bytes"],
                          "workload": "62 500 groups of 16 rows of a 1e6 x 750 uint64 per-atom contribution array (C5 shape)",
                          "d2h_bytes": k["d2h_bytes"], "traffic": (ncu_traffic("reduce_rows_500k", "traffic") or 0) * 2 or None,
                          "traffic_note": "ncu --set full capture at 500 000 rows (profiles/r01b_reduce_rows_ncu_full.txt), scaled x2"}
        except Exception as e:   # an extra: never fails the bench line
            hbm_kernel = {"error": str(e)[:200]}

    if rank == 0:
        line = {"metric": "frames/sec of full mddf (real + random phases + counters)", "value": value, "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_used / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 search + f64 finalisation",
                "data": "synthetic",
                "config": {"workload": w["desc"], "frames_per_step": fps, "frames_per_step_total": fps * world, "scale": args.scale,
                           "n_random_samples": opt.n_random_samples, "bulk_range": [opt.dbulk, opt.cutoff], "nbins": nbins,
                           "l2": f"distinct inputs per step = {fps * (xv[0].nbytes + (0 if w['auto'] else xs[0].nbytes)) / 1e6:.0f} MB (larger than the 126 MB L2; no flush needed)",
                           "frames_in_flight": args.streams or 8, "allreduce_per_step": world > 1, "hits_per_frame": hits / max(1, fps * args.steps * world),
                           "deferred_to_exact_per_frame": deferred},
                "device_ms_per_step": 1e3 * t_dev / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps,
                "host_submit_ms_per_step": host_submit_ms,
                "gpu_launches": int(launches), "clocks": clk.summary(), "e2e": e2e, "roofline": roof, "roofline_hbm_kernel": hbm_kernel,
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
