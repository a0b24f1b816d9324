#!/usr/bin/env python
"""bench.py -- frames/s of the full mddf hot path (real + random phases + counters) on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (CPU arm: Julia's ComplexMixtures if present, else the fp64 port)

The workload is BASELINE.json configs[3] = C4, the configuration the >= 100x target is quoted on (synthetic 1M-atom
protein in mixed solvent, mddf(protein, water), n_random_samples = 10); `--config` selects the others.  A "step" is
one pass of the hot path over one batch of `--frames-per-step` distinct synthetic frames.  ONE JSON line (rank 0):

  value     frames/s with the frames already resident in HBM (cmx_submit_frame_device), device-timed with CUDA events
            on the library's compute streams and cross-checked with the barrier-bracketed wall clock, max over ranks
  e2e       frames/s through the public C-ABI feed from PINNED HOST memory: a step is this GPU's share of the
            configuration's trajectory on a full box (C4: 5000 frames / 8 GPUs = 625 frames): H2D copy of every frame +
            kernels + ONE all-reduce + ONE cmx_finish (D2H of all counters) per step, as a real run does
  job       strong-scaling leg: the whole trajectory (C4: 5000 frames shared by the N ranks) through
            create -> feed -> all-reduce -> finish -> destroy, wall clock, fixed costs included
  roofline  dominant kernel (random-phase search) vs the measured HBM peak, algorithmic bytes per SURVEY.md 8(d)
  cpu_baseline  the CPU arm on the host cores, bounded sample, with the 8-array parity guard against the device
  secondary the C2 line (value + e2e), so that rounds stay comparable
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (builder, solute selection, solvent selection (None = self), bulk_range, trajectory frames, description)
    "C2": ("config_c2", "solute", "water", (10.0, 15.0), 1000,
           "C2 synthetic 100k-atom protein(6000)+urea(800x8)+water(29200x3), cubic 100 A: mddf(protein, water), per-atom contributions"),
    "C2urea": ("config_c2", "solute", "urea", (10.0, 15.0), 1000,
               "C2 synthetic 100k-atom system: mddf(protein, urea), per-atom contributions"),
    "C3": ("config_c3", "glycerol", None, (20.0, 25.0), 200,
           "C3 synthetic 200k-atom glycerol(5000x14)+water triclinic: glycerol self-MDDF"),
    "C4": ("config_c4", "solute", "water", (10.0, 15.0), 5000,
           "C4 synthetic 1M-atom protein(20000)+cosolvent(5000x14)+water(303333x3), cubic 216 A: mddf(protein, water)"),
    "C5": ("config_c5", "solute", "water", (10.0, 15.0), 2000,
           "C5 synthetic 5M-atom slab(1e6)+water(1333333x3): mddf(slab, water), per-atom contributions"),
}
COUNTER_KEYS = ("md_count", "md_count_random", "rdf_count", "rdf_count_random", "solute_group_count",
                "solute_group_count_random", "solvent_group_count", "solvent_group_count_random")
METRIC = "frames/sec of full mddf (real + random phases + counters)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=list(CONFIGS))
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the system (testing only; reported in config)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-job", action="store_true", help="skip the strong-scaling whole-trajectory leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 line")
    ap.add_argument("--n-random-samples", type=int, default=10)
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0, help="frames per kernel launch (0 = the library's choice)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-kernel measurement (tuning runs only)")
    ap.add_argument("--no-hbm-kernel", action="store_true",
                    help="skip the extra measurement of the HBM-bound kernel of the path's tail (cmx_reduce_groups, 6 GB)")
    return ap.parse_args()


def ncu_traffic(config, phase):
    """DRAM bytes (read + write) per FRAME of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json; the caller scales by its frames per launch)."""
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
    """SM clock + throttle reasons DURING the timed region.  NVML in a sampling thread (10 ms period, no process start-up:
    a region of a few tens of ms still gets samples); falls back to one long-running `nvidia-smi -lms 50` whose lines are
    time-stamped on arrival.  Only samples inside [start, stop] are summarised (the nearest ones if the region was shorter
    than a period)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]        # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index=0):
        self.index, self.samples, self.t0, self.t1, self.proc, self.nvml, self.stop = index, [], None, None, None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                u = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + u) if not u.startswith("GPU-") else u)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml, self.h = pynvml, h
            self._sample_nvml()                      # fails here, not in the thread, if the queries are not supported
            self.t = threading.Thread(target=self._run_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _sample_nvml(self):
        n = self.nvml
        clk = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        try:
            r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.samples.append((time.perf_counter(), [str(clk), str(self.smax)] + ["Active" if r & b else "Not Active" for b in self.BITS]))

    def _run_nvml(self):
        while not self.stop:
            try:
                self._sample_nvml()
            except Exception:
                pass
            time.sleep(0.01)

    def _run(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [t.strip() for t in line.strip().split(",")]))

    def __enter__(self):
        self.t0 = time.perf_counter(); return self

    def __exit__(self, *a):
        self.t1 = time.perf_counter()

    def close(self):
        self.stop = True
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        good = [(t, s) for t, s in self.samples if len(s) >= 6]
        inside = [s for t, s in good if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30)]
        if not inside and good and self.t0 is not None:   # region shorter than the sampling period: the samples nearest to it
            mid = 0.5 * (self.t0 + (self.t1 or self.t0))
            inside = [s for _, s in sorted(good, key=lambda ts: abs(ts[0] - mid))[:3]]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        sm = sorted(float(s[0]) for s in inside)
        reasons = [n for k, n in enumerate(self.NAMES) if any(s[2 + k].lower().startswith("active") for s in inside)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(inside[0][1]), "reasons": reasons, "samples": len(inside),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def build_workload(args, config, rank, world, fps_override=0):
    import cmx_b200 as cm
    from cmx_b200 import synthetic as syn
    builder, sol_name, solv_name, bulk_range, traj_frames, desc = CONFIGS[config]
    system = getattr(syn, builder)(args.scale)
    solute = system.selections[sol_name]
    solvent = system.selections[solv_name] if solv_name else solute
    auto = solv_name is None
    opt = cm.Options(bulk_range=bulk_range, n_random_samples=args.n_random_samples, seed=321, silent=True)
    fps = fps_override or args.frames_per_step
    if fps <= 0:
        in_bytes = 12 * (solvent.natoms if auto else solute.natoms + solvent.natoms)
        # > 126 MB of distinct input per step (larger than L2); small systems take half a trajectory (500 of
        # C2's 1000 frames) per step so that a step is not dominated by its fixed costs
        fps = max(int(np.ceil(150e6 / in_bytes)), min(512, int(600e6 / in_bytes)), 16)
    # weak scaling: every rank gets its own `fps` frames per step (frame ids interleaved as in the sharded driver)
    frame_ids = [1 + rank + world * k for k in range(fps)]
    return dict(cm=cm, system=system, solute=solute, solvent=solvent, auto=auto, opt=opt, fps=fps, frame_ids=frame_ids, desc=desc,
                traj_frames=traj_frames, config=config)


def frames_of(w, frame_ids):
    s = w["system"]
    xs, xv = [], []
    for fid in frame_ids:
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


def cpu_arm(w, xs, xv, frame_ids, nthreads):
    from oracle import cmx_oracle as orc
    o = orc.Oracle.from_problem(w["solute"], w["solvent"], w["opt"], irefatom_of(w, xv[0]), w["auto"])
    t0 = time.perf_counter()
    o.run_frames(xs, xv, w["system"].cell, frame_ids=frame_ids, use_clist=True, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return len(frame_ids) / dt, dt, o


# ---- the real reference, when the GPU box has it ----------------------------------------
JULIA_SCRIPT = r"""
using ComplexMixtures, PDBTools
# (the synthetic system is written by bench.py: DCD + a minimal PDB with the selections as segment names)
atoms = readPDB(ARGS[1])
solute = AtomSelection(select(atoms, "segname SOLU"), nmols = 1)
solvent = AtomSelection(select(atoms, "segname SOLV"), natomspermol = parse(Int, ARGS[3]))
opt = Options(bulk_range = (parse(Float64, ARGS[4]), parse(Float64, ARGS[5])), n_random_samples = parse(Int, ARGS[6]), seed = 321, silent = true)
traj = Trajectory(ARGS[2], solute, solvent)
mddf(traj, opt)                       # compile
t = @elapsed mddf(Trajectory(ARGS[2], solute, solvent), opt)
println("CMX_JULIA_SECONDS ", t)
"""


def julia_probe():
    """(usable, why): is `julia` with ComplexMixtures.jl on this box?  Never installs anything."""
    exe = shutil.which("julia")
    if not exe:
        return False, "julia not on PATH"
    try:
        r = subprocess.run([exe, f"-t{os.cpu_count()}", "-e", "using ComplexMixtures; println(\"ok\")"], capture_output=True, text=True, timeout=300)
        if r.returncode == 0 and "ok" in r.stdout:
            return True, exe
        return False, "julia present but `using ComplexMixtures` failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:160]
    except Exception as e:
        return False, f"julia probe failed: {e}"


def reference_arm(args, ncores):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  Julia's
    ComplexMixtures.mddf when the box has it (kind = "julia"); otherwise -- the case in this image, which has no Julia
    and no network -- the fp64 port of the same path (oracle/cmx_oracle.c, kind = "port"), frame-parallel over all
    cores like src/mddf.jl:285-338."""
    w = build_workload(args, args.config, 0, 1)
    nf = args.cpu_frames or max(ncores, 8)
    ids = [1 + k for k in range(nf)]
    xs, xv = frames_of(w, ids)
    have_julia, why = julia_probe()
    kind, note = "port", why
    times = []
    if have_julia and not w["auto"] and w["solute"].nmols == 1:
        try:
            import tempfile
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from common import write_dcd
            d = tempfile.mkdtemp()
            dcd, pdb = os.path.join(d, "t.dcd"), os.path.join(d, "t.pdb")
            write_dcd(dcd, np.concatenate([xs, xv], axis=1), w["system"].cell)
            with open(pdb, "w") as f:
                n = 0
                for seg, arr, apm in (("SOLU", xs[0], len(xs[0])), ("SOLV", xv[0], w["solvent"].natomspermol)):
                    for k, p in enumerate(arr):
                        n += 1
                        f.write("ATOM  %5d  C%-2d RES X%4d    %8.3f%8.3f%8.3f  1.00  0.00      %-4s\n" % (n % 100000, k % apm % 100, (k // apm) % 10000, p[0] % 1000, p[1] % 1000, p[2] % 1000, seg))
                f.write("END\n")
            js = os.path.join(d, "run.jl"); open(js, "w").write(JULIA_SCRIPT)
            for _ in range(max(1, args.steps)):
                r = subprocess.run([why, f"-t{ncores}", js, pdb, dcd, str(w["solvent"].natomspermol), str(w["opt"].dbulk), str(w["opt"].cutoff),
                                    str(args.n_random_samples)], capture_output=True, text=True, timeout=1500)
                t = [float(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("CMX_JULIA_SECONDS")]
                if not t:
                    raise RuntimeError((r.stderr.strip().splitlines() or ["no timing"])[-1][:160])
                times.append(t[0])
            kind, note = "julia", "ComplexMixtures.jl mddf() on the generated DCD"
        except Exception as e:
            times, note = [], f"julia run failed ({e}); fell back to the port"
    if not times:
        for _ in range(min(args.warmup, 1)):
            cpu_arm(w, xs[:min(nf, ncores)], xv[:min(nf, ncores)], ids[:min(nf, ncores)], ncores)
        for _ in range(args.steps):
            times.append(cpu_arm(w, xs, xv, ids, ncores)[1])
    val = nf * len(times) / sum(times)
    sample = f"{nf} frames/step of the same workload, {ncores} threads, frame-parallel as src/mddf.jl:285-338; {note}"
    return {"metric": METRIC, "value": val, "unit": "frames/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the b200 arm's configuration (same workload, same frames per step of the GPU arm); the CPU steps are a bounded
            # sample of it: `cpu_sample_frames_per_step` frames each, stated again in cpu_baseline.sample
            "config": {"workload": w["desc"], "frames_per_step": w["fps"], "frames_per_step_total": w["fps"] * max(1, int(os.environ.get("WORLD_SIZE", "1"))),
                       "scale": args.scale, "n_random_samples": args.n_random_samples, "bulk_range": [w["opt"].dbulk, w["opt"].cutoff],
                       "cpu_sample_frames_per_step": nf},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": ncores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


class Bench:
    """One configuration on this rank's GPU: engine, device-resident frames, pinned staging ring."""

    def __init__(self, args, config, rank, world, local_rank, fps_override=0):
        import ctypes as C
        import torch
        from cmx_b200.engine import Engine, cell_to_c
        self.C, self.torch, self.args, self.rank, self.world, self.local_rank = C, torch, args, rank, world, local_rank
        w = self.w = build_workload(args, config, rank, world, fps_override)
        self.fps = w["fps"]
        self.xs, self.xv = frames_of(w, w["frame_ids"])
        self.iref = irefatom_of(w, self.xv[0])
        self.cell = w["system"].cell
        self.eng = Engine(solute=w["solute"], solvent=w["solvent"], options=w["opt"], irefatom=self.iref, autocorrelation=w["auto"],
                          device=local_rank, ring_slots=self.fps, n_streams=args.streams, batch_frames=args.batch)
        self.lib, self.h = self.eng.lib, self.eng.h
        self.cellc = cell_to_c(self.cell)
        self.cellp = self.cellc.ctypes.data_as(C.POINTER(C.c_double))
        self.d_xv = torch.from_numpy(self.xv).cuda()
        self.d_xs = self.d_xv if w["auto"] else torch.from_numpy(self.xs).cuda()
        self.sv_stride, self.ss_stride = self.d_xv[0].numel() * 4, self.d_xs[0].numel() * 4
        self.pv, self.ps = self.d_xv.data_ptr(), self.d_xs.data_ptr()
        for k in range(self.fps):   # fill every pinned slot once (untimed): slot k holds frame k
            a_s, a_v = self.eng.acquire()
            a_v[...] = self.xv[k]
            if not w["auto"]:
                a_s[...] = self.xs[k]
            self.eng.submit(w["frame_ids"][k], 1.0, self.cell)
        self.eng.sync(); self.eng.reset()
        self.scratch = None
        self.null_s, self.null_v = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()

    # ---- plumbing ----
    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            self.torch.cuda.synchronize()

    def counters_tensor(self):
        ptr, n = self.eng.counters_device()

        class W_:
            pass
        wobj = W_()
        wobj.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}
        return self.torch.as_tensor(wobj, device=f"cuda:{self.local_rank}")

    def allreduce_counts(self):
        """the one exchange step of the path: sum of the integer blocks over the ranks -- into a SCRATCH copy, so that
        the live accumulators of this rank keep counting its own frames only"""
        if self.world == 1:
            return None
        import torch.distributed as dist
        t = self.counters_tensor()
        if self.scratch is None:
            self.scratch = self.torch.empty_like(t)
        self.scratch.copy_(t)
        dist.all_reduce(self.scratch, op=dist.ReduceOp.SUM)
        return self.scratch

    def max_over_ranks(self, *vals):
        tt = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(v) for v in tt]

    # ---- steps ----
    def step_device(self, collective=True, nframes=None):
        C = self.C
        for k in range(nframes or self.fps):
            rc = self.lib.cmx_submit_frame_device(self.h, C.c_void_p(self.ps + k * self.ss_stride), C.c_void_p(self.pv + k * self.sv_stride),
                                                  self.w["frame_ids"][k], 1.0, self.cellp)
            if rc:
                raise RuntimeError(self.lib.cmx_last_error(self.h).decode())
        self.eng.sync()
        if collective:
            self.allreduce_counts()

    def feed_pinned(self, nframes, first_id=None):
        """nframes through acquire/submit; the pinned slots are cycled (slot k % fps already holds frame k % fps)"""
        C, fps, ids = self.C, self.fps, self.w["frame_ids"]
        for k in range(nframes):
            rc = self.lib.cmx_acquire_frame_buffer(self.h, C.byref(self.null_s), C.byref(self.null_v))
            rc = rc or self.lib.cmx_submit_frame(self.h, ids[k % fps], 1.0, self.cellp)
            if rc:
                raise RuntimeError(self.lib.cmx_last_error(self.h).decode())

    def step_e2e(self, nframes):
        self.feed_pinned(nframes)
        self.eng.sync()
        self.allreduce_counts()
        return self.eng.finish(copy=False)     # D2H of all counters into the engine's pinned result arrays

    # ---- measurements ----
    def measure_value(self, steps, warmup):
        args = self.args
        for _ in range(max(warmup, 3)):
            self.step_device()
        self.eng.reset()
        self.barrier()
        st0 = self.eng.stats()
        sampler = ClockSampler(self.local_rank)
        time.sleep(0.15)
        with sampler as clk:
            t0 = time.perf_counter()
            for _ in range(steps):
                self.step_device()
            self.barrier()
            wall = time.perf_counter() - t0
        sampler.close()
        st1 = self.eng.stats()
        dev_ms = st1["gpu_ms_total"] - st0["gpu_ms_total"]
        t_dev, wall = self.max_over_ranks(max(dev_ms * 1e-3, 1e-9), wall)
        # device events cover first-kernel..last-kernel of each step; the wall bracket (barrier+sync both sides)
        # is what the job takes: report the slower of the two views
        t_used = max(t_dev, wall)
        total_frames = self.fps * steps * self.world
        c = self.eng.finish()
        nfr = max(1, self.fps * steps)
        return dict(value=total_frames / t_used, t_used=t_used, t_dev=t_dev, wall=wall, clocks=clk.summary(),
                    launches=int(st1["kernel_launches"] - st0["kernel_launches"]), batches=int(st1["batches"] - st0["batches"]),
                    host_submit_ms=(st1["host_submit_ms"] - st0["host_submit_ms"]) / steps,
                    host_wait_ms=(st1["host_wait_ms"] - st0["host_wait_ms"]) / steps,
                    deferred=(st1["deferred"] - st0["deferred"]) / nfr,
                    hits_real=float(c["md_count"].sum()) / nfr, hits_random=float(c["md_count_random"].sum()) / nfr)

    def measure_e2e(self, steps, warmup):
        nfe = max(self.fps, int(round(self.w["traj_frames"] / 8)))     # this GPU's share of the trajectory on a full 8-GPU box
        self.eng.reset()
        for _ in range(max(1, min(warmup, 2))):
            self.step_e2e(min(nfe, 2 * self.fps))
        self.eng.reset()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = self.step_e2e(nfe)
            self.eng.reset()
        self.barrier()
        (te,) = self.max_over_ranks(time.perf_counter() - t0)
        in_bytes = (self.xv[0].nbytes if self.w["auto"] else self.xs[0].nbytes + self.xv[0].nbytes)
        out_bytes = sum(v.nbytes for v in res.values() if isinstance(v, np.ndarray))
        return {"value": nfe * steps * self.world / te, "unit": "frames/s", "h2d_bytes_per_step": int(in_bytes * nfe),
                "d2h_bytes_per_step": int(out_bytes), "frames_per_step": nfe, "steps": steps,
                "step": f"{nfe} frames from pinned host memory (the share of one GPU of 8 of the {self.w['traj_frames']}-frame trajectory; "
                        f"{self.fps} distinct frames cycled), one all-reduce and one cmx_finish (D2H of all counters) per step"}

    def measure_h2d_ceiling(self, in_bytes_per_frame):
        """What the host side allows: every rank copies pinned host memory to its GPU AT THE SAME TIME (the traffic
        pattern of the e2e leg without any kernel) -- aggregate GB/s and the frames/s ceiling it puts on `e2e`."""
        torch = self.torch
        n = 256 << 20
        src = torch.empty(n, dtype=torch.uint8).pin_memory()
        dst = torch.empty(n, dtype=torch.uint8, device=f"cuda:{self.local_rank}")
        reps = 8
        dst.copy_(src, non_blocking=True)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        self.barrier()
        (dt,) = self.max_over_ranks(time.perf_counter() - t0)
        agg = reps * n * self.world / dt / 1e9
        return {"aggregate_GBps": agg, "per_gpu_GBps": agg / self.world, "frames_per_s_ceiling": agg * 1e9 / in_bytes_per_frame,
                "what": f"{self.world} rank(s) copying 8 x 256 MiB of pinned host memory to their GPUs simultaneously, wall clock, max over ranks"}

    def measure_roofline(self, value):
        eng, w, fps = self.eng, self.w, self.fps
        eng.reset(); eng.set_option("active_streams", 1); eng.set_option("profile", 1)   # one batch at a time: clean per-kernel times
        s0 = eng.stats()
        self.step_device(collective=False)
        s1 = eng.stats()
        eng.set_option("profile", 0)
        ms_rand = (s1["gpu_ms_search_random"] - s0["gpu_ms_search_random"]) / fps
        ms_real = (s1["gpu_ms_search_real"] - s0["gpu_ms_search_real"]) / fps
        ms_frame = (s1["gpu_ms_total"] - s0["gpu_ms_total"]) / fps
        nbatch = max(1, s1["batches"] - s0["batches"])
        b_in, b_rand = algorithmic_bytes(w)
        peak, peak_src = peaks()
        pair_path = w["solute"].nmols > 1 and w["solute"].natomspermol <= 64     # the library's auto rule (cmx_config.path = 0)
        kname = ("k_pair_random", "k_pairs") if pair_path else ("k_tile_search<RANDOM>", "k_tile_search<REAL>")
        dom_ms, dom_bytes, dom = (ms_rand, b_rand, "random-phase search") if ms_rand >= ms_real else (ms_real, b_in, "real-phase search")
        kernel_name = kname[0] if ms_rand >= ms_real else kname[1]
        ach = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        eng.reset(); eng.set_option("count_pairs", 1)
        self.step_device(collective=False)
        pe = eng.stats()["pair_evals"] / fps
        eng.set_option("count_pairs", 0); eng.set_option("active_streams", 0)
        fpl = 1.0 if pair_path else fps / nbatch      # frames per launch (batched launches on the grid path; the pair path launches per frame)
        tpf = ncu_traffic(self.w["config"], "random" if dom.startswith("random") else "real")
        return {"bound": "hbm", "kernel": f"{kernel_name} ({dom})", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": None if tpf is None else tpf * fpl, "traffic_per_frame": tpf, "peak_source": peak_src,
                "frames_per_launch": fpl, "algorithmic_bytes_per_launch": dom_bytes * fpl, "kernel_ms_per_launch": dom_ms * fpl,
                "algorithmic_bytes_per_frame": dom_bytes, "kernel_ms_per_frame": dom_ms,
                "kernel_share_of_frame": dom_ms / ms_frame if ms_frame > 0 else None,
                "frame_algorithmic_bytes": b_in + b_rand, "frame_achieved_GBps": (b_in + b_rand) * value / self.world / 1e9,
                "pair_evals_per_frame": pe, "pair_evals_per_s": pe * value,
                # second view (the kernel is instruction-bound, not HBM-bound): pair evaluations per second inside the
                # dominant kernel against lanes x clock / 8 issue slots per pair (the SASS sweep loop of k_tile_search)
                "alu_view": {"kernel_pair_evals_per_s": (pe * (ms_rand / max(ms_rand + ms_real, 1e-12)) / (dom_ms * 1e-3)) if dom_ms > 0 else None,
                             "inner_loop_peak_pair_evals_per_s": 148 * 128 * 1.965e9 / 8.0}}

    def measure_cpu(self, ncores):
        """the CPU arm on a bounded sample + the parity guard: the device counters of the same frames must equal the
        oracle's in ALL eight arrays"""
        w, args = self.w, self.args
        nfc = args.cpu_frames or min(self.fps, max(ncores, 8))
        ids = w["frame_ids"][:nfc]
        cpu_arm(w, self.xs[:2], self.xv[:2], ids[:2], ncores)   # warm the library
        val_cpu, dt, o = cpu_arm(w, self.xs[:nfc], self.xv[:nfc], ids, ncores)
        self.eng.reset()
        for k in range(nfc):
            self.eng.submit_device(self.ps + k * self.ss_stride, self.pv + k * self.sv_stride, self.cell, frame_index=ids[k])
        devc = self.eng.finish()
        ref = o.counters()
        bad = [k for k in COUNTER_KEYS if not np.array_equal(devc[k], ref[k])]
        return {"value": val_cpu, "unit": "frames/s", "cores": ncores, "kind": "port",
                "sample": f"{nfc} frames of the same workload, oracle/cmx_oracle.c cell-list path, {ncores} OpenMP threads, frame-parallel",
                "counts_equal_device": not bad, "arrays_compared": list(COUNTER_KEYS), "frames_compared": nfc, "arrays_differing": bad}

    def guard_multi_gpu(self):
        """N > 1: every rank submits its first two frames, the blocks are all-reduced (scratch), and rank 0 checks the sum
        against its own single-GPU run of the same 2 N frames (bit-exact: integer sums, frame-keyed Philox)."""
        import torch.distributed as dist
        w = self.w
        self.eng.reset()
        self.step_device(collective=False, nframes=2)
        summed = self.allreduce_counts().clone()
        ok = None
        if self.rank == 0:
            ids = [1 + r + self.world * k for r in range(self.world) for k in range(2)]
            xs, xv = frames_of(w, ids)
            self.eng.reset()
            self.eng.sync()
            for k, fid in enumerate(ids):
                self.eng.submit_arrays(xs[k], xv[k], self.cell, frame_index=fid)
            self.eng.sync()
            ok = bool(self.torch.equal(self.counters_tensor(), summed))
        self.eng.reset()
        dist.barrier()
        return ok

    def close(self):
        if self.eng is not None:
            self.eng.close()
            self.eng = None
        self.d_xv = self.d_xs = self.scratch = None


def job_leg(args, b):
    """Strong scaling of the actual job: the configuration's whole trajectory shared by the ranks, through the public
    sequence create -> acquire/submit (pinned host frames, cycled) -> sync -> all-reduce -> finish -> destroy."""
    from cmx_b200.engine import Engine
    w, world = b.w, b.world
    total = w["traj_frames"]
    mine = len(range(b.rank, total, world))
    b.barrier()
    t0 = time.perf_counter()
    eng = Engine(solute=w["solute"], solvent=w["solvent"], options=w["opt"], irefatom=b.iref, autocorrelation=w["auto"],
                 device=b.local_rank, n_streams=args.streams, batch_frames=args.batch)
    t_create = time.perf_counter() - t0
    C = b.C
    fps = b.fps
    lib, h = eng.lib, eng.h
    ps, pv = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
    done = 0
    seen = set()                                 # a slot of the ring is filled the first time it is handed out (host memcpy, as a
    for k in range(mine):                        # reader would); later acquisitions re-send the bytes it already holds: no file I/O in this leg
        rc = lib.cmx_acquire_frame_buffer(h, C.byref(ps), C.byref(pv))
        if rc:
            raise RuntimeError(lib.cmx_last_error(h).decode())
        addr = C.cast(pv, C.c_void_p).value
        if addr not in seen:
            seen.add(addr)
            C.memmove(pv, b.xv[k % fps].ctypes.data, b.xv[k % fps].nbytes)
            if not w["auto"]:
                C.memmove(ps, b.xs[k % fps].ctypes.data, b.xs[k % fps].nbytes)
        rc = lib.cmx_submit_frame(h, 1 + b.rank + world * k, 1.0, b.cellp)
        if rc:
            raise RuntimeError(lib.cmx_last_error(h).decode())
        done = k + 1
        if (k & 63) == 63 and time.perf_counter() - t0 > 120.0:      # never lets a sick run hold the ranks' collectives hostage
            break
    eng.sync()
    t_feed = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        ptr, n = eng.counters_device()

        class W_:
            pass
        wobj = W_()
        wobj.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}
        t = b.torch.as_tensor(wobj, device=f"cuda:{b.local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        b.torch.cuda.synchronize()
    # sum!(R, r_chunk) leaves ONE Result: after the all-reduce every rank holds the sums, rank 0 reads them back
    res = eng.finish(copy=False) if b.rank == 0 else None
    t_finish = time.perf_counter() - t0
    hits = float(res["md_count"].sum()) if res is not None else 0.0
    t1 = time.perf_counter()
    eng.close()
    t_destroy = time.perf_counter() - t1
    b.barrier()
    (wall,) = b.max_over_ranks(time.perf_counter() - t0)
    (done_all,) = b.max_over_ranks(float(mine - done))
    if done_all > 0:     # a rank stopped early (120 s guard): the leg is reported as truncated, not as a throughput
        return {"frames": total, "frames_per_rank": mine, "truncated_after_s": wall, "frames_per_s": None, "md_count_sum": hits}
    return {"frames": total, "frames_per_rank": mine, "wall_s": wall, "frames_per_s": total / wall, "create_s": t_create,
            "feed_s": t_feed - t_create, "allreduce_finish_s": t_finish - t_feed, "destroy_s": t_destroy, "md_count_sum": hits,
            "what": "whole trajectory, strong scaling: create -> acquire/submit from pinned host frames -> sync -> all-reduce -> finish (rank 0) -> destroy, wall clock"}


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args, ncores)))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"), timeout=datetime.timedelta(seconds=300))

    b = Bench(args, args.config, rank, world, local_rank)
    w, fps = b.w, b.fps
    v = b.measure_value(args.steps, args.warmup)
    e2e = None if args.no_e2e else b.measure_e2e(max(2, min(args.steps, 5)), args.warmup)
    if e2e is not None:
        e2e["h2d_ceiling"] = b.measure_h2d_ceiling(e2e["h2d_bytes_per_step"] / e2e["frames_per_step"])
    guard = b.guard_multi_gpu() if world > 1 else None
    roof = b.measure_roofline(v["value"]) if rank == 0 and not args.no_roofline else None
    if world > 1:
        dist.barrier()
    cpu = b.measure_cpu(ncores) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    nbins = b.eng.nbins
    job = None if args.no_job else job_leg(args, b)
    b.close()

    # ---- secondary: the C2 line of the previous rounds (value + e2e), same rank layout ----
    secondary = None
    if not args.no_secondary and args.config != "C2":
        b2 = Bench(args, "C2", rank, world, local_rank)
        v2 = b2.measure_value(max(3, args.steps // 2), 3)
        e2 = None if args.no_e2e else b2.measure_e2e(2, 1)
        secondary = {"metric": METRIC, "value": v2["value"], "unit": "frames/s", "ms_per_step": 1e3 * v2["t_used"] / max(3, args.steps // 2),
                     "config": {"workload": b2.w["desc"], "frames_per_step": b2.fps}, "e2e": e2, "gpu_launches": v2["launches"],
                     "launches_per_frame": v2["launches"] / max(1, b2.fps * max(3, args.steps // 2)),
                     "host_submit_ms_per_step": v2["host_submit_ms"], "host_wait_ms_per_step": v2["host_wait_ms"]}
        b2.close()

    # ---- the one HBM-bound kernel next to the path (rank 0, N=1): per-residue sums of a per-atom contribution array of
    # C5's shape (1e6 rows x 750 bins x 8 B = 6 GB), cmx_reduce_groups; reported beside the dominant kernel's roofline
    hbm_kernel = None
    if rank == 0 and world == 1 and not args.no_hbm_kernel and args.scale == 1.0:
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
        steps = args.steps
        line = {"metric": METRIC, "value": v["value"], "unit": "frames/s",
                "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * v["t_used"] / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 search + f64 finalisation",
                "data": "synthetic",
                "config": {"workload": w["desc"], "frames_per_step": fps, "frames_per_step_total": fps * world, "scale": args.scale,
                           "n_random_samples": w["opt"].n_random_samples, "bulk_range": [w["opt"].dbulk, w["opt"].cutoff], "nbins": nbins,
                           "l2": f"distinct inputs per step = {fps * (b.xv[0].nbytes + (0 if w['auto'] else b.xs[0].nbytes)) / 1e6:.0f} MB (larger than the 126 MB L2; no flush needed)",
                           "batch_frames": args.batch or "auto", "batches_in_flight": args.streams or "auto", "allreduce_per_step": world > 1,
                           "hits_per_frame": v["hits_real"], "hits_per_frame_random": v["hits_random"],
                           "deferred_to_exact_per_frame": v["deferred"]},
                "device_ms_per_step": 1e3 * v["t_dev"] / steps, "wall_ms_per_step": 1e3 * v["wall"] / steps,
                "host_submit_ms_per_step": v["host_submit_ms"], "host_wait_ms_per_step": v["host_wait_ms"],
                "host_cpu_ms_per_step": v["host_submit_ms"] - v["host_wait_ms"],
                "gpu_launches": v["launches"], "launches_per_frame": v["launches"] / max(1, fps * steps),
                "clocks": v["clocks"], "e2e": e2e, "job": job, "multi_gpu_sum_equals_single_gpu": guard,
                "roofline": roof, "roofline_hbm_kernel": hbm_kernel, "cpu_baseline": cpu, "secondary": secondary}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
