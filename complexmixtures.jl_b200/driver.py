"""mddf() / coordination_number() -- the drivers of the reference (src/mddf.jl:186-347, 539-574)
with the chunk loop (src/mddf.jl:263-339) replaced by calls into libcmx_b200.so.

What is kept from the reference: argument validation, TrajectoryMetaData, Result construction,
frame selection (firstframe/lastframe/stride, zero-weight frames skipped: goto_nextframe!,
src/mddf.jl:95-111), the cooperative stop file (src/mddf.jl:301-304) and finalresults!.
What changes (src/parallel_setup.jl): instead of nthreads chunk tasks with private Result copies,
computed frames are dealt round-robin to the ranks of a torch.distributed job (one process per
GPU); every rank runs its own engine and the integer counters are summed with ONE all-reduce.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from .engine import DcdFile, Engine, XtcFile
from .options import Options
from .results import Result, finalresults, new_result
from .selection import AtomSelection
from .trajectory import NamdDCD, Trajectory, XTCTraj, make_trajectory, trajectory_metadata


def frames_to_compute(options: Options, lastframe_read: int, frame_weights) -> list:
    """1-based frame numbers that are computed (to_compute_frames minus zero-weight frames)."""
    fw = np.asarray(frame_weights, dtype=np.float64).reshape(-1)
    out = []
    for iframe in range(options.firstframe, lastframe_read + 1, options.stride):
        w = 1.0 if fw.size == 0 else float(fw[iframe - 1])
        if w != 0.0:
            out.append((iframe, w))
    return out


def shard(frames: list, rank: int, world: int) -> list:
    """Frame f -> rank (position in the computed-frame list) mod world."""
    return [fw for k, fw in enumerate(frames) if k % world == rank]


def _dist_info():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def _device_tensor(ptr: int, n: int, typestr: str, device: int):
    import torch

    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(w, device=f"cuda:{device}")


def weights_agree(my_weights, allreduce_min) -> bool:
    """True iff every frame of EVERY rank carries one and the same weight.  ``allreduce_min(values)`` returns the
    element-wise minimum of ``values`` over the ranks (a collective: every rank must call this function, with or
    without frames of its own)."""
    ws = [float(x) for x in my_weights]
    lo, hi = (min(ws), max(ws)) if ws else (float("inf"), float("-inf"))
    glo, gneg_hi = allreduce_min([lo, -hi])            # global min weight and -(global max weight)
    return float(glo) == -float(gneg_hi)


def allreduce_counters(eng: Engine, my_weights) -> tuple:
    """The single exchange step: sum the accumulators over the GPUs (sum!, src/results.jl:629-649).

    Integer hits can only be summed when every frame of every rank had ONE and the same weight (each rank scales its
    integers by its own weight in cmx_finish); the ranks agree on that collectively BEFORE the collective, so no rank can
    take the other branch.  Otherwise every rank folds its counters to f64 with its weights applied and the f64 arrays
    are summed.  Returns the all-reduced (volume_total, sum_weights)."""
    import torch
    import torch.distributed as dist
    dev = eng.cfg.device
    st = eng.stats()

    def _min(values):
        t = torch.tensor(values, dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return t.tolist()
    if weights_agree(my_weights, _min):
        ptr, n = eng.counters_device()
        t = _device_tensor(ptr, n, "<i8", dev)
    else:
        ptr, n = eng.counters_device_f64()
        t = _device_tensor(ptr, n, "<f8", dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    s = torch.tensor([st["volume_total"], st["sum_weights"]], dtype=torch.float64, device=t.device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize(t.device)
    return float(s[0]), float(s[1])


def mddf(trajectory, solute: Optional[AtomSelection] = None, solvent: Optional[AtomSelection] = None,
         options: Optional[Options] = None, *, trajectory_format: str = "", frame_weights=(),
         coordination_number_only: bool = False, low_memory: bool = False, device: Optional[int] = None,
         devices=None, path: int = 0, feed: str = "auto", reader_threads: int = 0, group_arrays: bool = True,
         _engine_kw: Optional[dict] = None, _engine_cache: Optional[dict] = None, distributed: Optional[bool] = None) -> Result:
    """mddf(trajectory_file, solute, solvent, options; ...) or mddf(trajectory, options; ...).

    ``low_memory`` is accepted for compatibility and is a no-op: the device keeps ONE set of
    counters per GPU regardless of the thread count (src/parallel_setup.jl:21-54 does not apply).

    ``devices``: several GPUs behind ONE engine in this process (``cmx_config.n_devices``; the reference's single call
    uses the whole machine, src/parallel_setup.jl): frames are dealt to the devices in order and the per-device
    counters are summed on the first device inside ``cmx_finish``.  Under ``torchrun`` (one process per GPU) leave it
    unset: the ranks share the frames and sum with one all-reduce.

    ``group_arrays=False``: the group-count arrays (per atom / per custom group x nbins: 12 GB for a 5 M-atom system with
    per-atom contributions) STAY on the device.  The Result carries the O(nbins) vectors as usual and ``R.device``, a
    :class:`DeviceResult` whose ``contributions(side, groups, type)`` / ``reduce_groups`` / ``final_results`` evaluate
    ``contributions(R, group; type)`` (src/tools/contributions.jl:70-248) for any number of groups on the accumulators in
    HBM; call ``R.device.close()`` when done.

    ``feed``: "native" = the library's own DCD / XTC feed (``cmx_run_dcd`` / ``cmx_run_xtc``: reader threads -> pinned ring ->
    raw frame H2D -> device gather of the selections; the frames of this rank are read by offset, the
    others are never touched), "host" = this module's reader writing into the pinned staging slot
    (``cmx_acquire_frame_buffer`` / ``cmx_submit_frame``), "auto" = native for DCD and XTC files.  Both give
    the same counters.  The cooperative stop file (src/mddf.jl:301-304) is polled by both feeds (the native one checks it
    between ring refills inside ``cmx_run_dcd`` / ``cmx_run_xtc``).
    """
    if isinstance(trajectory, str):
        if isinstance(solvent, Options) and options is None:      # mddf(file, solute_and_solvent, options)
            options, solvent = solvent, None
        options = options or Options()
        trajectory = make_trajectory(trajectory, solute, solvent, format=trajectory_format, lastframe=options.lastframe)
    else:
        if isinstance(solute, Options) and options is None:       # mddf(trajectory, options)
            options = solute
        options = options or Options()
    assert isinstance(trajectory, Trajectory)
    if feed not in ("auto", "native", "host"):
        raise ValueError("feed must be 'auto', 'native' or 'host'")
    native = feed == "native" or (feed == "auto" and isinstance(trajectory, (NamdDCD, XTCTraj)))
    if native and not isinstance(trajectory, (NamdDCD, XTCTraj)):
        raise ValueError("feed='native' needs a DCD or XTC trajectory")
    tmeta = trajectory_metadata(trajectory, options)
    R = new_result(trajectory, options, tmeta, frame_weights)
    rank, world = (0, 1) if distributed is False else _dist_info()   # distributed=False: ignore an initialised process group
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0
    # mddf_many: one engine (device state, streams, staging ring) serves every trajectory of the batch
    key = (tmeta.irefatom, R.autocorrelation, coordination_number_only, device, path, tuple(devices or ()))
    eng = None if _engine_cache is None else _engine_cache.get(key)
    if eng is None:
        eng = Engine(solute=trajectory.solute, solvent=trajectory.solvent, options=options, irefatom=tmeta.irefatom,
                     autocorrelation=R.autocorrelation, coordination_number_only=coordination_number_only, device=device,
                     path=path, devices=devices, **(_engine_kw or {}))
        if _engine_cache is not None:
            _engine_cache[key] = eng
    else:
        eng.reset()
    todo = frames_to_compute(options, tmeta.lastframe_read, R.files[0].frame_weights)
    if native:
        my = shard(todo, rank, world)
        is_dcd = isinstance(trajectory, NamdDCD)
        src = DcdFile(trajectory.filename) if is_dcd else XtcFile(trajectory.filename)
        try:
            w = [wt for _, wt in my]
            (eng.run_dcd if is_dcd else eng.run_xtc)(src, trajectory.solute.indices, trajectory.solvent.indices, [f - 1 for f, _ in my],
                                                     None if all(v == 1.0 for v in w) else w, n_reader_threads=reader_threads)
        finally:
            eng.sync()
            src.close()
    else:
        mine = set(f for f, _ in shard(todo, rank, world))
        weights = dict(todo)
        trajectory.open()
        trajectory.firstframe()
        try:
            for iframe in range(1, tmeta.lastframe_read + 1):
                if os.path.isfile("stop_complexmixtures"):   # src/mddf.jl:301-304
                    break
                if iframe in mine:
                    xs, xv = eng.acquire()
                    trajectory.nextframe(xs, xv)             # reader writes fp32 straight into the pinned slot
                    eng.submit(iframe, weights[iframe], trajectory.getunitcell())
                else:
                    trajectory.nextframe()
        finally:
            trajectory.close()
    if world > 1:
        vol, sw = allreduce_counters(eng, [wt for _, wt in shard(todo, rank, world)])
        c = eng.finish(groups=group_arrays)   # ONE read-back of the (summed) counters
        c["volume_total"], c["sum_weights"] = vol, sw
    else:
        c = eng.finish(groups=group_arrays)
    R.engine_stats = eng.stats()
    for k in ("md_count", "md_count_random", "rdf_count", "rdf_count_random"):
        setattr(R, k, c[k])
    for k in ("solute_group_count", "solute_group_count_random", "solvent_group_count", "solvent_group_count_random"):
        setattr(R, k, c[k] if group_arrays else np.zeros((0, R.nbins)))
    R.volume.total = c["volume_total"]
    if not group_arrays:
        R.device = DeviceResult(eng, c["sum_weights"], c["volume_total"], owns=_engine_cache is None)
    elif _engine_cache is None:
        eng.close()
    return finalresults(R, options, coordination_number_only=coordination_number_only)


class DeviceResult:
    """The accumulators of a finished run, still in HBM (``mddf(..., group_arrays=False)``): the post-processing calls of
    the reference that read the group arrays, evaluated on the device for many groups at once."""

    def __init__(self, eng: Engine, sum_weights: float, volume_sum: float, owns: bool = True):
        self.engine, self.sum_weights, self.volume_sum, self._owns = eng, float(sum_weights), float(volume_sum), owns

    def contributions(self, side: str, groups, type: str = "mddf") -> np.ndarray:
        """[n_groups, nbins]: ``contributions(R, SoluteGroup|SolventGroup(rows); type)`` for every group of 0-based rows"""
        return self.engine.contributions(side, groups, type, self.sum_weights, self.volume_sum)

    def reduce_groups(self, which: str, groups) -> np.ndarray:
        return self.engine.reduce_groups(which, groups)

    def final_results(self) -> dict:
        return self.engine.final_results(self.sum_weights, self.volume_sum)

    def close(self):
        if self._owns and self.engine is not None:
            self.engine.close()
        self.engine = None


def coordination_number(trajectory, solute=None, solvent=None, options=None, **kw) -> Result:
    """coordination_number(...) == mddf(...; coordination_number_only=true), src/mddf.jl:539-574."""
    if "coordination_number_only" in kw:
        raise ValueError("The keyword argument `coordination_number_only` is not valid for this function. "
                         "It is, by definition, set to `true` in this function.")
    return mddf(trajectory, solute, solvent, options, coordination_number_only=True, **kw)


def mddf_many(trajectory_files, solute: AtomSelection, solvent: Optional[AtomSelection] = None,
              options: Optional[Options] = None, *, frame_weights=None, **kw):
    """The same analysis on several trajectories (or parts of one), the use case of ``merge``
    (src/tools/merge.jl:1-9): returns ``(results, merge(results))``.  Unlike calling ``mddf`` in a loop, ONE
    engine -- device buffers, streams, pinned staging ring -- serves the whole batch (``cmx_reset`` between
    files), so short trajectories do not pay the create/destroy cost again."""
    from .results import merge
    options = options or Options()
    cache: dict = {}
    results = []
    try:
        for k, f in enumerate(trajectory_files):
            fw = () if frame_weights is None else frame_weights[k]
            results.append(mddf(f, solute, solvent, options, frame_weights=fw, _engine_cache=cache, **kw))
    finally:
        for eng in cache.values():
            eng.close()
    return results, merge(results)
