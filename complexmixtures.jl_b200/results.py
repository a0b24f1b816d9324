"""Result model and the O(nbins) normalisation that stays on the host.

Reference: src/results.jl -- ``Result`` :61-118, constructor + frame-weight validation
:124-193, ``setbin`` :28, ``set_samples`` :230-237, ``shellradius`` :272-275,
``sphericalshellvolume`` :252-255, ``sum_frame_weights`` :288-297, ``_mddf_final_results!``
:320-376, ``renormalize!`` :378-428, ``_coordination_number_final_results!`` :430-469,
``save``/``load`` :533-588; units: src/io.jl:6-14.

The hot path (device) fills the *_count arrays and volume.total through ``cmx_finish``;
everything here is post-processing of those arrays, field-for-field what the reference does.
"""
from __future__ import annotations

import json
import math
import warnings
from dataclasses import dataclass, field
from typing import List

import numpy as np

from .options import Options
from .selection import AtomSelection

VERSION = "2.18.3-DEV+b200"
MOLE = 6.022140857e23
ANGS3_TO_CM3_PER_MOL = MOLE / 1e24  # units.Angs3tocm3permol, src/io.jl:10


def setbin(d: float, step: float) -> int:
    """1-based histogram bin, src/results.jl:28."""
    return max(1, math.ceil(d / step))


def sphericalshellvolume(i: int, step: float) -> float:
    rmin = (i - 1) * step
    return (4 * math.pi / 3) * ((rmin + step) ** 3 - rmin ** 3)


def shellradius(i, step):
    rmin = (np.asarray(i, dtype=np.float64) - 1) * step
    return (0.5 * ((rmin + step) ** 3 + rmin ** 3)) ** (1 / 3)


@dataclass
class Density:
    solute: float = 0.0
    solvent: float = 0.0
    solvent_bulk: float = 0.0


@dataclass
class Volume:
    total: float = 0.0
    bulk: float = 0.0
    domain: float = 0.0
    shell: np.ndarray = None


@dataclass
class TrajectoryFileOptions:
    """src/Options.jl:274-283."""
    filename: str
    options: Options
    irefatom: int
    lastframe_read: int
    nframes_read: int
    frame_weights: np.ndarray


@dataclass
class Result:
    nbins: int
    dbulk: float
    cutoff: float
    autocorrelation: bool
    solute: AtomSelection
    solvent: AtomSelection
    files: List[TrajectoryFileOptions]
    weights: List[float] = field(default_factory=lambda: [1.0])
    Version: str = VERSION

    def __post_init__(self):
        nb = self.nbins
        z = lambda: np.zeros(nb)
        self.d = z()
        self.md_count, self.md_count_random = z(), z()
        self.coordination_number, self.coordination_number_random = z(), z()
        self.mddf, self.kb = z(), z()
        self.solute_group_count = np.zeros((self.solute.n_groups, nb))
        self.solute_group_count_random = np.zeros((self.solute.n_groups, nb))
        self.solvent_group_count = np.zeros((self.solvent.n_groups, nb))
        self.solvent_group_count_random = np.zeros((self.solvent.n_groups, nb))
        self.rdf_count, self.rdf_count_random = z(), z()
        self.sum_rdf_count, self.sum_rdf_count_random = z(), z()
        self.rdf, self.kb_rdf = z(), z()
        self.density = Density()
        self.volume = Volume(shell=z())

    @property
    def options(self) -> Options:
        return self.files[0].options

    @property
    def irefatom(self) -> int:
        return self.files[0].irefatom


def new_result(trajectory, options: Options, trajectory_data, frame_weights=()) -> Result:
    """Result(trajectory, options; trajectory_data, frame_weights), src/results.jl:124-193."""
    nbins = setbin(options.cutoff, options.binstep)
    fw = np.asarray(frame_weights, dtype=np.float64)
    if fw.size > 0:
        if sum(1 for s in fw.shape if s > 1) > 1:
            raise ValueError(f"The frame_weights provided must be one-dimensional, but a {fw.shape}-dimensional array was given.")
        fw = fw.reshape(-1)
        if len(fw) < trajectory_data.lastframe_read:
            raise ValueError("The length of the frame_weights vector provided must at least the number of frames to be read.")
        rng = range(options.firstframe, trajectory_data.lastframe_read + 1, options.stride)
        if sum(fw[i - 1] for i in rng) <= 0.0:
            raise ValueError(f"The sum of the frame weights must be greater than zero for the frames that will be considered: {rng}")
    else:
        fw = np.ones(trajectory_data.lastframe_read)
    return Result(nbins=nbins, dbulk=options.dbulk, cutoff=options.cutoff,
                  autocorrelation=trajectory.autocorrelation, solute=trajectory.solute, solvent=trajectory.solvent,
                  files=[TrajectoryFileOptions(trajectory.filename, options, trajectory_data.irefatom,
                                               trajectory_data.lastframe_read, trajectory_data.nframes_read, fw)])


def set_samples(R: Result):
    """src/results.jl:230-237."""
    n = R.solvent.nmols - 1 if R.autocorrelation else R.solvent.nmols
    return n, R.options.n_random_samples


def sum_frame_weights(R: Result) -> float:
    """src/results.jl:288-297."""
    Q = 0.0
    for f in R.files:
        Q += float(sum(f.frame_weights[i - 1] for i in range(f.options.firstframe, f.lastframe_read + 1, f.options.stride)))
    return Q


def finalresults(R: Result, options: Options, *, coordination_number_only: bool = False) -> Result:
    """finalresults!, src/results.jl:311-318."""
    return _coordination_number_final_results(R, options) if coordination_number_only else _mddf_final_results(R, options)


def _mddf_final_results(R: Result, options: Options) -> Result:
    """src/results.jl:320-376."""
    R.d[:] = shellradius(np.arange(1, R.nbins + 1), options.binstep)
    nsolv, nrand = set_samples(R)
    Q = sum_frame_weights(R)
    R.md_count /= R.solute.nmols * Q
    R.solute_group_count /= R.solute.nmols * Q
    R.solute_group_count_random /= nrand * Q
    if R.autocorrelation:
        R.solvent_group_count = R.solute_group_count.copy()
        R.solvent_group_count_random = R.solute_group_count_random.copy()
    else:
        R.solvent_group_count /= R.solute.nmols * Q
        R.solvent_group_count_random /= nrand * Q
    R.md_count_random /= nrand * Q
    R.rdf_count /= R.solute.nmols * Q
    R.rdf_count_random /= nrand * Q
    R.volume.total = R.volume.total / Q
    R.volume.shell = R.volume.total * (R.rdf_count_random / nsolv)
    binstep = R.options.binstep
    ibulk = setbin(R.dbulk + 0.5 * binstep, binstep)
    R.volume.domain = float(np.sum(R.volume.shell[: ibulk - 1]))
    if not R.options.usecutoff:
        R.volume.bulk = R.volume.total - R.volume.domain
        n_solvent_in_bulk = nsolv - float(np.sum(R.rdf_count))
    else:
        n_solvent_in_bulk = float(np.sum(R.rdf_count[ibulk - 1: R.nbins]))
        R.volume.bulk = float(np.sum(R.volume.shell[ibulk - 1: R.nbins]))
    R.density.solvent = R.solvent.nmols / R.volume.total
    R.density.solute = R.solute.nmols / R.volume.total
    with np.errstate(divide="ignore", invalid="ignore"):   # IEEE semantics as in Julia (Inf/NaN, no exception)
        R.density.solvent_bulk = float(np.float64(n_solvent_in_bulk) / np.float64(R.volume.bulk))
    density_fix = np.float64(R.density.solvent_bulk) / np.float64(R.density.solvent)
    return renormalize_(R, density_fix, silent=options.silent)


def renormalize_(R: Result, density_fix: float, *, silent: bool = True) -> Result:
    """renormalize!, src/results.jl:378-428."""
    R.md_count_random *= density_fix
    R.rdf_count_random *= density_fix
    R.solute_group_count_random *= density_fix
    R.solvent_group_count_random *= density_fix
    R.coordination_number = np.cumsum(R.md_count)
    R.coordination_number_random = np.cumsum(R.md_count_random)
    pos = R.md_count_random > 0.0
    if not silent and not np.all(pos):
        warnings.warn("Ideal-gas histogram bins with zero samples. Increase n_random_samples, "
                      "number of trajectory frames, and/or bin size.")
    R.mddf = np.zeros(R.nbins)
    R.mddf[pos] = R.md_count[pos] / R.md_count_random[pos]
    with np.errstate(divide="ignore", invalid="ignore"):
        R.kb = ANGS3_TO_CM3_PER_MOL * (1 / np.float64(R.density.solvent_bulk)) * (R.coordination_number - R.coordination_number_random)
        posr = R.rdf_count_random > 0.0
        R.rdf = np.zeros(R.nbins)
        R.rdf[posr] = R.rdf_count[posr] / R.rdf_count_random[posr]
        R.sum_rdf_count = np.cumsum(R.rdf_count)
        R.sum_rdf_count_random = np.cumsum(R.rdf_count_random)
        R.kb_rdf = ANGS3_TO_CM3_PER_MOL * (1 / np.float64(R.density.solvent_bulk)) * (R.sum_rdf_count - R.sum_rdf_count_random)
    return R


def _coordination_number_final_results(R: Result, options: Options) -> Result:
    """src/results.jl:430-469."""
    if not options.silent:
        warnings.warn("coordination_number_only was set to true, so the MDDF and KB integrals were not computed.")
    R.d[:] = shellradius(np.arange(1, R.nbins + 1), options.binstep)
    Q = sum_frame_weights(R)
    R.md_count /= R.solute.nmols * Q
    R.solute_group_count /= R.solute.nmols * Q
    if R.autocorrelation:
        R.solvent_group_count = R.solute_group_count.copy()
    else:
        R.solvent_group_count /= R.solute.nmols * Q
    R.rdf_count /= R.solute.nmols * Q
    R.volume.total = R.volume.total / Q
    R.density.solvent = R.solvent.nmols / R.volume.total
    R.density.solute = R.solute.nmols / R.volume.total
    R.coordination_number = np.cumsum(R.md_count)
    R.sum_rdf_count = np.cumsum(R.rdf_count)
    return R


# ---------------------------------------------------------------------------------------
# JSON schema of the reference (src/results.jl:533-588; golden files under test/data)
# ---------------------------------------------------------------------------------------
_VEC = ["d", "md_count", "md_count_random", "coordination_number", "coordination_number_random", "mddf", "kb",
        "rdf_count", "rdf_count_random", "sum_rdf_count", "sum_rdf_count_random", "rdf", "kb_rdf"]
_MAT = ["solute_group_count", "solvent_group_count", "solute_group_count_random", "solvent_group_count_random"]


def _same_selection(a: AtomSelection, b: AtomSelection) -> bool:
    return (a.nmols == b.nmols and a.natomspermol == b.natomspermol and len(a.indices) == len(b.indices)
            and bool(np.array_equal(a.indices, b.indices)) and a.n_groups == b.n_groups)


def merge(results: List[Result]) -> Result:
    """merge(r::Vector{Result}), src/tools/merge.jl:10-148: averages of the functions and counters of
    the sets, weighted by the (weighted) number of frames of each set; ``weights`` of the merged
    Result = fraction of the frames read from each file."""
    results = list(results)
    if not results:
        raise ValueError("merge needs at least one Result")
    r0 = results[0]
    for i, ri in enumerate(results):
        for rj in results[i + 1:]:
            if ri.nbins != rj.nbins:
                raise ValueError("To merge Results, the number of bins of the histograms of the sets must be the same.")
            if not np.isclose(ri.cutoff, rj.cutoff):
                raise ValueError("To merge Results, cutoff distance of the of the histograms of the sets must be the same.")
            if not _same_selection(ri.solute, rj.solute) or not _same_selection(ri.solvent, rj.solvent):
                raise ValueError("To merge Results, the solute and solvent selections of the sets must be the same.")
    ntot_frames = sum(f.nframes_read for r in results for f in r.files)
    tot_frame_weight = sum(sum_frame_weights(r) for r in results)
    files = [f for r in results for f in r.files]
    R = Result(nbins=r0.nbins, dbulk=r0.dbulk, cutoff=r0.cutoff, autocorrelation=r0.autocorrelation, solute=r0.solute,
               solvent=r0.solvent, files=files, weights=[f.nframes_read / ntot_frames for f in files])
    R.d = r0.d.copy()
    arrays = ("mddf", "kb", "rdf", "kb_rdf", "md_count", "md_count_random", "coordination_number",
              "coordination_number_random", "solute_group_count", "solute_group_count_random", "solvent_group_count",
              "solvent_group_count_random", "rdf_count", "rdf_count_random", "sum_rdf_count", "sum_rdf_count_random")
    for r in results:
        w = sum_frame_weights(r) / tot_frame_weight
        for k in arrays:
            setattr(R, k, getattr(R, k) + w * np.asarray(getattr(r, k)))
        R.density.solute += w * r.density.solute
        R.density.solvent += w * r.density.solvent
        R.density.solvent_bulk += w * r.density.solvent_bulk
        R.volume.total += w * r.volume.total
        R.volume.bulk += w * r.volume.bulk
        R.volume.domain += w * r.volume.domain
        R.volume.shell = R.volume.shell + w * np.asarray(r.volume.shell)
    return R


def save(R: Result, filename: str) -> str:
    out = {"Version": R.Version, "nbins": R.nbins, "dbulk": R.dbulk, "cutoff": R.cutoff,
           "autocorrelation": R.autocorrelation, "solute": R.solute.to_dict(), "solvent": R.solvent.to_dict()}
    for k in _VEC:
        out[k] = np.asarray(getattr(R, k)).tolist()
    for k in _MAT:
        out[k] = [row.tolist() for row in np.asarray(getattr(R, k))]
    out["density"] = dict(solute=R.density.solute, solvent=R.density.solvent, solvent_bulk=R.density.solvent_bulk)
    out["volume"] = dict(total=R.volume.total, bulk=R.volume.bulk, domain=R.volume.domain, shell=np.asarray(R.volume.shell).tolist())
    out["files"] = [dict(filename=f.filename, options=f.options.to_dict(), irefatom=f.irefatom,
                         lastframe_read=f.lastframe_read, nframes_read=f.nframes_read,
                         frame_weights=np.asarray(f.frame_weights).tolist()) for f in R.files]
    out["weights"] = list(R.weights)
    with open(filename, "w") as f:
        json.dump(out, f)
    return filename


def _sel_from_dict(d) -> AtomSelection:
    import warnings as _w
    with _w.catch_warnings():
        _w.simplefilter("ignore")
        return AtomSelection(d["indices"], nmols=d["nmols"], natomspermol=d["natomspermol"],
                             group_atom_indices=d.get("group_atom_indices") or None,
                             group_names=d.get("group_names") or None)


def load(filename: str) -> Result:
    with open(filename) as f:
        d = json.load(f)
    files = []
    for fo in d["files"]:
        o = dict(fo["options"])
        opts = Options(firstframe=o["firstframe"], lastframe=o["lastframe"], stride=o["stride"], irefatom=o["irefatom"],
                       n_random_samples=o["n_random_samples"], binstep=o["binstep"], dbulk=o["dbulk"],
                       cutoff=o["cutoff"] if o["usecutoff"] else None, usecutoff=o["usecutoff"], lcell=o["lcell"],
                       GC=o["GC"], GC_threshold=o["GC_threshold"], seed=o["seed"], StableRNG=o["StableRNG"],
                       nthreads=o["nthreads"], silent=True)
        files.append(TrajectoryFileOptions(fo["filename"], opts, fo["irefatom"], fo["lastframe_read"],
                                           fo["nframes_read"], np.asarray(fo["frame_weights"], dtype=np.float64)))
    R = Result(nbins=d["nbins"], dbulk=d["dbulk"], cutoff=d["cutoff"], autocorrelation=d["autocorrelation"],
               solute=_sel_from_dict(d["solute"]), solvent=_sel_from_dict(d["solvent"]), files=files,
               weights=list(d["weights"]), Version=d["Version"])
    for k in _VEC:
        setattr(R, k, np.asarray(d[k], dtype=np.float64))
    for k in _MAT:
        setattr(R, k, np.asarray(d[k], dtype=np.float64).reshape(len(d[k]), -1))
    R.density = Density(**d["density"])
    v = d["volume"]
    R.volume = Volume(v["total"], v["bulk"], v["domain"], np.asarray(v["shell"], dtype=np.float64))
    return R
