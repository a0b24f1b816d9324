"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see the header of cmx_oracle.c for the parity status).  It restates, on
top of libcmx_oracle.so, the driver of the reference: ``mddf(trajectory, options)``
(src/mddf.jl:227-347) and ``finalresults!`` (src/results.jl:311-469) with plain numpy.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcmx_oracle.so")
    src = os.path.join(_HERE, "cmx_oracle.c")
    src2 = os.path.join(_HERE, "xtc_two_phase.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(src2)):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return so


class OrcConfig(C.Structure):
    _fields_ = [("nmols_solute", C.c_int32), ("napm_solute", C.c_int32), ("nmols_solvent", C.c_int32),
                ("napm_solvent", C.c_int32), ("autocorrelation", C.c_int32), ("irefatom", C.c_int32),
                ("usecutoff", C.c_int32), ("nbins", C.c_int32), ("n_random_samples", C.c_int32),
                ("coordination_number_only", C.c_int32), ("ngroups_solute", C.c_int32),
                ("ngroups_solvent", C.c_int32), ("custom_solute", C.c_int32), ("custom_solvent", C.c_int32),
                ("cutoff", C.c_double), ("dbulk", C.c_double), ("binstep", C.c_double), ("seed", C.c_uint64),
                ("solute_grp_off", C.c_void_p), ("solute_grp_ids", C.c_void_p),
                ("solvent_grp_off", C.c_void_p), ("solvent_grp_ids", C.c_void_p)]


class OrcCounters(C.Structure):
    _fields_ = [("md_count", C.c_void_p), ("md_count_random", C.c_void_p), ("rdf_count", C.c_void_p),
                ("rdf_count_random", C.c_void_p), ("solute_group", C.c_void_p), ("solute_group_random", C.c_void_p),
                ("solvent_group", C.c_void_p), ("solvent_group_random", C.c_void_p),
                ("volume_total", C.c_double), ("pair_evals", C.c_int64)]


MD_DTYPE = np.dtype([("within_cutoff", np.int32), ("i", np.int32), ("j", np.int32),
                     ("ref_atom_within_cutoff", np.int32), ("d", np.float64), ("d_ref_atom", np.float64)])


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_dist_pbc.restype = C.c_double
        _LIB.orc_minimum_distances.restype = C.c_int64
        _LIB.orc_setbin.restype = C.c_int
        _LIB.orc_ref_solute.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def setbin(d, step):
    return lib().orc_setbin(C.c_double(d), C.c_double(step))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32(c, k, o)
    return list(o)


def eulermat(beta, gamma, theta):
    A = np.zeros(9)
    lib().orc_eulermat(C.c_double(beta), C.c_double(gamma), C.c_double(theta), _p(A))
    return A.reshape(3, 3)


def move(x, newcm, beta, gamma, theta):
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    cm = np.ascontiguousarray(newcm, dtype=np.float64)
    lib().orc_move(_p(x), C.c_int(len(x)), _p(cm), C.c_double(beta), C.c_double(gamma), C.c_double(theta))
    return x


def random_move(cell, x, iref, seed, slot=0, sample=0, frame=0):
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    cell9 = cell_to_c(cell)
    lib().orc_random_move(_p(cell9), _p(x), C.c_int(len(x)), C.c_int(iref), C.c_uint64(seed), C.c_uint32(slot),
                          C.c_uint32(sample), C.c_uint32(frame))
    return x


def update_md(a, b):
    aa = np.array([a], dtype=MD_DTYPE); bb = np.array([b], dtype=MD_DTYPE); out = np.zeros(1, dtype=MD_DTYPE)
    lib().orc_update_md(_p(aa), _p(bb), _p(out))
    return tuple(out[0])


def cell_to_c(cell) -> np.ndarray:
    """3x3 matrix with lattice vectors as COLUMNS -> column-major double[9]."""
    cell = np.asarray(cell, dtype=np.float64)
    if cell.shape == (3,):
        cell = np.diag(cell)
    return np.ascontiguousarray(cell.T).reshape(9).copy()


def dist_pbc(cell, xi, xj):
    return lib().orc_dist_pbc(_p(cell_to_c(cell)), _p(np.asarray(xi, dtype=np.float64)), _p(np.asarray(xj, dtype=np.float64)))


class Oracle:
    """One mddf problem: selections + options flattened exactly like the C-ABI config."""

    def __init__(self, *, nmols_solute, napm_solute, nmols_solvent, napm_solvent, autocorrelation, irefatom,
                 cutoff, dbulk, usecutoff, binstep, n_random_samples, seed, coordination_number_only=False,
                 solute_groups=None, solvent_groups=None, n_groups_solute=None, n_groups_solvent=None):
        """irefatom is 1-based (as in the reference).  *_groups = (offsets, ids) CSR or None."""
        self.cfg = cfg = OrcConfig()
        cfg.nmols_solute, cfg.napm_solute = nmols_solute, napm_solute
        cfg.nmols_solvent, cfg.napm_solvent = nmols_solvent, napm_solvent
        cfg.autocorrelation = int(autocorrelation)
        cfg.irefatom = irefatom - 1
        cfg.usecutoff = int(usecutoff)
        cfg.cutoff, cfg.dbulk, cfg.binstep = cutoff, dbulk, binstep
        cfg.nbins = max(1, math.ceil(cutoff / binstep))
        cfg.n_random_samples = n_random_samples
        cfg.coordination_number_only = int(coordination_number_only)
        cfg.seed = seed
        self._keep = []
        for side, grp, ng in (("solute", solute_groups, n_groups_solute), ("solvent", solvent_groups, n_groups_solvent)):
            napm = napm_solute if side == "solute" else napm_solvent
            if grp is not None and grp[0] is not None:
                off = np.ascontiguousarray(grp[0], dtype=np.int32); ids = np.ascontiguousarray(grp[1], dtype=np.int32)
                self._keep += [off, ids]
                setattr(cfg, f"custom_{side}", 1)
                setattr(cfg, f"{side}_grp_off", off.ctypes.data); setattr(cfg, f"{side}_grp_ids", ids.ctypes.data)
                setattr(cfg, f"ngroups_{side}", int(ng))
            else:
                setattr(cfg, f"custom_{side}", 0)
                setattr(cfg, f"ngroups_{side}", napm)
        self.nbins = cfg.nbins
        self.reset()

    @classmethod
    def from_problem(cls, solute, solvent, options, irefatom, autocorrelation, coordination_number_only=False):
        return cls(nmols_solute=solute.nmols, napm_solute=solute.natomspermol, nmols_solvent=solvent.nmols,
                   napm_solvent=solvent.natomspermol, autocorrelation=autocorrelation, irefatom=irefatom,
                   cutoff=options.cutoff, dbulk=options.dbulk, usecutoff=options.usecutoff, binstep=options.binstep,
                   n_random_samples=options.n_random_samples, seed=options.seed,
                   coordination_number_only=coordination_number_only,
                   solute_groups=solute.group_csr(), solvent_groups=solvent.group_csr(),
                   n_groups_solute=solute.n_groups, n_groups_solvent=solvent.n_groups)

    def reset(self):
        cfg, nb = self.cfg, self.cfg.nbins
        self.md_count = np.zeros(nb); self.md_count_random = np.zeros(nb)
        self.rdf_count = np.zeros(nb); self.rdf_count_random = np.zeros(nb)
        self.solute_group_count = np.zeros((cfg.ngroups_solute, nb)); self.solute_group_count_random = np.zeros((cfg.ngroups_solute, nb))
        self.solvent_group_count = np.zeros((cfg.ngroups_solvent, nb)); self.solvent_group_count_random = np.zeros((cfg.ngroups_solvent, nb))
        self.ctr = ctr = OrcCounters()
        ctr.md_count, ctr.md_count_random = self.md_count.ctypes.data, self.md_count_random.ctypes.data
        ctr.rdf_count, ctr.rdf_count_random = self.rdf_count.ctypes.data, self.rdf_count_random.ctypes.data
        ctr.solute_group, ctr.solute_group_random = self.solute_group_count.ctypes.data, self.solute_group_count_random.ctypes.data
        ctr.solvent_group, ctr.solvent_group_random = self.solvent_group_count.ctypes.data, self.solvent_group_count_random.ctypes.data
        ctr.volume_total, ctr.pair_evals = 0.0, 0

    @property
    def volume_total(self): return self.ctr.volume_total
    @property
    def pair_evals(self): return self.ctr.pair_evals

    def frame(self, xsolute, xsolvent, cell, weight=1.0, frame_index=0, use_clist=True, want_lists=False):
        """mddf_frame!/coordination_number_frame! on one frame (coordinates fp32-valued)."""
        cfg = self.cfg
        xs = np.ascontiguousarray(xsolute, dtype=np.float64)
        xv = xs if cfg.autocorrelation else np.ascontiguousarray(xsolvent, dtype=np.float64)
        lists = rlists = None
        if want_lists:
            lists = np.zeros((cfg.nmols_solute, cfg.nmols_solvent), dtype=MD_DTYPE)
            rlists = np.zeros((max(cfg.n_random_samples, 1), cfg.nmols_solvent), dtype=MD_DTYPE)
        lib().orc_frame(C.byref(cfg), _p(xs), _p(xv), _p(cell_to_c(cell)), C.c_double(weight), C.c_uint32(frame_index),
                        C.c_int(int(use_clist)), C.byref(self.ctr), _p(lists) if want_lists else None,
                        _p(rlists) if want_lists else None)
        return lists, rlists

    def minimum_distances(self, x_molecule, xsolvent, cell, isolute=0, use_clist=False):
        cfg = self.cfg
        x = np.ascontiguousarray(x_molecule, dtype=np.float64); y = np.ascontiguousarray(xsolvent, dtype=np.float64)
        out = np.zeros(cfg.nmols_solvent, dtype=MD_DTYPE)
        lib().orc_minimum_distances(C.byref(cfg), _p(cell_to_c(cell)), _p(x), _p(y), C.c_int(isolute), C.c_int(int(use_clist)), _p(out))
        return out

    def ref_solute(self, frame_index, sample):
        return lib().orc_ref_solute(C.byref(self.cfg), C.c_uint32(frame_index), C.c_uint32(sample))

    def run_frames(self, xs, xv, cells, weights=None, frame_ids=None, use_clist=True, nthreads=1):
        """Frame-parallel chunk loop (src/mddf.jl:285-338) over fp32 arrays [nframes, n, 3]."""
        cfg = self.cfg
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        xv = xs if cfg.autocorrelation else np.ascontiguousarray(xv, dtype=np.float32)
        nf = xs.shape[0]
        cells = np.ascontiguousarray(np.stack([cell_to_c(c) for c in cells]) if np.ndim(cells) == 3 or isinstance(cells, list)
                                     else np.tile(cell_to_c(cells), (nf, 1)), dtype=np.float64)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        fid = None if frame_ids is None else np.ascontiguousarray(frame_ids, dtype=np.uint32)
        lib().orc_run_frames(C.byref(cfg), C.c_int(nf), _p(xs), _p(xv), _p(cells), _p(w) if w is not None else None,
                             _p(fid) if fid is not None else None, C.c_int(int(use_clist)), C.c_int(nthreads), C.byref(self.ctr))

    def counters(self):
        return dict(md_count=self.md_count.copy(), md_count_random=self.md_count_random.copy(),
                    rdf_count=self.rdf_count.copy(), rdf_count_random=self.rdf_count_random.copy(),
                    solute_group_count=self.solute_group_count.copy(), solute_group_count_random=self.solute_group_count_random.copy(),
                    solvent_group_count=self.solvent_group_count.copy(), solvent_group_count_random=self.solvent_group_count_random.copy(),
                    volume_total=self.volume_total)


# ---------------------------------------------------------------------------------------
# finalresults! restated on plain arrays (src/results.jl:320-469)
# ---------------------------------------------------------------------------------------
ANGS3_TO_CM3_PER_MOL = 6.022140857e23 / 1e24  # src/io.jl:6-10


def shellradius(i, step):
    rmin = (i - 1) * step
    return (0.5 * ((rmin + step) ** 3 + rmin ** 3)) ** (1.0 / 3.0)


def sphericalshellvolume(i, step):
    rmin = (i - 1) * step
    return (4 * math.pi / 3) * ((rmin + step) ** 3 - rmin ** 3)


def finalresults(c: dict, *, nmols_solute, nmols_solvent, autocorrelation, n_random_samples, binstep, dbulk, cutoff,
                 usecutoff, Q, coordination_number_only=False):
    """c = Oracle.counters() (raw weighted counts).  Returns a namespace with the Result fields."""
    nb = len(c["md_count"])
    r = SimpleNamespace(nbins=nb)
    r.d = np.array([shellradius(i, binstep) for i in range(1, nb + 1)])
    ns = nmols_solute * Q
    r.md_count = c["md_count"] / ns
    r.rdf_count = c["rdf_count"] / ns
    r.solute_group_count = c["solute_group_count"] / ns
    r.solvent_group_count = r.solute_group_count.copy() if autocorrelation else c["solvent_group_count"] / ns
    r.volume_total = c["volume_total"] / Q
    r.density_solvent = nmols_solvent / r.volume_total
    r.density_solute = nmols_solute / r.volume_total
    r.coordination_number = np.cumsum(r.md_count)
    r.sum_rdf_count = np.cumsum(r.rdf_count)
    if coordination_number_only:
        return r
    nr = n_random_samples * Q
    nsolv = nmols_solvent - 1 if autocorrelation else nmols_solvent
    r.md_count_random = c["md_count_random"] / nr
    r.rdf_count_random = c["rdf_count_random"] / nr
    r.solute_group_count_random = c["solute_group_count_random"] / nr
    r.solvent_group_count_random = r.solute_group_count_random.copy() if autocorrelation else c["solvent_group_count_random"] / nr
    r.volume_shell = r.volume_total * (r.rdf_count_random / nsolv)
    ibulk = max(1, math.ceil((dbulk + 0.5 * binstep) / binstep))
    r.volume_domain = float(np.sum(r.volume_shell[: ibulk - 1]))
    if not usecutoff:
        r.volume_bulk = r.volume_total - r.volume_domain
        n_bulk = nsolv - float(np.sum(r.rdf_count))
    else:
        n_bulk = float(np.sum(r.rdf_count[ibulk - 1:]))
        r.volume_bulk = float(np.sum(r.volume_shell[ibulk - 1:]))
    with np.errstate(divide="ignore", invalid="ignore"):
        r.density_solvent_bulk = np.float64(n_bulk) / np.float64(r.volume_bulk)   # IEEE semantics as in Julia (Inf/NaN, no exception)
    fix = r.density_solvent_bulk / r.density_solvent
    for k in ("md_count_random", "rdf_count_random", "solute_group_count_random", "solvent_group_count_random"):
        setattr(r, k, getattr(r, k) * fix)
    r.coordination_number_random = np.cumsum(r.md_count_random)
    r.mddf = np.where(r.md_count_random > 0, r.md_count / np.where(r.md_count_random > 0, r.md_count_random, 1.0), 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        r.kb = ANGS3_TO_CM3_PER_MOL * (1 / r.density_solvent_bulk) * (r.coordination_number - r.coordination_number_random)
        r.rdf = np.where(r.rdf_count_random > 0, r.rdf_count / np.where(r.rdf_count_random > 0, r.rdf_count_random, 1.0), 0.0)
        r.sum_rdf_count_random = np.cumsum(r.rdf_count_random)
        r.kb_rdf = ANGS3_TO_CM3_PER_MOL * (1 / r.density_solvent_bulk) * (r.sum_rdf_count - r.sum_rdf_count_random)
    return r


# ---- prototype of the planned device-side XTC decoder (oracle/xtc_two_phase.c) -------------------------------------
def xtc_two_phase_decode(block: bytes, natoms: int):
    """(xyz fp32 [natoms,3] in Angstrom, group bytes) of one compressed XTC coordinate block via the two-phase scheme:
    serial skeleton walk (1 byte per group) + independent per-group decode."""
    L = lib()
    buf = np.frombuffer(block, dtype=np.uint8)
    codes = np.zeros(natoms, dtype=np.uint8)
    L.xtc_skeleton.restype = C.c_int
    ng = L.xtc_skeleton(_p(buf), C.c_size_t(buf.size), C.c_int(natoms), _p(codes), C.c_int(codes.size))
    if ng <= 0:
        raise ValueError("malformed XTC coordinate block")
    xyz = np.zeros((natoms, 3), dtype=np.float32)
    L.xtc_decode_from_skeleton.restype = C.c_int
    ok = L.xtc_decode_from_skeleton(_p(buf), C.c_size_t(buf.size), C.c_int(natoms), _p(codes), C.c_int(ng), _p(xyz))
    if not ok:
        raise ValueError("skeleton does not match the block")
    return xyz, codes[:ng].copy()
