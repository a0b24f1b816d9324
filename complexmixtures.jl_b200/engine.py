"""ctypes binding of the C ABI (include/cmx_b200.h) -- the Python twin of the Julia ``ccall`` shim.

There is no CPU fallback: if libcmx_b200.so is missing or no CUDA device is present the
constructor raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcmx_b200.so")
_SRC = [os.path.join(_HERE, "csrc", f) for f in
        ("cmx_b200.cu", "cmx_kernels.cuh", "cmx_device.cuh", "cmx_pairs.cuh", "cmx_pairs_host.inl", "cmx_group.inl", "cmx_feed.inl", "cmx_xtc.inl")]
_HDR = os.path.join(os.path.dirname(_HERE), "include", "cmx_b200.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "63"]


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libcmx_b200.so (in-tree)."""
    newest = max(os.path.getmtime(p) for p in _SRC + [_HDR])
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("CMX_NVCC_EXTRA", "").split() + ["-o", LIB_PATH, _SRC[0]]   # (extra -D flags: tuning experiments)
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


class CmxConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("solute_nmols", C.c_int32),
                ("solute_natomspermol", C.c_int32), ("solvent_nmols", C.c_int32), ("solvent_natomspermol", C.c_int32),
                ("autocorrelation", C.c_int32), ("irefatom", C.c_int32), ("usecutoff", C.c_int32),
                ("n_random_samples", C.c_int32), ("coordination_number_only", C.c_int32), ("lcell", C.c_int32),
                ("n_groups_solute", C.c_int32), ("n_groups_solvent", C.c_int32), ("path", C.c_int32),
                ("ring_slots", C.c_int32), ("keep_lists", C.c_int32), ("group_lanes", C.c_int32),
                ("n_streams", C.c_int32), ("batch_frames", C.c_int32),
                ("cutoff", C.c_double), ("dbulk", C.c_double), ("binstep", C.c_double), ("seed", C.c_uint64),
                ("solute_group_offsets", C.c_void_p), ("solute_group_ids", C.c_void_p),
                ("solvent_group_offsets", C.c_void_p), ("solvent_group_ids", C.c_void_p),
                ("n_devices", C.c_int32), ("reserved1", C.c_int32), ("device_ids", C.c_void_p)]


class CmxCounters(C.Structure):
    _fields_ = [("nbins", C.c_int32), ("n_groups_solute", C.c_int32), ("n_groups_solvent", C.c_int32), ("reserved", C.c_int32),
                ("md_count", C.c_void_p), ("md_count_random", C.c_void_p), ("rdf_count", C.c_void_p),
                ("rdf_count_random", C.c_void_p), ("solute_group_count", C.c_void_p), ("solute_group_count_random", C.c_void_p),
                ("solvent_group_count", C.c_void_p), ("solvent_group_count_random", C.c_void_p),
                ("volume_total", C.c_double), ("sum_weights", C.c_double)]


class CmxStats(C.Structure):
    _fields_ = [("frames", C.c_int64), ("kernel_launches", C.c_int64), ("deferred", C.c_int64), ("pair_evals", C.c_int64),
                ("hits_real", C.c_int64), ("hits_random", C.c_int64), ("h2d_bytes", C.c_int64),
                ("gpu_ms_total", C.c_double), ("gpu_ms_main", C.c_double), ("gpu_ms_search_real", C.c_double),
                ("gpu_ms_search_random", C.c_double), ("gpu_ms_reduce", C.c_double),
                ("host_submit_ms", C.c_double), ("host_wait_ms", C.c_double), ("batches", C.c_int64),
                ("volume_total", C.c_double), ("sum_weights", C.c_double)]


FINAL_VECTORS = ("d", "md_count", "md_count_random", "coordination_number", "coordination_number_random", "mddf", "kb",
                 "rdf_count", "rdf_count_random", "sum_rdf_count", "sum_rdf_count_random", "rdf", "kb_rdf", "volume_shell")
FINAL_SCALARS = ("volume_total", "volume_bulk", "volume_domain", "density_solute", "density_solvent", "density_solvent_bulk",
                 "density_fix", "sum_weights")
CONTRIBUTION_TYPES = ("mddf", "coordination_number", "md_count", "kbi")


class CmxFinal(C.Structure):
    """cmx_final (include/cmx_b200.h): the O(nbins) part of Result after finalresults!"""
    _fields_ = ([("nbins", C.c_int32), ("reserved", C.c_int32)] + [(k, C.c_void_p) for k in FINAL_VECTORS]
                + [(k, C.c_double) for k in FINAL_SCALARS])


class CmxXtcInfo(C.Structure):
    _fields_ = [("natoms", C.c_int64), ("nframes", C.c_int64)]


class CmxDcdInfo(C.Structure):
    _fields_ = [("natoms", C.c_int64), ("nframes", C.c_int64), ("first_frame_offset", C.c_int64), ("frame_bytes", C.c_int64)]


MD_DTYPE = np.dtype([("within_cutoff", np.int32), ("i", np.int32), ("j", np.int32),
                     ("ref_atom_within_cutoff", np.int32), ("d", np.float64), ("d_ref_atom", np.float64)])

EXPORTS = ["cmx_version", "cmx_last_error", "cmx_create", "cmx_destroy", "cmx_acquire_frame_buffer", "cmx_submit_frame",
           "cmx_submit_frame_device", "cmx_sync", "cmx_counters_device", "cmx_counters_device_f64", "cmx_finish", "cmx_read_minimum_distances",
           "cmx_read_random_minimum_distances", "cmx_get_stats", "cmx_reset", "cmx_set_option", "cmx_alloc_pinned",
           "cmx_free_pinned", "cmx_dcd_last_error", "cmx_dcd_open", "cmx_dcd_close", "cmx_dcd_read_frame", "cmx_run_dcd",
           "cmx_reduce_groups", "cmx_final_results", "cmx_contributions", "cmx_xtc_open", "cmx_xtc_close", "cmx_xtc_read_frame", "cmx_xtc_read_frame_device", "cmx_run_xtc"]

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library and declare the signatures (no compute call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with cmx_b200.engine.build() (nvcc, sm_100a). "
                           "There is no CPU fallback for the minimum-distance path.")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.cmx_version.restype = C.c_char_p
    lib.cmx_last_error.restype = C.c_char_p; lib.cmx_last_error.argtypes = [vp]
    lib.cmx_create.argtypes = [C.POINTER(CmxConfig), C.POINTER(vp)]
    lib.cmx_destroy.argtypes = [vp]
    lib.cmx_acquire_frame_buffer.argtypes = [vp, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_float))]
    lib.cmx_submit_frame.argtypes = [vp, C.c_int64, C.c_double, C.POINTER(C.c_double)]
    lib.cmx_submit_frame_device.argtypes = [vp, vp, vp, C.c_int64, C.c_double, C.POINTER(C.c_double)]
    lib.cmx_sync.argtypes = [vp]
    lib.cmx_counters_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64)]
    lib.cmx_counters_device_f64.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64)]
    lib.cmx_finish.argtypes = [vp, C.POINTER(CmxCounters)]
    lib.cmx_read_minimum_distances.argtypes = [vp, C.c_int32, vp]
    lib.cmx_read_random_minimum_distances.argtypes = [vp, C.c_int32, vp]
    lib.cmx_get_stats.argtypes = [vp, C.POINTER(CmxStats)]
    lib.cmx_reset.argtypes = [vp]
    lib.cmx_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    lib.cmx_alloc_pinned.argtypes = [C.POINTER(vp), C.c_int64]
    lib.cmx_free_pinned.argtypes = [vp]
    lib.cmx_dcd_last_error.restype = C.c_char_p
    lib.cmx_dcd_open.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(CmxDcdInfo)]
    lib.cmx_dcd_close.argtypes = [vp]
    lib.cmx_dcd_read_frame.argtypes = [vp, C.c_int64, vp, vp, vp, C.POINTER(C.c_double)]
    lib.cmx_run_dcd.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int32]
    lib.cmx_reduce_groups.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp]
    lib.cmx_final_results.argtypes = [vp, C.c_double, C.c_double, C.POINTER(CmxFinal)]
    lib.cmx_contributions.argtypes = [vp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, vp, vp, vp]
    lib.cmx_xtc_open.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(CmxXtcInfo)]
    lib.cmx_xtc_close.argtypes = [vp]
    lib.cmx_run_xtc.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int32]
    lib.cmx_xtc_read_frame.argtypes = [vp, C.c_int64, vp, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    lib.cmx_xtc_read_frame_device.argtypes = [vp, C.c_int64, C.c_int32, vp, C.POINTER(C.c_double)]
    for name in EXPORTS:
        if name not in ("cmx_version", "cmx_last_error", "cmx_dcd_last_error"):
            getattr(lib, name).restype = C.c_int32
    _lib = lib
    return lib


class CmxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcmx_b200 error {code}: {msg}")
        self.code = code


def cell_to_c(cell) -> np.ndarray:
    """3x3 matrix, lattice vectors as COLUMNS (convert_unitcell) -> column-major double[9]."""
    cell = np.asarray(cell, dtype=np.float64)
    if cell.shape == (3,):
        cell = np.diag(cell)
    return np.ascontiguousarray(cell.T).reshape(9).copy()


class DcdFile:
    """Native DCD reader of the library (cmx_dcd_*): pure host code, usable without a GPU.  Mirrors
    NamdDCD (src/trajectory_formats/NamdDCD.jl): frame count from the file size, unit cell record
    [A, gamma, B, beta, alpha, C] -> matrix with the lattice vectors as columns."""

    def __init__(self, filename: str):
        self.lib = load_library()
        self.h = C.c_void_p()
        info = CmxDcdInfo()
        rc = self.lib.cmx_dcd_open(os.fsencode(filename), C.byref(self.h), C.byref(info))
        if rc:
            self.h = None
            raise CmxError(rc, self.lib.cmx_dcd_last_error().decode())
        self.filename = filename
        self.natoms, self.nframes = int(info.natoms), int(info.nframes)
        self.first_frame_offset, self.frame_bytes = int(info.first_frame_offset), int(info.frame_bytes)

    def read_frame(self, iframe: int):
        """(xyz fp32 [natoms,3], cell 3x3 with lattice vectors as columns) of 0-based frame ``iframe``."""
        x = np.empty((3, self.natoms), dtype=np.float32)
        cell = np.zeros(9)
        rc = self.lib.cmx_dcd_read_frame(self.h, int(iframe), x[0].ctypes.data, x[1].ctypes.data, x[2].ctypes.data,
                                         cell.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise CmxError(rc, self.lib.cmx_dcd_last_error().decode())
        return np.ascontiguousarray(x.T), cell.reshape(3, 3).T.copy()

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmx_dcd_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class XtcFile:
    """Native GROMACS XTC reader of the library (cmx_xtc_*): pure host code, usable without a GPU.  Frames are
    indexed at open time; ``read_frame`` decodes the compressed coordinate block to fp32 Angstrom."""

    def __init__(self, filename: str):
        self.lib = load_library()
        self.h = C.c_void_p()
        info = CmxXtcInfo()
        rc = self.lib.cmx_xtc_open(os.fsencode(filename), C.byref(self.h), C.byref(info))
        if rc:
            self.h = None
            raise CmxError(rc, self.lib.cmx_dcd_last_error().decode())
        self.filename, self.natoms, self.nframes = filename, int(info.natoms), int(info.nframes)

    def read_frame(self, iframe: int, out: Optional[np.ndarray] = None):
        """(xyz fp32 [natoms,3] in Angstrom, cell 3x3 with the lattice vectors as columns, step, time)."""
        xyz = np.empty((self.natoms, 3), dtype=np.float32) if out is None else out
        cell, step, time = np.zeros(9), C.c_int32(), C.c_float()
        rc = self.lib.cmx_xtc_read_frame(self.h, int(iframe), xyz.ctypes.data, cell.ctypes.data_as(C.POINTER(C.c_double)),
                                         C.byref(step), C.byref(time))
        if rc:
            raise CmxError(rc, self.lib.cmx_dcd_last_error().decode())
        return xyz, cell.reshape(3, 3).T.copy(), int(step.value), float(time.value)

    def read_frame_device(self, iframe: int, device: int = 0):
        """the same frame decoded on the GPU (the decoder of the native feed): (xyz fp32 [natoms,3] in Angstrom, cell)"""
        xyz = np.empty((self.natoms, 3), dtype=np.float32)
        cell = np.zeros(9)
        rc = self.lib.cmx_xtc_read_frame_device(self.h, int(iframe), int(device), xyz.ctypes.data, cell.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise CmxError(rc, self.lib.cmx_dcd_last_error().decode())
        return xyz, cell.reshape(3, 3).T.copy()

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmx_xtc_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One ``mddf`` problem on one GPU (= one chunk task of the reference, src/mddf.jl:288-337)."""

    def __init__(self, *, solute, solvent, options, irefatom: int, autocorrelation: bool,
                 coordination_number_only: bool = False, device: int = 0, path: int = 0, keep_lists: bool = False,
                 ring_slots: int = 0, group_lanes: int = 0, n_streams: int = 0, batch_frames: int = 0, devices=None):
        self.lib = load_library()
        cfg = CmxConfig()
        cfg.struct_size = C.sizeof(CmxConfig)
        cfg.device = device
        cfg.solute_nmols, cfg.solute_natomspermol = solute.nmols, solute.natomspermol
        cfg.solvent_nmols, cfg.solvent_natomspermol = solvent.nmols, solvent.natomspermol
        cfg.autocorrelation = int(autocorrelation)
        cfg.irefatom = irefatom
        cfg.usecutoff = int(options.usecutoff)
        cfg.n_random_samples = options.n_random_samples
        cfg.coordination_number_only = int(coordination_number_only)
        cfg.lcell = options.lcell
        cfg.n_groups_solute, cfg.n_groups_solvent = solute.n_groups, solvent.n_groups
        cfg.path, cfg.ring_slots, cfg.keep_lists, cfg.group_lanes = path, ring_slots, int(keep_lists), group_lanes
        cfg.n_streams = n_streams
        cfg.batch_frames = batch_frames
        self._keep = []
        if devices is not None and len(devices) > 1:      # several GPUs behind this one handle (frames dealt k mod n)
            ids = np.ascontiguousarray(devices, dtype=np.int32)
            self._keep.append(ids)
            cfg.n_devices, cfg.device_ids = len(ids), ids.ctypes.data
        cfg.cutoff, cfg.dbulk, cfg.binstep = options.cutoff, options.dbulk, options.binstep
        cfg.seed = options.seed if options.seed > 0 else 0
        for side, sel in (("solute", solute), ("solvent", solvent)):
            off, ids = sel.group_csr()
            if off is not None:
                off = np.ascontiguousarray(off, dtype=np.int32); ids = np.ascontiguousarray(ids, dtype=np.int32)
                self._keep += [off, ids]
                setattr(cfg, f"{side}_group_offsets", off.ctypes.data); setattr(cfg, f"{side}_group_ids", ids.ctypes.data)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.lib.cmx_create(C.byref(cfg), C.byref(self.h))
        if rc:
            raise CmxError(rc, self.lib.cmx_last_error(None).decode())
        self.nbins = max(1, int(np.ceil(options.cutoff / options.binstep)))
        self.n_solute_atoms = solute.nmols * solute.natomspermol
        self.n_solvent_atoms = solvent.nmols * solvent.natomspermol
        self.autocorrelation = bool(autocorrelation)
        self.nmols_solvent = solvent.nmols

    def _ck(self, rc):
        if rc:
            raise CmxError(rc, self.lib.cmx_last_error(self.h).decode())

    def close(self):
        for blk in getattr(self, "_pinned_blocks", []):
            self.lib.cmx_free_pinned(blk)
        self._pinned_blocks, self._out, self._out_head = [], None, None
        if getattr(self, "h", None) is not None and self.h:
            self.lib.cmx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- frame feed ---------------------------------------------------------------------------
    def acquire(self):
        """numpy views (fp32 [n,3]) over the next pinned staging slot."""
        ps, pv = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        self._ck(self.lib.cmx_acquire_frame_buffer(self.h, C.byref(ps), C.byref(pv)))
        xv = np.ctypeslib.as_array(pv, shape=(self.n_solvent_atoms, 3))
        xs = xv if self.autocorrelation else np.ctypeslib.as_array(ps, shape=(self.n_solute_atoms, 3))
        return xs, xv

    def submit(self, frame_index: int, weight: float, cell):
        c = cell_to_c(cell)
        self._ck(self.lib.cmx_submit_frame(self.h, int(frame_index), float(weight), c.ctypes.data_as(C.POINTER(C.c_double))))

    def submit_arrays(self, xsolute, xsolvent, cell, frame_index: int = 0, weight: float = 1.0):
        xs, xv = self.acquire()
        xv[...] = xsolvent
        if not self.autocorrelation:
            xs[...] = xsolute
        self.submit(frame_index, weight, cell)

    def submit_device(self, d_solute_ptr: int, d_solvent_ptr: int, cell, frame_index: int = 0, weight: float = 1.0):
        c = cell_to_c(cell)
        self._ck(self.lib.cmx_submit_frame_device(self.h, C.c_void_p(d_solute_ptr or 0), C.c_void_p(d_solvent_ptr), int(frame_index),
                                                  float(weight), c.ctypes.data_as(C.POINTER(C.c_double))))

    def run_dcd(self, dcd: "DcdFile", solute_indices, solvent_indices, frames, weights=None, n_reader_threads: int = 0):
        """cmx_run_dcd: the whole frame loop for a DCD file inside the library (reader threads ->
        pinned ring -> one H2D per raw frame -> device gather of the selections -> frame pipeline).
        ``frames``: 0-based frame numbers in the file; the Philox frame key is frame + 1."""
        fr = np.ascontiguousarray(frames, dtype=np.int64)
        sv = np.ascontiguousarray(solvent_indices, dtype=np.int32)
        ss = None if self.autocorrelation else np.ascontiguousarray(solute_indices, dtype=np.int32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._ck(self.lib.cmx_run_dcd(self.h, dcd.h, None if ss is None else ss.ctypes.data, sv.ctypes.data, fr.ctypes.data,
                                      None if w is None else w.ctypes.data, int(fr.size), int(n_reader_threads)))

    def run_xtc(self, xtc: "XtcFile", solute_indices, solvent_indices, frames, weights=None, n_reader_threads: int = 0):
        """cmx_run_xtc: as ``run_dcd`` for a GROMACS XTC file (the reader threads also decode)."""
        fr = np.ascontiguousarray(frames, dtype=np.int64)
        sv = np.ascontiguousarray(solvent_indices, dtype=np.int32)
        ss = None if self.autocorrelation else np.ascontiguousarray(solute_indices, dtype=np.int32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._ck(self.lib.cmx_run_xtc(self.h, xtc.h, None if ss is None else ss.ctypes.data, sv.ctypes.data, fr.ctypes.data,
                                      None if w is None else w.ctypes.data, int(fr.size), int(n_reader_threads)))

    def reduce_groups(self, which: str, groups) -> np.ndarray:
        """cmx_reduce_groups: per-group sums of the rows of one group-count array, on the device.
        ``which``: solute_group_count | solute_group_count_random | solvent_group_count |
        solvent_group_count_random; ``groups``: sequence of sequences of 0-based row ids.
        Returns f64 [n_groups, nbins] (what summing those rows of ``finish()`` would give)."""
        code = ["solute_group_count", "solute_group_count_random", "solvent_group_count", "solvent_group_count_random"].index(which)
        off = np.zeros(len(groups) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(g) for g in groups])
        rows = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.int32).reshape(-1) for g in groups])
                                    if len(groups) and off[-1] else np.zeros(0, dtype=np.int32), dtype=np.int32)
        out = np.zeros((len(groups), self.nbins))
        self._ck(self.lib.cmx_reduce_groups(self.h, code, len(groups), off.ctypes.data, rows.ctypes.data if rows.size else None,
                                            out.ctypes.data))
        return out

    @staticmethod
    def _groups_csr(groups):
        off = np.zeros(len(groups) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(g) for g in groups])
        rows = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.int32).reshape(-1) for g in groups])
                                    if len(groups) and off[-1] else np.zeros(0, dtype=np.int32), dtype=np.int32)
        return off, rows

    def final_results(self, sum_weights: float = 0.0, volume_sum: float = 0.0) -> dict:
        """cmx_final_results: finalresults! (src/results.jl:311-469) on the device -- the O(nbins) vectors and the
        Volume / Density scalars of Result.  ``sum_weights`` / ``volume_sum`` <= 0: the handle's own sums."""
        f = CmxFinal()
        out = {k: np.zeros(self.nbins) for k in FINAL_VECTORS}
        for k, v in out.items():
            setattr(f, k, v.ctypes.data)
        self._ck(self.lib.cmx_final_results(self.h, float(sum_weights), float(volume_sum), C.byref(f)))
        out.update({k: getattr(f, k) for k in FINAL_SCALARS})
        return out

    def contributions(self, side: str, groups, type: str = "mddf", sum_weights: float = 0.0, volume_sum: float = 0.0) -> np.ndarray:
        """cmx_contributions: contributions(R, SoluteGroup|SolventGroup; type) (src/tools/contributions.jl:70-248) for
        many groups at once, on the device.  ``side``: "solute" | "solvent"; ``groups``: sequence of sequences of 0-based
        rows of that side's group-count array; ``type``: mddf | coordination_number | md_count | kbi.
        Returns f64 [n_groups, nbins] (the matrix of ResidueContributions when the groups are residues)."""
        off, rows = self._groups_csr(groups)
        out = np.zeros((len(groups), self.nbins))
        self._ck(self.lib.cmx_contributions(self.h, ("solute", "solvent").index(side), CONTRIBUTION_TYPES.index(type), float(sum_weights),
                                            float(volume_sum), len(groups), off.ctypes.data, rows.ctypes.data if rows.size else None,
                                            out.ctypes.data))
        return out

    def sync(self):
        self._ck(self.lib.cmx_sync(self.h))

    def reset(self):
        self._ck(self.lib.cmx_reset(self.h))

    def set_option(self, name: str, value: float):
        self._ck(self.lib.cmx_set_option(self.h, name.encode(), float(value)))

    # ---- results --------------------------------------------------------------------------------
    def counters_device(self):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.cmx_counters_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def counters_device_f64(self):
        """device pointer / length of the f64 counters with the frame weights applied (all-reduce payload when the
        weights of the ranks are not one and the same number); the next finish() writes the array out as is"""
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.cmx_counters_device_f64(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def _result_arrays(self, groups: bool = True):
        """Result arrays in page-locked memory, allocated once per engine (the group arrays only when asked for)."""
        key = "_out" if groups else "_out_head"
        if getattr(self, key, None) is None:
            nb, gs, gv = self.nbins, self.cfg.n_groups_solute, self.cfg.n_groups_solvent
            shapes = dict(md_count=(nb,), md_count_random=(nb,), rdf_count=(nb,), rdf_count_random=(nb,))
            if groups:
                shapes.update(solute_group_count=(gs, nb), solute_group_count_random=(gs, nb),
                              solvent_group_count=(gv, nb), solvent_group_count_random=(gv, nb))
            total = sum(int(np.prod(sh)) for sh in shapes.values())
            p = C.c_void_p()
            if self.lib.cmx_alloc_pinned(C.byref(p), total * 8):
                raise CmxError(2, "cudaHostAlloc failed")
            self._pinned_blocks = getattr(self, "_pinned_blocks", []) + [p]
            flat = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(total,))
            out, off = {}, 0
            for k, sh in shapes.items():
                n = int(np.prod(sh))
                out[k] = flat[off:off + n].reshape(sh)
                off += n
            setattr(self, key, out)
        return getattr(self, key)

    def finish(self, copy: bool = True, groups: bool = True) -> dict:
        """cmx_finish: sync, (all-reduced) counters -> f64 arrays laid out like Result.  With
        copy=False the returned arrays are views of the engine's pinned buffers (valid until the
        next finish/close).  groups=False leaves the group arrays on the device (NULL members of cmx_counters):
        only md_count / rdf_count (+ random) come back -- for per-atom arrays of GBs, followed by
        ``contributions`` / ``reduce_groups`` / ``final_results`` on the device."""
        out = self._result_arrays(groups)
        c = CmxCounters()
        for k, v in out.items():
            setattr(c, k, v.ctypes.data)
        self._ck(self.lib.cmx_finish(self.h, C.byref(c)))
        res = {k: (v.copy() if copy else v) for k, v in out.items()}
        res["volume_total"] = c.volume_total
        res["sum_weights"] = c.sum_weights
        return res

    def minimum_distances(self, isolute: int = 0) -> np.ndarray:
        out = np.zeros(self.nmols_solvent, dtype=MD_DTYPE)
        self._ck(self.lib.cmx_read_minimum_distances(self.h, isolute, out.ctypes.data_as(C.c_void_p)))
        return out

    def random_minimum_distances(self, sample: int) -> np.ndarray:
        out = np.zeros(self.nmols_solvent, dtype=MD_DTYPE)
        self._ck(self.lib.cmx_read_random_minimum_distances(self.h, sample, out.ctypes.data_as(C.c_void_p)))
        return out

    def stats(self) -> dict:
        s = CmxStats()
        self._ck(self.lib.cmx_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in CmxStats._fields_}
