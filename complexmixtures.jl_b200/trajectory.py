"""Trajectory readers (host side; trajectory I/O stays on the host, BASELINE.json north_star).

Reference: src/Trajectory.jl:31-59 (format dispatch), :72-81 (convert_unitcell: unit cell as
3x3 matrix whose COLUMNS are the lattice vectors), :189-249 (TrajectoryMetaData, default
irefatom); src/trajectory_formats/NamdDCD.jl:141-188 (record layout, cell = [A,gamma,B,beta,
alpha,C]); src/trajectory_formats/PDBTraj.jl:120-154.

Difference to the reference that matters for the device feed: ``nextframe(dst_solute,
dst_solvent)`` gathers the selected atoms as **fp32** straight into caller-provided buffers
(the pinned staging slots of the engine) -- DCD/XTC coordinates are fp32 on disk
(NamdDCD.jl:41-43), so this is lossless.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .selection import AtomSelection


def cell_from_lengths_angles(A, B, C, alpha, beta, gamma) -> np.ndarray:
    """Lattice vectors as matrix columns, a along x, b in the xy plane (Chemfiles convention
    used by getunitcell, NamdDCD.jl:175-188).  Angles in degrees; any zero angle -> all 90."""
    if alpha == 0.0 or beta == 0.0 or gamma == 0.0:
        alpha = beta = gamma = 90.0
    m = np.zeros((3, 3))
    if alpha == 90.0 and beta == 90.0 and gamma == 90.0:
        m[0, 0], m[1, 1], m[2, 2] = A, B, C
        return m
    ca, cb, cg = (np.cos(np.deg2rad(v)) for v in (alpha, beta, gamma))
    sg = np.sin(np.deg2rad(gamma))
    m[:, 0] = (A, 0.0, 0.0)
    m[:, 1] = (B * cg, B * sg, 0.0)
    cx = C * cb
    cy = C * (ca - cb * cg) / sg
    m[:, 2] = (cx, cy, np.sqrt(max(C * C - cx * cx - cy * cy, 0.0)))
    return m


def is_orthorhombic(cell: np.ndarray, tol: float = 1e-10) -> bool:
    """convert_unitcell, src/Trajectory.jl:72-77."""
    s = abs(min(np.diag(cell)))
    off = cell - np.diag(np.diag(cell))
    return bool(np.all(np.abs(off) < tol * s))


class Trajectory:
    """Abstract trajectory: ``nframes``, ``solute``, ``solvent``, ``open/close/firstframe/
    nextframe/getunitcell`` (opentraj!/closetraj!/firstframe!/nextframe!/getunitcell)."""
    filename: str = ""
    nframes: int = 0
    solute: AtomSelection
    solvent: AtomSelection

    def open(self): ...
    def close(self): ...
    def firstframe(self): ...
    def nextframe(self, dst_solute: Optional[np.ndarray] = None, dst_solvent: Optional[np.ndarray] = None): ...
    def getunitcell(self) -> np.ndarray: ...

    @property
    def autocorrelation(self) -> bool:
        """isautocorrelation, src/results.jl:240-243."""
        a, b = self.solute.indices, self.solvent.indices
        return len(a) == len(b) and bool(np.array_equal(a, b))

    def _alloc(self):
        self.x_solute = np.zeros((self.solute.natoms, 3), dtype=np.float32)
        self.x_solvent = np.zeros((self.solvent.natoms, 3), dtype=np.float32)
        self._isol = self.solute.indices - 1
        self._isolv = self.solvent.indices - 1

    def _gather(self, xyz: np.ndarray, dst_solute, dst_solvent):
        ds = self.x_solute if dst_solute is None else dst_solute
        dv = self.x_solvent if dst_solvent is None else dst_solvent
        np.take(xyz, self._isol, axis=0, out=ds)
        np.take(xyz, self._isolv, axis=0, out=dv)
        return ds, dv


class NamdDCD(Trajectory):
    """NAMD/CHARMM DCD with unit-cell records (src/trajectory_formats/NamdDCD.jl)."""

    def __init__(self, filename: str, solute: AtomSelection, solvent: AtomSelection, lastframe: int = -1):
        self.filename, self.solute, self.solvent = filename, solute, solvent
        with open(filename, "rb") as f:
            def rec():
                n = struct.unpack("<i", f.read(4))[0]
                data = f.read(n)
                f.read(4)
                return data
            hdr = rec()
            if hdr[:4] != b"CORD":
                raise ValueError("not a little-endian DCD file")
            rec()                                  # title
            self.natoms_file = struct.unpack("<i", rec())[0]
            self._first = f.tell()
            first = rec()
            if len(first) != 48:                   # NamdDCD.jl:69-77
                raise ValueError("DCD file does not contain unit cell information.")
            f.seek(0, 2)
            size = f.tell()
        self._framebytes = (4 + 48 + 4) + 3 * (4 + 4 * self.natoms_file + 4)
        # "Sometimes the DCD files contains a wrong number of frames in the header" -> count
        # from the file size (NamdDCD.jl:211-230)
        self.nframes = (size - self._first) // self._framebytes
        if lastframe > 0:
            self.nframes = min(self.nframes, lastframe)
        self.lastatom = int(max(solute.indices.max(), solvent.indices.max()))
        self._alloc()
        self._f = None
        self._xyz = np.empty((self.natoms_file, 3), dtype=np.float32)
        self.unitcell_read = np.zeros(6)

    def open(self):
        self._f = open(self.filename, "rb")
        self.firstframe()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def firstframe(self):
        self._f.seek(self._first)

    def nextframe(self, dst_solute=None, dst_solvent=None):
        buf = self._f.read(self._framebytes)
        if len(buf) < self._framebytes:
            raise EOFError("end of DCD trajectory")
        self.unitcell_read[:] = np.frombuffer(buf, dtype="<f8", count=6, offset=4)
        n, off = self.natoms_file, 56
        for k in range(3):
            self._xyz[:, k] = np.frombuffer(buf, dtype="<f4", count=n, offset=off + 4)
            off += 8 + 4 * n
        return self._gather(self._xyz, dst_solute, dst_solvent)

    def getunitcell(self):
        A, g, B, b, a, C = self.unitcell_read
        return cell_from_lengths_angles(A, B, C, a, b, g)


class XTCTraj(Trajectory):
    """GROMACS XTC through the library's native decoder (``cmx_xtc_*``); the reference reads this format with
    Chemfiles (src/trajectory_formats/ChemFiles.jl): positions in Angstrom, unit cell with the lattice vectors as
    matrix columns (convert_unitcell)."""

    def __init__(self, filename: str, solute: AtomSelection, solvent: AtomSelection, lastframe: int = -1):
        from .engine import XtcFile
        self.filename, self.solute, self.solvent = filename, solute, solvent
        self._x = XtcFile(filename)
        self.natoms_file = self._x.natoms
        self.nframes = self._x.nframes if lastframe <= 0 else min(self._x.nframes, lastframe)
        if max(solute.indices.max(), solvent.indices.max()) > self.natoms_file:
            raise ValueError("selection index outside the atoms of the XTC file")
        self._alloc()
        self._xyz = np.empty((self.natoms_file, 3), dtype=np.float32)
        self._k, self._cell = 0, np.zeros((3, 3))

    def open(self): self._k = 0
    def close(self): pass
    def firstframe(self): self._k = 0

    def nextframe(self, dst_solute=None, dst_solvent=None):
        if self._k >= self._x.nframes:
            raise EOFError("end of XTC trajectory")
        _, self._cell, self.step, self.time = self._x.read_frame(self._k, out=self._xyz)
        self._k += 1
        return self._gather(self._xyz, dst_solute, dst_solvent)

    def getunitcell(self):
        return self._cell


class PDBTraj(Trajectory):
    """Multi-model PDB trajectory, one CRYST1 per frame (src/trajectory_formats/PDBTraj.jl)."""

    def __init__(self, filename: str, solute: AtomSelection, solvent: AtomSelection, lastframe: int = -1):
        self.filename, self.solute, self.solvent = filename, solute, solvent
        self._frames, self._cells = [], []
        cur, cell = [], None
        with open(filename) as f:
            for line in f:
                if line.startswith("CRYST1"):
                    v = [float(t) for t in line[6:].split()[:6]]
                    cell = cell_from_lengths_angles(*v[:6])
                elif line.startswith(("ATOM", "HETATM")):
                    cur.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
                elif line.startswith("END"):
                    if cur:
                        self._frames.append(np.array(cur, dtype=np.float32)); self._cells.append(cell)
                    cur = []
        if cur:
            self._frames.append(np.array(cur, dtype=np.float32)); self._cells.append(cell)
        self.nframes = len(self._frames) if lastframe < 0 else min(len(self._frames), lastframe)
        self._alloc()
        self._k = 0

    def open(self): self._k = 0
    def close(self): pass
    def firstframe(self): self._k = 0

    def nextframe(self, dst_solute=None, dst_solvent=None):
        self._cur = self._k
        self._k += 1
        return self._gather(self._frames[self._cur], dst_solute, dst_solvent)

    def getunitcell(self):
        return self._cells[self._cur]


class ArrayTrajectory(Trajectory):
    """In-memory trajectory (synthetic benchmark systems, tests).  ``frames`` is either an
    array [nframes, natoms, 3] (fp32) or a callable ``f(k) -> ([natoms,3] fp32, cell 3x3)``."""

    def __init__(self, frames, cells, solute: AtomSelection, solvent: AtomSelection, nframes: Optional[int] = None):
        self.filename = "<memory>"
        self.solute, self.solvent = solute, solvent
        self._frames, self._cells_in = frames, cells
        self.nframes = int(nframes if nframes is not None else len(frames))
        self._alloc()
        self._k = 0

    def open(self): self._k = 0
    def close(self): pass
    def firstframe(self): self._k = 0

    def nextframe(self, dst_solute=None, dst_solvent=None):
        k = self._k
        self._k += 1
        if callable(self._frames):
            xyz, cell = self._frames(k)
        else:
            xyz = self._frames[k]
            cell = self._cells_in[k] if np.ndim(self._cells_in) == 3 else self._cells_in
        self._cell = np.asarray(cell, dtype=np.float64)
        return self._gather(np.asarray(xyz, dtype=np.float32), dst_solute, dst_solvent)

    def getunitcell(self):
        return self._cell


def make_trajectory(filename: str, solute: AtomSelection, solvent: Optional[AtomSelection] = None, *,
                    format: str = "", lastframe: int = -1) -> Trajectory:
    """Trajectory(filename, solute, solvent; format), src/Trajectory.jl:31-63.  One selection
    only -> autocorrelation (:61-63)."""
    if solvent is None:
        solvent = solute
    fmt = format
    if not fmt:
        low = filename.lower()
        fmt = "dcd" if low.endswith(".dcd") else "PDBTraj" if low.endswith(".pdb") else "xtc" if low.endswith(".xtc") else ""
    if fmt == "dcd":
        return NamdDCD(filename, solute, solvent, lastframe=lastframe)
    if fmt == "xtc":
        return XTCTraj(filename, solute, solvent, lastframe=lastframe)
    if fmt == "PDBTraj":
        return PDBTraj(filename, solute, solvent, lastframe=lastframe)
    raise ValueError(f"Unsupported trajectory format for {filename!r}: the B200 host shim reads DCD, XTC "
                     "and PDB natively; other formats stay with the Julia/Chemfiles reader.")


@dataclass
class TrajectoryMetaData:
    """src/Trajectory.jl:180-249."""
    irefatom: int
    lastframe_read: int
    nframes_read: int
    n_groups_solute: int
    n_groups_solvent: int
    unitcell: np.ndarray


def trajectory_metadata(trajectory: Trajectory, options) -> TrajectoryMetaData:
    if options.irefatom > trajectory.solvent.natomspermol:
        raise ValueError(f"in MDDF options: Reference atom index {options.irefatom} is greater than number "
                         "of atoms of the solvent molecule. ")
    if options.lastframe > trajectory.nframes:
        raise ValueError("in MDDF options: lastframe is greater than trajectory.nframes. ")
    trajectory.open()
    trajectory.firstframe()
    _, xv = trajectory.nextframe()
    unitcell = np.array(trajectory.getunitcell(), dtype=np.float64)
    if options.irefatom == -1:
        # closest atom to the centre of coordinates of the first solvent molecule, :206-213
        first = xv[: trajectory.solvent.natomspermol].astype(np.float64)
        cm = first.mean(axis=0)
        irefatom = int(np.argmin(np.linalg.norm(first - cm, axis=1))) + 1
    else:
        irefatom = options.irefatom
    lastframe_read = trajectory.nframes if options.lastframe == -1 else options.lastframe
    nframes_read = len(range(options.firstframe, lastframe_read + 1, options.stride))
    trajectory.close()
    return TrajectoryMetaData(irefatom, lastframe_read, nframes_read, trajectory.solute.n_groups,
                              trajectory.solvent.n_groups, unitcell)
