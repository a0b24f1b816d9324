"""Synthetic solvated systems of the sizes named in BASELINE.json.

There is no network and most of the reference's trajectories are missing, so the benchmark
systems are generated: a compact "protein" (jittered lattice points in a sphere or slab), rigid
solvent templates (water 3 sites, urea 8, TMAO/glycerol-like 14) placed on a jittered lattice
outside the solute, random orientations.  Frame k is the base configuration with every solvent
molecule displaced by a Gaussian step (sigma 0.5 A) and translated by a random lattice vector
(coordinates are left UNWRAPPED, like real trajectories).  Everything is keyed by (seed, k) so any
rank can generate any frame.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .selection import AtomSelection
from .trajectory import ArrayTrajectory


def _template(kind: str) -> np.ndarray:
    if kind == "water":
        a = np.deg2rad(104.52 / 2)
        return np.array([[0, 0, 0], [0.9572 * np.sin(a), 0.9572 * np.cos(a), 0], [-0.9572 * np.sin(a), 0.9572 * np.cos(a), 0]])
    rng = np.random.default_rng({"urea": 11, "tmao": 12, "glycerol": 13, "cosolvent": 14}.get(kind, 15))
    n = {"urea": 8, "tmao": 14, "glycerol": 14, "cosolvent": 14}.get(kind, 14)
    # compact random cluster with ~1.2-1.5 A neighbour distances
    pts = [np.zeros(3)]
    while len(pts) < n:
        base = pts[rng.integers(len(pts))]
        v = rng.normal(size=3); v *= 1.4 / np.linalg.norm(v)
        p = base + v
        if min(np.linalg.norm(p - q) for q in pts) > 1.0:
            pts.append(p)
    return np.array(pts)


def _rotations(rng, n):
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


@dataclass
class SyntheticSystem:
    name: str
    cell: np.ndarray                 # 3x3, columns = lattice vectors
    base: np.ndarray                 # [natoms,3] float64 base configuration
    mol_first: np.ndarray            # first atom of every rigid unit that moves together
    mol_size: np.ndarray
    selections: dict                 # name -> AtomSelection
    seed: int
    n_solute_atoms: int              # atoms [0, n_solute_atoms) never move

    @property
    def natoms(self):
        return len(self.base)

    def frame(self, k: int):
        rng = np.random.default_rng([self.seed, k])
        nm = len(self.mol_first)
        step = rng.normal(scale=0.5, size=(nm, 3))
        shift = rng.integers(-2, 3, size=(nm, 3)).astype(np.float64) @ self.cell.T
        disp = step + shift
        xyz = self.base.copy()
        per_atom = np.repeat(disp, self.mol_size, axis=0)
        xyz[self.n_solute_atoms:] += per_atom
        return xyz.astype(np.float32), self.cell

    def trajectory(self, solute: str, solvent: Optional[str], nframes: int) -> ArrayTrajectory:
        s = self.selections[solute]
        v = self.selections[solvent] if solvent else s
        return ArrayTrajectory(lambda k: self.frame(k), None, s, v, nframes=nframes)


def make_system(name: str, *, cell, solute_atoms: int, solute_shape: str = "sphere", solute_mols: int = 1,
                solvents=(), seed: int = 2002, spacing_solute: float = 2.3) -> SyntheticSystem:
    """solvents: sequence of (selection_name, template_kind, n_molecules)."""
    rng = np.random.default_rng(seed)
    cell = np.asarray(cell, dtype=np.float64)
    if cell.shape == (3,):
        cell = np.diag(cell)
    inv = np.linalg.inv(cell)
    centre = cell @ np.array([0.5, 0.5, 0.5])
    vol = abs(np.linalg.det(cell))
    # ---- solute: jittered lattice inside a sphere / slab around the cell centre
    a = spacing_solute
    if solute_atoms > 0:
        if solute_shape == "sphere":
            R = (3 * solute_atoms * a ** 3 / (4 * np.pi)) ** (1 / 3) * 1.02
            while True:
                n = int(np.ceil(R / a)) + 1
                g = np.arange(-n, n + 1) * a
                P = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
                P = P[np.linalg.norm(P, axis=1) <= R]
                if len(P) >= solute_atoms:
                    break
                R *= 1.02
            d = np.linalg.norm(P, axis=1)
            P = P[np.argsort(d, kind="stable")[:solute_atoms]]
            Rsol = float(np.linalg.norm(P, axis=1).max())
            excl = lambda X: np.linalg.norm(X - centre, axis=1) < Rsol + 2.0
        else:  # slab spanning x and y, centred in z (orthorhombic cells)
            Lx, Ly = cell[0, 0], cell[1, 1]
            nx, ny = int(Lx // a), int(Ly // a)
            nz = int(np.ceil(solute_atoms / (nx * ny)))
            gx = (np.arange(nx) + 0.5) * (Lx / nx) - Lx / 2; gy = (np.arange(ny) + 0.5) * (Ly / ny) - Ly / 2
            gz = (np.arange(nz) - (nz - 1) / 2) * a
            P = np.stack(np.meshgrid(gx, gy, gz, indexing="ij"), -1).reshape(-1, 3)[:solute_atoms]
            half = nz * a / 2
            excl = lambda X: np.abs(X[:, 2] - centre[2]) < half + 2.0
        P = P + rng.uniform(-0.35, 0.35, size=P.shape) + centre
        # order the solute atoms along a space-filling-ish path so that "residues" are compact
        order = np.lexsort((P[:, 0] // 6, P[:, 1] // 6, P[:, 2] // 6))
        solute_xyz = P[order]
    else:
        solute_xyz = np.zeros((0, 3))
        excl = lambda X: np.zeros(len(X), dtype=bool)
    # ---- solvent sites: jittered lattice in fractional space
    nmol_total = sum(n for _, _, n in solvents)
    parts, mol_first, mol_size, selections = [solute_xyz], [], [], {}
    natoms = len(solute_xyz)
    if solute_atoms > 0:
        selections["solute"] = AtomSelection(np.arange(1, solute_atoms + 1), nmols=solute_mols)
    if nmol_total > 0:
        free_frac = 1.0
        if solute_atoms > 0:
            probe = (rng.uniform(size=(20000, 3)) @ cell.T)
            free_frac = max(0.05, 1.0 - excl(probe).mean())
        s = (vol * free_frac / (nmol_total * 1.15)) ** (1 / 3)
        nn = [max(1, int(np.floor(np.linalg.norm(cell[:, k]) / s))) for k in range(3)]
        while True:
            f = np.stack(np.meshgrid(*[(np.arange(n) + 0.5) / n for n in nn], indexing="ij"), -1).reshape(-1, 3)
            X = f @ cell.T
            X = X[~excl(X)]
            if len(X) >= nmol_total:
                break
            nn = [n + 1 for n in nn]
        X = X[rng.permutation(len(X))[:nmol_total]]
        X = X + rng.uniform(-0.25, 0.25, size=X.shape) * s
        off = 0
        for sel_name, kind, nmol in solvents:
            T = _template(kind); T = T - T.mean(axis=0)
            Rm = _rotations(rng, nmol)
            mol = np.einsum("nij,kj->nki", Rm, T) + X[off:off + nmol, None, :]
            off += nmol
            parts.append(mol.reshape(-1, 3))
            idx = np.arange(natoms + 1, natoms + nmol * len(T) + 1)
            selections[sel_name] = AtomSelection(idx, natomspermol=len(T))
            mol_first += list(range(natoms, natoms + nmol * len(T), len(T)))
            mol_size += [len(T)] * nmol
            natoms += nmol * len(T)
    base = np.concatenate(parts, axis=0)
    return SyntheticSystem(name, cell, base, np.array(mol_first, dtype=np.int64), np.array(mol_size, dtype=np.int64),
                           selections, seed, len(solute_xyz))


def residue_groups(solute: AtomSelection, atoms_per_residue: int = 16) -> AtomSelection:
    """The same solute selection with consecutive-atom "residue" custom groups (C2's 375-group run)."""
    idx = solute.indices
    groups = [idx[k:k + atoms_per_residue] for k in range(0, len(idx), atoms_per_residue)]
    names = [f"RES{k + 1}" for k in range(len(groups))]
    return AtomSelection(idx, nmols=solute.nmols, group_atom_indices=groups, group_names=names)


# ---- the named configurations (BASELINE.json configs[1..4]; sizes from SURVEY.md section 8d) ----
def config_c2(scale: float = 1.0) -> SyntheticSystem:
    """C2: 100 000 atoms, cubic 100 A: protein 6 000 atoms, urea 800 x 8, water 29 200 x 3."""
    L = 100.0 * scale ** (1 / 3)
    return make_system("C2", cell=[L, L, L], solute_atoms=int(6000 * scale),
                       solvents=[("urea", "urea", int(800 * scale)), ("water", "water", int(29200 * scale))], seed=2002)


def config_c3(scale: float = 1.0) -> SyntheticSystem:
    """C3: 199 999 atoms, triclinic a=(130,0,0) b=(30,125,0) c=(20,25,123): glycerol 5 000 x 14, water 43 333 x 3."""
    f = scale ** (1 / 3)
    cell = np.array([[130.0, 30.0, 20.0], [0.0, 125.0, 25.0], [0.0, 0.0, 123.0]]) * f
    return make_system("C3", cell=cell, solute_atoms=0,
                       solvents=[("glycerol", "glycerol", int(5000 * scale)), ("water", "water", int(43333 * scale))], seed=2003)


def config_c4(scale: float = 1.0) -> SyntheticSystem:
    """C4: 999 999 atoms, cubic 216 A: protein 20 000 atoms, cosolvent 5 000 x 14, water 303 333 x 3."""
    L = 216.0 * scale ** (1 / 3)
    return make_system("C4", cell=[L, L, L], solute_atoms=int(20000 * scale),
                       solvents=[("cosolvent", "cosolvent", int(5000 * scale)), ("water", "water", int(303333 * scale))], seed=2004)


def config_c5(scale: float = 1.0) -> SyntheticSystem:
    """C5: 5 000 000 atoms, 400 x 400 x 312.5 A: slab solute 1 000 000 atoms, water 1 333 333 x 3."""
    f = scale ** (1 / 3)
    return make_system("C5", cell=[400.0 * f, 400.0 * f, 312.5 * f], solute_atoms=int(1000000 * scale), solute_shape="slab",
                       solvents=[("water", "water", int(1333333 * scale))], seed=2005)
