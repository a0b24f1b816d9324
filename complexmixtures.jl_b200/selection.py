"""AtomSelection / SoluteGroup / SolventGroup -- host-side mirrors.

Reference: src/AtomSelection.jl:48-59 (struct), :290-372 (low-level constructor: group
indices are sorted :315-323, must be unique :326 and a subset of the selection :335),
:551-633 (group selectors).  Indices are 1-based atom numbers of the structure file, exactly
as in the reference; the device path only ever sees *positions within the selection*.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


@dataclass
class AtomSelection:
    indices: np.ndarray
    nmols: int = 0
    natomspermol: int = 0
    custom_groups: bool = False
    group_atom_indices: List[np.ndarray] = field(default_factory=list)
    group_names: List[str] = field(default_factory=list)

    def __init__(self, indices: Sequence[int], *, nmols: int = 0, natomspermol: int = 0,
                 group_atom_indices: Optional[Sequence[Sequence[int]]] = None,
                 group_names: Optional[Sequence[str]] = None):
        indices = np.asarray(indices, dtype=np.int64)
        natoms = len(indices)
        # set_nmols_natomspermol, src/AtomSelection.jl
        if nmols == 0 and natomspermol == 0:
            raise ValueError("Set nmols or natomspermol when defining a selection.")
        if natoms == 0:
            raise ValueError("Vector of atom indices provided is empty.")
        if nmols != 0:
            if natoms % nmols != 0:
                raise ValueError(f"Number of atoms in selection ({natoms}) is not a multiple of nmols ({nmols}).")
            natomspermol = natoms // nmols
        else:
            if natoms % natomspermol != 0:
                raise ValueError(f" Number of atoms in selection ({natoms}) is not a multiple of natomspermol ({natomspermol}).")
            nmols = natoms // natomspermol
        group_atom_indices = [np.asarray(g, dtype=np.int64) for g in (group_atom_indices or [])]
        group_names = list(group_names or [])
        custom = len(group_atom_indices) > 0
        if not custom and group_names and len(group_names) != natomspermol:
            raise ValueError("The length of the group_names vector does not correspond to the number of "
                             "atoms per molecule, but no atom groups vector was provided.")
        if custom:
            sel = set(indices.tolist())
            for k, g in enumerate(group_atom_indices):
                if not np.all(g[:-1] <= g[1:]):
                    warnings.warn("Group indices are not sorted. The array will be sorted for faster search.")
                    group_atom_indices[k] = g = np.sort(g)
                if len(np.unique(g)) != len(g):
                    raise ValueError("Found repeated indices in custom group atom indices.")
                if any(int(i) not in sel for i in g):
                    raise ValueError("Group atom indices not found in the the current AtomSelection main atomic indices.")
            if not group_names:
                warnings.warn("Vector of group atom indices was provided but vector of group names is empty.")
            elif len(group_names) != len(group_atom_indices):
                raise ValueError("The vector of group atom indices has a different number of elements than the vector of group names.")
        self.indices = indices
        self.nmols = int(nmols)
        self.natomspermol = int(natomspermol)
        self.custom_groups = custom
        self.group_atom_indices = group_atom_indices
        self.group_names = group_names

    @property
    def natoms(self) -> int:
        return self.nmols * self.natomspermol

    @property
    def n_groups(self) -> int:
        """TrajectoryMetaData.n_groups_*, src/Trajectory.jl:231-240."""
        return len(self.group_atom_indices) if self.custom_groups else self.natomspermol

    def group_csr(self):
        """CSR map "position in the selection -> groups containing indices[position]".

        Restates the per-hit search of update_group_count! (src/update_counters.jl:27-33) as a
        table built once: an atom credits *every* custom group that contains it.
        Returns (offsets[natoms+1], ids) as int32 arrays, or (None, None) without custom groups.
        """
        if not self.custom_groups:
            return None, None
        pos_of = {int(a): p for p, a in enumerate(self.indices.tolist())}
        per_pos = [[] for _ in range(len(self.indices))]
        for g, inds in enumerate(self.group_atom_indices):
            for a in inds.tolist():
                per_pos[pos_of[int(a)]].append(g)
        off = np.zeros(len(per_pos) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(p) for p in per_pos])
        ids = np.array([g for p in per_pos for g in p], dtype=np.int32)
        if len(ids) == 0:
            ids = np.zeros(1, dtype=np.int32)
        return off, ids

    def to_dict(self):
        return dict(nmols=self.nmols, natomspermol=self.natomspermol, indices=self.indices.tolist(),
                    custom_groups=self.custom_groups,
                    group_atom_indices=[g.tolist() for g in self.group_atom_indices],
                    group_names=list(self.group_names))


@dataclass
class SoluteGroup:
    """src/AtomSelection.jl:551-590 -- select by group index, group name, atom indices or atom names."""
    group_index: Optional[int] = None
    group_name: Optional[str] = None
    atom_indices: Optional[Sequence[int]] = None
    atom_names: Optional[Sequence[str]] = None

    def __init__(self, arg):
        self.group_index = self.group_name = self.atom_indices = self.atom_names = None
        if isinstance(arg, (int, np.integer)):
            self.group_index = int(arg)
        elif isinstance(arg, str):
            self.group_name = arg
        else:
            arg = list(arg)
            if len(arg) > 0 and isinstance(arg[0], str):
                self.atom_names = arg
            else:
                self.atom_indices = [int(a) for a in arg]


class SolventGroup(SoluteGroup):
    pass
