"""contributions / coordination_number accessors -- pure readers of the hot path's output.

Reference: src/tools/contributions.jl:70-248, src/tools/coordination_number.jl:105-106.
Kept so that a user of the reference finds the same post-processing calls working on the
``Result`` filled by the B200 engine.
"""
from __future__ import annotations

import numpy as np

from .results import ANGS3_TO_CM3_PER_MOL, Result
from .selection import SoluteGroup, SolventGroup

_TYPES = ("mddf", "coordination_number", "md_count", "kbi")


def contributions(R: Result, group, *, type: str = "mddf") -> np.ndarray:
    if type not in _TYPES:
        raise ValueError("type must be :mddf (default), :coordination_number, :md_count, :kbi")
    if isinstance(group, SolventGroup):
        atsel, gc, gcr = R.solvent, R.solvent_group_count, R.solvent_group_count_random
    elif isinstance(group, SoluteGroup):
        atsel, gc, gcr = R.solute, R.solute_group_count, R.solute_group_count_random
    else:
        raise TypeError("group must be a SoluteGroup or SolventGroup")
    if atsel.custom_groups:
        if group.group_index is None and group.group_name is None:
            raise ValueError("Custom groups are defined. Cannot retrieve general group contributions. "
                             "Please provide a group name or index.")
    elif group.atom_indices is None and group.atom_names is None:
        raise ValueError(f'The "{group.group_name}" string identifier of the group was set, but no custom '
                         "group names were defined. Please provide vectors of *atomic* indices or names.")
    sel, selr = np.zeros(R.nbins), np.zeros(R.nbins)
    if group.group_index is not None:
        ig = group.group_index
        if ig > len(gc):
            raise ValueError(f"Group {ig} greater than number of groups ({len(gc)}) of group contribution array.")
        sel, selr = gc[ig - 1].copy(), gcr[ig - 1].copy()
    if group.group_name is not None:
        if group.group_name not in atsel.group_names:
            raise ValueError(f"Group (or atom) name {group.group_name} not found in group names.")
        ig = atsel.group_names.index(group.group_name)
        sel, selr = gc[ig].copy(), gcr[ig].copy()
    if group.atom_indices is not None:
        ai = list(group.atom_indices)
        if not ai:
            raise ValueError("Group selection by group indices is empty.")
        if len(set(ai)) != len(ai):
            raise ValueError("Selection by atom indices contains repeated indices.")
        idx = atsel.indices.tolist()
        for iat in ai:
            if iat not in idx:
                raise ValueError(f"Atom index {iat} not found in the selection.")
            if atsel.nmols == 1:
                it = idx.index(iat)
                sel += gc[it]; selr += gcr[it]
            else:  # per-type arrays, scaled to one molecule (contributions.jl:186-193)
                it = (iat - idx[0]) % atsel.natomspermol
                sel += gc[it] / atsel.nmols; selr += gcr[it] / atsel.nmols
    if group.atom_names is not None:
        an = list(group.atom_names)
        if not an:
            raise ValueError("Selection by atom names is empty.")
        if len(set(an)) != len(an):
            raise ValueError("Selection by atom names contains repeated names.")
        for name in an:
            found = False
            for ig, gname in enumerate(atsel.group_names):
                if gname == name:
                    found = True
                    sel += gc[ig]; selr += gcr[ig]
            if not found:
                raise ValueError(f"Group (or atom) name {name} not found in group names.")
    if type == "mddf":
        out = np.zeros(R.nbins)
        pos = R.md_count_random != 0.0
        out[pos] = sel[pos] / R.md_count_random[pos]
        return out
    if type == "coordination_number":
        return np.cumsum(sel)
    if type == "md_count":
        return sel
    with np.errstate(divide="ignore", invalid="ignore"):
        return ANGS3_TO_CM3_PER_MOL * (1 / np.float64(R.density.solvent_bulk)) * (np.cumsum(sel) - np.cumsum(selr))


def coordination_number_of(R: Result, group=None) -> np.ndarray:
    """coordination_number(R[, group]), src/tools/coordination_number.jl:105-106."""
    return R.coordination_number if group is None else contributions(R, group, type="coordination_number")
