"""Options -- host-side mirror of the reference's immutable ``Options`` struct.

Reference: src/Options.jl:6-42 (fields), :68-192 (validating keyword constructor).  Same
field names, defaults and error behaviour (``ArgumentError`` -> ``ValueError``).  The fields
the device path reads are flattened into the C-ABI ``cmx_config`` (include/cmx_b200.h).
``nthreads``, ``GC``, ``GC_threshold`` and ``lcell`` are accepted for drop-in compatibility;
on the B200 path they are hints/no-ops.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass


@dataclass(frozen=True)
class Options:
    firstframe: int = 1
    lastframe: int = -1
    stride: int = 1
    irefatom: int = -1
    n_random_samples: int = 10
    binstep: float = 0.02
    dbulk: float = 10.0
    cutoff: float = 10.0
    usecutoff: bool = False
    lcell: int = 1
    GC: bool = True
    GC_threshold: float = 0.3
    seed: int = 321
    StableRNG: bool = False
    nthreads: int = 0
    silent: bool = False

    def __init__(self, *, firstframe=1, lastframe=-1, stride=1, irefatom=-1, n_random_samples=10,
                 binstep=0.02, dbulk=None, cutoff=None, usecutoff=None, bulk_range=None, lcell=1,
                 GC=True, GC_threshold=0.3, seed=321, StableRNG=False, nthreads=0, silent=False):
        warn = False
        # src/Options.jl:91-105
        if stride < 1:
            raise ValueError("in MDDF options: stride cannot be less than 1. ")
        if lastframe > 0 and lastframe < firstframe:
            raise ValueError("in MDDF options: lastframe must be greater or equal to firstframe. ")
        if n_random_samples < 1:
            raise ValueError("in MDDF options: n_random_samples must be greater than 0. "
                             "To skip the normalization of the distribution, use the "
                             "coordination_number(...) function instead of mddf(...).")
        # src/Options.jl:107-147
        if bulk_range is not None and any(v is not None for v in (dbulk, cutoff, usecutoff)):
            raise ValueError("The bulk_range argument implies that dbulk, cutoff, and usecutoff are not needed.")
        if all(v is None for v in (bulk_range, dbulk, cutoff, usecutoff)):
            dbulk, cutoff, usecutoff, warn = 10.0, 10.0, False, True
        elif bulk_range is not None:
            if len(bulk_range) != 2:
                raise ValueError("bulk_range must be a tuple or vector with two elements, "
                                 "corresponding to dbulk and cutoff.")
            dbulk, cutoff = bulk_range
            usecutoff = True
        else:
            if dbulk is None:
                dbulk, warn = 10.0, True
            if usecutoff is None:
                usecutoff, warn = False, True
            if cutoff is None:
                cutoff = dbulk + 4.0 if usecutoff else dbulk
                warn = True
            elif not usecutoff:
                raise ValueError("in MDDF options: cutoff was defined with usecutoff set to false")
        if warn and not silent:
            warnings.warn(f"Using default values for dbulk, cutoff and/or usecutoff: dbulk = {dbulk} "
                          f"cutoff = {cutoff} usecutoff = {usecutoff}. It is recommended to set "
                          "bulk_range manually, e.g. Options(bulk_range=(8.0, 12.0))", stacklevel=2)
        # src/Options.jl:162-170
        if usecutoff and dbulk >= cutoff:
            raise ValueError(" in MDDF options: The bulk volume is zero (dbulk must be smaller than cutoff). ")
        if (cutoff / binstep) % 1 > 1.0e-5:
            raise ValueError("in MDDF options: cutoff must be a multiple of binstep.")
        if (dbulk / binstep) % 1 > 1.0e-5:
            raise ValueError("in MDDF options: dbulk must be a multiple of binstep.")
        s = object.__setattr__
        s(self, "firstframe", int(firstframe)); s(self, "lastframe", int(lastframe)); s(self, "stride", int(stride))
        s(self, "irefatom", int(irefatom)); s(self, "n_random_samples", int(n_random_samples))
        s(self, "binstep", float(binstep)); s(self, "dbulk", float(dbulk)); s(self, "cutoff", float(cutoff))
        s(self, "usecutoff", bool(usecutoff)); s(self, "lcell", int(lcell)); s(self, "GC", bool(GC))
        s(self, "GC_threshold", float(GC_threshold)); s(self, "seed", int(seed)); s(self, "StableRNG", bool(StableRNG))
        s(self, "nthreads", int(nthreads)); s(self, "silent", bool(silent))

    def to_dict(self):
        return {k: getattr(self, k) for k in self.__dataclass_fields__}
