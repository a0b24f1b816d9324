"""complexmixtures.jl_b200 -- B200-native minimum-distance engine behind ComplexMixtures.jl's API.

The directory name carries a dot (it is the name the build contract fixes), so it cannot be
imported with a plain ``import``; ``cmx_b200.py`` at the repository root loads it under the
module name ``cmx_b200``.  Layout:

  csrc/            CUDA kernels (sm_100a) + the C-ABI (include/cmx_b200.h) -> libcmx_b200.so
  engine.py        ctypes binding of the C-ABI (what the Julia shim does with ccall)
  driver.py        mddf() / coordination_number() drivers (mirror of src/mddf.jl)
  options.py, selection.py, trajectory.py, results.py, contributions.py   host-side mirrors
  synthetic.py     generators of the synthetic benchmark systems named in BASELINE.json
"""
from .options import Options
from .selection import AtomSelection, SoluteGroup, SolventGroup
from .trajectory import (ArrayTrajectory, NamdDCD, PDBTraj, XTCTraj, Trajectory, make_trajectory,
                         trajectory_metadata, cell_from_lengths_angles)
from .results import (Result, finalresults, load, merge, save, setbin, shellradius, sphericalshellvolume)
from .contributions import contributions, coordination_number_of

__all__ = ["Options", "AtomSelection", "SoluteGroup", "SolventGroup", "Trajectory", "NamdDCD", "PDBTraj", "XTCTraj",
           "ArrayTrajectory", "make_trajectory", "trajectory_metadata", "Result", "finalresults", "load", "merge", "save",
           "setbin", "shellradius", "sphericalshellvolume", "contributions", "coordination_number_of",
           "cell_from_lengths_angles"]


from .driver import coordination_number, mddf, mddf_many  # noqa: E402  (the CUDA library itself is loaded lazily)
from . import engine, synthetic  # noqa: E402
from .engine import Engine  # noqa: E402

__all__ += ["mddf", "mddf_many", "coordination_number", "Engine", "engine", "synthetic"]
