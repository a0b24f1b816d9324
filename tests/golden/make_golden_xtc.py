"""Golden summary of the reference's XTC fixture (test/data/nucleic/trajectory.xtc) decoded by the library's native
reader: run in the build container (the fixture does not travel to the GPU box).  The decoder itself is pinned by
physics, not by this file: the last 1000 molecules of the fixture are TIP3P water and come out with O-H = 0.9572 A and
H-H = 1.5139 A to within the 0.01 A quantisation of the format (tests/test_host.py::test_native_xtc_reader_reference_fixture);
this file pins the decoded values against regressions.

    python tests/golden/make_golden_xtc.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cmx_b200.engine import XtcFile  # noqa: E402

SRC = "/root/reference/test/data/nucleic/trajectory.xtc"
f = XtcFile(SRC)
out = {"source": "test/data/nucleic/trajectory.xtc", "natoms": f.natoms, "nframes": f.nframes, "frames": []}
for k in range(f.nframes):
    x, cell, step, time = f.read_frame(k)
    out["frames"].append({"step": step, "time": time, "cell": cell.tolist(),
                          "first_atoms": x[:4].astype(float).tolist(), "last_atoms": x[-3:].astype(float).tolist(),
                          "sum": np.sum(x.astype(np.float64), axis=0).tolist(),
                          "sum_sq": float(np.sum(x.astype(np.float64) ** 2))})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "xtc_nucleic.json"), "w"), indent=1)
print("wrote xtc_nucleic.json", f.natoms, f.nframes)

# the first frame of the fixture, byte for byte (the frames of an XTC file are self-contained): a real GROMACS-written
# compressed frame that travels to the GPU box, where the DEVICE decoder is checked against the host decoder and against
# the summary above (tests/test_gpu_feed.py)
raw = open(SRC, "rb").read()
import struct
nbytes = struct.unpack(">i", raw[56 + 32:56 + 36])[0]          # header 56 B, then the 36-byte block header
first_len = 56 + 36 + ((nbytes + 3) & ~3)
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "nucleic_frame0.xtc"), "wb").write(raw[:first_len])
g = XtcFile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "nucleic_frame0.xtc"))
assert (g.natoms, g.nframes) == (f.natoms, 1) and np.array_equal(g.read_frame(0)[0], f.read_frame(0)[0])
print("wrote nucleic_frame0.xtc", first_len, "bytes")
