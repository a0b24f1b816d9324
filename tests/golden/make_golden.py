"""Generates the committed fixtures under tests/golden/ from the reference's own test data.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs
  namd_fixture.npz  coordinates (fp32, as stored in the DCD) of the atoms used by the reference's
                    NAMD tests, from test/data/NAMD/traj_duplicated_first_frame.dcd (3 frames; frames
                    1 and 2 are identical, src/mddf.jl:860-862): protein 1..1463, TMAO 1479..4012
                    (181 x 14) for all 3 frames, water 4013..62026 (19338 x 3) for frame 1; unit cells.
  toy.npz           the four toy PDB systems of src/mddf.jl:587-758 (test/data/toy/*.pdb).
  kat.json          known answers quoted from the reference's tests and golden JSON files.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cmx_b200 as cm  # noqa: E402

REF = "/root/reference/test/data"


def main():
    allsel = cm.AtomSelection(np.arange(1, 62027), nmols=1)
    t = cm.NamdDCD(f"{REF}/NAMD/traj_duplicated_first_frame.dcd", allsel, allsel)
    t.open()
    frames, cells = [], []
    for _ in range(t.nframes):
        x, _ = t.nextframe()
        frames.append(x.copy()); cells.append(t.getunitcell().copy())
    t.close()
    frames = np.stack(frames)
    np.savez_compressed(os.path.join(HERE, "namd_fixture.npz"),
                        protein=frames[:, 0:1463], tmao=frames[:, 1478:4012], water_frame1=frames[0, 4012:62026],
                        cells=np.stack(cells))
    toy = {}
    for name in ("cross", "self", "self_monoatomic", "self_monoatomic_duplicated_first_frame"):
        sel = None
        path = f"{REF}/toy/{name}.pdb"
        natoms = sum(1 for line in open(path).read().split("END")[0].splitlines() if line.startswith("ATOM"))
        sel = cm.AtomSelection(np.arange(1, natoms + 1), nmols=1)
        tr = cm.PDBTraj(path, sel, sel)
        tr.open()
        fr, ce = [], []
        for _ in range(tr.nframes):
            x, _ = tr.nextframe(); fr.append(x.copy()); ce.append(tr.getunitcell().copy())
        toy[name] = np.stack(fr); toy[name + "_cells"] = np.stack(ce)
    np.savez_compressed(os.path.join(HERE, "toy.npz"), **toy)
    kat = {
        "source": "reference tests; values quoted verbatim",
        "coordination_number.jl:126-134": {"cn_first_d_gt_3": 7.0, "cn_first_d_gt_5": 14.0, "sum_cn_O1": 1171.0,
                                            "options": {"lastframe": 1, "n_random_samples": 200}},
        "results.jl:277-281": {"shellradius(1,0.1)": 0.07937005259840998, "shellradius(5,0.3)": 1.3664650373440481},
        "mddf.jl:587-624 toy cross": {"volume_total": 27000.0, "sum_md_count": 1.0, "sum_coordination_number": 51.0,
                                      "n_random_samples": 100000},
        "minimum_distances.jl:204": {"unitcell": 84.42188262939453, "nmols_tmao": 181, "natoms_tmao": 2534, "natoms_protein": 1463},
        "irefatom": {},
        "golden_json_sums": {},
    }
    for f in ("tmao_tmao", "water_tmao", "water_water"):
        d = json.load(open(f"{REF}/NAMD/{f}.json"))
        kat["irefatom"][f] = d["files"][0]["irefatom"]
        kat["golden_json_sums"][f] = {"sum_md_count": float(np.sum(d["md_count"])), "sum_rdf_count": float(np.sum(d["rdf_count"])),
                                      "nbins": d["nbins"], "solute_nmols": d["solute"]["nmols"], "solvent_nmols": d["solvent"]["nmols"],
                                      "solvent_first_index": d["solvent"]["indices"][0], "solute_first_index": d["solute"]["indices"][0],
                                      "frames_used": [1, 6, 11, 16], "md_count": d["md_count"]}
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
