"""Loader for tests/golden/*.npz (written by tools/make_golden.py from the unmodified reference)."""
from __future__ import annotations

import glob
import os

import numpy as np

import refdump

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.vxa_text = bytes(z["vxa"]).decode("latin-1")
        self.lw = bool(z["lw"])
        self.n_steps, self.every = int(z["n_steps"]), int(z["every"])
        self.dump = refdump.RefDump(z["vox"].shape[0], z["bond"].shape[0], z["sim"], z["vox"], z["bond"], z["frame_hdr"],
                                    z["vstate"], z["bstate"], np.zeros(0))
        self.full_hdr = z["full_hdr"] if "full_hdr" in z else None
        self.full_vstate = z["full_vstate"] if "full_vstate" in z else None
        self.full_tail = z["full_tail"] if "full_tail" in z else None
        self.result_xml = bytes(z["result_xml"]).decode("latin-1") if "result_xml" in z else None
        self.fitness_file = bytes(z["fitness_file"]).decode() if "fitness_file" in z else None

    def robot(self):
        """Descriptor derived from the reference's own imported arrays (independent of the product importer)."""
        import re
        m = re.search(r"<CollisionHorizon>([^<]+)<", self.vxa_text)
        return refdump.desc_from_dump(self.dump, lw=self.lw, col_horizon=float(m.group(1)) if m else 2.0)

    @property
    def collide(self):
        return bool(self.dump.sim[15])

    @property
    def fluid(self):
        return self.lw and bool(self.dump.sim[22])


def flags(name: str):
    """(fluid, collide) of a fixture, read from its header (not from its name)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    lw, sim = bool(z["lw"]), z["sim"]
    return (lw and bool(sim[22])), bool(sim[15])


def parse_result(xml: str) -> dict:
    """tag -> float of a <Voxelyze_Sim_Result>, read the way evosoro does (line-wise <tag>value</tag>,
    read_write_voxelyze.py:27-35)."""
    import re
    return {m.group(1): float(m.group(2)) for m in re.finditer(r"<(\w+)>([^<]+)</\1>", xml)}


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
