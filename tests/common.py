"""Shared helpers for the tests: synthetic inputs and the two checkers (reference dump, C oracle)."""
from __future__ import annotations

import os
import tempfile

import numpy as np

import refdump
from evosoro_b200 import abi, vxa

TMP = os.path.join(tempfile.gettempdir(), "vx_tests")
os.makedirs(TMP, exist_ok=True)


def write_vxa(name: str, structure, **kw) -> str:
    path = os.path.join(TMP, name + ".vxa")
    with open(path, "w") as f:
        f.write(vxa.vxa_text(structure, **kw))
    return path


def dense_vxa(n: int, seed: int, **kw) -> str:
    return write_vxa(f"dense{n}_s{seed}", vxa.dense_lattice(n, seed), **kw)


def robot_via_reference(vxa_path: str, lw: bool = False) -> abi.HostRobot:
    """Import a .vxa with the UNMODIFIED reference (ref_dump, 0 steps) and flatten its arrays."""
    out = vxa_path + ".imp.bin"
    d = refdump.run_ref_dump(vxa_path, out, 0, 1, nocol=True, lw=lw)
    return refdump.desc_from_dump(d, lw=lw)


def max_abs(a, b) -> float:
    return float(np.abs(np.asarray(a) - np.asarray(b)).max())
