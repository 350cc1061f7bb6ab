"""Reader for the binary dumps written by oracle/ref_dump.cpp (the unmodified reference engine) and a
descriptor builder that derives the engine's flat model from the dump header with the closed forms of
SURVEY.md Appendix A.4 -- independent of the product importer, so the two can be checked against
each other.  Test infrastructure only."""
from __future__ import annotations

import math
import os
import struct
import subprocess
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from evosoro_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_DUMP = os.path.join(REF_DIR, "ref_dump")
REF_DUMP_LW = os.path.join(REF_DIR, "ref_dump_lw")


@dataclass
class RefDump:
    n_vox: int
    n_bond: int
    sim: np.ndarray        # [N_SIM]
    vox: np.ndarray        # [n_vox, N_VD]
    bond: np.ndarray       # [n_bond, N_BD]
    frame_hdr: np.ndarray  # [n_frames, 10] time, dt, step, cm xyz, inicm xyz, ncol
    vstate: np.ndarray     # [n_frames, n_vox, N_VS]
    bstate: np.ndarray     # [n_frames, n_bond, N_BS]
    tail: np.ndarray       # trailing doubles ("full" mode fitness block)

    # column helpers
    @property
    def pos(self): return self.vstate[:, :, 0:3]
    @property
    def quat(self): return self.vstate[:, :, 3:7]
    @property
    def vel(self): return self.vstate[:, :, 7:10]
    @property
    def angvel(self): return self.vstate[:, :, 10:13]
    @property
    def scale(self): return self.vstate[:, :, 13]
    @property
    def sfric(self): return self.vstate[:, :, 14]
    @property
    def time(self): return self.frame_hdr[:, 0]
    @property
    def com(self): return self.frame_hdr[:, 3:6]
    @property
    def ini_com(self): return self.frame_hdr[:, 6:9]


def load_dump(path: str) -> RefDump:
    raw = open(path, "rb").read()
    assert raw[:8] == b"VXDUMP02", raw[:8]
    nv, nb, nf, n_sim, n_vd, n_bd, n_vs, n_bs = struct.unpack_from("<8q", raw, 8)
    off = 8 + 64
    a = np.frombuffer(raw, dtype="<f8", offset=off)
    p = 0
    sim = a[p:p + n_sim]; p += n_sim
    vox = a[p:p + nv * n_vd].reshape(nv, n_vd); p += nv * n_vd
    bond = a[p:p + nb * n_bd].reshape(nb, n_bd); p += nb * n_bd
    fsz = 10 + nv * n_vs + nb * n_bs
    nf_avail = (len(a) - p) // fsz
    nf = min(nf, nf_avail)
    fr = a[p:p + nf * fsz].reshape(nf, fsz); p += nf * fsz
    hdr = fr[:, :10]
    vs = fr[:, 10:10 + nv * n_vs].reshape(nf, nv, n_vs)
    bs = fr[:, 10 + nv * n_vs:].reshape(nf, nb, n_bs)
    return RefDump(nv, nb, sim.copy(), vox.copy(), bond.copy(), hdr.copy(), vs.copy(), bs.copy(), a[p:].copy())


def have_ref(lw: bool = False) -> bool:
    return os.access(REF_DUMP_LW if lw else REF_DUMP, os.X_OK)


def run_ref_dump(vxa_path: str, out_path: str, n_steps: int, every: int = 1, nocol: bool = True, lw: bool = False, suffix: str = "") -> RefDump:
    exe = REF_DUMP_LW if lw else REF_DUMP + suffix   # suffix "_zlib": the build that reads Compression="ZLIB" (oracle/Makefile)
    cmd = [exe, "dump", vxa_path, out_path, str(n_steps), str(every)] + (["nocol"] if nocol else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return load_dump(out_path)


def run_ref_full(vxa_path: str, out_path: str, nocol: bool = True, lw: bool = False, suffix: str = "") -> RefDump:
    exe = REF_DUMP_LW if lw else REF_DUMP + suffix
    cmd = [exe, "full", vxa_path, out_path] + (["nocol"] if nocol else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return load_dump(out_path)


def f32(x: float) -> float:
    return float(np.float32(x))


def desc_from_dump(d: RefDump, lw: bool = False, col_horizon: float = 2.0) -> abi.HostRobot:
    """Flat model from the reference's own imported arrays (dump header), constants by closed form:
    CVX_Voxel::SetMaterial VX_Voxel.cpp:109-125, CVX_Bond::LinkVoxels/UpdateConstants VX_Bond.cpp:65-173."""
    S = d.sim
    sp = abi.SimParams()
    sp.dt = S[0] * S[1]
    sp.bond_damp_z, sp.col_damp_z, sp.slow_damp_z = S[2], S[3], S[4]
    sp.grav_acc = S[6] if S[5] else 0.0
    sp.init_cm_time = S[11]
    sp.temp_period = f32(S[10])
    sp.temp_amplitude = f32(S[9])
    sp.min_scale_fact = S[19]
    sp.max_scale_fact = 1.0 + S[20]
    sp.lattice_dim = S[14]
    if not lw:
        sp.growth_amplitude, sp.stop_condition_value = S[20], S[12]
        sp.midlife_freeze_time = S[22]
    sp.col_horizon = col_horizon   # CollisionHorizon is not reachable through the reference's public API: the caller reads it from the .vxa
    flags = 0
    if S[15]: flags |= abi.VXG_F_COLLISIONS
    if S[7]: flags |= abi.VXG_F_FLOOR
    if S[8]: flags |= abi.VXG_F_TEMPERATURE
    if lw:
        flags |= abi.VXG_F_LW_SCALE
        if S[22]:
            flags |= abi.VXG_F_FLUID
        sp.drag_coeff = S[23]
    sp.flags = flags

    vkeys: Dict[tuple, int] = {}
    vclasses: List[abi.VoxClass] = []
    vcls = np.zeros(d.n_vox, dtype=np.int32)
    for i in range(d.n_vox):
        v = d.vox[i]
        s, mat, E, nu, cte, mass, inertia, fm = v[3], int(v[4]), v[5], v[6], v[7], v[8], v[9], v[10]
        bc = tuple(v[24:36]) if len(v) >= 36 else (0.0,) * 12   # boundary-condition columns (dumps written before them: none)
        dev = tuple(v[36:42]) if (len(v) >= 42 and not lw) else (0.0,) * 6   # development columns
        key = (s, mat, E, nu, cte, mass, v[11], v[12], v[16], v[21], int(v[19]), v[14], v[22]) + bc + dev
        if key not in vkeys:
            c = abi.VoxClass()
            c.mass = mass; c.mass_inv = 1 / mass
            c.inertia = inertia; c.inertia_inv = 1 / inertia
            # _voxcad: SetMaterial caches the damping terms with the palette modulus (VX_Voxel.cpp:116-125) BEFORE an evolved
            # <Stiffness> overrides Vox_E (VX_Sim.cpp:983-988); GetLinearStiffness reads Vox_E live
            Em = v[5] if lw else v[14]   # the land/water fork's SetEMod refreshes the cached terms (LW/VX_Voxel.h:61)
            c.damp_lin = 2 * math.sqrt(mass * Em * s)
            c.damp_ang = 2 * math.sqrt(inertia * Em * s * s * s)
            c.k_floor = v[23]
            c.mu_static, c.mu_dynamic = v[11], v[12]
            c.nom_size = s
            c.cte = cte
            c.temp_amp_damp = v[21]
            c.phase_offset = v[16]
            c.e_mod = E
            c.init_size = v[22]
            c.dof_fixed = int(v[19])
            c.final_size, c.growth_time, c.start_growth_time, c.final_phase_offset, c.final_temp_amp_damp = dev[0:5]
            c.dev_flags = int(dev[5])
            for a in range(3):
                c.ext_force[a], c.ext_torque[a], c.ext_disp[a], c.ext_tdisp[a] = bc[a], bc[3 + a], bc[6 + a], bc[9 + a]
            c.mat_index = mat
            vkeys[key] = len(vclasses)
            vclasses.append(c)
        vcls[i] = vkeys[key]

    bkeys: Dict[tuple, int] = {}
    bclasses: List[abi.BondClass] = []
    bcls = np.zeros(d.n_bond, dtype=np.int32)
    b1 = d.bond[:, 0].astype(np.int32); b2 = d.bond[:, 1].astype(np.int32); bax = d.bond[:, 2].astype(np.int32) - 1  # reference Axis enum: AXIS_X = 1 (VX_Enums.h:66-70)
    for j in range(d.n_bond):
        v1, v2 = d.vox[b1[j]], d.vox[b2[j]]
        ax = int(bax[j])
        nomd = v2[ax] - v1[ax]
        # stress model modulus: _voxcad uses the palette material's Elastic_Mod (VX_Object.cpp:1477);
        # the land/water fork passes the voxel's own (possibly evolved) modulus (LW/VX_Object.cpp:1474)
        # (through a `float voxelEmod` parameter, LW/VX_Object.cpp:1472-1474)
        e1s, e2s = (f32(v1[5]), f32(v2[5])) if lw else (v1[14], v2[14])
        homog = int(v1[4]) == int(v2[4]) and (not lw or v1[5] == v2[5])
        key = (int(v1[4]), int(v2[4]), v1[5], v2[5], v1[6], v2[6], v1[3], v2[3], v1[8], v2[8], round(nomd / v1[3], 9))
        if key not in bkeys:
            # LinkVoxels reads Vox_E (VX_Bond.cpp:83): in _voxcad the internal bonds are linked during Import, before the reset
            # that applies an evolved <Stiffness>, so they keep the palette modulus; the land/water fork overrides Vox_E first
            E1, E2 = (v1[5], v2[5]) if lw else (v1[14], v2[14])
            u1, u2 = v1[6], v2[6]
            E = (E1 * E2 / (E1 + E2)) * 2
            u = 0.0 if (u1 == 0 and u2 == 0) else (u1 * u2 / (u1 + u2)) * 2
            L = (v1[3] + v2[3]) * 0.5
            G = E / (2 * (1 + u))
            A = L * L
            Iy = L * L * L * L / 12
            J = L * L * (L * L + L * L) / 12
            c = abi.BondClass()
            c.a1 = E * A / L
            c.a2 = G * J / L
            c.b1 = 12 * E * Iy / (L * L * L)
            c.b2 = 6 * E * Iy / (L * L)
            c.b3 = 2 * E * Iy / L
            c.len = L
            c.area = L * L
            c.e1, c.e2 = e1s, e2s
            c.nom_dist = nomd
            M1, M2, FM1, FM2, I1, I2 = v1[8], v2[8], v1[10], v2[10], v1[9], v2[9]
            c.c_a1m1, c.c_a1m2 = 2 * math.sqrt(c.a1 * M1), 2 * math.sqrt(c.a1 * M2)
            c.c_a2i1, c.c_a2i2 = 2 * math.sqrt(c.a2 * I1), 2 * math.sqrt(c.a2 * I2)
            c.c_b1m1, c.c_b1m2 = 2 * math.sqrt(c.b1 * M1), 2 * math.sqrt(c.b1 * M2)
            c.c_b2fm1, c.c_b2fm2 = 2 * math.sqrt(c.b2 * FM1), 2 * math.sqrt(c.b2 * FM2)
            c.c_b3i1, c.c_b3i2 = 2 * math.sqrt(c.b3 * I1), 2 * math.sqrt(c.b3 * I2)
            c.homogeneous = 1 if homog else 0
            assert abs(c.a1 - d.bond[j, 3]) <= 1e-12 * abs(c.a1), (c.a1, d.bond[j, 3])
            bkeys[key] = len(bclasses)
            bclasses.append(c)
        bcls[j] = bkeys[key]
    return abi.HostRobot(d.vox[:, 0:3].copy(), vcls, b1, b2, bax, bcls, vclasses, bclasses, sp)
