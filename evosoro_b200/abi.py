"""ctypes mirror of ``include/vxgpu.h`` (the C-ABI structs) plus a numpy-backed robot descriptor.

The layout here must match the header field for field; ``tests/test_abi.py`` checks the sizes against
the values the compiled library reports.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List

import numpy as np

VXG_F_FLOOR = 0x0001
VXG_F_TEMPERATURE = 0x0002
VXG_F_LW_SCALE = 0x0004
VXG_F_STICKY_FLOOR = 0x0008
VXG_F_FLUID = 0x0010
VXG_F_COLLISIONS = 0x0020

VXG_PREC_F32 = 0
VXG_PREC_F64 = 1

VXG_ST_DIVERGED = 0x1
VXG_ST_FINISHED = 0x2
VXG_ST_COL_OVERFLOW = 0x4

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class VoxClass(C.Structure):
    _fields_ = [
        ("mass", C.c_double), ("mass_inv", C.c_double), ("inertia", C.c_double), ("inertia_inv", C.c_double),
        ("damp_lin", C.c_double), ("damp_ang", C.c_double), ("k_floor", C.c_double),
        ("mu_static", C.c_double), ("mu_dynamic", C.c_double), ("nom_size", C.c_double),
        ("cte", C.c_double), ("temp_amp_damp", C.c_double), ("phase_offset", C.c_double), ("e_mod", C.c_double), ("init_size", C.c_double),
        ("ext_force", C.c_double * 3), ("ext_torque", C.c_double * 3),
        ("ext_disp", C.c_double * 3), ("ext_tdisp", C.c_double * 3),
        ("dof_fixed", C.c_int32), ("mat_index", C.c_int32),
        ("final_size", C.c_double), ("growth_time", C.c_double), ("start_growth_time", C.c_double),
        ("final_phase_offset", C.c_double), ("final_temp_amp_damp", C.c_double),
        ("dev_flags", C.c_int32), ("reserved", C.c_int32),
    ]


class BondClass(C.Structure):
    _fields_ = [
        ("a1", C.c_double), ("a2", C.c_double), ("b1", C.c_double), ("b2", C.c_double), ("b3", C.c_double),
        ("len", C.c_double), ("area", C.c_double), ("e1", C.c_double), ("e2", C.c_double), ("nom_dist", C.c_double),
        ("c_a1m1", C.c_double), ("c_a1m2", C.c_double), ("c_a2i1", C.c_double), ("c_a2i2", C.c_double),
        ("c_b1m1", C.c_double), ("c_b1m2", C.c_double), ("c_b2fm1", C.c_double), ("c_b2fm2", C.c_double),
        ("c_b3i1", C.c_double), ("c_b3i2", C.c_double),
        ("homogeneous", C.c_int32), ("reserved", C.c_int32),
    ]


class SimParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("bond_damp_z", C.c_double), ("col_damp_z", C.c_double), ("slow_damp_z", C.c_double),
        ("grav_acc", C.c_double), ("init_cm_time", C.c_double), ("temp_period", C.c_double),
        ("temp_amplitude", C.c_double), ("min_scale_fact", C.c_double), ("max_scale_fact", C.c_double),
        ("lattice_dim", C.c_double), ("drag_coeff", C.c_double), ("col_horizon", C.c_double),
        ("flags", C.c_uint32), ("reserved", C.c_uint32),
        ("growth_amplitude", C.c_double), ("midlife_freeze_time", C.c_double), ("stop_condition_value", C.c_double),
    ]


class RobotDesc(C.Structure):
    _fields_ = [
        ("n_vox", C.c_int32), ("n_bond", C.c_int32), ("n_vclass", C.c_int32), ("n_bclass", C.c_int32),
        ("nom_pos", c_double_p), ("vox_class", c_int32_p),
        ("bond_v1", c_int32_p), ("bond_v2", c_int32_p), ("bond_axis", c_int32_p), ("bond_class", c_int32_p),
        ("vclasses", C.POINTER(VoxClass)), ("bclasses", C.POINTER(BondClass)),
        ("sim", SimParams),
    ]


class RobotState(C.Structure):
    _fields_ = [
        ("pos", c_double_p), ("quat", c_double_p), ("lin_mom", c_double_p), ("ang_mom", c_double_p),
        ("scale", c_double_p), ("static_fric", c_int32_p),
        ("bond_force1", c_double_p), ("bond_force2", c_double_p), ("bond_moment1", c_double_p),
        ("bond_moment2", c_double_p), ("bond_strain", c_double_p), ("bond_small_angle", c_int32_p),
        ("vox_strain", c_double_p),
    ]


class Limits(C.Structure):
    _fields_ = [
        ("max_vox", C.c_int32), ("max_vclass", C.c_int32), ("max_bclass", C.c_int32), ("precision", C.c_int32),
        ("max_col_pairs", C.c_int32), ("max_mesh_verts", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


class PopOptions(C.Structure):
    _fields_ = [("precision", C.c_int32), ("chunk_robots", C.c_int32), ("want_mesh", C.c_int32), ("col_pairs_per_vox", C.c_int32),
                ("reserved", C.c_int32 * 4)]


VXG_FAM_COMPACT, VXG_FAM_COLLIDE, VXG_FAM_MESH = 1, 2, 4


class RobotStats(C.Structure):
    _fields_ = [
        ("com", C.c_double * 3), ("ini_com", C.c_double * 3), ("max_vel", C.c_double),
        ("min_pos", C.c_double * 3), ("max_pos", C.c_double * 3), ("time", C.c_double),
        ("steps", C.c_int64), ("n_touching_floor", C.c_int32), ("ini_com_latched", C.c_int32),
        ("anterior_dist", C.c_double), ("posterior_dist", C.c_double), ("anterior_y", C.c_double),
        ("posterior_y", C.c_double), ("eol_posterior_y", C.c_double),
        ("n_nonfeet_touching_floor", C.c_int32), ("reserved", C.c_int32),
        ("mesh_volume", C.c_double),
    ]


class VxaInfo(C.Structure):
    """vxh_info (include/vxhost.h)"""
    _fields_ = [
        ("n_vox", C.c_int32), ("n_bond", C.c_int32),
        ("x_voxels", C.c_int32), ("y_voxels", C.c_int32), ("z_voxels", C.c_int32),
        ("stop_condition_type", C.c_int32), ("self_col_enabled", C.c_int32), ("write_fitness_file", C.c_int32),
        ("stop_condition_value", C.c_double), ("init_cm_time", C.c_double), ("afterlife_time", C.c_double),
        ("dt_frac", C.c_double), ("optimal_dt", C.c_double), ("lattice_dim", C.c_double),
        ("col_horizon", C.c_double),
        ("fitness_file_name", C.c_char * 512),
    ]


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(c_int32_p)


@dataclass
class HostRobot:
    """A robot descriptor whose arrays are owned by numpy (keeps them alive for the ctypes view)."""
    nom_pos: np.ndarray      # [n_vox, 3] float64
    vox_class: np.ndarray    # [n_vox] int32
    bond_v1: np.ndarray
    bond_v2: np.ndarray
    bond_axis: np.ndarray
    bond_class: np.ndarray
    vclasses: List[VoxClass]
    bclasses: List[BondClass]
    sim: SimParams

    def __post_init__(self):
        self.nom_pos = np.ascontiguousarray(self.nom_pos, dtype=np.float64).reshape(-1, 3)
        for k in ("vox_class", "bond_v1", "bond_v2", "bond_axis", "bond_class"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.int32))
        self._vc = (VoxClass * max(1, len(self.vclasses)))(*self.vclasses)
        self._bc = (BondClass * max(1, len(self.bclasses)))(*self.bclasses)

    @property
    def n_vox(self) -> int:
        return int(self.nom_pos.shape[0])

    @property
    def n_bond(self) -> int:
        return int(self.bond_v1.shape[0])

    def desc(self) -> RobotDesc:
        d = RobotDesc()
        d.n_vox, d.n_bond = self.n_vox, self.n_bond
        d.n_vclass, d.n_bclass = len(self.vclasses), len(self.bclasses)
        d.nom_pos = _dp(self.nom_pos)
        d.vox_class = _ip(self.vox_class)
        d.bond_v1, d.bond_v2 = _ip(self.bond_v1), _ip(self.bond_v2)
        d.bond_axis, d.bond_class = _ip(self.bond_axis), _ip(self.bond_class)
        d.vclasses = C.cast(self._vc, C.POINTER(VoxClass))
        d.bclasses = C.cast(self._bc, C.POINTER(BondClass))
        d.sim = self.sim
        return d

    @classmethod
    def from_desc(cls, d: RobotDesc) -> "HostRobot":
        """Deep-copy a descriptor the C side owns (e.g. one returned by the importer)."""
        nv, nb = d.n_vox, d.n_bond

        def arr(p, n, dt):
            return np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n else np.zeros(0, dt)

        vcs, bcs = [], []
        for i in range(d.n_vclass):
            v = VoxClass()
            C.memmove(C.byref(v), C.byref(d.vclasses[i]), C.sizeof(VoxClass))
            vcs.append(v)
        for i in range(d.n_bclass):
            b = BondClass()
            C.memmove(C.byref(b), C.byref(d.bclasses[i]), C.sizeof(BondClass))
            bcs.append(b)
        sim = SimParams()
        C.memmove(C.byref(sim), C.byref(d.sim), C.sizeof(SimParams))
        return cls(arr(d.nom_pos, nv * 3, np.float64).reshape(-1, 3), arr(d.vox_class, nv, np.int32),
                   arr(d.bond_v1, nb, np.int32), arr(d.bond_v2, nb, np.int32), arr(d.bond_axis, nb, np.int32),
                   arr(d.bond_class, nb, np.int32), vcs, bcs, sim)


class StateBuffers:
    """numpy buffers + the RobotState view over them."""

    def __init__(self, n_vox: int, n_bond: int):
        nb = max(1, n_bond)
        self.pos = np.zeros((n_vox, 3))
        self.quat = np.zeros((n_vox, 4))
        self.lin_mom = np.zeros((n_vox, 3))
        self.ang_mom = np.zeros((n_vox, 3))
        self.scale = np.zeros(n_vox)
        self.static_fric = np.zeros(n_vox, dtype=np.int32)
        self.bond_force1 = np.zeros((nb, 3))
        self.bond_force2 = np.zeros((nb, 3))
        self.bond_moment1 = np.zeros((nb, 3))
        self.bond_moment2 = np.zeros((nb, 3))
        self.bond_strain = np.zeros(nb)
        self.bond_small_angle = np.zeros(nb, dtype=np.int32)
        self.n_bond = n_bond

    def view(self) -> RobotState:
        s = RobotState()
        s.pos, s.quat, s.lin_mom, s.ang_mom = _dp(self.pos), _dp(self.quat), _dp(self.lin_mom), _dp(self.ang_mom)
        s.scale, s.static_fric = _dp(self.scale), _ip(self.static_fric)
        s.bond_force1, s.bond_force2 = _dp(self.bond_force1), _dp(self.bond_force2)
        s.bond_moment1, s.bond_moment2 = _dp(self.bond_moment1), _dp(self.bond_moment2)
        s.bond_strain, s.bond_small_angle = _dp(self.bond_strain), _ip(self.bond_small_angle)
        return s
