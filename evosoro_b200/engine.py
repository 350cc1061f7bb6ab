"""ctypes binding of ``libvxgpu.so`` (``include/vxgpu.h``) -- the Python host side of the engine.

Thin by design: every call maps 1:1 onto a C-ABI entry point.  There is no CPU compute path: if the
CUDA library is missing or no GPU is visible the calls raise ``EngineError`` (loudly) rather than
falling back to anything.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# VXG_LIB points the binding at another build of the same library (kernel tuning builds, `make OUT=../lib_x EXTRA=-D...`)
LIB_PATH = os.environ.get("VXG_LIB") or os.path.join(_HERE, "lib", "libvxgpu.so")
CSRC = os.path.join(_HERE, "csrc")
_LIB = None


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vxgpu error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile libvxgpu.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(_HERE), "include", "vxgpu.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", CSRC, "-j8"] + (["-B"] if force else []), check=True)
    return LIB_PATH


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(-100, f"{LIB_PATH} is missing: run evosoro_b200.engine.build() / __graft_entry__.build(); "
                                    "the engine has no CPU path")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.vxg_abi_version.restype = C.c_int
        L.vxg_device_count.restype = C.c_int
        L.vxg_last_error.restype = C.c_char_p
        L.vxg_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(abi.Limits)]
        L.vxg_destroy.argtypes = [vp]
        L.vxg_destroy.restype = None
        L.vxg_upload.argtypes = [vp, C.c_int, C.POINTER(abi.RobotDesc)]
        L.vxg_set_step_budget.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.vxg_step.argtypes = [vp, C.c_int]
        L.vxg_sync.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.vxg_last_step_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.vxg_launch_count.argtypes = [vp]
        L.vxg_launch_count.restype = C.c_int64
        L.vxg_download_state.argtypes = [vp, C.c_int, C.POINTER(abi.RobotState)]
        L.vxg_upload_state.argtypes = [vp, C.c_int, C.POINTER(abi.RobotState), C.c_double, C.c_int64]
        L.vxg_reduce_stats.argtypes = [vp, C.POINTER(abi.RobotStats)]
        L.vxg_enable_bond_record.argtypes = [vp, C.c_int]
        L.vxg_kernel_config.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.vxg_struct_sizes.argtypes = [C.POINTER(C.c_int), C.c_int]
        L.vxg_set_collision_pairs.argtypes = [vp, C.c_int, C.c_int, abi.c_int32_p, abi.c_int32_p, abi.c_double_p]
        L.vxg_run_batch_host.argtypes = [vp, C.POINTER(abi.RobotDesc), C.c_int, C.c_int, abi.c_double_p]
        L.vxg_mesh_vertex_count.argtypes = [C.POINTER(abi.RobotDesc)]
        L.vxg_upload_bytes.argtypes = [vp]
        L.vxg_upload_bytes.restype = C.c_int64
        L.vxg_mark.argtypes = [vp, C.c_int]
        L.vxg_mark_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
        # host side (include/vxhost.h)
        L.vxh_last_error.restype = C.c_char_p
        L.vxh_load_vxa_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
        L.vxh_load_vxa_string.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
        L.vxh_free.argtypes = [vp]
        L.vxh_free.restype = None
        L.vxh_desc.argtypes = [vp]
        L.vxh_desc.restype = C.POINTER(abi.RobotDesc)
        L.vxh_get_info.argtypes = [vp, C.POINTER(abi.VxaInfo)]
        L.vxh_cell_index.argtypes = [vp, abi.c_int32_p]
        L.vxh_step_plan.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.vxh_step_plan_ex.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.vxh_format_result.argtypes = [vp, C.POINTER(abi.RobotStats), C.c_char_p, C.c_int, C.c_char_p]
        L.vxg_set_eol_step.argtypes = [vp, C.POINTER(C.c_int64)]
        i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        L.vxg_pop_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(abi.RobotDesc), C.c_int, C.POINTER(abi.PopOptions)]
        L.vxg_pop_destroy.argtypes = [vp]
        L.vxg_pop_destroy.restype = None
        L.vxg_pop_upload.argtypes = [vp, C.POINTER(abi.RobotDesc)]
        L.vxg_pop_set_plan.argtypes = [vp, i64p, i64p, i64p]
        L.vxg_pop_step.argtypes = [vp, C.c_int64]
        L.vxg_pop_run.argtypes = [vp]
        L.vxg_pop_sync.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.vxg_pop_reduce_stats.argtypes = [vp, C.POINTER(abi.RobotStats)]
        L.vxg_pop_download_state.argtypes = [vp, C.c_int, C.POINTER(abi.RobotState)]
        L.vxg_pop_run_host.argtypes = [vp, C.POINTER(abi.RobotDesc), C.c_int, abi.c_double_p, C.c_int]
        L.vxg_pop_n_batches.argtypes = [vp]
        L.vxg_pop_batch_info.argtypes = [vp, C.c_int, i32p, i32p, i32p, i32p]
        L.vxg_pop_locate.argtypes = [vp, C.c_int, i32p, i32p]
        L.vxg_pop_mark.argtypes = [vp, C.c_int]
        L.vxg_pop_mark_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
        for f in ("vxg_pop_launch_count", "vxg_pop_upload_bytes", "vxg_pop_download_bytes"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_int64
        _LIB = L
    return _LIB


def _ck(rc: int):
    if rc != 0:
        raise EngineError(rc, lib().vxg_last_error().decode(errors="replace"))


def device_count() -> int:
    return lib().vxg_device_count()


def set_pooling(on: bool) -> None:
    """``vxg_set_pooling``: destroyed batches leave their device / pinned blocks on a free list for the next ``Batch`` (a resident
    process that builds a population per GA generation); ``False`` releases the list."""
    _ck(lib().vxg_set_pooling(1 if on else 0))


VARIANT_VOXCAD = 0
VARIANT_LAND_WATER = 1


class Model:
    """An imported ``.vxa`` (``vxh_model``): the flat descriptor plus the CLI's step plan.

    Mirrors ``CVX_Sim::LoadVXAFile`` + ``Import`` (VX_Sim.cpp:159-165, :488-717); raises ``EngineError``
    with a positive code when the file enables a feature the kernels do not implement."""

    def __init__(self, path: Optional[str] = None, text: Optional[str] = None, variant: int = VARIANT_VOXCAD,
                 ignore_collisions: bool = False):
        self.h = C.c_void_p()
        L = lib()
        if text is not None:
            rc = L.vxh_load_vxa_string(text.encode("latin-1"), variant, int(ignore_collisions), C.byref(self.h))
        else:
            rc = L.vxh_load_vxa_file(os.fsencode(path), variant, int(ignore_collisions), C.byref(self.h))
        if rc != 0:
            raise EngineError(rc, L.vxh_last_error().decode(errors="replace"))
        self.info = abi.VxaInfo()
        L.vxh_get_info(self.h, C.byref(self.info))
        self.robot = abi.HostRobot.from_desc(L.vxh_desc(self.h).contents)

    def step_plan(self):
        """(total_steps, latch_step, final_time) of the reference CLI loop for this file."""
        t, l, ft = C.c_int64(), C.c_int64(), C.c_double()
        rc = lib().vxh_step_plan(self.h, C.byref(t), C.byref(l), C.byref(ft))
        if rc != 0:
            raise EngineError(rc, lib().vxh_last_error().decode(errors="replace"))
        return t.value, l.value, ft.value

    def step_plan_ex(self):
        """(total_steps, latch_step, eol_step, final_time)"""
        t, l, e, ft = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        rc = lib().vxh_step_plan_ex(self.h, C.byref(t), C.byref(l), C.byref(e), C.byref(ft))
        if rc != 0:
            raise EngineError(rc, lib().vxh_last_error().decode(errors="replace"))
        return t.value, l.value, e.value, ft.value

    def format_result(self, stats: "abi.RobotStats", path: Optional[str] = None) -> str:
        """The result XML of CVX_SimGA::WriteResultFile for this robot's reductions."""
        buf = C.create_string_buffer(4096)
        n = lib().vxh_format_result(self.h, C.byref(stats), buf, 4096, os.fsencode(path) if path else None)
        if n < 0:
            raise EngineError(n, lib().vxh_last_error().decode(errors="replace"))
        return buf.value.decode()

    def cell_index(self) -> np.ndarray:
        out = np.zeros(self.info.n_vox, dtype=np.int32)
        lib().vxh_cell_index(self.h, out.ctypes.data_as(abi.c_int32_p))
        return out

    def close(self):
        if self.h:
            lib().vxh_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """A population of robots resident on one GPU (``vxg_batch``)."""

    def __init__(self, robots: Sequence[abi.HostRobot], device: int = 0, precision: int = abi.VXG_PREC_F32,
                 max_vox: Optional[int] = None, max_vclass: Optional[int] = None, max_bclass: Optional[int] = None,
                 upload: bool = True, max_col_pairs: Optional[int] = None):
        self.robots: List[abi.HostRobot] = list(robots)
        if not self.robots:
            raise ValueError("empty batch")
        lim = abi.Limits()
        lim.max_vox = max_vox or max(r.n_vox for r in self.robots)
        lim.max_vclass = max_vclass or max(len(r.vclasses) for r in self.robots)
        lim.max_bclass = max_bclass or max(1, max(len(r.bclasses) for r in self.robots))
        lim.precision = precision
        # land/water fluid drag needs the surface-mesh arrays: size them for the largest mesh of the batch
        if any(r.sim.flags & abi.VXG_F_FLUID for r in self.robots):
            lim.max_mesh_verts = max(lib().vxg_mesh_vertex_count(C.byref(r.desc())) for r in self.robots)
        # self-collision: room for the contact-bond list (a handful in practice; 4 bytes each in shared memory)
        if max_col_pairs is not None:
            lim.max_col_pairs = max_col_pairs
        elif any(r.sim.flags & abi.VXG_F_COLLISIONS for r in self.robots):
            lim.max_col_pairs = 512
        self.limits = lim
        self.h = C.c_void_p()
        _ck(lib().vxg_create(C.byref(self.h), device, len(self.robots), C.byref(lim)))
        self._descs = (abi.RobotDesc * len(self.robots))(*[r.desc() for r in self.robots])
        if upload:
            for i in range(len(self.robots)):
                _ck(lib().vxg_upload(self.h, i, C.byref(self._descs[i])))

    def reupload(self, robot: int, desc: Optional[abi.HostRobot] = None):
        """``vxg_upload`` of one slot: the robot (or a new one of the same limits) restarts from its reset state; the other
        slots keep the state the device computed."""
        if desc is not None:
            self.robots[robot] = desc
            self._descs[robot] = desc.desc()
        _ck(lib().vxg_upload(self.h, robot, C.byref(self._descs[robot])))

    @property
    def n_robots(self) -> int:
        return len(self.robots)

    def set_step_budget(self, total_steps=None, latch_step=None):
        def arr(x):
            if x is None:
                return None, None
            a = np.ascontiguousarray(x, dtype=np.int64)
            assert a.shape == (self.n_robots,)
            return a, a.ctypes.data_as(C.POINTER(C.c_int64))
        ta, tp = arr(total_steps)
        la, lp = arr(latch_step)
        _ck(lib().vxg_set_step_budget(self.h, tp, lp))

    def set_eol_step(self, eol_step):
        a = np.ascontiguousarray(eol_step, dtype=np.int64)
        assert a.shape == (self.n_robots,)
        _ck(lib().vxg_set_eol_step(self.h, a.ctypes.data_as(C.POINTER(C.c_int64))))

    def step(self, n: int):
        _ck(lib().vxg_step(self.h, int(n)))

    def sync(self) -> np.ndarray:
        st = np.zeros(self.n_robots, dtype=np.uint32)
        _ck(lib().vxg_sync(self.h, st.ctypes.data_as(C.POINTER(C.c_uint32))))
        return st

    def last_step_ms(self) -> float:
        ms = C.c_float()
        _ck(lib().vxg_last_step_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        return int(lib().vxg_launch_count(self.h))

    def mark(self, slot: int):
        _ck(lib().vxg_mark(self.h, slot))

    def mark_elapsed_ms(self, slot0: int, slot1: int) -> float:
        ms = C.c_float()
        _ck(lib().vxg_mark_elapsed_ms(self.h, slot0, slot1, C.byref(ms)))
        return float(ms.value)

    def enable_bond_record(self, on: bool = True):
        _ck(lib().vxg_enable_bond_record(self.h, 1 if on else 0))

    def kernel_config(self) -> dict:
        t, v, s = C.c_int(), C.c_int(), C.c_int()
        sb, bb, db = C.c_int64(), C.c_int64(), C.c_int64()
        _ck(lib().vxg_kernel_config(self.h, C.byref(t), C.byref(v), C.byref(s), C.byref(sb), C.byref(bb), C.byref(db)))
        return {"threads": t.value, "voxels_per_thread": v.value, "smem_resident": bool(s.value),
                "smem_bytes": sb.value, "blob_bytes": bb.value, "dyn_bytes": db.value}

    def download(self, robot: int, bonds: bool = False) -> abi.StateBuffers:
        r = self.robots[robot]
        buf = abi.StateBuffers(r.n_vox, r.n_bond)
        v = buf.view()
        if not bonds:
            v.bond_force1 = v.bond_force2 = v.bond_moment1 = v.bond_moment2 = v.bond_strain = None
        _ck(lib().vxg_download_state(self.h, robot, C.byref(v)))
        return buf

    def upload_state(self, robot: int, buf: abi.StateBuffers, cur_time: float, cur_step: int):
        v = buf.view()
        _ck(lib().vxg_upload_state(self.h, robot, C.byref(v), float(cur_time), int(cur_step)))

    def stats(self):
        out = (abi.RobotStats * self.n_robots)()
        _ck(lib().vxg_reduce_stats(self.h, out))
        return out

    def run_batch_host(self, n_steps: int, pos_out: Optional[np.ndarray] = None) -> np.ndarray:
        """End-to-end call with host buffers: upload every descriptor, step, bring positions back."""
        if pos_out is None:
            pos_out = np.zeros((self.n_robots, self.limits.max_vox, 3))
        _ck(lib().vxg_run_batch_host(self.h, self._descs, self.n_robots, int(n_steps), pos_out.ctypes.data_as(abi.c_double_p)))
        return pos_out

    def h2d_bytes_per_upload(self) -> int:
        return int(lib().vxg_upload_bytes(self.h))

    def close(self):
        if self.h:
            lib().vxg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Population:
    """A ragged population on one GPU (``vxg_population``): robots of any mix of sizes and features, bucketed by the library
    into homogeneous batches that run concurrently.  Robot indices are positions in ``robots``."""

    def __init__(self, robots: Sequence[abi.HostRobot], device: int = 0, precision: int = abi.VXG_PREC_F32, chunk_robots: int = 0,
                 want_mesh: bool = False, col_pairs_per_vox: int = 0):
        self.robots: List[abi.HostRobot] = list(robots)
        if not self.robots:
            raise ValueError("empty population")
        self._descs = (abi.RobotDesc * len(self.robots))(*[r.desc() for r in self.robots])
        opt = abi.PopOptions()
        opt.precision, opt.chunk_robots, opt.want_mesh, opt.col_pairs_per_vox = precision, chunk_robots, int(want_mesh), col_pairs_per_vox
        self.h = C.c_void_p()
        _ck(lib().vxg_pop_create(C.byref(self.h), device, self._descs, len(self.robots), C.byref(opt)))
        self.max_vox = max(r.n_vox for r in self.robots)

    @property
    def n_robots(self) -> int:
        return len(self.robots)

    def upload(self):
        _ck(lib().vxg_pop_upload(self.h, self._descs))

    def set_plan(self, total_steps=None, latch_step=None, eol_step=None):
        keep = []

        def arr(x):
            if x is None:
                return None
            a = np.ascontiguousarray(x, dtype=np.int64)
            assert a.shape == (self.n_robots,)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int64))
        _ck(lib().vxg_pop_set_plan(self.h, arr(total_steps), arr(latch_step), arr(eol_step)))

    def step(self, n: int):
        _ck(lib().vxg_pop_step(self.h, int(n)))

    def run(self):
        _ck(lib().vxg_pop_run(self.h))

    def sync(self) -> np.ndarray:
        st = np.zeros(self.n_robots, dtype=np.uint32)
        _ck(lib().vxg_pop_sync(self.h, st.ctypes.data_as(C.POINTER(C.c_uint32))))
        return st

    def stats(self):
        out = (abi.RobotStats * self.n_robots)()
        _ck(lib().vxg_pop_reduce_stats(self.h, out))
        return out

    def download(self, robot: int) -> abi.StateBuffers:
        r = self.robots[robot]
        buf = abi.StateBuffers(r.n_vox, r.n_bond)
        v = buf.view()
        v.bond_force1 = v.bond_force2 = v.bond_moment1 = v.bond_moment2 = v.bond_strain = None
        _ck(lib().vxg_pop_download_state(self.h, robot, C.byref(v)))
        return buf

    def run_host(self, n_steps: int, pos_out: Optional[np.ndarray] = None) -> np.ndarray:
        """End-to-end call with host buffers: flatten + upload every descriptor, step, positions back (pipelined per batch)."""
        if pos_out is None:
            pos_out = np.zeros((self.n_robots, self.max_vox, 3))
        _ck(lib().vxg_pop_run_host(self.h, self._descs, int(n_steps), pos_out.ctypes.data_as(abi.c_double_p), pos_out.shape[1]))
        return pos_out

    def batches(self):
        """[(n_robots, stride, threads, family bits)] of the homogeneous batches the population was bucketed into."""
        out = []
        for i in range(lib().vxg_pop_n_batches(self.h)):
            n, s, t, f = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            _ck(lib().vxg_pop_batch_info(self.h, i, C.byref(n), C.byref(s), C.byref(t), C.byref(f)))
            out.append((n.value, s.value, t.value, f.value))
        return out

    def locate(self, robot: int):
        b, s = C.c_int32(), C.c_int32()
        _ck(lib().vxg_pop_locate(self.h, robot, C.byref(b), C.byref(s)))
        return b.value, s.value

    def launch_count(self) -> int:
        return int(lib().vxg_pop_launch_count(self.h))

    def mark(self, slot: int):
        _ck(lib().vxg_pop_mark(self.h, slot))

    def mark_elapsed_ms(self, slot0: int, slot1: int) -> float:
        ms = C.c_float()
        _ck(lib().vxg_pop_mark_elapsed_ms(self.h, slot0, slot1, C.byref(ms)))
        return float(ms.value)

    def h2d_bytes(self) -> int:
        return int(lib().vxg_pop_upload_bytes(self.h))

    def d2h_bytes(self) -> int:
        return int(lib().vxg_pop_download_bytes(self.h))

    def close(self):
        if self.h:
            lib().vxg_pop_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
