"""ctypes binding of oracle/libvxoracle.so (the plain-C restatement).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from evosoro_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "libvxoracle.so")
        srcs = [os.path.join(ORACLE_DIR, "vx_oracle.c"), os.path.join(ORACLE_DIR, "vx_oracle.h"), os.path.join(ROOT, "include", "vxgpu.h")]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
            subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.vxo_create.restype = C.c_void_p
        L.vxo_create.argtypes = [C.POINTER(abi.RobotDesc)]
        L.vxo_destroy.argtypes = [C.c_void_p]
        L.vxo_step.restype = C.c_int64
        L.vxo_step.argtypes = [C.c_void_p, C.c_int64]
        L.vxo_diverged.argtypes = [C.c_void_p]
        L.vxo_time.restype = C.c_double
        L.vxo_time.argtypes = [C.c_void_p]
        L.vxo_step_count.restype = C.c_int64
        L.vxo_step_count.argtypes = [C.c_void_p]
        L.vxo_get_state.argtypes = [C.c_void_p, C.POINTER(abi.RobotState)]
        L.vxo_get_com.argtypes = [C.c_void_p, abi.c_double_p]
        L.vxo_num_touching_floor.argtypes = [C.c_void_p]
        L.vxo_max_vel.restype = C.c_double
        L.vxo_max_vel.argtypes = [C.c_void_p]
        L.vxo_set_collision_pairs.argtypes = [C.c_void_p, C.c_int, abi.c_int32_p, abi.c_int32_p, abi.c_double_p]
        _LIB = L
    return _LIB


class OracleSim:
    def __init__(self, robot: abi.HostRobot):
        self.robot = robot
        self._d = robot.desc()
        self.h = lib().vxo_create(C.byref(self._d))
        self.buf = abi.StateBuffers(robot.n_vox, robot.n_bond)

    def step(self, n: int = 1) -> int:
        return lib().vxo_step(self.h, n)

    def state(self) -> abi.StateBuffers:
        v = self.buf.view()
        lib().vxo_get_state(self.h, C.byref(v))
        return self.buf

    def com(self) -> np.ndarray:
        out = np.zeros(3)
        lib().vxo_get_com(self.h, out.ctypes.data_as(abi.c_double_p))
        return out

    @property
    def time(self) -> float:
        return lib().vxo_time(self.h)

    @property
    def diverged(self) -> bool:
        return bool(lib().vxo_diverged(self.h))

    def __del__(self):
        try:
            if self.h:
                lib().vxo_destroy(self.h)
                self.h = None
        except Exception:
            pass
