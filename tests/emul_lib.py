"""ctypes binding of tests/_build/libvxemul.so: the CPU emulation of the device arithmetic
(tests/emul/vx_emul.cpp).  Test infrastructure only -- lets the CPU test-suite check the kernels'
fp32/fp64 "delta formulation" against the oracle without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from evosoro_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "_build", "libvxemul.so")
        srcs = [os.path.join(HERE, "emul", "vx_emul.cpp")] + [
            os.path.join(ROOT, "evosoro_b200", "csrc", f) for f in ("vx_math.cuh", "vx_robot.cuh", "vx_pack.h")] + [
            os.path.join(ROOT, "include", "vxgpu.h")]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, srcs[0]], check=True)
        L = C.CDLL(so)
        L.vxe_create.restype = C.c_void_p
        L.vxe_create.argtypes = [C.POINTER(abi.RobotDesc), C.c_int]
        L.vxe_destroy.argtypes = [C.c_void_p]
        L.vxe_step.restype = C.c_int64
        L.vxe_step.argtypes = [C.c_void_p, C.c_int64]
        L.vxe_get_state.argtypes = [C.c_void_p, C.POINTER(abi.RobotState)]
        L.vxe_time.restype = C.c_double
        L.vxe_time.argtypes = [C.c_void_p]
        L.vxe_diverged.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class EmulSim:
    def __init__(self, robot: abi.HostRobot, prec: int = abi.VXG_PREC_F32):
        self.robot = robot
        self._d = robot.desc()
        self.h = lib().vxe_create(C.byref(self._d), prec)
        self.buf = abi.StateBuffers(robot.n_vox, robot.n_bond)

    def step(self, n: int = 1) -> int:
        return lib().vxe_step(self.h, n)

    def state(self) -> abi.StateBuffers:
        v = self.buf.view()
        lib().vxe_get_state(self.h, C.byref(v))
        return self.buf

    @property
    def time(self) -> float:
        return lib().vxe_time(self.h)

    @property
    def diverged(self) -> bool:
        return bool(lib().vxe_diverged(self.h))

    def __del__(self):
        try:
            if self.h:
                lib().vxe_destroy(self.h)
                self.h = None
        except Exception:
            pass
