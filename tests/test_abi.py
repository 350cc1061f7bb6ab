"""The C-ABI library: loads without a GPU, exports every symbol include/*.h declares, the ctypes mirror has
the compiled struct sizes, and the compute entry points fail loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import pytest

from evosoro_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vx[gh]_[a-z_0-9]+)\s*\(", txt)))


@pytest.mark.parametrize("header", ["vxgpu.h", "vxhost.h"])
def test_every_declared_symbol_is_exported(header):
    lib = engine.lib()
    names = _declared(header)
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/{header} but not exported by libvxgpu.so"


def test_struct_sizes_match_ctypes_mirror():
    lib = engine.lib()
    out = (C.c_int * 16)()
    n = lib.vxg_struct_sizes(out, 16)
    want = [abi.VoxClass, abi.BondClass, abi.SimParams, abi.RobotDesc, abi.RobotState, abi.Limits, abi.RobotStats]
    assert n == len(want)
    for i, t in enumerate(want):
        assert out[i] == C.sizeof(t), (t.__name__, out[i], C.sizeof(t))
    assert lib.vxg_abi_version() == 5


def test_headers_compile_as_plain_c(tmp_path):
    """include/*.h are plain C (no C++ / CUDA / torch types in the signatures)."""
    src = tmp_path / "t.c"
    src.write_text('#include "vxgpu.h"\n#include "vxhost.h"\nint main(void){ vxg_limits l; (void)l; return sizeof(vxg_robot_desc) > 0 ? 0 : 1; }\n')
    import subprocess
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_no_cpu_fallback_without_device():
    if engine.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    lim = abi.Limits()
    lim.max_vox, lim.max_vclass, lim.max_bclass = 8, 1, 1
    h = C.c_void_p()
    rc = engine.lib().vxg_create(C.byref(h), 0, 1, C.byref(lim))
    assert rc == -5  # VXG_ERR_NO_DEVICE
    assert b"no CPU path" in engine.lib().vxg_last_error()
    with pytest.raises(engine.EngineError):
        from evosoro_b200 import vxa
        m = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(2, 1)))
        engine.Batch([m.robot])


def test_product_does_not_reference_the_oracle():
    """Nothing under evosoro_b200/ or include/ may import, link or execute oracle/."""
    bad = []
    for base in ("evosoro_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    t = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"vx_oracle|libvxoracle|oracle_lib|oracle/_ref|ref_dump", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_class_api_adapter_has_no_cpu_path(tmp_path):
    """Seam B2 binary (reference CLI + library, Integrate() swapped): without a CUDA device it must stop with the engine's
    message and exit code 3 -- never fall through to the reference's CPU loop it is linked with."""
    import os
    import subprocess
    from evosoro_b200 import engine, vxa
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "voxelyze_gpu_adapter")
    if not os.access(exe, os.X_OK):
        pytest.skip("oracle/_ref/voxelyze_gpu_adapter not built")
    if engine.device_count() > 0:
        pytest.skip("a GPU is visible")
    p = tmp_path / "r.vxa"
    p.write_text(vxa.vxa_text(vxa.dense_lattice(2, 1), sim=vxa.SimParams(init_cm_time=0.001, stop_condition_value=0.01, fitness_file=str(tmp_path / "f.xml"))))
    r = subprocess.run([exe, "-f", str(p)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr
    assert not (tmp_path / "f.xml").exists()
