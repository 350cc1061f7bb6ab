"""The DEVICE arithmetic (evosoro_b200/csrc/vx_math.cuh + vx_robot.cuh, compiled as host C++ by
tests/emul) against the reference's golden trajectories, on CPU.  This is the kernels' own code path
(fp32 production instantiation and the fp64 one), so formulation errors show up without a GPU.
Tolerances: fp64 <= 1e-9 lattice pitch (different but algebraically equal formulation); fp32 <= 2e-5 pitch
on positions over the fixture horizon (<= 1500 steps).

Bond forces: the reference's ToRotationVector (Vec3D.h:270-285) evaluates 2 sqrt((2-2w)/(1-w^2)) and
returns 0 when w rounds to 1, so for rotations below ~1e-7 rad its bending/torsion terms are rounding
noise (relative error ~ 1e-16/(1-w)); the kernels use the cancellation-free form.  While the lattice is
still unrotated (first few dozen steps) the two therefore differ by up to ~1e-3 of the largest bond
force -- through the damping term that is 1e-12 m on positions, which is what parity is stated on."""
import numpy as np
import pytest

import emul_lib
import golden_lib
from evosoro_b200 import abi

LAT = 0.01
CASES = [n for n in golden_lib.names() if not n.endswith("_full")]
FLUID = lambda n: golden_lib.flags(n)[0]
COLLIDE = lambda n: golden_lib.flags(n)[1]


@pytest.mark.parametrize("name", CASES)
def test_emulated_f64_kernel_arithmetic(name):
    """For the lw_fluid_* fixtures this is the logic-parity pin of the fluid-drag phase (mesh vertices, facet
    normals/areas, drag law): the kernels' own mesh_vertex / voxel_drag functions in fp64 vs the reference."""
    g = golden_lib.Golden(name)
    d = g.dump
    LAT = float(d.sim[14])   # the fixture's own lattice pitch
    # 4 = fp64 with the land/water drag phase, 5 = fp64 with the self-collision list
    e = emul_lib.EmulSim(g.robot(), (7 if COLLIDE(name) else 4) if FLUID(name) else (5 if COLLIDE(name) else abi.VXG_PREC_F64))   # 7 = both
    ptol = 1e-7 if COLLIDE(name) else 1e-9   # contact (overlap * stiffness) amplifies rounding differences
    for f in range(1, d.vstate.shape[0]):
        assert e.step(g.every) == g.every
        s = e.state()
        assert np.abs(s.pos - d.pos[f]).max() <= ptol * LAT, (name, f)
        assert np.abs(s.quat - d.quat[f]).max() <= (1e-6 if COLLIDE(name) else 1e-8)
        assert np.abs(s.scale - d.scale[f]).max() <= 1e-15
        assert np.array_equal(s.static_fric, d.sfric[f].astype(np.int32))
        if d.n_bond and not COLLIDE(name):
            assert np.array_equal(s.bond_small_angle, d.bstate[f, :, 13].astype(np.int32))
            fmax = max(1e-9, np.abs(d.bstate[f, :, 0:3]).max())
            ftol = 2e-3 if g.every * f <= 100 else 1e-6   # see the module docstring
            if FLUID(name):
                ftol = 2e-3   # neutrally buoyant, nothing settles the lattice: rotations stay in the noise regime longer
            assert np.abs(s.bond_force1 - d.bstate[f, :, 0:3]).max() <= ftol * fmax
            assert np.abs(s.bond_force2 - d.bstate[f, :, 3:6]).max() <= ftol * fmax


@pytest.mark.parametrize("prec", [abi.VXG_PREC_F32, 2], ids=["wide", "compact"])
@pytest.mark.parametrize("name", CASES)
def test_emulated_f32_kernel_arithmetic(name, prec):
    """prec 2 = fp32 with the compact bond-class table (BondKc): every class of the reference's importer
    must qualify for it, and it must meet the same tolerance as the wide table."""
    g = golden_lib.Golden(name)
    d = g.dump
    LAT = float(d.sim[14])   # the fixture's own lattice pitch
    if FLUID(name):   # fp32 with the land/water drag phase (8 / 10: and the self-collision list); 9 / 10 = the compact form
        prec = (10 if COLLIDE(name) else 9) if prec == 2 else (8 if COLLIDE(name) else 3)
    elif COLLIDE(name):
        prec = 11 if prec == 2 else 6   # fp32 with the self-collision list (11: compact form)
    e = emul_lib.EmulSim(g.robot(), prec)
    if prec in (2, 11) and (name.startswith("bc_") or name == "dense3_density" or name == "ex_sticky_floor_bc"):
        assert not e.h   # robots with fixed / forced voxels or bonds between voxels of different mass are kept out of the production family
        return
    assert e.h, "bond classes did not qualify for the compact form"
    for f in range(1, d.vstate.shape[0]):
        assert e.step(g.every) == g.every
        s = e.state()
        assert np.abs(s.pos - d.pos[f]).max() <= 2e-5 * LAT, (name, f, np.abs(s.pos - d.pos[f]).max() / LAT)
        assert np.abs(s.quat - d.quat[f]).max() <= 1e-5
        assert np.abs(s.scale - d.scale[f]).max() <= 1e-15


@pytest.mark.parametrize("name,prec", [("ex_example_1_full", 5), ("ex_example_1_full", 11), ("ex_phase_offset_full", 11),
                                       ("ex_lw_hopper_full", 5), ("ex_lw_hexapus_full", 10)])
def test_emulated_whole_runs_of_the_reference_example_inputs(name, prec):
    """The inputs the reference ships, as shipped (SelfColEnabled=1, evolved morphologies, fluid), over their WHOLE lifetimes
    (5 059 .. 68 319 steps) on the kernels' own arithmetic: final centre of mass within 1e-6 pitch (fp64: 5 / 7) or 1e-3 pitch
    (fp32 production forms: 10 / 11) of the unmodified reference's.  The GPU suite runs all 13 through the CLI."""
    from evosoro_b200 import engine
    g = golden_lib.Golden(name)
    m = engine.Model(text=g.vxa_text, variant=engine.VARIANT_LAND_WATER if g.lw else engine.VARIANT_VOXCAD)
    total, latch, _ = m.step_plan()
    assert total == int(g.full_tail[0])
    rob = m.robot
    LAT = float(g.dump.sim[14])
    mass = np.array([c.mass for c in rob.vclasses])[rob.vox_class][:, None]
    e = emul_lib.EmulSim(rob, prec)
    assert e.h
    e.step(latch)
    ini = (e.state().pos * mass).sum(0) / mass.sum()
    e.step(total - latch)
    com = (e.state().pos * mass).sum(0) / mass.sum()
    assert not e.diverged
    tol = 1e-6 if prec in (5, 7) else 1e-3
    assert np.abs(ini - g.full_hdr[-1, 6:9]).max() <= tol * LAT
    assert np.abs(com - g.full_hdr[-1, 3:6]).max() <= tol * LAT
