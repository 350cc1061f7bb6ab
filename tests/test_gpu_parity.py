"""GPU parity: the CUDA engine (through the C-ABI, libvxgpu.so) against
  (a) golden trajectories of the UNMODIFIED reference (tests/golden/*.npz, tools/make_golden.py),
  (b) the reference itself run on this box (oracle/_ref/ref_dump, when it travelled with the snapshot),
  (c) the C oracle (oracle/vx_oracle.c) on seeded inputs, up to BASELINE.json's full sizes.
Tolerances are stated per test: positions in units of the lattice pitch."""
import os
import subprocess

import numpy as np
import pytest

import common
import golden_lib
import oracle_lib
import refdump
from evosoro_b200 import abi, engine, vxa

pytestmark = pytest.mark.gpu

LAT = 0.01  # lattice pitch of all synthetic inputs (m)
TRAJ = [n for n in golden_lib.names() if not n.endswith("_full")]
needs_ref = pytest.mark.skipif(not refdump.have_ref(), reason="oracle/_ref/ref_dump did not travel with the snapshot")


def _variant(g):
    return engine.VARIANT_LAND_WATER if g.lw else engine.VARIANT_VOXCAD


def _run_gpu(rob, prec, frames, every, bonds=True):
    b = engine.Batch([rob], precision=prec)
    if bonds:
        b.enable_bond_record(True)
    out = []
    for _ in range(frames):
        b.step(every)
        st = b.sync()
        assert st[0] == 0
        out.append(b.download(0, bonds=bonds))
    b.close()
    return out


# ---- (a) golden fixtures, imported by the product importer ---------------------------------------
# (sample_vxa/full_cross.vxa has a class per voxel -- 339 voxel and 706 bond classes: its fp64 class tables alone exceed shared memory)
@pytest.mark.parametrize("name", [n for n in TRAJ if n != "ex_lw_full_cross"])
def test_golden_f64_logic_parity(name):
    """P1: fp64 instantiation vs the reference, per-frame state: positions <= 1e-9 pitch (<= 1e-7 with live contacts: overlap x
    stiffness amplifies the rounding differences of two fp64 evaluation orders).  Covers the fluid-drag and self-collision phases
    too (fp64 on the device, blob in global memory), including the reference's own shipped example inputs."""
    g = golden_lib.Golden(name)
    d = g.dump
    LAT = float(d.sim[14])   # the fixture's own lattice pitch
    rob = engine.Model(text=g.vxa_text, variant=_variant(g)).robot
    fluid, collide = golden_lib.flags(name)
    got = _run_gpu(rob, abi.VXG_PREC_F64, d.vstate.shape[0] - 1, g.every, bonds=not (fluid or collide))
    for f, s in enumerate(got, start=1):
        assert common.max_abs(s.pos, d.pos[f]) <= (1e-7 if collide else 1e-9) * LAT, (name, f, common.max_abs(s.pos, d.pos[f]) / LAT)
        assert common.max_abs(s.quat, d.quat[f]) <= (1e-6 if collide else 1e-8)
        assert common.max_abs(s.scale, d.scale[f]) <= 1e-15
        assert np.array_equal(s.static_fric, d.sfric[f].astype(np.int32))
        if d.n_bond and not (fluid or collide):
            assert np.array_equal(s.bond_small_angle, d.bstate[f, :, 13].astype(np.int32))
            if g.every * f > 100:   # before that the reference's rotation vectors are rounding noise (tests/test_emul_golden.py)
                fmax = np.abs(d.bstate[f, :, 0:3]).max()
                assert common.max_abs(s.bond_force1, d.bstate[f, :, 0:3]) <= 1e-6 * fmax
                assert common.max_abs(s.bond_force2, d.bstate[f, :, 3:6]) <= 1e-6 * fmax


@pytest.mark.parametrize("name", TRAJ)
def test_golden_f32_short_horizon(name):
    """P2: fp32 production kernel vs the reference over the fixture horizon (<= 1500 steps):
    max |dpos| <= 2e-5 pitch, quaternions <= 1e-5."""
    g = golden_lib.Golden(name)
    d = g.dump
    LAT = float(d.sim[14])   # the fixture's own lattice pitch
    rob = engine.Model(text=g.vxa_text, variant=_variant(g)).robot
    got = _run_gpu(rob, abi.VXG_PREC_F32, d.vstate.shape[0] - 1, g.every, bonds=False)
    for f, s in enumerate(got, start=1):
        assert common.max_abs(s.pos, d.pos[f]) <= 2e-5 * LAT, (name, f, common.max_abs(s.pos, d.pos[f]) / LAT)
        assert common.max_abs(s.quat, d.quat[f]) <= 1e-5
        assert common.max_abs(s.scale, d.scale[f]) <= 1e-15


@pytest.mark.parametrize("prec,tol", [(abi.VXG_PREC_F32, 1e-5), (abi.VXG_PREC_F64, 1e-9)])
def test_cantilever_rests_at_beam_theory_deflection(prec, tol):
    """P4: 20x1x1 cantilever (VX_Benchmark.cpp:36-79), clamped at x = 0, transverse tip load, 80 000 steps to rest:
    tip deflection within `tol` pitch of the reference's and within 0.3 % of 4 P (n-1)^3 / (E s)."""
    g = golden_lib.Golden("bc_cantilever20")
    d = g.dump
    n, P, E, s = 20, 3e-5, 1e6, 1e-3
    b = engine.Batch([engine.Model(text=g.vxa_text).robot], precision=prec)
    b.step(g.n_steps)
    assert b.sync()[0] == 0
    st = b.download(0)
    b.close()
    tip, ref_tip = d.vox[-1, 2] - st.pos[-1, 2], d.vox[-1, 2] - d.pos[-1, -1, 2]
    assert abs(tip - ref_tip) <= tol * s
    assert abs(tip / (4 * P * (n - 1) ** 3 / (E * s)) - 1) < 3e-3
    assert np.array_equal(st.pos[0], d.vox[0, 0:3])   # the clamped voxel never moves


@pytest.mark.parametrize("name,prec,tol", [("dense3_full", abi.VXG_PREC_F32, 1e-2), ("dense3_full", abi.VXG_PREC_F64, 1e-6),
                                           ("dense6_s1_full", abi.VXG_PREC_F32, 1e-2), ("dense10_s3_full", abi.VXG_PREC_F32, 1e-2)])
def test_full_cli_run_fitness(name, prec, tol):
    """P3: the whole `voxelyze -f` run (step budget and IniCM latch from vxh_step_plan, all on the device):
    NormFinalDist / finalDistY / final COM within `tol` voxel lengths of the reference's result."""
    g = golden_lib.Golden(name)
    m = engine.Model(text=g.vxa_text)
    total, latch, t_end = m.step_plan()
    assert total == int(g.full_tail[0])
    b = engine.Batch([m.robot], precision=prec)
    b.set_step_budget([total], [latch])
    b.step(total + 1000)            # more than the budget: the device stops at `total`
    st = b.sync()
    assert st[0] == abi.VXG_ST_FINISHED
    s = b.stats()[0]
    assert s.steps == total and s.time == t_end and s.ini_com_latched == 1
    ref_ini, ref_com = g.full_hdr[-1, 6:9], g.full_hdr[-1, 3:6]
    ini, com = np.array(s.ini_com[:]), np.array(s.com[:])
    assert np.abs(ini - ref_ini).max() <= 1e-4 * LAT
    assert np.abs(com - ref_com).max() <= tol * LAT
    norm_final = np.hypot(com[0] - ini[0], com[1] - ini[1]) / LAT
    assert abs(norm_final - g.full_tail[1]) <= tol
    assert abs((com[1] - ini[1]) / LAT - g.full_tail[2]) <= tol
    assert s.n_touching_floor == int(g.full_tail[12])
    b.close()


# ---- (b) the reference itself, run here -------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("n,seed", [(6, 1), (4, 7)])
def test_dense_lattice_f64_vs_reference_run(n, seed):
    path = common.dense_vxa(n, seed)
    d = refdump.run_ref_dump(path, path + ".bin", 2000, 200)
    rob = engine.Model(path=path).robot
    got = _run_gpu(rob, abi.VXG_PREC_F64, 10, 200)
    for f, s in enumerate(got, start=1):
        assert common.max_abs(s.pos, d.pos[f]) <= 1e-9 * LAT
        assert common.max_abs(s.quat, d.quat[f]) <= 1e-8
        assert np.array_equal(s.bond_small_angle, d.bstate[f, :, 13].astype(np.int32))
        fmax = np.abs(d.bstate[f, :, 0:3]).max()
        assert common.max_abs(s.bond_force1, d.bstate[f, :, 0:3]) <= 1e-7 * fmax


@needs_ref
@pytest.mark.parametrize("n,seed", [(6, 1), (10, 3)])
def test_dense_lattice_f32_vs_reference_run(n, seed):
    """BASELINE lattice sizes (6^3, 10^3), 2000 steps: fp32 kernel within 1e-5 pitch of the reference."""
    path = common.dense_vxa(n, seed)
    d = refdump.run_ref_dump(path, path + ".bin", 2000, 500)
    rob = engine.Model(path=path).robot
    got = _run_gpu(rob, abi.VXG_PREC_F32, 4, 500, bonds=False)
    for f, s in enumerate(got, start=1):
        assert common.max_abs(s.pos, d.pos[f]) <= 1e-5 * LAT
        assert common.max_abs(s.quat, d.quat[f]) <= 5e-6


# ---- (c) oracle + size-independent properties, up to full size -------------------------------------
def test_batch_of_identical_robots_is_uniform():
    """Every slot of a batch runs the same arithmetic: 64 copies give bitwise identical state."""
    rob = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(5, 3))).robot
    b = engine.Batch([rob] * 64)
    b.step(500)
    b.sync()
    ref = b.download(0)
    for r in (1, 17, 63):
        s = b.download(r)
        assert np.array_equal(s.pos, ref.pos) and np.array_equal(s.quat, ref.quat)
    b.close()


@pytest.mark.parametrize("n,steps", [(3, 700), (6, 2500), (8, 2500), (10, 3000), (11, 1500)])
def test_gather_form_is_bit_identical_to_the_push_form(n, steps, monkeypatch):
    """The production kernel's two forms (vxg_gather.cuh: damping history in tensor memory, shares gathered by the -axis
    neighbour, with CTA-wide barriers or as a per-warp dataflow; vxg_kernels.cuh: shares pushed into shared-memory accumulators) run
    the same arithmetic in the same order: every word of the state agrees after thousands of actuated steps, in one launch or many."""
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=1.0)
    # (n = 11: sparse only, ~800 voxels each -- a dense 11^3 would move the whole batch to stride 2048)
    robots = [engine.Model(text=vxa.vxa_text(vxa.dense_lattice(n, 80 + s) if s % 2 == 0 and n < 11 else vxa.sparse_lattice(n, 80 + s, fill=0.6), sim=sim)).robot for s in range(4)]
    latch = [steps // 3] * 4    # IniCM latch in the middle of a launch: the block reduction shares scratch with other users
    monkeypatch.setenv("VXG_KERNEL_FORM", "push")
    a = engine.Batch(robots); a.set_step_budget(None, latch); a.step(steps); assert not a.sync().any()
    # ("cluster": stride-1024 robots split over a two-CTA cluster, vxg_cluster.cuh -- the dense 10^3 ones, cut between two z layers, and
    # the sparse 11^3 ones, cut wherever voxel 512 happens to be; other strides run the single-CTA kernel under that name too)
    for form in ("barrier", "flow", "cluster"):
        monkeypatch.setenv("VXG_KERNEL_FORM", form)
        b = engine.Batch(robots); b.set_step_budget(None, latch)
        for i in range(5):
            b.step(steps // 5 + (3 if i == 0 else -3 if i == 4 else 0))      # launch lengths that are not multiples of the window
        assert not b.sync().any()
        # (IniCM is a sum over the voxels: the 1024-thread gather kernel adds them up in another order than the push form)
        assert np.allclose([list(x.ini_com) for x in a.stats()], [list(x.ini_com) for x in b.stats()], rtol=1e-13, atol=0)
        for r in range(4):
            sa, sb = a.download(r), b.download(r)
            for f in ("pos", "quat", "lin_mom", "ang_mom", "scale", "static_fric", "bond_small_angle"):
                assert np.array_equal(getattr(sa, f), getattr(sb, f)), (form, r, f)
        assert [int(s.steps) for s in b.stats()] == [steps] * 4 and [s.time for s in b.stats()] == [s.time for s in a.stats()]
        b.close()
    a.close()


@pytest.mark.parametrize("name,steps", [("col_pincer", 800), ("ex_example_2", 1500), ("ex_example_1", 3000), ("lw_fluid_dense3", 1200),
                                        ("lw_fluid_col_pincer", 800), ("ex_lw_hexapus", 3000), ("ex_lw_quadrupus", 2000)])
def test_gather_form_with_self_collision_or_fluid_drag_is_bit_identical_to_the_push_form(name, steps, monkeypatch):
    """The same for the self-collision and land/water variants: contact-bond list rebuilds, live contacts (the pincer's 22,
    Example_2's folds) and the per-voxel contact lists of the gather form against the list scan of the push form; the surface
    mesh, facet drag and per-voxel strains of the fluid environment (hexapus / quadrupus: fluid AND self-collision, as shipped)."""
    g = golden_lib.Golden(name)
    rob = engine.Model(text=g.vxa_text, variant=_variant(g)).robot
    monkeypatch.setenv("VXG_KERNEL_FORM", "push")
    a = engine.Batch([rob, rob]); a.step(steps); assert not a.sync().any()
    monkeypatch.setenv("VXG_KERNEL_FORM", "barrier")
    b = engine.Batch([rob, rob])
    for i in range(4):
        b.step(steps // 4)
    assert not b.sync().any()
    sa, sb = a.download(1), b.download(1)
    for f in ("pos", "quat", "lin_mom", "ang_mom", "scale", "static_fric", "bond_small_angle"):
        assert np.array_equal(getattr(sa, f), getattr(sb, f)), f
    assert a.stats()[0].max_vel == b.stats()[0].max_vel


def test_reuploading_one_slot_leaves_the_others_running():
    """vxg_upload of a single slot after the batch has advanced: that robot restarts from rest, the others keep the state
    the device computed (the upload path moves the slots the host changed, nothing else)."""
    sim = vxa.SimParams(init_cm_time=0.005, stop_condition_value=0.2)
    robots = [engine.Model(text=vxa.vxa_text(vxa.dense_lattice(4, 60 + s), sim=sim)).robot for s in range(3)]
    a = engine.Batch(robots); a.step(300); assert not a.sync().any()
    a.reupload(1)
    a.step(200); assert not a.sync().any()
    ref = engine.Batch(robots); ref.step(500); ref.sync()
    fresh = engine.Batch(robots[1:2]); fresh.step(200); fresh.sync()
    for r in (0, 2):
        assert np.array_equal(a.download(r).pos, ref.download(r).pos) and np.array_equal(a.download(r).scale, ref.download(r).scale)
    assert np.array_equal(a.download(1).pos, fresh.download(0).pos)
    st = a.stats()
    assert [int(s.steps) for s in st] == [500, 200, 500]
    for b in (a, ref, fresh):
        b.close()


def test_reuploading_one_slot_keeps_the_live_contact_lists_of_the_others():
    """The same with SelfColEnabled=1 and live contact bonds: the contact-bond lists are device-managed state, so re-uploading
    slot 1 must not replace the lists of slots 0 and 2 with the host's (empty) mirror."""
    g = golden_lib.Golden("col_pincer")
    rob = engine.Model(text=g.vxa_text).robot
    a = engine.Batch([rob] * 3); a.step(500); assert not a.sync().any()      # contacts are live from ~step 300 on
    a.reupload(1)
    a.step(300); assert not a.sync().any()
    ref = engine.Batch([rob]); ref.step(800); ref.sync()
    fresh = engine.Batch([rob]); fresh.step(300); fresh.sync()
    for r in (0, 2):
        assert np.array_equal(a.download(r).pos, ref.download(0).pos)
    assert np.array_equal(a.download(1).pos, fresh.download(0).pos)
    assert np.isfinite(a.download(0).pos).all()
    for b in (a, ref, fresh):
        b.close()


def test_large_robot_with_a_class_per_voxel_runs_from_global_memory():
    """A 10^3 robot whose every voxel has its own phase offset (continuous CPPN output) needs 1000 voxel classes: the class
    tables take the shared memory the blob would use, so the batch runs the same kernels on the blob in global memory.  Checked
    against the fp64 C oracle after 150 steps through actuation onset."""
    rng = np.random.RandomState(5)
    text = vxa.vxa_text(vxa.dense_lattice(10, 77, materials=(1, 3, 4)), sim=vxa.SimParams(init_cm_time=0.02, stop_condition_value=0.2),
                        phase_offset=rng.uniform(-0.5, 0.5, (10, 10, 10)))
    rob = engine.Model(text=text).robot
    assert len(rob.vclasses) == 1000
    b = engine.Batch([rob, rob])
    cfg = b.kernel_config()
    assert not cfg["smem_resident"]
    b.step(150)
    assert not b.sync().any()
    o = oracle_lib.OracleSim(rob)
    o.step(150)
    got, want = b.download(0), o.state()
    assert common.max_abs(got.pos, want.pos) <= 1e-5 * LAT
    assert common.max_abs(got.scale, want.scale) <= 1e-12 and np.ptp(want.scale) > 1e-4   # actuation is on, per voxel
    assert np.array_equal(got.pos, b.download(1).pos)
    b.close()


def test_largest_supported_robot_and_the_size_limit():
    """A 12^3 robot (1728 voxels, stride 2048) does not fit in shared memory and runs from global memory; it must agree with the
    fp64 C oracle like the resident path.  Above 2048 voxels (the reference has no limit, VX_Sim.cpp:530-600) the strides are 8192 and
    32768 at 16 / 64 voxels per thread: 13^3 and 21^3 robots against the oracle, in fp32 and in fp64.  Voxel indices are 16-bit:
    32^3 = 32768 voxels is refused, not truncated."""
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=0.2)
    rob = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(12, 3), sim=sim)).robot
    assert rob.n_vox == 1728
    b = engine.Batch([rob])
    assert not b.kernel_config()["smem_resident"]
    b.step(200)
    assert not b.sync().any()
    o = oracle_lib.OracleSim(rob)
    o.step(200)
    assert common.max_abs(b.download(0).pos, o.state().pos) <= 1e-5 * LAT
    b.close()
    for n, steps in ((13, 200), (21, 120)):
        big = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(n, 3), sim=sim)).robot
        assert big.n_vox == n ** 3
        o = oracle_lib.OracleSim(big)
        o.step(steps)
        for prec, tol in ((abi.VXG_PREC_F32, 1e-5), (abi.VXG_PREC_F64, 1e-9)):
            b = engine.Batch([big, big], precision=prec)
            assert b.kernel_config()["voxels_per_thread"] == (16 if n == 13 else 64) and not b.kernel_config()["smem_resident"]
            b.step(steps)
            assert not b.sync().any()
            assert common.max_abs(b.download(1).pos, o.state().pos) <= tol * LAT, (n, prec)
            assert [int(x.steps) for x in b.stats()] == [steps, steps]
            b.close()
    huge = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(32, 3), sim=sim)).robot
    with pytest.raises(engine.EngineError) as ei:
        engine.Batch([huge])
    assert ei.value.code > 0   # unsupported size: the caller must abort


def test_robots_outside_the_production_family_run_the_wide_one():
    """The production kernels take equal end-voxel masses, no boundary conditions and no sticky floor for granted; a batch with
    one robot outside that set (here: materials of different density) must run the full-feature family -- visible in the launch
    shape of a 10^3 batch (512 x 2 instead of 1024 x 1) -- and still agree with the oracle."""
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=0.2)
    pal = vxa.four_material_palette()
    pal[1].density, pal[2].density = 3e6, 5e5
    plain = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(10, 5), sim=sim)).robot
    heavy = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(10, 5), sim=sim, palette=pal)).robot
    a = engine.Batch([plain, plain])
    assert a.kernel_config()["threads"] == 1024
    a.close()
    b = engine.Batch([plain, heavy], max_bclass=max(len(plain.bclasses), len(heavy.bclasses)))
    cfg = b.kernel_config()
    assert cfg["threads"] == 512 and cfg["voxels_per_thread"] == 2 and cfg["smem_resident"]
    b.step(200)
    assert not b.sync().any()
    for r, rob in enumerate((plain, heavy)):
        o = oracle_lib.OracleSim(rob)
        o.step(200)
        assert common.max_abs(b.download(r).pos, o.state().pos) <= 1e-5 * LAT
    b.close()


@pytest.mark.parametrize("n,threads", [(7, 352), (9, 736)])
def test_strides_between_the_powers_of_two(n, threads):
    """7^3 and 9^3 robots run with 352 / 736 voxel slots (not 512 / 1024): same results as the oracle, and a batch of mixed
    seeds stays independent."""
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=0.2)
    robots = [engine.Model(text=vxa.vxa_text(vxa.dense_lattice(n, 80 + s), sim=sim)).robot for s in range(3)]
    b = engine.Batch(robots + robots[:1])
    assert b.kernel_config()["threads"] == threads
    b.step(250)
    assert not b.sync().any()
    for r in (0, 2):
        o = oracle_lib.OracleSim(robots[r])
        o.step(250)
        assert common.max_abs(b.download(r).pos, o.state().pos) <= 1e-5 * LAT
    assert np.array_equal(b.download(0).pos, b.download(3).pos)
    b.close()


def test_step_chunking_is_exact():
    """K steps in one launch == K single-step launches (state round-trips through HBM exactly)."""
    rob = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(4, 11))).robot
    b1 = engine.Batch([rob]); b2 = engine.Batch([rob])
    b1.step(300)
    for _ in range(300):
        b2.step(1)
    b1.sync(); b2.sync()
    s1, s2 = b1.download(0), b2.download(0)
    assert np.array_equal(s1.pos, s2.pos) and np.array_equal(s1.quat, s2.quat)
    assert np.array_equal(s1.lin_mom, s2.lin_mom) and np.array_equal(s1.ang_mom, s2.ang_mom)
    b1.close(); b2.close()


def test_gpu_matches_c_oracle_with_actuation():
    """Through the actuation onset (t > InitCmTime): fp32 kernel vs the C oracle, COM within 1e-5 pitch."""
    sim = vxa.SimParams(init_cm_time=0.02, stop_condition_value=0.2)
    rob = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(5, 2), sim=sim)).robot
    o = oracle_lib.OracleSim(rob)
    b = engine.Batch([rob])
    for _ in range(5):
        o.step(400)
        b.step(400)
        b.sync()
        s = b.download(0)
        so = o.state()
        assert common.max_abs(s.pos, so.pos) <= 1e-5 * LAT
        assert common.max_abs(s.scale, so.scale) <= 1e-12
    st = b.stats()[0]
    assert np.abs(np.array(st.com[:]) - o.com()).max() <= 1e-5 * LAT
    assert st.steps == 2000 and abs(st.time - o.time) < 1e-15
    assert st.n_touching_floor == oracle_lib.lib().vxo_num_touching_floor(o.h)
    b.close()


def test_full_size_population_1024_x_10cubed():
    """BASELINE.json's headline shape: 1024 robots x 10^3 voxels on one GPU.  Slots r and r+512 hold the same
    robot (512 distinct seeds): the two halves must be bitwise identical whichever SM / wave ran them, and
    sampled robots must agree with the fp64 C oracle to 1e-5 pitch after 300 steps through actuation onset."""
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=0.2)
    distinct = [engine.Model(text=vxa.vxa_text(vxa.dense_lattice(10, s), sim=sim)).robot for s in range(512)]
    b = engine.Batch(distinct + distinct)
    cfg = b.kernel_config()
    assert cfg["smem_resident"] and cfg["threads"] == 1024 and cfg["voxels_per_thread"] == 1
    b.step(300)
    st = b.sync()
    assert not st.any()
    for r in (0, 77, 511):
        a, c = b.download(r), b.download(r + 512)
        assert np.array_equal(a.pos, c.pos) and np.array_equal(a.quat, c.quat) and np.array_equal(a.lin_mom, c.lin_mom)
    for r in (0, 511):
        o = oracle_lib.OracleSim(distinct[r])
        o.step(300)
        assert common.max_abs(b.download(r).pos, o.state().pos) <= 1e-5 * LAT
    stats = b.stats()
    assert all(s.steps == 300 for s in stats)
    b.close()


def test_ragged_batch_mixed_sizes_and_budgets():
    """One batch holding robots of different sizes (1, 27, 62 and 216 voxels) with different step budgets:
    each stops at its own budget and matches a batch-of-one run bit for bit."""
    one = np.zeros((2, 2, 2), dtype=np.int32); one[0, 0, 0] = 1
    structs = [one, vxa.dense_lattice(3, 5), vxa.sparse_lattice(5, 3, fill=0.5), vxa.dense_lattice(6, 2, materials=(1, 3, 4))]
    robots = [engine.Model(text=vxa.vxa_text(s)).robot for s in structs]
    budgets = [50, 400, 125, 300]
    b = engine.Batch(robots)
    b.set_step_budget(budgets, None)
    b.step(400)
    st = b.sync()
    assert all(x == abi.VXG_ST_FINISHED for x in st)
    stats = b.stats()
    for r, rob in enumerate(robots):
        assert stats[r].steps == budgets[r]
        solo = engine.Batch([rob])
        solo.step(budgets[r]); solo.sync()
        a, c = b.download(r), solo.download(0)
        assert np.array_equal(a.pos, c.pos) and np.array_equal(a.quat, c.quat)
        solo.close()
    b.close()


def test_divergence_is_reported_not_hidden():
    """A bond strain > 100 makes Integrate() return false (VX_Sim.cpp:1775-1777): status bit 0, robot frozen."""
    rob = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(2, 1))).robot
    b = engine.Batch([rob])
    b.step(5); b.sync()
    s = b.download(0)
    s.pos[0, 0] += 5.0          # tear one voxel 500 pitches away
    b.upload_state(0, s, 5 * rob.sim.dt, 5)
    b.step(10)
    st = b.sync()
    assert st[0] & abi.VXG_ST_DIVERGED
    assert b.stats()[0].steps == 5
    b.close()


def test_host_buffer_entry_point_matches_resident_path():
    """vxg_run_batch_host (the e2e call the benchmark times) == upload + step + download."""
    robots = [engine.Model(text=vxa.vxa_text(vxa.dense_lattice(4, s))).robot for s in range(6)]
    e = engine.Batch(robots, upload=False)
    pos = e.run_batch_host(250)
    b = engine.Batch(robots)
    b.step(250); b.sync()
    for r in range(6):
        assert np.array_equal(pos[r, : robots[r].n_vox], b.download(r).pos)
    e.close(); b.close()


# ---- P5: the CLI contract (seam B1) ----------------------------------------------------------------
FULL_XML = [n for n in golden_lib.names() if n.endswith("_full") and golden_lib.Golden(n).result_xml]
CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "evosoro_b200", "lib", "voxelyze")


def _check_result(ours_xml, g):
    ours, ref = golden_lib.parse_result(ours_xml), golden_lib.parse_result(g.result_xml)
    assert list(ours.keys()) == list(ref.keys())
    for k, v in ref.items():
        if k == "RobotVolumeEnd":   # volume of the deformed surface mesh at the end (m^3): relative, it is ~1e-5
            assert ours[k] == pytest.approx(v, rel=1e-4), k
            continue
        if k in ("NumTouchingFloor", "NumNonFeetTouchingFloor", "VoxelNumber", "FoundNeedleInHaystack"):
            assert ours[k] == v, k
        else:   # distances are in voxel lengths: P3 tolerance 1e-2, relative 1e-3 for the O(1) ones
            assert ours[k] == pytest.approx(v, rel=1e-3, abs=1e-2), k


@pytest.mark.parametrize("name", FULL_XML)
def test_cli_drop_in_writes_reference_result_file(name, tmp_path):
    """`voxelyze -f file.vxa`: exit code 1, result XML at <FitnessFileName> with the reference's tags/values."""
    g = golden_lib.Golden(name)
    out = tmp_path / "fit.xml"
    vxa_path = tmp_path / (name + ".vxa")
    vxa_path.write_text(g.vxa_text.replace(g.fitness_file, str(out)))
    cmd = [CLI, "-f", str(vxa_path)] + (["--lw"] if g.lw else [])
    rc = subprocess.run(cmd).returncode
    assert rc == 1
    _check_result(out.read_text(), g)


def test_cli_reference_example_inputs_hit_the_recorded_reference_values(tmp_path):
    """BASELINE.md section 4 (values recorded from the reference before any of this existed): Example_1.vxa -> NormFinalDist
    0.0795605 after 49 366 steps; sample_vxa/hexapus.vxa -> normAbsoluteDisplacement 17.1761.  Both files as shipped
    (SelfColEnabled=1), one batch per fork through the drop-in CLI."""
    for name, lw, tag, want in (("ex_example_1_full", False, "NormFinalDist", 0.0795605), ("ex_lw_hexapus_full", True, "normAbsoluteDisplacement", 17.1761)):
        g = golden_lib.Golden(name)
        out = tmp_path / (name + ".xml")
        p = tmp_path / (name + ".vxa")
        p.write_text(g.vxa_text.replace(g.fitness_file, str(out)))
        assert subprocess.run([CLI, "-f", str(p)] + (["--lw"] if lw else [])).returncode == 1
        got = golden_lib.parse_result(out.read_text())
        assert got[tag] == pytest.approx(want, abs=1e-2), (name, got[tag])
        assert golden_lib.parse_result(g.result_xml)[tag] == pytest.approx(want, abs=5e-5)   # the fixture is that same run
    assert int(golden_lib.Golden("ex_example_1_full").full_tail[0]) == 49366


def test_cli_all_reference_example_inputs_in_one_batch(tmp_path):
    """Every input the reference ships for one fork as ONE ragged batch (98 .. 743 voxels, 5 520 .. 49 366 steps, live
    self-collision): each robot's result file equals the one it gets alone (checked against the reference's by the test above)."""
    names = [n for n in FULL_XML if n.startswith("ex_") and not golden_lib.Golden(n).lw]
    args, outs = [], []
    for n in names:
        g = golden_lib.Golden(n)
        out = tmp_path / (n + ".xml"); p = tmp_path / (n + ".vxa")
        p.write_text(g.vxa_text.replace(g.fitness_file, str(out)))
        args += ["-f", str(p)]; outs.append((out, g))
    assert subprocess.run([CLI] + args).returncode == 1
    for out, g in outs:
        _check_result(out.read_text(), g)


def test_cli_batch_mode_many_files(tmp_path):
    """The batched form: several -f files in one process = one GPU batch; every robot gets its own result file,
    identical to running it alone."""
    names = ["dense3_full", "dense3_afterlife_full"]
    outs, args = [], []
    for n in names * 2:
        g = golden_lib.Golden(n)
        out = tmp_path / f"fit_{len(outs)}.xml"
        p = tmp_path / f"r_{len(outs)}.vxa"
        p.write_text(g.vxa_text.replace(g.fitness_file, str(out)))
        outs.append((out, g)); args += ["-f", str(p)]
    assert subprocess.run([CLI] + args).returncode == 1
    for out, g in outs:
        _check_result(out.read_text(), g)
    assert outs[0][0].read_text() == outs[2][0].read_text()     # same robot in two slots: identical file


def _diverging_vxa(fit):
    """DtFrac 4: the explicit integrator is unstable, a bond passes strain 100 after ~90 steps.  The reference CLI never returns
    on this input (TimeStep keeps failing, CurTime stands still) and writes no result -- checked with oracle/_ref/voxelyze_ref."""
    return vxa.vxa_text(vxa.dense_lattice(3, 7), sim=vxa.SimParams(dt_frac=4.0, init_cm_time=0.01, stop_condition_value=0.1, fitness_file=str(fit)))


def test_cli_diverged_robot_writes_no_result_and_spares_its_batch_mates(tmp_path):
    """VX_Sim.cpp:1147-1150 + main.cpp:89: a diverged simulation never reaches its stop condition, so no result file appears and
    evosoro keeps the worst fitness for it (evaluation.py:107-110, 213-215).  Here: no file for that robot (a stale one is
    removed), exit code 0, and the other robots of the batch get exactly the result they get alone."""
    g = golden_lib.Golden("dense3_full")
    good, bad = tmp_path / "good.xml", tmp_path / "bad.xml"
    bad.write_text("<stale/>")
    (tmp_path / "a.vxa").write_text(g.vxa_text.replace(g.fitness_file, str(good)))
    (tmp_path / "b.vxa").write_text(_diverging_vxa(bad))
    r = subprocess.run([CLI, "-f", str(tmp_path / "b.vxa"), "-f", str(tmp_path / "a.vxa")], capture_output=True, text=True)
    assert r.returncode == 0 and "Simulation Diverged" in r.stderr
    assert not bad.exists()
    _check_result(good.read_text(), g)
    r = subprocess.run([CLI, "-f", str(tmp_path / "b.vxa")], capture_output=True, text=True)   # alone: the reference hangs, we report
    assert r.returncode == 0 and not bad.exists()


def test_cli_bad_file_does_not_cost_the_others_their_results(tmp_path):
    """One unreadable and one unsupported file in a batch: both are reported and skipped, the rest run."""
    g = golden_lib.Golden("dense3_full")
    good = tmp_path / "good.xml"
    (tmp_path / "a.vxa").write_text(g.vxa_text.replace(g.fitness_file, str(good)))
    (tmp_path / "eq.vxa").write_text(vxa.vxa_text(vxa.dense_lattice(2, 1), sim=vxa.SimParams(equilibrium_mode=1, fitness_file=str(tmp_path / "eq.xml"))))
    (tmp_path / "junk.vxa").write_text("<VXA><Simulator>")
    r = subprocess.run([CLI, "-f", str(tmp_path / "junk.vxa"), "-f", str(tmp_path / "a.vxa"), "-f", str(tmp_path / "eq.vxa"),
                        "-f", str(tmp_path / "missing.vxa")], capture_output=True, text=True)
    assert r.returncode == 0
    assert r.stderr.count("Problem importing VXA file") == 3 and "not supported" in r.stderr
    _check_result(good.read_text(), g)
    assert not (tmp_path / "eq.xml").exists()


def test_cli_rejects_unsupported_instead_of_falling_back(tmp_path):
    p = tmp_path / "eq.vxa"
    p.write_text(vxa.vxa_text(vxa.dense_lattice(2, 1), sim=vxa.SimParams(equilibrium_mode=1, fitness_file=str(tmp_path / "f.xml"))))
    r = subprocess.run([CLI, "-f", str(p)], capture_output=True, text=True)
    assert r.returncode == 0 and "not supported" in r.stderr
    assert not (tmp_path / "f.xml").exists()


# ---- seam B2: the reference's own CLI + library with CVX_Sim::Integrate() swapped for the engine ------
ADAPTER = os.path.join(refdump.REF_DIR, "voxelyze_gpu_adapter")
needs_adapter = pytest.mark.skipif(not os.access(ADAPTER, os.X_OK), reason="oracle/_ref/voxelyze_gpu_adapter did not travel with the snapshot")


@needs_adapter
@pytest.mark.parametrize("name,f64", [("dense3_full", False), ("dense3_full", True), ("dense3_afterlife_full", False), ("col_pincer_full", False),
                                      ("dev_growth3_full", False), ("dev_growth3_full", True), ("evosoro_basic_id1_full", False),
                                      ("ex_example_1_full", False), ("ex_phase_offset_full", False)])
def test_class_api_adapter_reference_cli_on_the_engine(name, f64, tmp_path):
    """voxelyzeMain/main.cpp, LoadVXAFile/Import, TimeStep's latches, UpdateCollisions (host-side contact list ->
    vxg_set_collision_pairs), UpdateStats and WriteResultFile are all the UNMODIFIED reference; only Integrate() is
    evosoro_b200/adapter/vx_integrate_gpu.cpp (one device step + write-back into VoxArray per call).  The result file
    the reference writes from the written-back state must be the one it writes on its own CPU path."""
    g = golden_lib.Golden(name)
    out = tmp_path / "fit.xml"
    vxa_path = tmp_path / (name + ".vxa")
    vxa_path.write_text(g.vxa_text.replace(g.fitness_file, str(out)))
    env = dict(os.environ, **({"VXG_ADAPTER_F64": "1"} if f64 else {}))
    r = subprocess.run([ADAPTER, "-f", str(vxa_path)], env=env, capture_output=True, text=True)
    assert r.returncode == 1, r.stderr
    _check_result(out.read_text(), g)
    if f64:   # logic parity: the reference's own getters over fp64 device state reproduce its CPU run to the printed digits
        ours, ref = golden_lib.parse_result(out.read_text()), golden_lib.parse_result(g.result_xml)
        for k, v in ref.items():
            assert ours[k] == pytest.approx(v, rel=1e-6, abs=1e-9), k


ADAPTER_LW = os.path.join(refdump.REF_DIR, "voxelyze_gpu_adapter_lw")


@pytest.mark.skipif(not os.access(ADAPTER_LW, os.X_OK), reason="oracle/_ref/voxelyze_gpu_adapter_lw did not travel with the snapshot")
@pytest.mark.parametrize("name", ["lw_dense3_full", "lw_fluid_dense3_full", "ex_lw_hopper_full"])
def test_class_api_adapter_land_water_fork(name, tmp_path):
    """The same seam under the land/water fork's CLI + library: its Integrate() (bonds, surface-mesh fluid drag, voxels) runs on
    the device; the per-voxel strains are written back too, so the reference's own deformable mesh yields RobotVolumeEnd."""
    g = golden_lib.Golden(name)
    out = tmp_path / "fit.xml"
    vxa_path = tmp_path / (name + ".vxa")
    vxa_path.write_text(g.vxa_text.replace(g.fitness_file, str(out)))
    r = subprocess.run([ADAPTER_LW, "-f", str(vxa_path)], capture_output=True, text=True)
    assert r.returncode == 1, r.stderr
    _check_result(out.read_text(), g)


@needs_adapter
def test_class_api_adapter_rejects_unsupported(tmp_path):
    p = tmp_path / "eq.vxa"
    p.write_text(vxa.vxa_text(vxa.dense_lattice(2, 1), sim=vxa.SimParams(equilibrium_mode=1, fitness_file=str(tmp_path / "f.xml"))))
    r = subprocess.run([ADAPTER, "-f", str(p)], capture_output=True, text=True)
    assert r.returncode == 3 and "EquilibriumModeEnabled" in r.stderr
    assert not (tmp_path / "f.xml").exists()


# ---- BASELINE config 5: land/water fork, 8^3 swimming robots with fluid drag -------------------------
def _lw_fluid_text(n, seed):
    env = vxa.EnvParams(fluid_environment=1, aggregate_drag_coefficient=750.0)
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=0.2)
    return vxa.vxa_text(vxa.dense_lattice(n, seed, materials=(1, 3, 4)), sim=sim, env=env)


def test_config5_land_water_fluid_batch_8cubed(tmp_path):
    """Dense 8^3 land/water robots in fluid (surface-mesh drag on 386 vertices / 768 facets each): duplicated slots
    are bitwise identical, and -- when the reference build travelled -- robot 0 tracks the unmodified LW engine to
    1e-5 pitch over 1000 steps."""
    texts = [_lw_fluid_text(8, s) for s in range(8)]
    robots = [engine.Model(text=t, variant=engine.VARIANT_LAND_WATER).robot for t in texts]
    assert all(r.sim.flags & abi.VXG_F_FLUID for r in robots)
    b = engine.Batch(robots + robots)
    assert b.limits.max_mesh_verts == 9 ** 3 - 7 ** 3
    b.step(1000)
    st = b.sync()
    assert not st.any()
    a, c = b.download(3), b.download(11)
    assert np.array_equal(a.pos, c.pos) and np.array_equal(a.quat, c.quat)
    if refdump.have_ref(lw=True):
        p = tmp_path / "lw8.vxa"
        p.write_text(texts[0])
        d = refdump.run_ref_dump(str(p), str(p) + ".bin", 1000, 500, lw=True)
        assert common.max_abs(b.download(0).pos, d.pos[2]) <= 1e-5 * LAT
    # drag must actually act: the same robots without the fluid flag end up somewhere else
    dry = [engine.Model(text=t.replace("<FluidEnvironment>1", "<FluidEnvironment>0"), variant=engine.VARIANT_LAND_WATER).robot for t in texts[:1]]
    b2 = engine.Batch(dry)
    b2.step(1000); b2.sync()
    assert common.max_abs(b2.download(0).pos, b.download(0).pos) > 1e-4 * LAT
    b.close(); b2.close()


def test_fluid_robot_needs_mesh_capacity():
    rob = engine.Model(text=_lw_fluid_text(3, 1), variant=engine.VARIANT_LAND_WATER).robot
    lim_batch = engine.Batch([engine.Model(text=vxa.vxa_text(vxa.dense_lattice(3, 1))).robot])   # created without mesh capacity
    d = rob.desc()
    import ctypes as C
    rc = engine.lib().vxg_upload(lim_batch.h, 0, C.byref(d))
    assert rc == 2 and b"max_mesh_verts" in engine.lib().vxg_last_error()
    lim_batch.close()


# ---- self-collision: the device-side list rebuild and contact forces -------------------
def test_collision_list_overflow_is_an_error_not_a_truncation():
    g = golden_lib.Golden("col_pincer")
    rob = engine.Model(text=g.vxa_text).robot
    b = engine.Batch([rob], max_col_pairs=4)      # the pincer needs up to 22 contact bonds
    b.step(800)
    st = b.sync()
    assert st[0] & abi.VXG_ST_COL_OVERFLOW and st[0] & abi.VXG_ST_DIVERGED
    assert b.stats()[0].steps < 800
    b.close()


def test_collision_state_survives_launch_boundaries():
    """The contact-bond list, MaxDispSinceLastBondUpdate and MaxVoxVel persist between launches: 800 single-step
    launches == one 800-step launch, bit for bit, through the contact episode."""
    g = golden_lib.Golden("col_pincer")
    rob = engine.Model(text=g.vxa_text).robot
    b1 = engine.Batch([rob]); b2 = engine.Batch([rob])
    b1.step(800)
    for _ in range(16):
        b2.step(50)
    b1.sync(); b2.sync()
    s1, s2 = b1.download(0), b2.download(0)
    assert np.array_equal(s1.pos, s2.pos) and np.array_equal(s1.quat, s2.quat)
    assert common.max_abs(s1.pos, g.dump.pos[-1]) <= 2e-5 * LAT
    b1.close(); b2.close()


@needs_ref
def test_large_folding_robot_with_real_contacts_vs_reference_run(tmp_path):
    """A 660-voxel robot whose limbs press against each other (over a hundred contact bonds, trajectories with and
    without collisions differ by ~1 pitch): fp32 kernel with the device-side list vs the unmodified reference
    with SelfColEnabled=1, 2000 steps.  The input is built here, not copied: two actuated slabs on a
    passive base plate that fold onto each other."""
    st = np.zeros((9, 10, 8), dtype=np.int32)
    st[:, :, 0:2] = 1                       # passive base plate
    st[0:2, :, 2:] = 3; st[2:4, :, 2:] = 4  # left slab bends right
    st[5:7, :, 2:] = 4; st[7:9, :, 2:] = 3  # right slab bends left, one empty column between them
    sim = vxa.SimParams(init_cm_time=0.005, stop_condition_value=0.5, self_col_enabled=1)
    p = tmp_path / "fold.vxa"
    p.write_text(vxa.vxa_text(st, sim=sim))
    d = refdump.run_ref_dump(str(p), str(p) + ".bin", 1000, 250, nocol=False)
    dn = refdump.run_ref_dump(str(p), str(p) + ".n.bin", 1000, 1000, nocol=True)
    assert d.frame_hdr[:, 9].max() > 10                                  # contact bonds exist ...
    assert common.max_abs(d.pos[-1], dn.pos[-1]) > 1e-2 * LAT            # ... and matter
    rob = engine.Model(path=str(p)).robot
    b = engine.Batch([rob] * 4)
    for f in range(1, 5):
        b.step(250); assert not b.sync().any()
        assert common.max_abs(b.download(2).pos, d.pos[f]) <= 5e-5 * LAT, f
    b.close()


# ---- randomized robots: mixed densities / Poisson ratios / friction, ragged shapes ---------------------
def _random_robot_text(seed):
    rng = np.random.RandomState(seed)
    n = int(rng.randint(3, 7))
    st = vxa.sparse_lattice(n, seed, fill=float(rng.uniform(0.4, 0.95)))
    pal = [vxa.Material("m%d" % i, elastic_mod=float(10 ** rng.uniform(6.3, 7.7)), cte=float(rng.choice([0, 0.01, -0.01, 0.02])),
                        density=float(10 ** rng.uniform(5.7, 6.3)), poisson=float(rng.uniform(0.0, 0.45)),
                        u_static=float(rng.uniform(0.3, 1.5)), u_dynamic=float(rng.uniform(0.1, 0.8))) for i in range(4)]
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=0.3, dt_frac=float(rng.uniform(0.5, 0.95)),
                        bond_damping_z=float(rng.uniform(0.5, 1.0)), slow_damping_z=float(rng.uniform(0.0, 0.05)))
    env = vxa.EnvParams(temp_period=float(rng.uniform(0.05, 0.3)), growth_amplitude=float(rng.choice([0.0, 0.2])))
    return vxa.vxa_text(st, palette=pal, sim=sim, env=env)


def test_randomized_population_vs_c_oracle():
    """48 random robots in ONE ragged batch (3^3..6^3 sparse shapes, 4 random materials each with different
    densities -- so Force2 != -Force1 --, Poisson ratios, friction, damping, dt fractions, actuation periods):
    every robot within 2e-5 pitch of the fp64 C oracle after 1500 steps."""
    robots = [engine.Model(text=_random_robot_text(1000 + s)).robot for s in range(48)]
    b = engine.Batch(robots)
    b.step(1500)
    assert not b.sync().any()
    worst = 0.0
    for r, rob in enumerate(robots):
        o = oracle_lib.OracleSim(rob)
        o.step(1500)
        worst = max(worst, common.max_abs(b.download(r).pos, o.state().pos) / LAT)
    assert worst <= 2e-5, worst
    b.close()


def test_cli_reruns_a_robot_that_outgrew_its_contact_list(tmp_path):
    """The reference's contact-bond list grows without bound (VX_Sim.cpp:2357-2413); the engine's has a capacity.  A robot that
    outgrows it is not failed: the CLI runs it again with more room.  Forced here by a first pass with room for 8 bonds on the
    reference's Example_2.vxa (743 voxels, folds onto itself: the list outgrows 8 bonds at step 494): same result file as with the
    default capacity, which is the reference's."""
    g = golden_lib.Golden("ex_example_2_full")
    outs = []
    for tag, env in (("a", {}), ("b", {"VXG_COL_PAIRS_PER_VOX": "-8"})):
        out = tmp_path / (tag + ".xml"); p = tmp_path / (tag + ".vxa")
        p.write_text(g.vxa_text.replace(g.fitness_file, str(out)))
        r = subprocess.run([CLI, "-f", str(p)], env=dict(os.environ, **env), capture_output=True, text=True)
        assert r.returncode == 1, r.stderr
        assert ("outgrew the contact-bond list" in r.stderr) == bool(env)
        outs.append(golden_lib.parse_result(out.read_text()))
    assert outs[0] == outs[1]
    _check_result((tmp_path / "b.xml").read_text(), g)


def test_gather_form_of_an_8_cubed_swimmer_is_bit_identical_to_the_push_form(monkeypatch):
    """BASELINE config 5's shape (dense 8^3 in the fluid environment) runs the land/water kernel whose class-table and surface-mesh
    capacities are compile-time constants (two robots per SM); a sparser swimmer in the same stride has more mesh vertices than that
    instantiation holds and runs the run-time one: both against the push form, every word of the state."""
    env = vxa.EnvParams(fluid_environment=1, aggregate_drag_coefficient=750.0)
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=1.0)
    for st in (vxa.dense_lattice(8, 21), vxa.sparse_lattice(8, 22, fill=0.75)):
        rob = engine.Model(text=vxa.vxa_text(st, sim=sim, env=env), variant=engine.VARIANT_LAND_WATER).robot
        monkeypatch.setenv("VXG_KERNEL_FORM", "push")
        a = engine.Batch([rob, rob]); a.step(900); assert not a.sync().any()
        monkeypatch.setenv("VXG_KERNEL_FORM", "barrier")
        b = engine.Batch([rob, rob])
        for i in range(3):
            b.step(300)
        assert not b.sync().any()
        assert b.kernel_config()["threads"] == 512
        sa, sb = a.download(1), b.download(1)
        for f in ("pos", "quat", "lin_mom", "ang_mom", "scale", "static_fric", "bond_small_angle"):
            assert np.array_equal(getattr(sa, f), getattr(sb, f)), f
        a.close(); b.close()


def test_block_cache_recycles_a_destroyed_batch_without_changing_results():
    """vxg_set_pooling (the resident engine's allocator): batches created from recycled device / pinned blocks -- same sizes, smaller,
    other kernel families -- compute what freshly allocated ones compute (recycled blocks are zero-filled like new ones)."""
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=1.0)
    robots = [engine.Model(text=vxa.vxa_text(vxa.dense_lattice(5, 30 + s), sim=sim)).robot for s in range(3)]
    small = [engine.Model(text=vxa.vxa_text(vxa.sparse_lattice(5, 40, fill=0.5), sim=vxa.SimParams(init_cm_time=0.01, stop_condition_value=1.0, self_col_enabled=1))).robot]

    def run(rs):
        b = engine.Batch(rs); b.step(700); assert not b.sync().any()
        out = [b.download(i).pos.copy() for i in range(len(rs))]
        b.close()
        return out

    ref_a, ref_b = run(robots), run(small)
    engine.set_pooling(True)
    try:
        for _ in range(3):
            for got, want in zip(run(robots), ref_a):
                assert np.array_equal(got, want)
            for got, want in zip(run(small), ref_b):
                assert np.array_equal(got, want)
    finally:
        engine.set_pooling(False)


def test_cli_shards_a_population_over_devices_by_cost(tmp_path):
    """`voxelyze --devices a,b`: robots are dealt out by timesteps x voxels, largest first onto the least loaded device, one population
    per device, all started before any is waited for.  With one GPU in the box the same device is named twice:
    two populations side by side.  Every robot's result file equals the one a single-device run writes -- through the one-shot CLI
    and through the resident engine."""
    from evosoro_b200 import evaluation
    sims = [vxa.SimParams(init_cm_time=0.01, stop_condition_value=0.05 + 0.01 * (i % 3)) for i in range(7)]
    mats = [(1, 3, 4), (1, 2, 3, 4)]          # without / with the stiff material: ten times the timesteps
    runs = {}
    for tag, devs, resident in (("one", (0,), False), ("two", (0, 0), False), ("srv", (0, 0), True)):
        paths = []
        for i, sim in enumerate(sims):
            sim.fitness_file = str(tmp_path / ("%s_%d.xml" % (tag, i)))
            p = tmp_path / ("%s_%d.vxa" % (tag, i))
            p.write_text(vxa.vxa_text(vxa.dense_lattice(3 + i % 3, 60 + i, materials=mats[i % 2]), sim=sim))
            paths.append(str(p))
        assert evaluation.run_batch(paths, devices=devs, resident=resident) == [1]
        runs[tag] = [golden_lib.parse_result((tmp_path / ("%s_%d.xml" % (tag, i))).read_text()) for i in range(len(sims))]
    assert runs["two"] == runs["one"] and runs["srv"] == runs["one"]
    evaluation._resident_engine([0, 0], evaluation.VOXELYZE, []).close()
