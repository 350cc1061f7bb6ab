"""The .vxa importer (vxh_*, evosoro_b200/csrc/vxh_import.cpp) against the reference's own Import, pinned
by the golden fixtures' dump headers (the arrays the unmodified reference built from the same .vxa text)."""
import numpy as np
import pytest

import golden_lib
from evosoro_b200 import abi, engine, vxa


def _expand_v(rob):
    return [rob.vclasses[c] for c in rob.vox_class]


def _expand_b(rob):
    return [rob.bclasses[c] for c in rob.bond_class]


@pytest.mark.parametrize("name", golden_lib.names())
def test_importer_matches_reference_import(name):
    g = golden_lib.Golden(name)
    ref = g.robot()
    m = engine.Model(text=g.vxa_text, variant=engine.VARIANT_LAND_WATER if g.lw else engine.VARIANT_VOXCAD)
    r = m.robot
    assert r.n_vox == ref.n_vox and r.n_bond == ref.n_bond
    assert np.array_equal(r.nom_pos, ref.nom_pos)                      # bit-exact lattice positions
    assert np.array_equal(r.bond_v1, ref.bond_v1) and np.array_equal(r.bond_v2, ref.bond_v2)
    assert np.array_equal(r.bond_axis, ref.bond_axis)
    assert np.array_equal(m.cell_index(), g.dump.vox[:, 20].astype(np.int32))   # StoXIndexMap
    for a, b in zip(_expand_v(r), _expand_v(ref)):
        for f, _ in abi.VoxClass._fields_:
            va, vb = getattr(a, f), getattr(b, f)
            if hasattr(va, "__len__"):
                va, vb = list(va), list(vb)
            assert va == vb, (f, va, vb)
    for a, b in zip(_expand_b(r), _expand_b(ref)):
        for f, _ in abi.BondClass._fields_:
            va, vb = getattr(a, f), getattr(b, f)
            if f == "nom_dist":
                assert abs(va - vb) <= 1e-15
            else:
                assert va == vb, (f, va, vb)
    for f, _ in abi.SimParams._fields_:
        assert getattr(r.sim, f) == getattr(ref.sim, f), f
    # dt = DtFrac * OptimalDt exactly as the reference computed it (CalcMaxDt)
    assert r.sim.dt == g.dump.sim[0] * g.dump.sim[1]


def test_step_plan_replays_cli_loop():
    g = golden_lib.Golden("dense3_full")
    m = engine.Model(text=g.vxa_text)
    total, latch, t_end = m.step_plan()
    assert total == int(g.full_tail[0])
    assert t_end == g.full_hdr[-1, 0]          # CurTime, bit-exact fp64 accumulation
    assert 0 < latch < total


def test_class_deduplication_is_compact():
    m = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(6, 1)))
    assert len(m.robot.vclasses) == 4 and len(m.robot.bclasses) <= 16
    assert m.info.n_vox == 216 and m.info.n_bond == 540


@pytest.mark.parametrize("edit,code", [
    (lambda t: t.replace("<EquilibriumModeEnabled>0", "<EquilibriumModeEnabled>1"), 1),
    (lambda t: t.replace("<FluidDampEnabled>0</FluidDampEnabled>", "<VolumeEffectsEnabled>1</VolumeEffectsEnabled>"), 1),
    (lambda t: t.replace("<MatModel>0", "<MatModel>2"), 1),
    (lambda t: t.replace("<NumFixed>0", "<NumFixed>1"), -1),                                   # count without its <FRegion>
    (lambda t: t.replace("<NumFixed>0</NumFixed>", "<NumFixed>1</NumFixed><FRegion><PrimType>3</PrimType></FRegion>"), 1),   # mesh region
    (lambda t: t.replace('Compression="ASCII_READABLE"', 'Compression="QT_ZLIB"'), 1),         # Qt's deprecated encoding (refused upstream too)
    (lambda t: t.replace('Compression="ASCII_READABLE"', 'Compression="ZLIB"'), -1),           # digits are not deflated data
    (lambda t: t.replace("</VXA>", ""), -1),
    (lambda t: t.replace("<X_Voxels>3", "<X_Voxels>4"), -1),
])
def test_unsupported_or_malformed_input_is_rejected(edit, code):
    """No CPU fallback exists, so anything outside the implemented path must be an error, not a silent difference."""
    text = edit(vxa.vxa_text(vxa.dense_lattice(3, 1)))
    with pytest.raises(engine.EngineError) as ei:
        engine.Model(text=text)
    assert ei.value.code == code, str(ei.value)


def test_selfcol_can_be_switched_off_like_enablefeature():
    text = vxa.vxa_text(vxa.dense_lattice(3, 1), sim=vxa.SimParams(self_col_enabled=1))
    m = engine.Model(text=text, ignore_collisions=True)
    assert m.info.self_col_enabled == 1 and not (m.robot.sim.flags & abi.VXG_F_COLLISIONS)
    m = engine.Model(text=text)
    assert (m.robot.sim.flags & abi.VXG_F_COLLISIONS) and m.robot.sim.col_horizon == 2.0


def test_empty_structure_is_an_error():
    with pytest.raises(engine.EngineError):
        engine.Model(text=vxa.vxa_text(np.zeros((2, 2, 2), dtype=np.int32)))


def test_file_and_string_entry_points_agree(tmp_path):
    text = vxa.vxa_text(vxa.sparse_lattice(4, 2))
    p = tmp_path / "r.vxa"
    p.write_text(text)
    a, b = engine.Model(path=str(p)).robot, engine.Model(text=text).robot
    assert np.array_equal(a.nom_pos, b.nom_pos) and np.array_equal(a.bond_class, b.bond_class)


def test_boundary_conditions_legacy_and_current_layouts_agree():
    """<Fixed_Regions>/<Forced_Regions> (legacy, VX_Environment.cpp:140-169) and <Boundary_Conditions> give the same model;
    a pre-DofFixed <Fixed>1</Fixed> means all six DOFs."""
    st = vxa.dense_lattice(3, 2)
    cur = vxa.vxa_text(st, boundary_conditions=[vxa.Region(0, (0, 0, 0), (0.3, 1, 1), dof_fixed=0x3F),
                                                vxa.Region(0, (0.7, 0, 0), (0.3, 1, 1), force=(0, 0, -1.0))])
    reg = lambda body: f"<FRegion><PrimType>0</PrimType>{body}</FRegion>"
    fixed = reg("<X>0</X><Y>0</Y><Z>0</Z><dX>0.3</dX><dY>1</dY><dZ>1</dZ><Fixed>1</Fixed>")
    forced = reg("<X>0.7</X><Y>0</Y><Z>0</Z><dX>0.3</dX><dY>1</dY><dZ>1</dZ><ForceZ>-1.0</ForceZ>")
    old = vxa.vxa_text(st).replace("<NumFixed>0</NumFixed>", "<NumFixed>1</NumFixed>" + fixed).replace("<NumForced>0</NumForced>", "<NumForced>1</NumForced>" + forced)
    a, b = engine.Model(text=cur).robot, engine.Model(text=old).robot
    assert [c.dof_fixed for c in _expand_v(a)] == [c.dof_fixed for c in _expand_v(b)]
    assert [list(c.ext_force) for c in _expand_v(a)] == [list(c.ext_force) for c in _expand_v(b)]
    dof = np.array([c.dof_fixed for c in _expand_v(a)])
    assert (dof == 0x3F).sum() == 9 and a.n_bond == 54 - 12      # the 12 bonds inside the clamped face are not created
    fz = np.array([c.ext_force[2] for c in _expand_v(a)])
    assert (fz != 0).sum() == 9 and abs(fz.sum() + 1.0) < 1e-15   # the region's force is shared by the voxels it touches


def test_fluid_environment_file_is_rejected_by_the_voxcad_variant():
    """_voxcad has no fluid environment: a swimming experiment's file imported with the wrong fork would be simulated dry and
    yield a plausible but wrong fitness.  The importer rejects it; the land/water variant takes it."""
    text = vxa.vxa_text(vxa.dense_lattice(2, 3), env=vxa.EnvParams(fluid_environment=1, aggregate_drag_coefficient=750.0))
    with pytest.raises(engine.EngineError) as e:
        engine.Model(text=text, variant=engine.VARIANT_VOXCAD)
    assert e.value.code == 1 and "land/water" in str(e.value)   # 1 = VXG_UNSUPPORTED_FEATURE
    m = engine.Model(text=text, variant=engine.VARIANT_LAND_WATER)
    assert m.robot.sim.flags & abi.VXG_F_FLUID
    dry = vxa.vxa_text(vxa.dense_lattice(2, 3), env=vxa.EnvParams(fluid_environment=0))
    assert not (engine.Model(text=dry, variant=engine.VARIANT_VOXCAD).robot.sim.flags & abi.VXG_F_FLUID)


def test_structure_layer_encodings_decode_to_the_same_robot():
    """CVXC_Structure::ReadXML (VX_Object.cpp:1755-1798): ASCII_READABLE, BASE64 and ZLIB layers of the same lattice import to the
    same model (the reference ships one ZLIB input, voxelyzeMain/Example_stickyFloorBC.vxa: fixture ex_sticky_floor_bc)."""
    import base64, re, zlib
    text = vxa.vxa_text(vxa.dense_lattice(4, 3))
    ref = engine.Model(text=text).robot

    def recode(enc, how):
        t = text.replace('Compression="ASCII_READABLE"', 'Compression="%s"' % enc)
        return re.sub(r"<Layer><!\[CDATA\[([0-9]+)\]\]></Layer>", lambda m: "<Layer><![CDATA[" + how(bytes(int(c) for c in m.group(1))) + "]]></Layer>", t)

    for enc, how in (("BASE64", lambda b: base64.b64encode(b).decode()), ("ZLIB", lambda b: base64.b64encode(zlib.compress(b)).decode())):
        t = recode(enc, how)
        assert t != text.replace("ASCII_READABLE", enc)   # the layers were rewritten
        r = engine.Model(text=t).robot
        assert r.n_vox == ref.n_vox and r.n_bond == ref.n_bond
        assert np.array_equal(np.asarray(r.vox_class), np.asarray(ref.vox_class)) and np.array_equal(np.asarray(r.nom_pos), np.asarray(ref.nom_pos))
