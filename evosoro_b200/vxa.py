"""Writer for Voxelyze ``.vxa`` input files (the wire format between evosoro's GA and the engine).

Emits the same XML layout as the reference's writer
(``/root/reference/evosoro/tools/read_write_voxelyze.py:53-401``): ``<Simulator>`` /
``<Environment>`` / ``<VXC>`` with an ``ASCII_READABLE`` structure whose ``<Layer>`` lines hold one
material digit per cell, x fastest, z ascending.  The same text feeds the reference CPU engine
(``voxelyze -f``) and this repo's importer, which is how the parity tests and the benchmark make sure
both engines see identical inputs.

Only the generator lives here -- no simulation logic.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence

import numpy as np


@dataclass
class Material:
    """One ``<Material>`` of the palette (reference: read_write_voxelyze.py:175-344)."""
    name: str
    elastic_mod: float
    cte: float = 0.0
    density: float = 1e6
    poisson: float = 0.35
    u_static: float = 1.0
    u_dynamic: float = 0.5
    mat_model: int = 0


def evosoro_palette(fat=5e6, bone=5e8, muscle=5e6) -> List[Material]:
    """The 7-material palette evosoro hard-codes (read_write_voxelyze.py:175-344)."""
    return [
        Material("Passive_Soft", fat),
        Material("Passive_Hard", bone),
        Material("Active_+", muscle, cte=0.01),
        Material("Active_-", muscle, cte=-0.01),
        Material("Obstacle", 5e7),
        Material("Head_Active_+", fat, cte=0.01),
        Material("Food", muscle),
    ]


def four_material_palette() -> List[Material]:
    """The reduced palette of SURVEY.md Appendix C (IDs 1-4 of the evosoro palette)."""
    return evosoro_palette()[:4]


@dataclass
class SimParams:
    """``<Simulator>`` block (reference defaults: evosoro/base.py:24-27, read_write_voxelyze.py:62-121)."""
    dt_frac: float = 0.9
    bond_damping_z: float = 1.0
    col_damping_z: float = 0.8
    slow_damping_z: float = 0.01
    self_col_enabled: int = 0
    col_system: int = 3
    collision_horizon: float = 2
    stop_condition_type: int = 2
    stop_condition_value: float = 5.0
    afterlife_time: float = 0.0
    mid_life_freeze_time: float = 0.0
    init_cm_time: float = 1.0
    equilibrium_mode: int = 0
    min_temp_fact: float = 0.1
    fitness_file: str = ""
    extra_tags: dict = field(default_factory=dict)


@dataclass
class EnvParams:
    """``<Environment>`` block (reference defaults: evosoro/base.py:53-55)."""
    grav_enabled: int = 1
    grav_acc: float = -9.81
    floor_enabled: int = 1
    temp_enabled: int = 1
    temp_amp: float = 39.0
    temp_base: float = 25.0
    vary_temp_enabled: int = 1
    temp_period: float = 0.25
    sticky_floor: int = 0
    lattice_dim: float = 0.01
    growth_amplitude: Optional[float] = None
    fluid_environment: Optional[int] = None            # land/water fork only
    aggregate_drag_coefficient: Optional[float] = None  # land/water fork only
    extra_tags: dict = field(default_factory=dict)


@dataclass
class Region:
    """One ``<FRegion>`` boundary condition (reference: VX_FRegion.cpp:122-163).  Position / size are fractions
    of the workspace; ``dof_fixed`` is the bit mask X=1 Y=2 Z=4 TX=8 TY=16 TZ=32 (VX_FRegion.h:30-37)."""
    prim_type: int = 0                      # 0 box, 1 cylinder, 2 sphere
    pos: Sequence[float] = (0.0, 0.0, 0.0)
    size: Sequence[float] = (0.0, 0.0, 0.0)
    radius: float = 0.0
    dof_fixed: int = 0
    force: Sequence[float] = (0.0, 0.0, 0.0)
    torque: Sequence[float] = (0.0, 0.0, 0.0)
    displace: Sequence[float] = (0.0, 0.0, 0.0)
    ang_displace: Sequence[float] = (0.0, 0.0, 0.0)


def _fmt(v) -> str:
    if isinstance(v, (bool, np.bool_)):
        return str(int(v))
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    return repr(float(v))


def vxa_text(structure: np.ndarray, palette: Sequence[Material] = None, sim: SimParams = None,
             env: EnvParams = None, phase_offset: Optional[np.ndarray] = None,
             boundary_conditions: Optional[Sequence[Region]] = None, per_voxel: Optional[dict] = None) -> str:
    """Return the ``.vxa`` text for one robot.

    ``structure`` is an integer array indexed ``[x, y, z]`` of material IDs (0 = empty), the same
    orientation evosoro's phenotype arrays use (read_write_voxelyze.py:380-386).
    ``phase_offset`` (same shape, floats) is written as a ``<PhaseOffset>`` block when given.
    ``boundary_conditions`` (fixed / forced regions) are written as a ``<Boundary_Conditions>`` block; evosoro
    itself always writes the empty legacy ``<Fixed_Regions>`` / ``<Forced_Regions>`` pair.
    ``per_voxel`` maps further per-voxel block names (``TempAmpDamp``, ``FinalPhaseOffset``, ``InitialVoxelSize``,
    ``FinalVoxelSize``, ``GrowthTime``, ``StartGrowthTime``, ``Stiffness`` ...; reference: read_write_voxelyze.py:403-520,
    VX_Object.cpp:1879-2322) to float arrays of the structure's shape, written like ``<PhaseOffset>``.
    """
    palette = list(palette) if palette is not None else four_material_palette()
    sim = sim or SimParams()
    env = env or EnvParams()
    st = np.asarray(structure)
    nx, ny, nz = st.shape
    o: List[str] = []
    w = o.append
    w('<?xml version="1.0" encoding="ISO-8859-1"?>')
    w('<VXA Version="1.0">')
    w("<Simulator>")
    for tag, val in sim.extra_tags.items():
        w(f"<{tag}>{_fmt(val)}</{tag}>")
    w("<Integration>")
    w("<Integrator>0</Integrator>")
    w(f"<DtFrac>{_fmt(sim.dt_frac)}</DtFrac>")
    w("</Integration>")
    w("<Damping>")
    w(f"<BondDampingZ>{_fmt(sim.bond_damping_z)}</BondDampingZ>")
    w(f"<ColDampingZ>{_fmt(sim.col_damping_z)}</ColDampingZ>")
    w(f"<SlowDampingZ>{_fmt(sim.slow_damping_z)}</SlowDampingZ>")
    w("</Damping>")
    w("<Collisions>")
    w(f"<SelfColEnabled>{int(sim.self_col_enabled)}</SelfColEnabled>")
    w(f"<ColSystem>{int(sim.col_system)}</ColSystem>")
    w(f"<CollisionHorizon>{_fmt(sim.collision_horizon)}</CollisionHorizon>")
    w("</Collisions>")
    w("<Features>")
    w("<FluidDampEnabled>0</FluidDampEnabled>")
    w("<PoissonKickBackEnabled>0</PoissonKickBackEnabled>")
    w("<EnforceLatticeEnabled>0</EnforceLatticeEnabled>")
    w("</Features>")
    w("<StopCondition>")
    w(f"<StopConditionType>{int(sim.stop_condition_type)}</StopConditionType>")
    w(f"<StopConditionValue>{_fmt(sim.stop_condition_value)}</StopConditionValue>")
    w(f"<AfterlifeTime>{_fmt(sim.afterlife_time)}</AfterlifeTime>")
    w(f"<MidLifeFreezeTime>{_fmt(sim.mid_life_freeze_time)}</MidLifeFreezeTime>")
    w(f"<InitCmTime>{_fmt(sim.init_cm_time)}</InitCmTime>")
    w("</StopCondition>")
    w("<EquilibriumMode>")
    w(f"<EquilibriumModeEnabled>{int(sim.equilibrium_mode)}</EquilibriumModeEnabled>")
    w("</EquilibriumMode>")
    w("<GA>")
    w(f"<WriteFitnessFile>{1 if sim.fitness_file else 0}</WriteFitnessFile>")
    if sim.fitness_file:  # an empty element crashes the reference loader (XML_Rip FindLoadElement on NULL text)
        w(f"<FitnessFileName>{sim.fitness_file}</FitnessFileName>")
    w("</GA>")
    w(f"<MinTempFact>{_fmt(sim.min_temp_fact)}</MinTempFact>")
    w("</Simulator>")
    w("<Environment>")
    for tag, val in env.extra_tags.items():
        w(f"<{tag}>{_fmt(val)}</{tag}>")
    if env.fluid_environment is not None:
        w(f"<FluidEnvironment>{int(env.fluid_environment)}</FluidEnvironment>")
    if env.aggregate_drag_coefficient is not None:
        w(f"<AggregateDragCoefficient>{_fmt(env.aggregate_drag_coefficient)}</AggregateDragCoefficient>")
    if boundary_conditions:
        w("<Boundary_Conditions>")
        w(f"<NumBCs>{len(boundary_conditions)}</NumBCs>")
        for g in boundary_conditions:
            w("<FRegion>")
            w(f"<PrimType>{int(g.prim_type)}</PrimType>")
            for t, v in zip(("X", "Y", "Z"), g.pos):
                w(f"<{t}>{_fmt(float(v))}</{t}>")
            for t, v in zip(("dX", "dY", "dZ"), g.size):
                w(f"<{t}>{_fmt(float(v))}</{t}>")
            w(f"<Radius>{_fmt(float(g.radius))}</Radius>")
            w("<R>0.4</R><G>0.6</G><B>0.4</B><alpha>1</alpha>")
            w(f"<DofFixed>{int(g.dof_fixed)}</DofFixed>")
            for pre, vec in (("Force", g.force), ("Torque", g.torque), ("Displace", g.displace), ("AngDisplace", g.ang_displace)):
                for t, v in zip("XYZ", vec):
                    w(f"<{pre}{t}>{_fmt(float(v))}</{pre}{t}>")
            w("</FRegion>")
        w("</Boundary_Conditions>")
    else:
        w("<Fixed_Regions>")
        w("<NumFixed>0</NumFixed>")
        w("</Fixed_Regions>")
        w("<Forced_Regions>")
        w("<NumForced>0</NumForced>")
        w("</Forced_Regions>")
    w("<Gravity>")
    w(f"<GravEnabled>{int(env.grav_enabled)}</GravEnabled>")
    w(f"<GravAcc>{_fmt(env.grav_acc)}</GravAcc>")
    w(f"<FloorEnabled>{int(env.floor_enabled)}</FloorEnabled>")
    w("</Gravity>")
    w("<Thermal>")
    w(f"<TempEnabled>{int(env.temp_enabled)}</TempEnabled>")
    w(f"<TempAmp>{_fmt(env.temp_amp)}</TempAmp>")
    w(f"<TempBase>{_fmt(env.temp_base)}</TempBase>")
    w(f"<VaryTempEnabled>{int(env.vary_temp_enabled)}</VaryTempEnabled>")
    w(f"<TempPeriod>{_fmt(env.temp_period)}</TempPeriod>")
    w("</Thermal>")
    if env.growth_amplitude is not None:
        w(f"<GrowthAmplitude>{_fmt(env.growth_amplitude)}</GrowthAmplitude>")
    w(f"<StickyFloor>{int(env.sticky_floor)}</StickyFloor>")
    w("</Environment>")
    w('<VXC Version="0.93">')
    w("<Lattice>")
    w(f"<Lattice_Dim>{_fmt(env.lattice_dim)}</Lattice_Dim>")
    for t in ("X_Dim_Adj", "Y_Dim_Adj", "Z_Dim_Adj"):
        w(f"<{t}>1</{t}>")
    for t in ("X_Line_Offset", "Y_Line_Offset", "X_Layer_Offset", "Y_Layer_Offset"):
        w(f"<{t}>0</{t}>")
    w("</Lattice>")
    w("<Voxel>")
    w("<Vox_Name>BOX</Vox_Name>")
    for t in ("X_Squeeze", "Y_Squeeze", "Z_Squeeze"):
        w(f"<{t}>1</{t}>")
    w("</Voxel>")
    w("<Palette>")
    for i, m in enumerate(palette):
        w(f'<Material ID="{i + 1}">')
        w("<MatType>0</MatType>")
        w(f"<Name>{m.name}</Name>")
        w("<Display><Red>0</Red><Green>1</Green><Blue>1</Blue><Alpha>1</Alpha></Display>")
        w("<Mechanical>")
        w(f"<MatModel>{int(m.mat_model)}</MatModel>")
        w(f"<Elastic_Mod>{_fmt(m.elastic_mod)}</Elastic_Mod>")
        w("<Plastic_Mod>0</Plastic_Mod>")
        w("<Yield_Stress>0</Yield_Stress>")
        w("<FailModel>0</FailModel>")
        w("<Fail_Stress>0</Fail_Stress>")
        w("<Fail_Strain>0</Fail_Strain>")
        w(f"<Density>{_fmt(m.density)}</Density>")
        w(f"<Poissons_Ratio>{_fmt(m.poisson)}</Poissons_Ratio>")
        w(f"<CTE>{_fmt(m.cte)}</CTE>")
        w(f"<uStatic>{_fmt(m.u_static)}</uStatic>")
        w(f"<uDynamic>{_fmt(m.u_dynamic)}</uDynamic>")
        w("</Mechanical>")
        w("</Material>")
    w("</Palette>")
    w('<Structure Compression="ASCII_READABLE">')
    w(f"<X_Voxels>{nx}</X_Voxels>")
    w(f"<Y_Voxels>{ny}</Y_Voxels>")
    w(f"<Z_Voxels>{nz}</Z_Voxels>")
    w("<Data>")
    for z in range(nz):
        layer = "".join(str(int(st[x, y, z])) for y in range(ny) for x in range(nx))
        w(f"<Layer><![CDATA[{layer}]]></Layer>")
    w("</Data>")
    blocks = {} if phase_offset is None else {"PhaseOffset": phase_offset}
    blocks.update(per_voxel or {})
    for tag, arr in blocks.items():
        po = np.asarray(arr, dtype=float)
        w(f"<{tag}>")
        for z in range(nz):
            layer = "".join(f"{float(po[x, y, z])!r}, " for y in range(ny) for x in range(nx))   # plain decimal text: the engines read it with atof
            w(f"<Layer><![CDATA[{layer}]]></Layer>")
        w(f"</{tag}>")
    w("</Structure>")
    w("</VXC>")
    w("</VXA>")
    return "\n".join(o) + "\n"


def dense_lattice(n: int, seed: int, materials: Iterable[int] = (1, 2, 3, 4)) -> np.ndarray:
    """Dense n^3 lattice of SURVEY.md Appendix C: ``random.seed(seed)`` once, then one
    ``random.choice`` per cell in file order (z slowest, x fastest).  Returned indexed [x, y, z]."""
    rng = random.Random(seed)
    mats = list(materials)
    st = np.zeros((n, n, n), dtype=np.int32)
    for z in range(n):
        for y in range(n):
            for x in range(n):
                st[x, y, z] = rng.choice(mats)
    return st


def sparse_lattice(n: int, seed: int, fill: float = 0.6, materials: Iterable[int] = (1, 2, 3, 4)) -> np.ndarray:
    """A connected CPPN-like blob inside an n^3 box (robustness inputs):
    cells are grown from the centre by randomised flood fill until ``fill`` of the box is occupied."""
    rng = random.Random(seed)
    mats = list(materials)
    st = np.zeros((n, n, n), dtype=np.int32)
    target = max(1, int(fill * n ** 3))
    c = (n // 2, n // 2, 0)
    st[c] = rng.choice(mats)
    frontier = [c]
    count = 1
    while count < target and frontier:
        x, y, z = frontier[rng.randrange(len(frontier))]
        nbrs = [(x + 1, y, z), (x - 1, y, z), (x, y + 1, z), (x, y - 1, z), (x, y, z + 1), (x, y, z - 1)]
        free = [p for p in nbrs if all(0 <= p[i] < n for i in range(3)) and st[p] == 0]
        if not free:
            frontier.remove((x, y, z))
            continue
        p = free[rng.randrange(len(free))]
        st[p] = rng.choice(mats)
        frontier.append(p)
        count += 1
    return st
