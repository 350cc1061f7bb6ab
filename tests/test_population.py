"""Ragged populations (vxg_pop_*, evosoro_b200/csrc/vxg_population.cu): robots of mixed size and features are bucketed into
homogeneous batches that run concurrently, each in its own kernel family at its own stride, with results bit-identical to
running every robot alone (replaces evosoro's one-process-per-robot fan-out, evaluation.py:89-90)."""
import numpy as np
import pytest

import golden_lib
from evosoro_b200 import abi, engine, vxa

pytestmark = pytest.mark.gpu


def _mixed_population():
    sim = vxa.SimParams(init_cm_time=0.01, stop_condition_value=1.0)
    col = vxa.SimParams(init_cm_time=0.01, stop_condition_value=1.0, self_col_enabled=1)
    pal = vxa.four_material_palette()
    pal[0].density, pal[1].density = 1e6, 3e6            # bonds between voxels of different mass: wide family
    texts = [vxa.vxa_text(vxa.dense_lattice(6, 1), sim=sim),                       # 0  6^3 plain (stride 224, production family)
             vxa.vxa_text(vxa.sparse_lattice(6, 2, fill=0.5), sim=sim),            # 1  108 voxels (stride 128)
             vxa.vxa_text(vxa.dense_lattice(10, 3), sim=sim),                      # 2  10^3 (stride 1024)
             golden_lib.Golden("col_pincer").vxa_text,                             # 3  self-collision with live contacts
             vxa.vxa_text(vxa.dense_lattice(6, 4), sim=sim),                       # 4  6^3 plain again: same bucket as 0
             vxa.vxa_text(vxa.dense_lattice(3, 5), sim=sim, palette=pal),          # 5  wide family
             vxa.vxa_text(vxa.dense_lattice(6, 6), sim=col),                       # 6  6^3 with SelfColEnabled=1
             vxa.vxa_text(vxa.dense_lattice(6, 7), sim=sim)]                       # 7  bucket of 0 and 4
    return [engine.Model(text=t).robot for t in texts]


def test_mixed_population_runs_each_bucket_in_its_own_family_bit_identical_to_solo_runs():
    robots = _mixed_population()
    pop = engine.Population(robots)
    info = pop.batches()
    where = [pop.locate(r) for r in range(len(robots))]
    fam = lambda r: info[where[r][0]]
    # small robots of one family share a batch at the widest stride any of them needs (one kernel image per family) ...
    assert where[0][0] == where[1][0] == where[4][0] == where[7][0] and fam(0)[0] == 4 and fam(0)[1] == 224 and fam(0)[3] == abi.VXG_FAM_COMPACT
    assert where[3][0] == where[6][0] and fam(3)[3] & abi.VXG_FAM_COLLIDE and fam(6)[1] == 224      # ... the self-colliding ones among themselves
    assert fam(2)[0] == 1 and fam(2)[1] == 1024 and fam(2)[2] == 1024       # the 10^3 robot: its own batch on the 1024 x 1 production kernel
    assert fam(5)[0] == 1 and not (fam(5)[3] & abi.VXG_FAM_COMPACT)         # different voxel masses: the wide family, alone
    assert len(info) == 4
    total = np.array([700, 650, 600, 800, 700, 500, 640, 333], dtype=np.int64)
    pop.set_plan(total, np.full(8, 200, dtype=np.int64), None)
    pop.run()
    st = pop.sync()
    assert (st == abi.VXG_ST_FINISHED).all()
    stats = pop.stats()
    for r, rob in enumerate(robots):
        solo = engine.Batch([rob])
        solo.set_step_budget([int(total[r])], [200])
        solo.step(int(total[r]))
        assert solo.sync()[0] == abi.VXG_ST_FINISHED
        a, b = pop.download(r), solo.download(0)
        for f in ("pos", "quat", "lin_mom", "ang_mom", "scale", "static_fric"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (r, f)
        s = solo.stats()[0]
        assert stats[r].steps == total[r] and list(stats[r].com) == list(s.com) and list(stats[r].ini_com) == list(s.ini_com)
        solo.close()
    pop.close()


def test_pipelined_host_buffer_call_matches_the_resident_path():
    """vxg_pop_run_host in chunks of 5 robots (flatten / upload / simulate / download overlapped across chunks) == one
    resident batch stepped the same number of steps; a second call starts from the reset state again."""
    robots = [engine.Model(text=vxa.vxa_text(vxa.dense_lattice(4, s))).robot for s in range(17)]
    pop = engine.Population(robots, chunk_robots=5)
    assert [b[0] for b in pop.batches()] == [5, 4, 4, 4]                    # even chunks
    pos = pop.run_host(250)
    ref = engine.Batch(robots); ref.step(250); ref.sync()
    for r in range(17):
        assert np.array_equal(pos[r, : robots[r].n_vox], ref.download(r).pos)
    pos2 = pop.run_host(250)
    assert np.array_equal(pos, pos2)
    assert pop.h2d_bytes() > 0 and pop.d2h_bytes() >= 17 * 64 * 3 * 8
    pop.close(); ref.close()


def test_population_reports_a_diverged_member_and_finishes_the_rest():
    good = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(3, 5), sim=vxa.SimParams(init_cm_time=0.01, stop_condition_value=0.1))).robot
    bad = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(3, 7), sim=vxa.SimParams(dt_frac=4.0, init_cm_time=0.01, stop_condition_value=0.1))).robot
    pop = engine.Population([good, bad, good])
    pop.set_plan(np.array([400, 400, 400], dtype=np.int64), None, None)
    pop.run()
    st = pop.sync()
    assert st[0] == abi.VXG_ST_FINISHED and st[2] == abi.VXG_ST_FINISHED and st[1] & abi.VXG_ST_DIVERGED
    assert np.array_equal(pop.download(0).pos, pop.download(2).pos)
    pop.close()
