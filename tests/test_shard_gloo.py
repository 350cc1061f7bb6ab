"""Multi-rank host logic on CPU: world_size 2 over gloo (the N>1 path of bench.py uses the same helpers
over NCCL).  Each rank imports its contiguous shard of a small population, advances it with the CPU oracle
(there is no GPU here; the sharding / gather logic is what is under test) and all-gathers the fitness."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_robots, q):
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    import torch.distributed as dist
    import oracle_lib
    from evosoro_b200 import engine, shard, vxa
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_robots, world, rank)
    rows = []
    for r in range(lo, hi):
        m = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(2, r), sim=vxa.SimParams(init_cm_time=0.005)))
        o = oracle_lib.OracleSim(m.robot)
        o.step(50)
        rows.append(np.concatenate([[r], o.com()]))
    local = np.array(rows).reshape(-1, 4)
    allf = shard.gather_fitness(local, n_robots, dist)
    tmax = shard.max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, allf, tmax))


@pytest.mark.parametrize("n_robots", [5, 4])
def test_world2_gloo_shard_and_gather(n_robots):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_robots, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(30)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    a, b = res[0][1], res[1][1]
    assert a.shape == (n_robots, 4) and np.array_equal(a, b)          # every rank holds the whole population's fitness
    assert np.array_equal(a[:, 0], np.arange(n_robots))                # in population order, ragged last shard handled
    assert res[0][2] == res[1][2] == 2.0                               # max over ranks
    # the gathered values equal a single-process evaluation
    sys.path.insert(0, HERE)
    import oracle_lib
    from evosoro_b200 import engine, vxa
    m = engine.Model(text=vxa.vxa_text(vxa.dense_lattice(2, n_robots - 1), sim=vxa.SimParams(init_cm_time=0.005)))
    o = oracle_lib.OracleSim(m.robot); o.step(50)
    assert np.array_equal(a[-1, 1:], o.com())


def test_shard_ranges_cover_population():
    from evosoro_b200 import shard
    for n in (0, 1, 7, 8, 1024, 8191):
        for w in (1, 2, 4, 8):
            spans = [shard.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= (n + w - 1) // w
