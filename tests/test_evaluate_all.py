"""The batched evaluate_all hook (evosoro_b200/evaluation.py) driven the way evosoro's Optimizer drives the
reference's (evosoro/tools/algorithms.py:15, evaluation.py:18-219), with stand-ins for the GA objects and for
evosoro's two I/O functions (the package itself is Python 2 and cannot be imported here).  The stand-in reader
is a line-for-line restatement of read_voxlyze_results (read_write_voxelyze.py:7-37)."""
import hashlib
import os
import types

import numpy as np
import pytest

import golden_lib
from evosoro_b200 import evaluation, vxa


class Log:
    def __init__(self): self.lines = []
    def message(self, s): self.lines.append(s)


class Phenotype:
    def __init__(self, st): self.st = st
    def is_valid(self): return bool((self.st > 0).any())


class Ind:
    def __init__(self, idx, st):
        self.id, self.phenotype, self.fitness, self.md5 = idx, Phenotype(st), -1e9, None
        self.genotype = types.SimpleNamespace(to_phenotype_mapping={})


class Pop(list):
    objective_dict = {0: {"name": "fitness", "tag": "<NormFinalDist>", "worst_value": -1e9, "output_node_name": None, "node_func": None},
                      1: {"name": "touching", "tag": "<NumTouchingFloor>", "worst_value": 0, "output_node_name": None, "node_func": None}}
    gen, pop_size, total_evaluations, best_fit_so_far = 0, 0, 0, -1e9

    def __init__(self, inds):
        super().__init__(inds)
        self.already_evaluated, self.all_evaluated_individuals_ids = {}, []
        self.pop_size = len(inds)


def make_io(sim_params):
    def write_voxelyze_file(sim, env, ind, run_directory, run_name):
        s = vxa.SimParams(**{**sim_params, **getattr(ind, "sim_overrides", {}),
                             "fitness_file": run_directory + "/fitnessFiles/softbotsOutput--id_%05i.xml" % ind.id})
        text = vxa.vxa_text(ind.phenotype.st if ind.phenotype.is_valid() else np.ones((1, 1, 1), dtype=np.int32), sim=s)
        with open(run_directory + "/voxelyzeFiles/" + run_name + "--id_%05i.vxa" % ind.id, "w") as f:
            f.write(text)
        return hashlib.md5(ind.phenotype.st.tobytes()).hexdigest()

    def read_voxlyze_results(population, print_log, filename):
        results = {rank: None for rank in range(len(population.objective_dict))}
        for rank, details in population.objective_dict.items():
            tag = details["tag"]
            if tag is not None:
                for line in open(filename):
                    if tag in line:
                        results[rank] = float(line[line.find(tag) + len(tag):line.find("</" + tag[1:])])
        return results
    return types.SimpleNamespace(write_voxelyze_file=write_voxelyze_file, read_voxlyze_results=read_voxlyze_results)


def test_run_batch_with_nothing_to_do():
    assert evaluation.run_batch([]) == []


@pytest.mark.gpu
def test_evaluate_all_hook_end_to_end(tmp_path):
    run = str(tmp_path)
    for d in ("voxelyzeFiles", "fitnessFiles", "bestSoFar/fitOnly", "Gen_0000", "ancestors"):
        os.makedirs(os.path.join(run, d))
    g = golden_lib.Golden("dense3_full")
    import re
    st = np.zeros((3, 3, 3), dtype=np.int32)
    layers = re.findall(r"CDATA\[(\d+)\]", g.vxa_text)
    for z, row in enumerate(layers):
        for k, ch in enumerate(row):
            st[k % 3, k // 3, z] = int(ch)
    inds = [Ind(0, st), Ind(1, vxa.dense_lattice(3, 5)), Ind(2, st.copy()), Ind(3, np.zeros((3, 3, 3), dtype=np.int32))]
    pop = Pop(inds)
    env = types.SimpleNamespace(actuation_variance=0)
    log = Log()
    io = make_io(dict(init_cm_time=0.05, stop_condition_value=0.25))
    evaluation.evaluate_all(None, env, pop, log, 0, run, "t", io=io)
    ref = golden_lib.parse_result(g.result_xml)
    assert inds[0].fitness == pytest.approx(ref["NormFinalDist"], abs=1e-2)      # same robot as the golden CLI run
    assert inds[0].touching == ref["NumTouchingFloor"]
    assert inds[2].fitness == inds[0].fitness                                    # identical phenotype evaluated in the same batch
    assert inds[3].fitness == -1e9 and "Skipping invalid individual" in log.lines
    assert inds[1].fitness > -1e9 and pop.total_evaluations == 3
    assert not os.listdir(os.path.join(run, "fitnessFiles"))                     # result files consumed like the reference does
    assert len(os.listdir(os.path.join(run, "bestSoFar/fitOnly"))) >= 1
    # second generation: cached phenotypes are not re-simulated
    pop2 = Pop([Ind(10, st)])
    pop2.already_evaluated = pop.already_evaluated
    evaluation.evaluate_all(None, env, pop2, log, 0, run, "t", io=io)
    assert pop2[0].fitness == inds[0].fitness and pop2.total_evaluations == 0


@pytest.mark.gpu
def test_evaluate_all_leaves_the_worst_fitness_on_a_diverged_individual(tmp_path):
    """evaluation.py:107-110, 213-215 of the reference: an individual whose simulation diverges never yields a result file and keeps
    the objective's worst value; it is not cached as evaluated, and the rest of the population is unaffected."""
    run = str(tmp_path)
    for d in ("voxelyzeFiles", "fitnessFiles", "bestSoFar/fitOnly", "Gen_0000", "ancestors"):
        os.makedirs(os.path.join(run, d))
    inds = [Ind(0, vxa.dense_lattice(3, 5)), Ind(1, vxa.dense_lattice(3, 7)), Ind(2, vxa.dense_lattice(3, 6))]
    inds[1].sim_overrides = {"dt_frac": 4.0}          # unstable integration: strain > 100 after ~90 steps
    pop = Pop(inds)
    log = Log()
    io = make_io(dict(init_cm_time=0.01, stop_condition_value=0.1))
    evaluation.evaluate_all(None, types.SimpleNamespace(actuation_variance=0), pop, log, 0, run, "t", io=io)
    assert inds[1].fitness == -1e9 and inds[1].md5 not in pop.already_evaluated
    assert inds[0].fitness > -1e9 and inds[2].fitness > -1e9
    assert pop.all_evaluated_individuals_ids == [0, 2]
    assert any("did not produce a result" in l for l in log.lines)
    assert not os.listdir(os.path.join(run, "fitnessFiles"))


def test_cli_plan_mode_reports_cost_without_a_gpu(tmp_path):
    """`voxelyze --plan`: import only (host code), one "<file> <voxels> <bonds> <timesteps>" line per input -- what cost-balanced
    sharding over several GPUs is based on (timesteps x voxels: a bone-bone bond makes a robot take ten times the steps)."""
    import subprocess
    soft = tmp_path / "soft.vxa"; hard = tmp_path / "hard.vxa"
    soft.write_text(vxa.vxa_text(vxa.dense_lattice(4, 0, materials=(1, 3, 4))))
    hard.write_text(vxa.vxa_text(vxa.dense_lattice(5, 1, materials=(1, 2, 3, 4))))
    r = subprocess.run([evaluation.VOXELYZE, "--plan", "-f", str(soft), "-f", str(tmp_path / "missing.vxa"), "-f", str(hard)], capture_output=True, text=True)
    rows = [l.split() for l in r.stdout.strip().splitlines()]
    assert r.returncode == 0 and "missing.vxa" in r.stderr          # one input could not be imported: reported, the others planned
    assert [row[0] for row in rows] == [str(soft), str(hard)]
    assert [int(row[1]) for row in rows] == [64, 125] and int(rows[1][3]) == 78054 and 9.9 < int(rows[1][3]) / int(rows[0][3]) < 10.1


@pytest.mark.gpu
def test_evaluate_all_through_the_resident_engine(tmp_path):
    """`voxelyze --serve` (evaluation.ResidentEngine): two generations through ONE engine process give the fitness the one-shot
    call gives, a diverging individual of the second generation costs only itself, and the process is still up afterwards."""
    run = str(tmp_path)
    for d in ("voxelyzeFiles", "fitnessFiles", "bestSoFar/fitOnly", "Gen_0000", "ancestors"):
        os.makedirs(os.path.join(run, d))
    env = types.SimpleNamespace(actuation_variance=0)
    io = make_io(dict(init_cm_time=0.01, stop_condition_value=0.1))
    a = Pop([Ind(0, vxa.dense_lattice(3, 5)), Ind(1, vxa.dense_lattice(3, 6))])
    b = Pop([Ind(0, vxa.dense_lattice(3, 5)), Ind(1, vxa.dense_lattice(3, 6))])
    evaluation.evaluate_all(None, env, a, Log(), 0, run, "t", io=io)
    evaluation.evaluate_all(None, env, b, Log(), 0, run, "t", io=io, resident=True)
    assert [i.fitness for i in b] == [i.fitness for i in a]
    eng = evaluation._resident_engine([0], evaluation.VOXELYZE, [])
    pid = eng.proc.pid
    inds = [Ind(10, vxa.dense_lattice(3, 7)), Ind(11, vxa.dense_lattice(3, 8)), Ind(12, vxa.dense_lattice(3, 5))]
    inds[1].sim_overrides = {"dt_frac": 4.0}
    c = Pop(inds); log = Log()
    evaluation.evaluate_all(None, env, c, log, 0, run, "t", io=io, resident=True)
    assert inds[1].fitness == -1e9 and inds[0].fitness > -1e9 and inds[2].fitness == a[0].fitness
    assert eng.proc.pid == pid and eng.proc.poll() is None          # same process, still serving
    eng.close()
    assert eng.proc is None


def test_resident_engine_refuses_to_start_without_a_gpu():
    """No CPU path behind the server either: without a device `voxelyze --serve` exits with a message instead of answering "ready"."""
    from evosoro_b200 import engine
    if engine.device_count() > 0:
        pytest.skip("a GPU is visible")
    eng = evaluation.ResidentEngine()
    with pytest.raises(RuntimeError):
        eng.run(["/nonexistent.vxa"])
