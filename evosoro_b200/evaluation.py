"""Batched replacement for evosoro's ``evaluate_all`` (evosoro/tools/evaluation.py:18-219).

The reference fires one ``./voxelyze -f <id>.vxa`` subprocess per individual and polls ``fitnessFiles/`` until
every result file has appeared.  Here the whole population goes to the GPU engine as ONE batch per device
(``evosoro_b200/lib/voxelyze -f a.vxa -f b.vxa ...``, the drop-in CLI of ``csrc/voxelyze_main.cpp``), which
writes exactly the per-individual result files the reference's ``read_voxlyze_results`` parses.  Everything
else -- writing the .vxa files, the md5 cache of already evaluated phenotypes, invalid individuals, objective
assignment, file bookkeeping -- follows the reference step for step, so ``Optimizer(evaluation_func=...)``
(evosoro/tools/algorithms.py:15) can take this function unchanged.

Written to run under evosoro's own interpreter as well (Python 2.7 syntax subset, no third-party imports); the
two evosoro I/O functions are imported lazily from the caller's environment or injected for tests.
"""
from __future__ import print_function

import os
import shutil
import subprocess as sub
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
VOXELYZE = os.path.join(_HERE, "lib", "voxelyze")


class ResidentEngine(object):
    """``voxelyze --serve`` kept alive between generations: the process start-up of the one-shot form -- above all the CUDA context,
    1 - 1.5 s of a 2.6 s generation of 1024 robots -- is paid once per GA run instead of once per generation.  One request = one
    batch, simulated exactly like the one-shot call (same importer, same kernels, same result files)."""

    def __init__(self, devices=(0,), voxelyze=VOXELYZE, extra_args=()):
        self.cmd = [voxelyze, "--serve", "--devices", ",".join(str(d) for d in (list(devices) or [0]))] + list(extra_args)
        self.proc = None

    def _start(self):
        self.proc = sub.Popen(self.cmd, stdin=sub.PIPE, stdout=sub.PIPE, universal_newlines=True, bufsize=1)
        if self.proc.stdout.readline().strip() != "ready":
            self.close()
            raise RuntimeError("voxelyze --serve did not come up (no GPU visible?)")

    def run(self, vxa_paths):
        """Returns 1 when every robot wrote its result file, else 0 (the CLI's exit-code convention)."""
        import tempfile
        vxa_paths = list(vxa_paths)
        if not vxa_paths:
            return 1
        if self.proc is None or self.proc.poll() is not None:
            self._start()
        fd, lst = tempfile.mkstemp(suffix=".lst")
        try:
            with os.fdopen(fd, "w") as f:
                f.write("\n".join(vxa_paths) + "\n")
            self.proc.stdin.write("run " + lst + "\n")
            self.proc.stdin.flush()
            while True:
                line = self.proc.stdout.readline()
                if not line:                      # the engine process died: every robot of the request is without a result
                    self.close()
                    return 0
                f = line.split()
                if f and f[0] == "done":
                    return 1 if int(f[2]) == 0 else 0
        finally:
            os.remove(lst)

    def close(self):
        if self.proc is not None:
            try:
                if self.proc.poll() is None:
                    self.proc.stdin.write("quit\n")
                    self.proc.stdin.flush()
                    self.proc.stdin.close()
                    self.proc.wait()
            except Exception:
                pass
            self.proc = None


_resident = {}


def _resident_engine(devices, voxelyze, extra_args):
    import atexit
    key = (voxelyze, tuple(devices), tuple(extra_args))
    if key not in _resident:
        _resident[key] = ResidentEngine(devices, voxelyze, extra_args)
        atexit.register(_resident[key].close)
    return _resident[key]


def run_batch(vxa_paths, devices=(0,), voxelyze=VOXELYZE, extra_args=(), resident=False):
    """Simulate every .vxa in ``vxa_paths`` (each writes its own <FitnessFileName>) as ONE call of the drop-in CLI: it imports
    the files on the host cores, deals the robots out over ``devices`` by cost (timesteps x voxels, largest first onto the
    least loaded GPU -- not by file count: a robot with a bone-bone bond takes ten times the steps of an all-tissue one) and
    runs each device's share as a ragged population.  Returns the list with the exit code (1 = every robot wrote its
    result, the reference's convention).  ``resident``: hand the batch to a ``voxelyze --serve`` process that stays up until the
    interpreter exits (ResidentEngine) instead of starting a process -- and a CUDA context -- per call."""
    vxa_paths = list(vxa_paths)
    if not vxa_paths:
        return []
    devices = list(devices) or [0]
    if resident:
        return [_resident_engine(devices, voxelyze, extra_args).run(vxa_paths)]
    cmd = [voxelyze, "--devices", ",".join(str(d) for d in devices)] + list(extra_args)
    for p in vxa_paths:
        cmd += ["-f", p]
    return [sub.call(cmd)]


def evaluate_all(sim, env, pop, print_log, save_vxa_every, run_directory, run_name, max_eval_time=60,
                 time_to_try_again=10, save_lineages=False, devices=(0,), io=None, variant="voxcad", extra_args=(), resident=False):
    """Same signature and side effects as evosoro.tools.evaluation.evaluate_all (the two timing arguments are
    accepted and unused: there is nothing to poll for).  ``io`` may carry ``write_voxelyze_file`` and
    ``read_voxlyze_results``; by default they come from evosoro's own ``read_write_voxelyze`` module.

    ``variant`` names the engine fork whose binary the experiment would have copied next to it: ``"voxcad"``
    (evosoro/_voxcad) or ``"land_water"`` (evosoro/_voxcad_land_water: the swimming_* and land/water experiments; adds
    ``--lw``).  A file with <FluidEnvironment>1 handed to the voxcad variant is rejected by the importer, not simulated dry.
    An individual whose simulation produces no result file (diverged, unsupported, unreadable) keeps the objective's worst
    value, as in the reference (evaluation.py:107-110, 213-215).  ``resident``: see run_batch (one engine process per GA run)."""
    if variant not in ("voxcad", "land_water"):
        raise ValueError("variant must be 'voxcad' or 'land_water'")
    extra_args = list(extra_args) + (["--lw"] if variant == "land_water" else [])
    if io is None:
        import read_write_voxelyze as io   # evosoro/tools/read_write_voxelyze.py, on the GA's import path
    start_time = time.time()
    to_run = []

    for ind in pop:
        ind.md5 = io.write_voxelyze_file(sim, env, ind, run_directory, run_name)
        vxa = run_directory + "/voxelyzeFiles/" + run_name + "--id_%05i.vxa" % ind.id

        if not ind.phenotype.is_valid():                                   # evaluation.py:63-68
            for rank, goal in pop.objective_dict.items():
                if goal["name"] != "age":
                    setattr(ind, goal["name"], goal["worst_value"])
            print_log.message("Skipping invalid individual")

        elif env.actuation_variance == 0 and ind.md5 in pop.already_evaluated:   # evaluation.py:70-80
            for rank, goal in pop.objective_dict.items():
                if goal["tag"] is not None:
                    setattr(ind, goal["name"], pop.already_evaluated[ind.md5][rank])
            if save_vxa_every > 0 and pop.gen % save_vxa_every == 0:
                shutil.copy(vxa, run_directory + "/Gen_%04i/" % pop.gen + run_name +
                            "--Gen_%04i--fit_%.08f--id_%05i.vxa" % (pop.gen, ind.fitness, ind.id))
        else:                                                              # evaluation.py:82-90
            pop.total_evaluations += 1
            to_run.append(ind)

    print_log.message("Launched one batched voxelyze call for {0} out of {1} individuals".format(len(to_run), len(pop)))
    run_batch([run_directory + "/voxelyzeFiles/" + run_name + "--id_%05i.vxa" % ind.id for ind in to_run], devices,
              extra_args=extra_args, resident=resident)
    missing = [ind.id for ind in to_run
               if not os.path.exists(run_directory + "/fitnessFiles/softbotsOutput--id_%05i.xml" % ind.id)]
    if missing:
        print_log.message("WARNING: {0} individuals did not produce a result (ids {1}). "
                          "The min fitness was assigned for these individuals".format(len(missing), missing))

    for ind in to_run:                                                     # evaluation.py:128-207
        result = run_directory + "/fitnessFiles/softbotsOutput--id_%05i.xml" % ind.id
        vxa = run_directory + "/voxelyzeFiles/" + run_name + "--id_%05i.vxa" % ind.id
        if not os.path.exists(result):
            continue
        objective_values_dict = io.read_voxlyze_results(pop, print_log, result)
        print_log.message("softbotsOutput--id_%05i.xml fit = %s" % (ind.id, objective_values_dict[0]))
        os.remove(result)
        for rank, details in pop.objective_dict.items():
            if objective_values_dict[rank] is not None:
                setattr(ind, details["name"], objective_values_dict[rank])
            else:
                for name, details_phenotype in ind.genotype.to_phenotype_mapping.items():
                    if name == details["output_node_name"]:
                        setattr(ind, details["name"], details["node_func"](details_phenotype["state"]))
        pop.already_evaluated[ind.md5] = [getattr(ind, details["name"]) for rank, details in pop.objective_dict.items()]
        pop.all_evaluated_individuals_ids += [ind.id]
        if ind.fitness > pop.best_fit_so_far:
            pop.best_fit_so_far = ind.fitness
            shutil.copy(vxa, run_directory + "/bestSoFar/fitOnly/" + run_name +
                        "--Gen_%04i--fit_%.08f--id_%05i.vxa" % (pop.gen, ind.fitness, ind.id))
        if save_lineages:
            shutil.copy(vxa, run_directory + "/ancestors/")
        if save_vxa_every > 0 and pop.gen % save_vxa_every == 0:
            shutil.move(vxa, run_directory + "/Gen_%04i/" % pop.gen + run_name +
                        "--Gen_%04i--fit_%.08f--id_%05i.vxa" % (pop.gen, ind.fitness, ind.id))
        else:
            os.remove(vxa)

    print_log.message("\nAll Voxelyze evals finished in {} seconds".format(time.time() - start_time))
    print_log.message("num_evaluated_this_gen: {0}".format(len(to_run)))
    print_log.message("total_evaluations: {}".format(pop.total_evaluations))
