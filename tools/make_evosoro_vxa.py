#!/usr/bin/env python
"""Write the .vxa files evosoro's own stock experiment (evosoro/examples/basic.py) hands to ./voxelyze for its first population,
with evosoro's OWN code: CPPN genotype -> phenotype mapping (evosoro/softbot.py, networks.py) and its writer
(evosoro/tools/read_write_voxelyze.py:40-401).  They are the "wire format" fixtures of tools/make_golden.py evosoro.

evosoro is Python 2: the package is copied to a scratch directory and made importable under Python 3 there (print statements,
three removed networkx spellings, one str -> bytes for hashlib); nothing of it is copied into this repository and no line that
decides what is written is touched.

    python tools/make_evosoro_vxa.py <out_dir> [pop_size]
"""
import os
import random
import re
import shutil
import sys
import tempfile

import numpy as np

REF = "/root/reference/evosoro"


def scratch_copy():
    root = tempfile.mkdtemp(prefix="evosoro_py3_")
    dst = os.path.join(root, "evosoro")
    shutil.copytree(REF, dst, ignore=shutil.ignore_patterns("_voxcad*", "*.pyc", "data", "examples"))
    for sub in ("softbot.py", "networks.py", "tools"):
        p = os.path.join(dst, sub)
        files = [os.path.join(p, f) for f in os.listdir(p) if f.endswith(".py")] if os.path.isdir(p) else [p]
        for f in files:
            s = open(f).read()
            s = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", s, flags=re.M)
            s = re.sub(r"random\.choice\(self\.graph\.(edges|nodes)\(\)\)", r"random.choice(list(self.graph.\1()))", s)
            s = re.sub(r"for (\w+) in self\.graph\.(nodes|edges)\(\):", r"for \1 in list(self.graph.\2()):", s)
            s = s.replace("m.update(string_for_md5)", "m.update(string_for_md5.encode())")
            open(f, "w").write(s)
    return root


def main(out_dir, pop_size):
    root = scratch_copy()
    sys.path.insert(0, root)
    np.product = np.prod
    import networkx as nx
    nx.Graph.node = property(lambda s: s.nodes)
    nx.Graph.edge = property(lambda s: s.adj)
    from evosoro.base import Sim, Env, ObjectiveDict
    from evosoro.networks import CPPN
    from evosoro.softbot import Genotype, Phenotype, Population
    from evosoro.tools.utils import count_occurrences, make_material_tree
    from evosoro.tools.read_write_voxelyze import write_voxelyze_file

    # The genotype -> phenotype mapping and the validity rule of the experiment are taken from evosoro's own example at run time
    # (the class definitions of examples/basic.py, executed here; the rest of that script launches the GA and is not run).
    src = open(os.path.join(REF, "examples", "basic.py")).read()
    block = src[src.index("class MyGenotype(Genotype):"):src.index("# Setting up the simulation object")]
    ns = dict(Genotype=Genotype, Phenotype=Phenotype, CPPN=CPPN, make_material_tree=make_material_tree, count_occurrences=count_occurrences,
              np=np, IND_SIZE=(6, 6, 6))
    exec(block, ns)
    MyGenotype, MyPhenotype = ns["MyGenotype"], ns["MyPhenotype"]
    random.seed(1); np.random.seed(1)           # SEED = 1, examples/basic.py:78-80

    sim = Sim(dt_frac=0.9, simulation_time=5, fitness_eval_init_time=1)     # examples/basic.py:135
    env = Env(sticky_floor=0, time_between_traces=0)                        # :138
    od = ObjectiveDict()
    od.add_objective(name="fitness", maximize=True, tag="<NormFinalDist>")
    od.add_objective(name="age", maximize=False, tag=None)
    pop = Population(od, MyGenotype, MyPhenotype, pop_size=pop_size)
    os.makedirs(os.path.join(out_dir, "voxelyzeFiles"), exist_ok=True)
    for ind in pop:
        write_voxelyze_file(sim, env, ind, out_dir, "Basic")
    shutil.rmtree(root, ignore_errors=True)
    return sorted(os.path.join(out_dir, "voxelyzeFiles", f) for f in os.listdir(os.path.join(out_dir, "voxelyzeFiles")))


if __name__ == "__main__":
    for p in main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 6):
        print(p)
