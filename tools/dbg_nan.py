import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import golden_lib
from evosoro_b200 import engine, abi
for name in sys.argv[1:]:
    g = golden_lib.Golden(name)
    m = engine.Model(text=g.vxa_text, variant=engine.VARIANT_LAND_WATER if g.lw else engine.VARIANT_VOXCAD)
    total, latch, eol, t_end = m.step_plan_ex()
    for form in ("push", "barrier"):
        os.environ["VXG_KERNEL_FORM"] = form
        for use_plan in (False, True):
            p = engine.Population([m.robot], want_mesh=g.lw)
            if use_plan:
                p.set_plan(np.array([total]), np.array([latch]), np.array([eol]))
            done, bad = 0, None
            while done < total:
                n = min(500, total - done)
                p.step(n); st = p.sync(); done += n
                pos = p.download(0).pos
                if not np.isfinite(pos).all():
                    bad = done; break
            s = p.stats()[0]
            print(name, form, "plan" if use_plan else "noplan", "batches", p.batches(), "total", total, "latch", latch, "eol", eol, "first non-finite after", bad, "status", st, "ini", list(s.ini_com), "com", list(s.com))
            p.close()
