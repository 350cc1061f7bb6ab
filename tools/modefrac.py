import sys; sys.path.insert(0,'/root/repo')
import numpy as np, bench
from evosoro_b200 import engine
robots=[engine.Model(text=bench.synthetic_vxa(10, s)).robot for s in range(32)]
b=engine.Batch(robots)
for tgt in (300, 1500, 3000, 5000, 8000, 12000):
    st=b.stats()[0].steps
    b.step(tgt-st); b.sync()
    fr=[]; wf=[]
    for r in range(32):
        s=b.download(r)
        sm=s.bond_small_angle[:robots[r].n_bond]
        fr.append(1-sm.mean())
    print(tgt, "large-angle fraction mean %.3f min %.3f max %.3f" % (np.mean(fr), np.min(fr), np.max(fr)))
