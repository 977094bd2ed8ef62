import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
D, k, n = 4, 3, 6
plan = g.Plan(D, k, n)
plan.set_flat(1)
x = np.random.default_rng(1).standard_normal(plan.size)
for d in (1, 4):
    y = plan.apply_D(d, x)
    print(d, float(np.abs(y).max()), flush=True)
print("done")
