import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
for (D, k, n, scheme) in [(3, 2, 4, "full"), (3, 2, 4, "sparse"), (2, 2, 3, "full")]:
    plan = g.Plan(D, k, n, scheme)
    x = np.random.default_rng(1).standard_normal(plan.size)
    KD = k ** D
    for d in range(1, D + 1):
        plan.set_flat(0); yt = plan.apply_D(d, x)
        plan.set_flat(1); yf = plan.apply_D(d, x)
        bad = np.abs(yt - yf).reshape(-1, KD).max(axis=1) > 1e-9 * np.abs(yt).max()
        idx = np.nonzero(bad)[0]
        print((D, k, n, scheme), "d", d, "ncells", bad.size, "bad cells", idx.size, idx[:12], idx[-5:] if idx.size else "", flush=True)
        if idx.size:
            c = idx[0]
            print("  first bad cell", c, "tiled", yt.reshape(-1, KD)[c], "flat", yf.reshape(-1, KD)[c])
