"""Round-2 kernels under compute-sanitizer: the row-tile kernel (single-tile class and, with GSG_RT_BUDGET_KB=60,
subtree tiles with partial rows; persistent CTAs walking several tiles), through every directional sweep, the
accumulate form, the gradient (concurrent right-hand side) and RK4; then the flat kernel at a small index set."""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
for (D, k, n) in [(5, 3, 4), (5, 3, 5)]:
    plan = g.Plan(D, k, n)
    plan.set_flat(0)
    print(plan.describe().splitlines()[0], flush=True)
    v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
    u0 = g.tensor_construct(D, k, n, [v1] * D)
    for d in (1, 3, D):
        y = plan.apply_D(d, u0)
    gr = plan.apply_grad(np.linspace(0.5, 1.5, D), u0)
    u = plan.rk4_advect(np.ones(D), u0, 1e-4, 2)
    print(D, k, n, float(np.abs(y).max()), float(np.abs(gr).max()), float(np.abs(u).max()), flush=True)
    plan.close()
plan = g.Plan(2, 3, 6)
plan.set_flat(1)
v1 = g.vcoeffs_DG(1, 3, 6, lambda x: math.sin(2 * math.pi * x))
u0 = g.tensor_construct(2, 3, 6, [v1] * 2)
print("flat", float(np.abs(plan.apply_laplacian(u0)).max()), float(np.abs(plan.rk4_advect(np.ones(2), u0, 1e-4, 2)).max()))
print("done")
