"""Small sweeps / RK4 / Laplacian / reconstruct for compute-sanitizer (memcheck, racecheck): exercises every kernel
family (stream, constant-bank, register-tiled long with and without column passes, generic, and the flat kernel) at
sizes that finish in seconds.  GSG_SAN_MODES="0,1" selects the sweep paths (0 = tiled, 1 = flat)."""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
modes = [int(m) for m in os.environ.get("GSG_SAN_MODES", "0,1").split(",")]
for mode in modes:
    for (D, k, n) in [(2, 3, 6), (3, 3, 5), (2, 4, 4), (2, 2, 6)]:
        plan = g.Plan(D, k, n)
        plan.set_flat(mode)
        v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
        u0 = g.tensor_construct(D, k, n, [v1] * D)
        for d in range(1, D + 1):
            y = plan.apply_D(d, u0)
        lap = plan.apply_laplacian(u0)
        u = plan.rk4_advect(np.ones(D), u0, 1e-4, 3)
        plan.set_rk4_mode(1)
        u = plan.rk4_advect(np.ones(D), u0, 1e-4, 2)
        uw, vw = plan.rk4_wave(u0, np.zeros_like(u0), 1e-4, 2)
        pts = np.random.default_rng(0).random((64, D))
        r = plan.reconstruct(u0, pts)
        print(mode, D, k, n, float(np.abs(u).max()), float(np.abs(lap).max()), float(np.abs(r).max()), flush=True)
print("done")
