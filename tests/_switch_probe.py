"""Worker of tests/test_gpu_switches.py: one process per environment variant.  Prints one JSON line with the relative
errors of (a) the gradient and a 3-step staged + Taylor RK4 on a plan whose long-pole classes take the row-tile kernel
by default, (b) the Laplacian and a 5-step RK4 (graph replay when the flat path is on) on a small index set, all
against the C oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import scipy.sparse as sp

import cbaseline as cb
import gsg_b200 as g
import gsg_oracle as o
from helpers import f_gauss, f_sin, product_state, relerr


def scipy_of(H):
    return sp.csc_matrix((H.nzval, H.rowval, H.colptr), shape=(H.m, H.n))


errors = {}
# (a) D = 5, k = 3, n = 5: 81 poles per item, classes p = 4, 5 long
D, k, n = 5, 3, 5
H = o.periodic_DLF_matrix(k, n)
plan = g.Plan(D, k, n, "sparse", H=scipy_of(H))
plan.set_flat(0)                 # N * D is small enough for the automatic flat path: this part is about the tiled kernels
desc = plan.describe()
u0 = product_state(o, D, k, n, f_gauss) + 0.3 * product_state(o, D, k, n, f_sin)
a = np.array([1.0, -0.5, 0.25, 2.0, -1.5])
ref = sum(a[d - 1] * cb.apply_D_poles(D, d, k, n, H, u0) for d in range(1, D + 1))
errors["grad"] = relerr(plan.apply_grad(a, u0), ref)
ref_rk = cb.rk4_advect(D, k, n, H, a, u0, 1e-4, 3)
errors["rk4_taylor"] = relerr(plan.rk4_advect(a, u0, 1e-4, 3), ref_rk)
plan.set_rk4_mode(1)
errors["rk4_staged"] = relerr(plan.rk4_advect(a, u0, 1e-4, 3), ref_rk)
plan.close()
# (b) D = 2, k = 3, n = 6: the flat path by default
D, k, n = 2, 3, 6
H = o.periodic_DLF_matrix(k, n)
plan = g.Plan(D, k, n, "sparse", H=scipy_of(H))
u0 = product_state(o, D, k, n, f_gauss)
lap_ref = sum(cb.apply_D_poles(D, d, k, n, H, cb.apply_D_poles(D, d, k, n, H, u0)) for d in (1, 2))
errors["laplacian"] = relerr(plan.apply_laplacian(u0), lap_ref)
a = np.array([1.0, -0.5])
errors["rk4_small"] = relerr(plan.rk4_advect(a, u0, 1e-4, 5), cb.rk4_advect(D, k, n, H, a, u0, 1e-4, 5))
flat = bool(plan.flat_active)
plan.close()
print(json.dumps({"rowtile": "rowtile(" in desc, "flat": flat, "errors": errors}))
