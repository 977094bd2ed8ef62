"""Writes tests/golden/oracle_k3.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference itself cannot be executed in this image
(Julia is absent) and ships no golden vectors; these fixtures freeze the oracle's own outputs so
later rounds notice any drift."""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import gsg_oracle as o  # noqa: E402

H = o.periodic_DLF_matrix(3, 4)
v1 = o.coeffs_1d(3, 4, lambda x: math.sin(2 * math.pi * x))
u = o.tensor_construct(3, 3, 4, [v1] * 3)
y = o.apply_D_poles(3, 2, 3, 4, u, H=H)
pts = np.random.default_rng(20240).random((16, 3))
r = np.array([o.reconstruct_DG(3, 3, 4, u, list(p)) for p in pts])
np.savez_compressed(os.path.join(HERE, "oracle_k3.npz"), H34_colptr=H.colptr, H34_rowval=H.rowval,
                    H34_nzval=H.nzval, u_sin_334=u, D2u_sin_334=y, pts=pts, recon_sin_334=r)
print("wrote oracle_k3.npz")
