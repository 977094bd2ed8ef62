"""GPU parity of the row-tile kernel for the long-pole classes (sweep_rowtile_kernel, the default for plans whose
items have >= 64 poles), deep enough that several tile shapes occur (single-tile classes, subtree tiles with partial
rows, recursion), for 1 / 2 / 4 poles per lane, against the C pole oracle.  Tolerance 1e-12 relative (BASELINE.json
north_star)."""
import numpy as np
import pytest

from helpers import random_state, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _scipy(H):
    import scipy.sparse as sp
    return sp.csc_matrix((H.nzval, H.rowval, H.colptr), shape=(H.m, H.n))


@pytest.fixture(scope="module")
def cb():
    import cbaseline
    return cbaseline


@pytest.mark.parametrize("D,k,n,budget,C", [(5, 3, 7, None, None), (6, 3, 5, None, None), (4, 4, 6, None, None), (5, 3, 6, 60, None),
                                            (6, 3, 6, 112, 2), (6, 3, 6, 150, 1), (4, 5, 5, None, None)])
def test_rowtile_every_axis_beta_and_gradient(gsg, oracle, cb, monkeypatch, D, k, n, budget, C):
    monkeypatch.setenv("GSG_ROWTILE", "1")
    if budget:
        monkeypatch.setenv("GSG_RT_BUDGET_KB", str(budget))
    if C:
        monkeypatch.setenv("GSG_RT_C", str(C))
    H = oracle.periodic_DLF_matrix(k, n)
    plan = gsg.Plan(D, k, n, "sparse", H=_scipy(H))
    plan.set_flat(0)
    assert "rowtile(" in plan.describe(), plan.describe()
    x = random_state(plan.size, seed=D * 10 + n)
    refs = {}
    for d in (1, D // 2 + 1, D):
        refs[d] = cb.apply_D_poles(D, d, k, n, H, x)
        err = relerr(plan.apply_D(d, x), refs[d])
        print(f"rowtile ({D},{k},{n}) d={d}: {err:.3e}")
        assert err <= TOL
    # accumulate form (RED.ADD into y) through the device-pointer entry point
    y0 = random_state(plan.size, seed=99)
    xd, yd = plan.to_device(x), plan.to_device(y0)
    plan.apply_D_dev(D, xd, yd, alpha=-0.5, beta=1.0)
    assert relerr(plan.to_host(yd), y0 - 0.5 * refs[D]) <= TOL
    # gradient through the concurrent right-hand side (phase 1 beta = 0 incl. the zeroed partial rows, then accumulation)
    a = np.linspace(0.5, 1.5, D)
    ref = sum(a[d - 1] * (refs[d] if d in refs else cb.apply_D_poles(D, d, k, n, H, x)) for d in range(1, D + 1))
    assert relerr(plan.apply_grad(a, x), ref) <= TOL
    # a second sweep into the same output (stale partial rows must not survive a beta = 0 sweep)
    assert relerr(plan.apply_D(1, x), refs[1]) <= TOL
    plan.close()
