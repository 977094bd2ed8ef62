"""Pin the CPU oracle against every known-answer / property assertion the reference's own tests
hold for the hot path.  The reference ships no golden vectors and Julia is
not installed, so these property pins plus the committed fixtures under tests/golden/ (produced by
the oracle itself, see tests/golden/make_golden.py) are what anchor it: "parity unpinned" at the
bit level, as the oracle header says."""
import math
import os

import numpy as np
import pytest

from helpers import f_cos, f_gauss, f_sin, product_state, relerr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_cell_index_table(oracle):
    """test/elementary.jl:17-27."""
    for l in range(1, 6):
        assert oracle.cell_index(1.3, l) == (1 << (l - 1))
        assert oracle.cell_index(0.01, l) == 1
    assert [oracle.cell_index(0.3, l) for l in range(1, 6)] == [1, 1, 2, 3, 5]


def test_get_size_and_configs(oracle):
    """get_size (src/dg_vmethods.jl:35-45) at the BASELINE configs."""
    assert oracle.get_size(1, 3, 5) == 96
    assert oracle.get_size(2, 3, 8) == 11520
    assert oracle.get_size(4, 3, 7) == 327888
    assert oracle.get_size(4, 4, 8) == 2686976
    import cbaseline
    assert cbaseline.get_size(6, 3, 8) == 34455456


@pytest.mark.parametrize("scheme", ["full", "sparse"])
def test_V2D_D2V_round_trip_exact(oracle, scheme):
    """test/vhier_DG.jl:36,44,57,65: V2D(D2V(dict)) == dict and D2V(V2D(vect)) == vect, exactly."""
    rng = np.random.default_rng(1)
    for k in range(1, 5):
        for l in range(1, 4):
            vect = rng.standard_normal(oracle.get_size(2, k, l, scheme))
            d = oracle.V2D(2, k, l, vect, scheme)
            assert np.array_equal(oracle.D2V(2, k, l, d, scheme), vect)
            d2 = oracle.V2D(2, k, l, oracle.D2V(2, k, l, d, scheme), scheme)
            assert all(np.array_equal(d[key], d2[key]) for key in d)


def test_ref_tables_consistent(oracle):
    """D2Vref / V2Dref are inverse of each other and follow get_index_1D in 1-D
    (src/dg_vmethods.jl:102-142, src/additional_tools.jl:18-20)."""
    VD = oracle.V2Dref(2, 2, 3)
    DV = oracle.D2Vref(2, 2, 3)
    assert all(DV[lcm] == j + 1 for j, lcm in enumerate(VD))
    for j, ((l,), (c,), (m,)) in enumerate(oracle.V2Dref(1, 3, 4)):
        assert oracle.get_index_1D(3, l, c, m) == j + 1


def test_basis_orthonormal(oracle):
    """The hierarchical functions v(k, l, c, m) are L2-orthonormal (property behind hier2pos' Q'Q = I)."""
    k, n = 3, 3
    idx = oracle.hier_index_list(k, n)
    nc = 1 << n
    G = np.zeros((len(idx), len(idx)))
    for a, (l1, c1, m1) in enumerate(idx):
        for b, (l2, c2, m2) in enumerate(idx):
            if b < a:
                continue
            s = 0.0
            for cell in range(nc):
                s += oracle.quad(lambda x: oracle.v(k, l1, c1, m1, x) * oracle.v(k, l2, c2, m2, x),
                                 cell / nc, (cell + 1) / nc, 8)
            G[a, b] = G[b, a] = s
    assert np.abs(G - np.eye(len(idx))).max() < 1e-12


def test_hier2pos_orthogonal_and_literal_shortcuts(oracle):
    Q = oracle.hier2pos(3, 3).toarray()
    assert np.abs(Q.T @ Q - np.eye(Q.shape[0])).max() < 1e-13
    # skipped exact-zero pairs: shortcut == full literal loops of the reference
    H_lit = oracle.periodic_DLF_matrix(3, 3, literal=True)
    H_fast = oracle.periodic_DLF_matrix(3, 3)
    assert np.array_equal(H_lit.colptr, H_fast.colptr) and np.array_equal(H_lit.rowval, H_fast.rowval)
    assert np.array_equal(H_lit.nzval, H_fast.nzval)


def test_H_structure_k3_n8(oracle):
    """Facts about periodic_DLF_matrix(3, 8) derived in SURVEY.md (8a a9, A.1)."""
    H = oracle.periodic_DLF_matrix(3, 8)
    Hd = H.toarray()
    assert H.m == 768
    assert abs(np.abs(Hd).max() - 1551.9175) < 1e-3               # 6.06 * 2^8
    assert np.count_nonzero(np.abs(Hd) > 1e-9 * np.abs(Hd).max()) == 25974
    assert 1e-10 < np.abs(Hd + Hd.T).max() < 1e-8                # the `tiny` trick breaks skew symmetry at 1.9e-9


@pytest.mark.parametrize("l", [2, 3, 4, 5])
def test_differentiation_1d_bound(oracle, l):
    """test/differentiation.jl:14-25: D_op * coeffs(cos 2 pi x) reconstructs to -2 pi sin 2 pi x,
    L2 error^2 < 2^-(k+l-2)."""
    k = 3
    H = oracle.periodic_DLF_matrix(k, l)
    dv = H.matvec(oracle.coeffs_1d(k, l, f_cos))
    nc = 1 << l
    err = sum(oracle.quad(lambda x: (oracle.reconstruct_DG(1, k, l, dv, [x]) + 2 * math.pi * math.sin(2 * math.pi * x)) ** 2,
                          c / nc, (c + 1) / nc, 6) for c in range(nc))
    assert err < 1.0 / (1 << (k + l - 2))


@pytest.mark.parametrize("l", [2, 3])
def test_differentiation_2d_full_bound(oracle, l):
    """test/differentiation.jl:29-40: 2-D FULL scheme, axis 1."""
    D, k = 2, 3
    v1 = oracle.coeffs_1d(k, l, f_cos)
    vc = oracle.tensor_construct(D, k, l, [v1, v1], scheme="full")
    dv = oracle.apply_D_poles(D, 1, k, l, vc, scheme="full")
    nc = 1 << l
    xs, ws = np.polynomial.legendre.leggauss(4)
    err = 0.0
    for cx in range(nc):
        for cy in range(nc):
            for xi, wi in zip(xs, ws):
                for yi, wj in zip(xs, ws):
                    x = (cx + 0.5 + 0.5 * xi) / nc
                    y = (cy + 0.5 + 0.5 * yi) / nc
                    val = oracle.reconstruct_DG(D, k, l, dv, [x, y], scheme="full")
                    err += wi * wj * (0.5 / nc) ** 2 * (val + 2 * math.pi * math.sin(2 * math.pi * x) * math.cos(2 * math.pi * y)) ** 2
    assert err < 1.0 / (1 << (k + l - 2))


@pytest.mark.parametrize("k,l", [(1, 3), (2, 4), (3, 3), (4, 2), (5, 2)])
def test_reconstruction_1d_bound(oracle, k, l):
    """test/hier_DG.jl:14-29: L2 error^2 of reconstruct(coeffs(sin 4x)) < 2^-(l+k-1)."""
    vc = oracle.coeffs_1d(k, l, lambda x: math.sin(4 * x))
    nc = 1 << l
    err = sum(oracle.quad(lambda x: (oracle.reconstruct_DG(1, k, l, vc, [x]) - math.sin(4 * x)) ** 2,
                          c / nc, (c + 1) / nc, 8) for c in range(nc))
    assert err < 1.0 / (1 << (l + k - 1))


@pytest.mark.parametrize("k,l", [(1, 4), (2, 3), (2, 6), (3, 5), (4, 3), (5, 2), (5, 4)])
def test_reconstruction_2d_sparse_bound(oracle, k, l):
    """test/hier_DG.jl:59-67: L2 error^2 of the SPARSE 2-D reconstruction of sin(4x + y) < 2^-(l+k-2).  The projection of
    a product function onto the tensor basis is the product of the 1-D projections, so the coefficients are
    tensor_construct(sin 4x, cos y) + tensor_construct(cos 4x, sin y) (the reference integrates the 2-D closure with
    HCubature); the error integral is a tensor Gauss-Legendre rule on the finest cells through the batched reconstruct."""
    D = 2
    c1 = {name: oracle.coeffs_1d(k, l, f) for name, f in
          [("s4", lambda x: math.sin(4 * x)), ("c4", lambda x: math.cos(4 * x)), ("s", math.sin), ("c", math.cos)]}
    vect = oracle.tensor_construct(D, k, l, [c1["s4"], c1["c"]]) + oracle.tensor_construct(D, k, l, [c1["c4"], c1["s"]])
    nc = 1 << l
    gx, gw = np.polynomial.legendre.leggauss(8)
    xs = np.concatenate([(c + 0.5 + 0.5 * gx) / nc for c in range(nc)])
    ws = np.concatenate([0.5 * gw / nc for _ in range(nc)])
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    W = np.outer(ws, ws)
    vals = oracle.reconstruct_DG_batch(D, k, l, vect, np.stack([X.ravel(), Y.ravel()], axis=1))
    err = float(np.sum(W.ravel() * (vals - np.sin(4 * X.ravel() + Y.ravel())) ** 2))
    print(f"2-D sparse reconstruction k={k} l={l}: L2 error^2 {err:.3e} < {1.0 / (1 << (l + k - 2)):.3e}")
    assert err < 1.0 / (1 << (l + k - 2))


def test_wave_energy_2d_sparse(oracle):
    """test/solvers.jl:54-77 (2-D sparse, k=3): sqrt(E) ~ sqrt(2) pi, energy non-increasing and
    conserved to 1e-8 -- here with fixed-step RK4 at n=4 instead of ODE.jl's adaptive ode45/78."""
    D, k, n = 2, 3, 4
    u0 = product_state(oracle, D, k, n, f_sin)
    y0 = np.concatenate([u0, np.zeros_like(u0)])
    mats = [oracle.D_matrix_poles(D, d, k, n) for d in (1, 2)]
    e0 = oracle.energy(mats, y0)
    y1 = oracle.rk4(oracle.wave_rhs(mats), y0, 5e-4, 40)
    e1 = oracle.energy(mats, y1)
    assert abs(math.sqrt(e0) - math.sqrt(2) * math.pi) < 1e-3
    assert abs(e0 - e1) < 1e-8 * e0 * 10


def test_assembly_identities(oracle):
    """literal Dict-loop assembly (src/multidim_derivative.jl:19-58) == per-pole principal
    sub-block identity == C helper, entry for entry."""
    import cbaseline
    for D, k, n, scheme in [(2, 3, 3, "sparse"), (3, 2, 3, "sparse"), (2, 2, 2, "full")]:
        H = oracle.periodic_DLF_matrix(k, n)
        x = np.random.default_rng(0).standard_normal(oracle.get_size(D, k, n, scheme))
        for d in range(1, D + 1):
            A = oracle.D_matrix_literal(D, d, k, n, scheme=scheme, H=H)
            B = oracle.D_matrix_poles(D, d, k, n, scheme=scheme, H=H)
            Cm = cbaseline.D_matrix(D, d, k, n, H, scheme=scheme)
            assert np.array_equal(A.toarray(), B.toarray())
            assert np.array_equal(A.toarray(), Cm.toarray())
            # CSC column scatter == per-pole ascending-column accumulation, bit for bit
            assert np.array_equal(A.matvec(x), oracle.apply_D_poles(D, d, k, n, x, scheme=scheme, H=H))


def test_c_pole_oracle_and_batch_reconstruct(oracle):
    """The C pole apply / slab-assembled CSC apply (oracle/csc_kernels.c) == the literal assembly's column
    scatter, bit for bit; the numpy-batched reconstruct == the scalar restatement, bit for bit."""
    import cbaseline
    for D, k, n, scheme in [(2, 3, 4, "sparse"), (3, 2, 3, "sparse"), (2, 2, 3, "full"), (4, 2, 2, "sparse")]:
        H = oracle.periodic_DLF_matrix(k, n)
        x = np.random.default_rng(1).standard_normal(oracle.get_size(D, k, n, scheme))
        for d in range(1, D + 1):
            lit = oracle.D_matrix_literal(D, d, k, n, scheme=scheme, H=H)
            ref = lit.matvec(x)
            assert np.array_equal(cbaseline.apply_D_poles(D, d, k, n, H, x, scheme=scheme), ref)
            ya, nnz = cbaseline.apply_D_assembled(D, d, k, n, H, x, scheme=scheme, slab_cols=97)
            assert nnz == lit.nnz and np.array_equal(ya, ref)
    for D, k, n, scheme in [(1, 3, 5, "sparse"), (2, 3, 4, "sparse"), (2, 2, 3, "full"), (3, 4, 3, "sparse")]:
        vect = product_state(oracle, D, k, n, f_sin, scheme=scheme) + 0.1 * product_state(oracle, D, k, n, f_gauss, scheme=scheme)
        pts = np.random.default_rng(2).random((24, D))
        pts[0], pts[1], pts[2], pts[3] = 0.0, 1.0, 0.5, 0.25
        a = np.array([oracle.reconstruct_DG(D, k, n, vect, list(p), scheme=scheme) for p in pts])
        assert np.array_equal(a, oracle.reconstruct_DG_batch(D, k, n, vect, pts, scheme=scheme))


def test_laplacian_reference_form(oracle):
    """laplacian_matrix = sum_d D_d*D_d as explicit sparse products (src/multidim_derivative.jl:71-79): the scipy
    product used at mid sizes == the literal Gustavson product, and (D*D)x vs D(Dx) differ only by rounding."""
    D, k, n = 2, 3, 3
    H = oracle.periodic_DLF_matrix(k, n)
    x = np.random.default_rng(4).standard_normal(oracle.get_size(D, k, n))
    mats = [oracle.D_matrix_poles(D, d, k, n, H=H) for d in (1, 2)]
    L = oracle.laplacian_matrix_ref(mats)
    lit = np.zeros_like(x)
    for d in (1, 2):
        A = oracle.D_matrix_literal(D, d, k, n, H=H)
        lit += oracle.spmatmul(A, A).matvec(x)
    assert relerr(L @ x, lit) < 1e-14
    assert relerr(sum(A @ (A @ x) for A in mats), lit) < 1e-13


def test_tensor_construct_matches_projection(oracle):
    """tensor_construct (src/tensor_construct.jl:19-63) of 1-D coefficients == coefficients of the
    product function evaluated by reconstruct at a point."""
    D, k, n = 2, 3, 3
    u = product_state(oracle, D, k, n, f_sin)
    for pt in ([0.3, 0.7], [0.11, 0.52]):
        exact = math.sin(2 * math.pi * pt[0]) * math.sin(2 * math.pi * pt[1])
        assert abs(oracle.reconstruct_DG(D, k, n, u, pt) - exact) < 2e-2      # sparse n=3 truncation


def test_golden_fixtures(oracle):
    """Committed fixtures (tests/golden/*.npz, written by make_golden.py from this oracle): guards
    the oracle against silent drift between rounds."""
    z = np.load(os.path.join(GOLDEN, "oracle_k3.npz"))
    H = oracle.periodic_DLF_matrix(3, 4)
    assert np.array_equal(z["H34_colptr"], H.colptr) and np.array_equal(z["H34_rowval"], H.rowval)
    assert relerr(H.nzval, z["H34_nzval"]) < 1e-15
    u = product_state(oracle, 3, 3, 4, f_sin)
    assert relerr(u, z["u_sin_334"]) < 1e-15
    y = oracle.apply_D_poles(3, 2, 3, 4, u, H=H)
    assert relerr(y, z["D2u_sin_334"]) < 1e-14
    pts = z["pts"]
    r = np.array([oracle.reconstruct_DG(3, 3, 4, u, list(p)) for p in pts])
    assert np.abs(r - z["recon_sin_334"]).max() < 1e-14


# ---------------------------------------------------------------------------------------------------------
# ODE.jl restatement (oracle/ode_oracle.py): tableau order conditions + the reference's own solver assertions
# ---------------------------------------------------------------------------------------------------------
def test_ode_tableaus_order_conditions():
    import ode_oracle as oo
    for bt, (p_main, p_emb) in ((oo.DOPRI5, (5, 4)), (oo.FEH78, (7, 8))):
        a, c = np.array(bt["a"]), np.array(bt["c"])
        assert np.abs(a.sum(axis=1) - c).max() < 1e-14                      # row sums
        assert np.allclose(np.triu(a), 0.0)                                  # explicit
        for row, order in zip(bt["b"], (p_main, p_emb)):
            b = np.array(row)
            # quadrature conditions sum b c^(j-1) = 1/j and the first tree conditions up to order 4
            for j in range(1, order + 1):
                assert abs(b @ c ** (j - 1) - 1.0 / j) < 1e-14, (bt["name"], order, j)
            assert abs(b @ (a @ c) - 1.0 / 6) < 1e-14
            assert abs(b @ (a @ c ** 2) - 1.0 / 12) < 1e-14
            assert abs(b @ (c * (a @ c)) - 1.0 / 8) < 1e-14
            assert abs(b @ (a @ (a @ c)) - 1.0 / 24) < 1e-14
    assert oo.is_fsal(oo.DOPRI5) and not oo.is_fsal(oo.FEH78)


def test_ode_empirical_orders():
    """one embedded step of size h on y' = -y + sin t: local error of the main solution ~ h^(p+1)"""
    import ode_oracle as oo
    F = lambda t, y: -y + np.sin(t)
    def exact(t):           # y(0) = 1
        return 1.5 * np.exp(-t) + 0.5 * (np.sin(t) - np.cos(t))
    for bt, p in ((oo.DOPRI5, 5), (oo.FEH78, 7)):
        errs = []
        for h in (0.4, 0.2):
            ks = [None] * len(bt["c"])
            y0 = np.array([1.0])
            ks[0] = F(0.0, y0)
            yt, _ = oo.rk_embedded_step(ks, y0, F, 0.0, h, bt)
            errs.append(abs(yt[0] - exact(h)))
        slope = math.log2(errs[0] / errs[1])
        assert p + 0.5 < slope < p + 1.8, (bt["name"], slope)


def test_ode_reference_solver_assertions(oracle):
    """test/solvers.jl:19-77 against the restated ode45 / ode78 + the oracle operators: sqrt(E) ~ 2 pi (1e-7) for the
    1-D wave in the position and hierarchical bases, sqrt(E) ~ sqrt(2) pi (1e-4) and an energy drop in (0, 1e-8)
    for the 2-D sparse k=3 n=5 wave, with both integrators."""
    import scipy.sparse as sp

    import ode_oracle as oo
    k, level = 4, 4
    f0 = lambda x: math.sin(2 * math.pi * x)
    v0 = lambda x: 2 * math.pi * math.cos(2 * math.pi * x)

    def csr(M):
        return sp.csc_matrix((M.nzval, M.rowval, M.colptr), shape=(M.m, M.n)).tocsr()

    for basis, order in (("pos", "45"), ("hier", "45"), ("hier", "78")):
        D_op = csr(oracle.periodic_DLF_matrix(k, level, basis))
        if basis == "pos":
            fc, vc = np.array(oracle.pos_vcoeffs_DG(k, level, f0)), np.array(oracle.pos_vcoeffs_DG(k, level, v0))
        else:
            fc, vc = oracle.coeffs_1d(k, level, f0), oracle.coeffs_1d(k, level, v0)
        L = (D_op @ D_op).tocsr()                                   # laplac = *(D_op, D_op), src/pdes.jl:110
        N = fc.size
        F = lambda t, y: np.concatenate([y[N:], L @ y[:N]])         # RHS = [[0 I];[L 0]], src/pdes.jl:22-49
        tout, yout = oo.oderk_adapt(F, np.concatenate([fc, vc]), [0.0, 1.0], oo.TABLEAUS[order])
        assert tout[-1] == 1.0 and len(tout) > 10
        for y in yout:                                              # energy_func_1D, src/pdes.jl:239-255
            E = np.sum((D_op @ y[:N]) ** 2) + np.sum(y[N:] ** 2)
            assert abs(math.sqrt(E) - 2 * math.pi) < 1.0e-7, (basis, order, math.sqrt(E) - 2 * math.pi)
    D, k2, n2 = 2, 3, 5
    mats = [oracle.D_matrix_poles(D, d, k2, n2).tocsr() for d in (1, 2)]
    L = oracle.laplacian_matrix_ref(mats).tocsr()
    u0 = product_state(oracle, D, k2, n2, f_sin)
    N = u0.size
    F = lambda t, y: np.concatenate([y[N:], L @ y[:N]])
    for order in ("78", "45"):
        tout, yout = oo.oderk_adapt(F, np.concatenate([u0, np.zeros(N)]), [0.0, 1.0], oo.TABLEAUS[order])
        E = [sum(np.sum((A @ y[:N]) ** 2) for A in mats) + np.sum(y[N:] ** 2) for y in yout]
        assert 0 < E[0] - E[-1] < 1.0e-8, (order, E[0] - E[-1])
        assert all(abs(math.sqrt(e) - math.sqrt(2) * math.pi) < 1.0e-4 for e in E)


# ---------------------------------------------------------------------------------------------------------
# nodal / point transforms (oracle/nodal_oracle.py): the reference's own assertions, test/transformations.jl:15-79
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 3, 5])
def test_transform_1D_mutual_inverses(k):
    import nodal_oracle as no
    n = 5
    I = np.eye(k << n)
    for a, b in (("points", "nodal"), ("points", "modal"), ("nodal", "modal"), ("pos", "modal")):
        A, B = no.transform_1D(k, n, a, b), no.transform_1D(k, n, b, a)
        assert np.linalg.norm(A @ B - I) < 1e-15 * 10 ** k, (k, a, b)
        assert np.linalg.norm(B @ A - I) < 1e-15 * 10 ** k, (k, a, b)


def test_transform_2D_mutual_inverses_and_literal_loop():
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    import nodal_oracle as no
    D, k, n = 2, 3, 5
    I = sp.identity(3072 if False else no.o.get_size(D, k, n))
    for a, b in (("points", "nodal"), ("nodal", "modal")):
        A = no.transform(D, k, n, no.transform_1D(k, n, a, b))
        B = no.transform(D, k, n, no.transform_1D(k, n, b, a))
        assert spl.norm(A @ B - I) < 1e-10 and spl.norm(B @ A - I) < 1e-10
    # block-Kronecker assembly == the reference's scalar column loop (make_column / inner_loop), entry for entry
    for a, b in (("modal", "nodal"), ("nodal", "points")):
        m1 = no.transform_1D(3, 2, a, b)
        A, B = no.transform(2, 3, 2, m1), no.transform_literal(2, 3, 2, m1)
        assert A.nnz == B.nnz and abs(A - B).max() == 0.0
