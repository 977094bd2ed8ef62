"""GPU parity tests: the CUDA path (through the C ABI of libgsgb200.so) against the CPU oracle
on the same seeded inputs.  Tolerance: 1e-12 relative in Float64 (BASELINE.json north_star),
measured as ||y_gpu - y_oracle||_2 / ||y_oracle||_2."""
import math

import numpy as np
import pytest

from helpers import f_cos, f_gauss, f_sin, product_state, random_state, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-12

# (D, k, n, scheme): small enough for the oracle to finish in seconds; covers every kernel
# class (register-resident short poles, generic block-CSR long poles, run-time-k fallback),
# both schemes, k = 1..6 and D = 1..6.
CASES = [
    (1, 3, 5, "sparse"),      # BASELINE config 1 size
    (2, 3, 6, "sparse"),
    (2, 3, 4, "full"),
    (3, 3, 5, "sparse"),
    (4, 3, 4, "sparse"),
    (6, 3, 3, "sparse"),
    (2, 1, 5, "sparse"),
    (2, 2, 6, "sparse"),
    (3, 2, 4, "full"),
    (2, 4, 5, "sparse"),
    (3, 4, 3, "sparse"),
    (2, 5, 4, "sparse"),
    (2, 6, 3, "sparse"),      # run-time-k fallback kernel
    (5, 2, 3, "sparse"),
]


# every test that takes `plans` runs twice: through the tiled class kernels (streaming / constant-bank /
# register-tiled / generic) and through the flat kernel (one launch per right-hand side; what small index sets
# take by default).  k = 6 has no flat instantiation: its "flat" run exercises the fall-through to the tiled path.
@pytest.fixture(scope="module", params=[0, 1], ids=["tiled", "flat"])
def plans(gsg, oracle, request):
    cache = {}
    mode = request.param

    def get(D, k, n, scheme):
        key = (D, k, n, scheme)
        if key not in cache:
            H = oracle.periodic_DLF_matrix(k, n)
            import scipy.sparse as sp
            Hs = sp.csc_matrix((H.nzval, H.rowval, H.colptr), shape=(H.m, H.n))
            # the SAME 1-D matrix is handed to the GPU plan and to the oracle
            plan = gsg.Plan(D, k, n, scheme, H=Hs)
            plan.set_flat(mode)
            cache[key] = (plan, H)
        return cache[key]

    get.mode = mode
    return get


@pytest.mark.parametrize("D,k,n,scheme", CASES)
def test_apply_D_every_axis(plans, oracle, D, k, n, scheme):
    plan, H = plans(D, k, n, scheme)
    N = oracle.get_size(D, k, n, scheme)
    assert plan.size == N
    x = random_state(N, seed=D * 100 + k * 10 + n)
    for d in range(1, D + 1):
        y_ref = oracle.apply_D_poles(D, d, k, n, x, scheme=scheme, H=H)
        y = plan.apply_D(d, x)
        assert relerr(y, y_ref) <= TOL, (d, relerr(y, y_ref))


@pytest.mark.parametrize("D,k,n", [(2, 3, 6), (3, 3, 5), (4, 3, 4)])
@pytest.mark.parametrize("f", [f_sin, f_gauss])
def test_apply_D_synthetic_initial_conditions(plans, oracle, D, k, n, f):
    """sin-product and Gaussian data, as north_star names them."""
    plan, H = plans(D, k, n, "sparse")
    x = product_state(oracle, D, k, n, f)
    for d in (1, D):
        y_ref = oracle.apply_D_poles(D, d, k, n, x, H=H)
        assert relerr(plan.apply_D(d, x), y_ref) <= TOL


def test_apply_D_matches_literal_assembly(plans, oracle):
    """Cross-check against the literal Dict-loop assembly of
    src/multidim_derivative.jl:19-58 applied by CSC column scatter (src/pdes.jl:63)."""
    D, k, n = 2, 3, 4
    plan, H = plans(D, k, n, "sparse")
    x = random_state(plan.size, seed=7)
    for d in (1, 2):
        A = oracle.D_matrix_literal(D, d, k, n, H=H)
        assert relerr(plan.apply_D(d, x), A.matvec(x)) <= TOL


@pytest.mark.parametrize("D,k,n,scheme", [(2, 3, 6, "sparse"), (4, 3, 4, "sparse"), (3, 2, 4, "full")])
def test_apply_grad_and_laplacian(plans, oracle, D, k, n, scheme):
    plan, H = plans(D, k, n, scheme)
    x = random_state(plan.size, seed=11)
    a = np.linspace(0.5, 1.5, D)
    mats = [oracle.D_matrix_poles(D, d, k, n, scheme=scheme, H=H) for d in range(1, D + 1)]
    g_ref = sum(ad * (A @ x) for ad, A in zip(a, mats))
    assert relerr(plan.apply_grad(a, x), g_ref) <= TOL
    l_ref = sum(A @ (A @ x) for A in mats)
    # (D*D)x vs D(Dx): rounding floor ~8e-13 relative at n=8; small n here
    assert relerr(plan.apply_laplacian(x), l_ref) <= TOL


@pytest.mark.parametrize("D,k,n", [(2, 3, 5), (4, 3, 4)])
@pytest.mark.parametrize("f", [f_sin, f_gauss])
def test_rk4_advect_fixed_steps(plans, oracle, D, k, n, f):
    """RK4 state after a FIXED number of steps at fixed dt."""
    plan, H = plans(D, k, n, "sparse")
    u0 = product_state(oracle, D, k, n, f)
    a = np.ones(D)
    mats = [oracle.D_matrix_poles(D, d, k, n, H=H) for d in range(1, D + 1)]
    dt, nsteps = 1.0e-4, 16
    ref = oracle.rk4(oracle.advect_rhs(mats, a), u0, dt, nsteps)
    out = plan.rk4_advect(a, u0, dt, nsteps)
    assert relerr(out, ref) <= TOL
    assert relerr(out, u0) > 1e-6          # the state actually moved


@pytest.mark.parametrize("mode", [0, 1])
def test_rk4_modes_agree_with_oracle(plans, oracle, mode):
    """Both RK4 drivers -- Taylor form for linear right-hand sides (mode 0) and the staged form (mode 1),
    each replayed from a captured CUDA graph after the first step -- against the oracle's classical RK4."""
    D, k, n = 3, 3, 4
    plan, H = plans(D, k, n, "sparse")
    u0 = product_state(oracle, D, k, n, f_gauss)
    a = np.array([1.0, -0.5, 0.25])
    mats = [oracle.D_matrix_poles(D, d, k, n, H=H) for d in range(1, D + 1)]
    dt, nsteps = 1.0e-4, 24
    ref = oracle.rk4(oracle.advect_rhs(mats, a), u0, dt, nsteps)
    plan.set_rk4_mode(mode)
    try:
        out = plan.rk4_advect(a, u0, dt, nsteps)
    finally:
        plan.set_rk4_mode(0)
    assert relerr(out, ref) <= TOL


def test_rk4_wave_fixed_steps(plans, oracle):
    D, k, n = 2, 3, 5
    plan, H = plans(D, k, n, "sparse")
    u0 = product_state(oracle, D, k, n, f_sin)
    v0 = np.zeros_like(u0)
    mats = [oracle.D_matrix_poles(D, d, k, n, H=H) for d in range(1, D + 1)]
    dt, nsteps = 2.0e-4, 16
    ref = oracle.rk4(oracle.wave_rhs(mats), np.concatenate([u0, v0]), dt, nsteps)
    u, v = plan.rk4_wave(u0, v0, dt, nsteps)
    assert relerr(np.concatenate([u, v]), ref) <= TOL
    e0 = oracle.energy(mats, np.concatenate([u0, v0]))
    assert abs(plan.energy(u0, v0) - e0) <= 1e-12 * e0
    # test/solvers.jl:64-75: sqrt(E) ~ sqrt(2) pi for sin(2 pi x) sin(2 pi y)
    assert abs(math.sqrt(plan.energy(u, v)) - math.sqrt(2) * math.pi) < 1e-4


@pytest.mark.parametrize("D,k,n,scheme", [(1, 3, 5, "sparse"), (2, 3, 4, "sparse"), (2, 2, 3, "full"),
                                          (3, 4, 3, "sparse"), (4, 3, 3, "sparse")])
def test_reconstruct(plans, oracle, D, k, n, scheme):
    if plans.mode == 1:
        pytest.skip("reconstruct does not depend on the sweep path")
    plan, _ = plans(D, k, n, scheme)
    vect = product_state(oracle, D, k, n, f_sin, scheme=scheme) + 0.1 * product_state(oracle, D, k, n, f_gauss, scheme=scheme)
    rng = np.random.default_rng(20240)
    pts = rng.random((64, D))
    pts[0, :] = 0.0          # edges of the domain and cell boundaries
    pts[1, :] = 1.0
    pts[2, :] = 0.5
    pts[3, :] = 0.25
    ref = np.array([oracle.reconstruct_DG(D, k, n, vect, list(p), scheme=scheme) for p in pts])
    out = plan.reconstruct(vect, pts)
    scale = max(np.abs(ref).max(), 1e-300)
    assert np.abs(out - ref).max() / scale <= TOL


def test_reconstruct_empty_and_single(plans, oracle):
    if plans.mode == 1:
        pytest.skip("reconstruct does not depend on the sweep path")
    plan, _ = plans(2, 3, 4, "sparse")
    vect = product_state(oracle, 2, 3, 4, f_cos)
    assert plan.reconstruct(vect, np.empty((0, 2))).shape == (0,)
    one = plan.reconstruct(vect, np.array([[0.3, 0.7]]))
    assert abs(one[0] - oracle.reconstruct_DG(2, 3, 4, vect, [0.3, 0.7])) <= 1e-12


def test_spmv_cross_check(gsg, plans, oracle):
    """CSR SpMV on the reference-style assembled matrix agrees with the matrix-free sweep."""
    D, k, n = 3, 3, 4
    plan, H = plans(D, k, n, "sparse")
    x = random_state(plan.size, seed=5)
    for d in (1, 2, 3):
        A = oracle.D_matrix_poles(D, d, k, n, H=H)
        y_spmv = gsg.spmv_csc(A, x)
        assert relerr(y_spmv, A @ x) <= TOL
        assert relerr(y_spmv, plan.apply_D(d, x)) <= TOL
    R = gsg.CsrMatrix(A)
    assert relerr(R @ x, A @ x) <= TOL


def test_linearity_and_skew_symmetry_mid_size(plans, oracle):
    """Size-independent properties at a size the oracle does not assemble: linearity and
    <x, D y> = -<D x, y> up to the reference's own ||H + H'|| ~ 1e-9 * max|H| noise."""
    D, k, n = 4, 3, 6
    plan, H = plans(D, k, n, "sparse")
    x = random_state(plan.size, seed=1)
    y = random_state(plan.size, seed=2)
    for d in (1, 4):
        Dx, Dy = plan.apply_D(d, x), plan.apply_D(d, y)
        Dxy = plan.apply_D(d, 2.0 * x - 3.0 * y)
        assert relerr(Dxy, 2.0 * Dx - 3.0 * Dy) <= 1e-13
        skew = abs(np.dot(x, Dy) + np.dot(Dx, y)) / (np.linalg.norm(x) * np.linalg.norm(Dy))
        assert skew < 1e-9


def test_error_behaviour(gsg, plans):
    plan, _ = plans(2, 3, 4, "sparse")
    with pytest.raises(ValueError):
        plan.apply_D(1, np.zeros(3))
    with pytest.raises(gsg.GsgError):
        plan.apply_D(3, np.zeros(plan.size))
    with pytest.raises(gsg.GsgError):
        gsg.Plan(2, 11, 3)          # DomainError: k > K_max (src/1d_dg_functions.jl:39)
    with pytest.raises(ValueError):
        gsg.get_size(2, 3, 3, scheme="energy")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_block_partition_matches_single_gpu(oracle, world):
    """The multi-GPU block partition, run as `world` virtual ranks (threads, in-process mailboxes) on one
    GPU: every rank sweeps only its pole groups, exchanges the level-0 blocks along the partition
    dimensions, and the owned parts add up to the unpartitioned RK4 state."""
    import threading

    import torch
    from gsg_b200.distributed import PartitionedRK4, ThreadComm
    D, k, n = 4, 3, 4
    H = oracle.periodic_DLF_matrix(k, n)
    import scipy.sparse as sp
    Hs = sp.csc_matrix((H.nzval, H.rowval, H.colptr), shape=(H.m, H.n))
    import gsg_b200 as g
    u0 = product_state(oracle, D, k, n, f_gauss)
    a = np.array([1.0, -0.5, 0.25, 2.0])
    dt, nsteps = 1.0e-4, 6
    ref_plan = g.Plan(D, k, n, "sparse", H=Hs, device=0)
    ref = ref_plan.rk4_advect(a, u0, dt, nsteps)
    comm = ThreadComm(world)
    full = ref_plan.to_device(u0)
    outs, errs = [None] * world, []

    def worker(r):
        try:
            plan = g.Plan(D, k, n, "sparse", H=Hs, device=0)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                plan.set_stream(stream)
                drv = PartitionedRK4(plan, a, r, world, torch.device("cuda", 0), comm.endpoint(r))
                drv.set_state(full)
                drv.step(dt, nsteps)
                stream.synchronize()
                outs[r] = drv.owned_state()
                stream.synchronize()
        except Exception as e:      # surface worker failures in the main thread
            errs.append(repr(e))

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    total = sum(outs)
    out = ref_plan.to_host(total)
    assert relerr(out, ref) <= TOL
    # ownership is a partition: every cell owned exactly once is implied by the sum matching;
    # and no rank owns everything
    assert all(float(o.abs().sum()) > 0 for o in outs)


def _sample_pole_indices(oracle, D, d, k, n, rng, npoles):
    """Global (reference-layout) indices of `npoles` random poles along axis d, each in 1-D layout order
    (level, cell, mode), computed from the block table alone (src/dg_vmethods.jl:53-73 strides)."""
    blocks, N = oracle.block_table(D, k, n)
    by_level = {lv: (off, ks) for lv, off, ks in blocks}
    starts = [lv for lv, _, _ in blocks if lv[d - 1] == 0]
    kD = k ** D
    out = []
    for _ in range(npoles):
        lv0 = starts[rng.integers(len(starts))]
        p = n - sum(lv0)
        ks0 = by_level[lv0][1]
        cells = [int(rng.integers(ks0[j])) for j in range(D)]
        modes = [int(rng.integers(k)) for j in range(D)]
        idx = []
        for ld in range(p + 1):
            lv = lv0[:d - 1] + (ld,) + lv0[d:]
            off, ks = by_level[lv]
            cstr = [kD * int(np.prod(ks[:j], dtype=np.int64)) for j in range(D)]
            base = off + sum(cells[j] * cstr[j] + modes[j] * k ** j for j in range(D) if j != d - 1)
            for c in range(ks[d - 1]):
                for m in range(k):
                    idx.append(base + c * cstr[d - 1] + m * k ** (d - 1))
        out.append(np.array(idx, dtype=np.int64))
    return out


def test_full_size_baseline_config(gsg, oracle):
    """BASELINE config 4 at full size (D=6, k=3, n=8, 34.5 M DOFs): sampled poles of every direction against
    the oracle's H (every pole-length class, N' = 3 ... 768), linearity, skew-symmetry."""
    import scipy.sparse as sp
    D, k, n = 6, 3, 8
    H = gsg.periodic_DLF_matrix(k, n)                 # the library's own H (test_abi pins it to the oracle's)
    plan = gsg.Plan(D, k, n, "sparse", H=H)
    assert plan.size == 34455456
    Hd = sp.csr_matrix(H)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(plan.size)
    y = rng.standard_normal(plan.size)
    coef = np.array([1.0, -0.5, 0.25, 2.0, -1.5, 0.75])
    grad_ref = np.zeros_like(x)
    for d in range(1, D + 1):
        Dx = plan.apply_D(d, x)
        grad_ref += coef[d - 1] * Dx
        worst, lengths = 0.0, set()
        for idx in _sample_pole_indices(oracle, D, d, k, n, rng, 120):
            Np = idx.size
            lengths.add(Np)
            ref = Hd[:Np, :Np] @ x[idx]
            worst = max(worst, float(np.abs(Dx[idx] - ref).max() / np.abs(ref).max()))
        assert worst <= 1e-12, (d, worst)
        assert len(lengths) >= 5                      # several pole-length classes were hit
        if d in (1, 6):
            Dy = plan.apply_D(d, y)
            assert relerr(plan.apply_D(d, 2.0 * x - 3.0 * y), 2.0 * Dx - 3.0 * Dy) <= 1e-13
            skew = abs(np.dot(x, Dy) + np.dot(Dx, y)) / (np.linalg.norm(x) * np.linalg.norm(Dy))
            assert skew < 1e-9
    # the fused gradient (direction-pair tiles + reduced sweeps) against the sum of the single-direction applies
    assert relerr(plan.apply_grad(coef, x), grad_ref) <= 1e-13
    # the two RK4 drivers (Taylor form / staged form) agree at full size, and the state moves
    a = np.ones(D)
    u0 = gsg.tensor_construct(D, k, n, [gsg.vcoeffs_DG(1, k, n, f_sin)] * D)
    u_taylor = plan.rk4_advect(a, u0, 1.0e-4, 3)
    plan.set_rk4_mode(1)
    u_staged = plan.rk4_advect(a, u0, 1.0e-4, 3)
    plan.set_rk4_mode(0)
    assert relerr(u_taylor, u_staged) <= 1e-12
    assert relerr(u_taylor, u0) > 1e-4
    # advection by a = (1,...,1) of prod sin(2 pi x_d) conserves the L2 norm to time-stepping accuracy
    assert abs(np.linalg.norm(u_taylor) / np.linalg.norm(u0) - 1.0) < 1e-6
    del plan


@pytest.mark.parametrize("D,k,n", [(2, 3, 6), (4, 3, 5), (3, 2, 5), (2, 4, 4), (4, 1, 4), (2, 5, 3), (6, 3, 3)])
def test_fused_gradient_matches_unfused(plans, oracle, D, k, n):
    """gsg_apply_grad takes the direction-pair path when every coefficient is non-zero (PAIR tiles for the 2-D
    sub-planes with n' <= 2 / 1 / 0 at k <= 3 / 4 / 5, reduced sweeps for the rest; odd D leaves the last
    direction unpaired); a zero coefficient forces the plain per-direction path.  Both against the oracle."""
    plan, H = plans(D, k, n, "sparse")
    x = random_state(plan.size, seed=11 * D + k)
    a = np.array([1.0, -0.5, 0.25, 2.0, -1.5, 0.75][:D])
    ref = sum(a[d - 1] * oracle.apply_D_poles(D, d, k, n, x, H=H) for d in range(1, D + 1))
    assert relerr(plan.apply_grad(a, x), ref) <= TOL
    a0 = a.copy()
    a0[0] = 0.0
    ref0 = sum(a0[d - 1] * oracle.apply_D_poles(D, d, k, n, x, H=H) for d in range(1, D + 1))
    assert relerr(plan.apply_grad(a0, x), ref0) <= TOL


@pytest.mark.parametrize("D,k,n,scheme", [(3, 3, 4, "sparse"), (2, 2, 5, "sparse"), (4, 3, 3, "sparse"), (2, 4, 3, "full")])
def test_device_pointer_applies_alpha_beta_masks(plans, oracle, D, k, n, scheme):
    """gsg_apply_D_dev with general alpha / beta, gsg_apply_dirs_dev with direction masks and beta = 1, and the
    padding slots of the device layout staying zero -- on both sweep paths."""
    import torch
    plan, H = plans(D, k, n, scheme)
    x = random_state(plan.size, seed=31)
    y0 = random_state(plan.size, seed=32)
    xd, yd = plan.to_device(x), plan.to_device(y0)
    Dx = [oracle.apply_D_poles(D, d, k, n, x, scheme=scheme, H=H) for d in range(1, D + 1)]
    plan.apply_D_dev(D, xd, yd, alpha=-0.75, beta=0.5)
    assert relerr(plan.to_host(yd), -0.75 * Dx[D - 1] + 0.5 * y0) <= TOL
    yd = plan.to_device(y0)
    plan.apply_D_dev(1, xd, yd, alpha=2.0, beta=1.0)
    assert relerr(plan.to_host(yd), 2.0 * Dx[0] + y0) <= TOL
    c = np.linspace(-1.0, 2.0, D)
    dirs = [d for d in range(1, D + 1) if d != 2]
    yd = plan.to_device(y0)
    plan.apply_dirs_dev(c, dirs, xd, yd, beta=1.0)
    ref = y0 + sum(c[d - 1] * Dx[d - 1] for d in dirs)
    assert relerr(plan.to_host(yd), ref) <= TOL
    plan.apply_dirs_dev(c, range(1, D + 1), xd, yd, beta=0.0)
    assert relerr(plan.to_host(yd), sum(c[d - 1] * Dx[d - 1] for d in range(1, D + 1))) <= TOL
    plan.sync()
    KD, KDp = k ** D, plan.cell_stride
    if KDp > KD:
        assert float(torch.abs(yd.view(-1, KDp)[:, KD:]).max()) == 0.0
