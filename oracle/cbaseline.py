"""ctypes front-end of oracle/csc_kernels.c -- TEST / BASELINE INFRASTRUCTURE ONLY.

Provides (a) the reference-style assembled matrix D_d (literal column loop of
src/multidim_derivative.jl:32-55 restated in C) for mid-size cross-checks, and (b) the CPU
baseline of the hot path: Julia's serial CSC column-scatter `A*x` (src/pdes.jl:63,179-180) and an
OpenMP CSR analogue of the optional MKLSparse path (src/GalerkinSparseGrids.jl:5-7), timed on a
BOUNDED, strided sample of columns / rows of the real operator and extrapolated by nnz.
Imported only by tests/, bench.py's cpu_baseline / --impl reference legs and smoke()."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csc_kernels.c")
_LIB = os.path.join(_HERE, "libgsg_oracle_c.so")


def _build():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
                        _SRC, "-o", _LIB], check=True)


def _load():
    _build()
    lib = C.CDLL(_LIB)
    i64, f64, vp = C.c_int64, C.c_double, C.c_void_p
    lib.gsgo_get_size.restype = i64
    lib.gsgo_get_size.argtypes = [C.c_int] * 4
    lib.gsgo_assemble_cols.restype = i64
    lib.gsgo_assemble_cols.argtypes = [C.c_int] * 5 + [vp, vp, vp, i64, i64, vp, vp, vp, i64]
    lib.gsgo_spmv_csc.restype = None
    lib.gsgo_spmv_csc.argtypes = [i64, vp, vp, vp, vp, vp]
    lib.gsgo_spmv_csr_omp.restype = None
    lib.gsgo_spmv_csr_omp.argtypes = [i64, vp, vp, vp, vp, vp]
    lib.gsgo_axpy.restype = None
    lib.gsgo_axpy.argtypes = [i64, f64, vp, vp]
    lib.gsgo_max_threads.restype = C.c_int
    lib.gsgo_set_threads.restype = None
    lib.gsgo_set_threads.argtypes = [C.c_int]
    lib.gsgo_apply_D_poles.restype = C.c_int
    lib.gsgo_apply_D_poles.argtypes = [C.c_int] * 5 + [vp, vp, vp, vp, vp]
    lib.gsgo_apply_D_assembled.restype = i64
    lib.gsgo_apply_D_assembled.argtypes = [C.c_int] * 5 + [vp, vp, vp, vp, vp, i64]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_SCHEME = {"sparse": 0, "full": 1}


def _H_arrays(H):
    """H: gsg_oracle.CSC or scipy csc -> 0-based int64 colptr/rowval + values."""
    if hasattr(H, "colptr"):
        return (np.ascontiguousarray(H.colptr, dtype=np.int64), np.ascontiguousarray(H.rowval, dtype=np.int64),
                np.ascontiguousarray(H.nzval, dtype=np.float64))
    import scipy.sparse as sp
    H = sp.csc_matrix(H)
    H.sort_indices()
    return (H.indptr.astype(np.int64), H.indices.astype(np.int64), np.ascontiguousarray(H.data, dtype=np.float64))


def get_size(D, k, n, scheme="sparse"):
    return int(lib().gsgo_get_size(D, k, n, _SCHEME[scheme]))


def assemble_cols(D, d, k, n, H, col_begin, col_end, scheme="sparse"):
    """Columns [col_begin, col_end) of D_d as (colptr, rowval, nzval), 0-based int64."""
    hc, hr, hv = _H_arrays(H)
    ncols = col_end - col_begin
    colptr = np.empty(ncols + 1, dtype=np.int64)
    nnz = lib().gsgo_assemble_cols(D, k, n, _SCHEME[scheme], d, _p(hc), _p(hr), _p(hv), col_begin, col_end,
                                   _p(colptr), None, None, 0)
    if nnz < 0:
        raise RuntimeError("gsgo_assemble_cols failed")
    rowval = np.empty(nnz, dtype=np.int64)
    nzval = np.empty(nnz, dtype=np.float64)
    lib().gsgo_assemble_cols(D, k, n, _SCHEME[scheme], d, _p(hc), _p(hr), _p(hv), col_begin, col_end,
                             _p(colptr), _p(rowval), _p(nzval), nnz)
    return colptr, rowval, nzval


def D_matrix(D, d, k, n, H, scheme="sparse"):
    """Full assembled D_d as scipy CSC (what `D_matrix(D, d, k, n)` returns in the reference)."""
    import scipy.sparse as sp
    N = get_size(D, k, n, scheme)
    colptr, rowval, nzval = assemble_cols(D, d, k, n, H, 0, N, scheme)
    return sp.csc_matrix((nzval, rowval, colptr), shape=(N, N))


def spmv_csc(colptr, rowval, nzval, x, y):
    lib().gsgo_spmv_csc(colptr.size - 1, _p(colptr), _p(rowval), _p(nzval), _p(x), _p(y))


def apply_D_poles(D, d, k, n, H, x, scheme="sparse"):
    """y = D_d x, pole by pole in C (OpenMP over pole groups), per-row summation order of the CSC scatter."""
    hc, hr, hv = _H_arrays(H)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    if lib().gsgo_apply_D_poles(D, k, n, _SCHEME[scheme], d, _p(hc), _p(hr), _p(hv), _p(x), _p(y)) != 0:
        raise RuntimeError("gsgo_apply_D_poles failed")
    return y


def apply_D_assembled(D, d, k, n, H, x, scheme="sparse", slab_cols=1 << 20):
    """y = D_matrix(D, d, k, n) * x with the matrix assembled slab by slab as the reference's column loop does
    and applied by CSC column scatter (full-size stand-in for the 7.5 GB assembled matrix); returns (y, nnz)."""
    hc, hr, hv = _H_arrays(H)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    nnz = lib().gsgo_apply_D_assembled(D, k, n, _SCHEME[scheme], d, _p(hc), _p(hr), _p(hv), _p(x), _p(y), slab_cols)
    if nnz < 0:
        raise RuntimeError("gsgo_apply_D_assembled failed")
    return y, int(nnz)


def rk4_advect(D, k, n, H, a, u0, dt, nsteps, scheme="sparse"):
    """Classical RK4 (gsg_oracle.rk4 stage order) of u' = -sum_d a_d D_d u with the C pole apply."""
    def rhs(u):
        acc = np.zeros_like(u)
        for d in range(1, D + 1):
            if a[d - 1] != 0.0:
                acc += a[d - 1] * apply_D_poles(D, d, k, n, H, u, scheme)
        return -acc
    y = np.array(u0, dtype=np.float64)
    for _ in range(nsteps):
        k1 = rhs(y)
        k2 = rhs(y + (0.5 * dt) * k1)
        k3 = rhs(y + (0.5 * dt) * k2)
        k4 = rhs(y + dt * k3)
        y = y + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    return y


def total_nnz(D, k, n, H, scheme="sparse"):
    """nnz of one assembled D_d: sum over pole groups of poles * nnz(H[:N', :N'])."""
    import gsg_oracle as o
    hc, hr, hv = _H_arrays(H)
    N1 = hc.size - 1
    # nnz of each principal sub-block
    cols = np.repeat(np.arange(N1), np.diff(hc))
    nnz_p = {}
    for p in range(n + 1):
        Np = k << p
        nnz_p[p] = int(np.count_nonzero((cols < Np) & (hr < Np)))
    blocks, _ = o.block_table(D - 1, k, n, scheme) if D > 1 else ([((), 0, ())], 0)
    total = 0
    for lv, _off, ks in blocks:
        p = n if scheme == "full" else n - sum(lv)
        npoles = k ** (D - 1) * int(np.prod(ks, dtype=np.int64)) if D > 1 else 1
        total += npoles * nnz_p[p]
    return total


def time_direction(D, d, k, n, H, x, threads=1, slab=1 << 18, scheme="sparse", workers=None):
    """Time ONE COMPLETE product D_d x of the reference's CPU path: the whole assembled D_d, slab by slab (the
    slabs are assembled by a thread pool, untimed -- the reference assembles once, outside its time loop -- and
    each slab's SpMV is timed).  threads == 1: Julia's serial CSC column scatter (stock SparseArrays, `A*x` of
    src/pdes.jl:63); threads > 1: OpenMP row-parallel CSR on the same entries (the optional MKLSparse path,
    src/GalerkinSparseGrids.jl:5-7).  Returns (seconds of SpMV, nnz, y)."""
    from concurrent.futures import ThreadPoolExecutor
    N = get_size(D, k, n, scheme)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(N)
    y += 0.0                                        # touch the pages: no first-write faults inside the timed SpMV
    if threads > 1:
        import scipy.sparse as sp
        hc, hr, hv = _H_arrays(H)
        Hs = sp.csc_matrix((hv, hr, hc), shape=(k << n, k << n)).T.tocsc()
        Hs.sort_indices()
        lib().gsgo_set_threads(int(threads))
    else:
        Hs = H
    get_size(D, k, n, scheme)
    assemble_cols(D, d, k, n, Hs, 0, 1, scheme)                      # builds the C side's cached index set
    bounds = [(b, min(N, b + slab)) for b in range(0, N, slab)]
    workers = workers or max(1, min(16, (os.cpu_count() or 1)))
    t_spmv, nnz = 0.0, 0
    with ThreadPoolExecutor(max_workers=workers) as pool:
        group = 2 * workers
        for g0 in range(0, len(bounds), group):
            part = bounds[g0:g0 + group]
            slabs = list(pool.map(lambda be: assemble_cols(D, d, k, n, Hs, be[0], be[1], scheme), part))
            for (b, e), (colptr, rowval, nzval) in zip(part, slabs):
                if threads == 1:
                    xs = x[b:e]
                    t0 = time.perf_counter()
                    spmv_csc(colptr, rowval, nzval, xs, y)
                    t_spmv += time.perf_counter() - t0
                else:
                    ys = y[b:e]
                    t0 = time.perf_counter()
                    lib().gsgo_spmv_csr_omp(e - b, _p(colptr), _p(rowval), _p(nzval), _p(x), _p(ys))
                    t_spmv += time.perf_counter() - t0
                nnz += nzval.size
    return t_spmv, nnz, y


def time_axpy(N, threads=1):
    x = np.ones(N)
    y = np.zeros(N)
    lib().gsgo_axpy(N, 0.5, _p(x), _p(y))          # first pass touches the pages
    t0 = time.perf_counter()
    lib().gsgo_axpy(N, 0.5, _p(x), _p(y))
    return time.perf_counter() - t0


def rk4_cpu_baseline(D, k, n, H, budget_s=15.0, threads=1, chunk=4096, scheme="sparse", seed=0):
    """Estimate RK4 DOF-updates/s of the reference's CPU path at (D, k, n).

    threads == 1: Julia's serial CSC column scatter (stock SparseArrays).
    threads  > 1: OpenMP row-parallel CSR (MKLSparse analogue), all host threads.
    For every direction d a strided sample of `chunk`-wide column (row) slabs of the REAL
    operator is assembled and multiplied against full-length vectors; the measured time is
    scaled by nnz_total / nnz_sampled.  One RK4 step = 4 RHS x D products + vector updates."""
    N = get_size(D, k, n, scheme)
    nnz_total = total_nnz(D, k, n, H, scheme)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(N)
    y = np.zeros(N)
    # a few large CONTIGUOUS slabs per direction (realistic cache reuse inside a slab), sized
    # from a rough assembly+product rate so that the whole measurement fits the budget
    nslab = 8
    cols_budget = budget_s * 1.0e6 / D                    # ~1e6 columns/s assembled and multiplied
    slab = int(max(chunk, min(N // nslab, cols_budget // nslab)))
    if threads > 1:
        import scipy.sparse as sp
        hc, hr, hv = _H_arrays(H)
        Ht = sp.csc_matrix((hv, hr, hc), shape=(k << n, k << n)).T.tocsc()
        Ht.sort_indices()
    t_products = 0.0
    sampled = 0
    per_dir = []
    for d in range(1, D + 1):
        t_d, nnz_d = 0.0, 0
        for si in range(nslab):
            b = int((N - slab) * (si + 0.5 * ((d - 1) / D)) / nslab)      # de-phase the directions
            e = min(N, b + slab)
            if threads == 1:
                colptr, rowval, nzval = assemble_cols(D, d, k, n, H, b, e, scheme)
                xs = np.ascontiguousarray(x[b:e])
                t0 = time.perf_counter()
                spmv_csc(colptr, rowval, nzval, xs, y)
                t_d += time.perf_counter() - t0
            else:
                rowptr, colidx, nzval = assemble_cols(D, d, k, n, Ht, b, e, scheme)   # rows of D_d
                ys = np.empty(e - b)
                t0 = time.perf_counter()
                lib().gsgo_spmv_csr_omp(e - b, _p(rowptr), _p(colidx), _p(nzval), _p(x), _p(ys))
                t_d += time.perf_counter() - t0
            nnz_d += nzval.size
        per_dir.append(t_d * nnz_total / max(nnz_d, 1))
        t_products += t_d
        sampled += nnz_d
    # vector arithmetic of one RK4 step: sum of D product vectors per RHS + stage updates
    t0 = time.perf_counter()
    lib().gsgo_axpy(N, 0.5, _p(x), _p(y))
    t_axpy = time.perf_counter() - t0
    n_axpy = 4 * D + 10
    t_step = 4.0 * sum(per_dir) + n_axpy * t_axpy
    return {
        "N": N, "nnz_per_direction": nnz_total, "nnz_sampled": sampled, "sample_fraction": sampled / (D * nnz_total),
        "t_products_measured_s": t_products, "t_spmv_full_est_s": per_dir, "t_axpy_s": t_axpy,
        "t_step_est_s": t_step, "dof_updates_per_s": N / t_step, "threads": threads,
        "nnz_per_s": sampled / max(t_products, 1e-12),
    }
