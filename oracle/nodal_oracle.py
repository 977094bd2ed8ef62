"""CPU restatement of the reference's nodal / point basis transforms -- TEST INFRASTRUCTURE ONLY.

The Vlasov right-hand side (src/pdes.jl:174-192) applies four sparse matrices m2n, n2p, p2n, n2m that the HOST
builds once (`make_modal2point_matrices` / `make_point2modal_matrices`, src/multidim_nodal_basis.jl:127-141) and
hands to `vlasov_evolve`; the product takes them over the C ABI exactly like that.  The tests need real ones, so
this module restates their construction:
  src/1d_nodal_basis.jl:46-115   lag_nodal / h_nodal / v_nodal (k in {2, 3, 5})
  src/1d_nodal_basis.jl:160-173  eval_v_nodal_right / eval_v_nodal_left (x +- 1e-16)
  src/1d_nodal_basis.jl:183-241  eval_points_1D, nodal2points_1D, points2nodal_1D (dense inverse)
  src/1d_nodal_basis.jl:281-362  nodal2pos_1D (hquadrature over half cells), pos2nodal_1D (dense inverse)
  src/1d_nodal_basis.jl:383-414  transform_1D
  src/multidim_nodal_basis.jl:19-125  relevant_cell, inner_loop / make_column / transform (tensor product of one
                                      1-D matrix restricted to the sparse index set, entries below atol dropped)
  src/additional_tools.jl:23-42  threshold
Deviations: hquadrature (HCubature 1.4.0) is replaced by a Gauss-Legendre rule that is exact for the polynomial
integrands (degree <= 2k-2 on each half cell); the D-dimensional matrix is assembled block pair by block pair
from Kronecker products of masked 1-D sub-blocks instead of the reference's column-by-column scalar loop (same
entries: the relevance mask and the product factorise over the dimensions; validated against the literal loop
for a small case in tests/test_oracle_pins.py).
PARITY STATUS: unpinned against Julia; pinned by the reference's own assertions test/transformations.jl:15-79
(mutual inverses: < 1e-15 * 10^k in 1-D for k in {2,3,5}, < 1e-10 in 2-D at k=3, n=5).
"""
from __future__ import annotations

from fractions import Fraction as Fr

import numpy as np

import gsg_oracle as o


# ---- src/1d_nodal_basis.jl:46-115 -----------------------------------------------------------------------------
def lag_nodal(k: int, mode: int, x: float) -> float:
    if x < 0 or x > 1:
        return 0.0
    if k == 2:
        return x if mode == 1 else 1 - x
    if k == 3:
        if mode == 1:
            return 2 * (x - 0.5) * (x - 1)
        if mode == 2:
            return -4 * x * (x - 1)
        return 2 * x * (x - 0.5)
    if k == 5:
        if mode == 1:
            return 32 / 3 * (x - 0.25) * (x - 0.5) * (x - 0.75) * (x - 1)
        if mode == 2:
            return -128 / 3 * x * (x - 0.5) * (x - 0.75) * (x - 1)
        if mode == 3:
            return 64 * x * (x - 0.25) * (x - 0.75) * (x - 1)
        if mode == 4:
            return -128 / 3 * x * (x - 0.25) * (x - 0.5) * (x - 1)
        return 32 / 3 * x * (x - 0.25) * (x - 0.5) * (x - 0.75)
    raise ValueError("k must be 2, 3 or 5")


def h_nodal(k: int, mode: int, x: float) -> float:
    if k == 2:
        if mode == 1:
            return max(0.0, 1 - abs(x))
        if abs(x) > 1 or abs(x) == 0:
            return 0.0
        return (x + 1) if x < 0 else 0.0
    if k == 3:
        if mode == 1:
            return max(0.0, -4 * x * (x + 1))
        if mode == 2:
            return max(0.0, -4 * x * (x - 1))
        if abs(x) > 1 or abs(x) == 0:
            return 0.0
        return 2 * (x + 1) * (x + 0.5) if x < 0 else 0.0
    if k == 5:
        if mode == 1:
            return 0.0 if (x < -1 or x > 0) else -128 / 3 * x * (x + 1) * (x + 0.5) * (x + 0.25)
        if mode == 2:
            return 0.0 if (x < -1 or x > 0) else -128 / 3 * x * (x + 1) * (x + 0.5) * (x + 0.75)
        if mode == 3:
            return 0.0 if (x > 1 or x < 0) else -128 / 3 * x * (x - 1) * (x - 0.5) * (x - 0.75)
        if mode == 4:
            return 0.0 if (x > 1 or x < 0) else -128 / 3 * x * (x - 1) * (x - 0.5) * (x - 0.25)
        if abs(x) > 1 or abs(x) == 0:
            return 0.0
        return 32 / 3 * (x + 1) * (x + 0.75) * (x + 0.5) * (x + 0.25) if x < 0 else 0.0
    raise ValueError("k must be 2, 3 or 5")


def v_nodal(k: int, level: int, cell: int, mode: int, x: float) -> float:
    if level == 0:
        return lag_nodal(k, mode, x)
    return h_nodal(k, mode, (1 << level) * x - (2 * cell - 1))


# ---- src/1d_nodal_basis.jl:160-241 ----------------------------------------------------------------------------
def eval_points_1D(k: int, max_level: int, level: int, cell: int, mode: int):
    """values of one nodal function on all collocation points (dense vector of length k * 2^max_level)"""
    out = []
    for l in range(max_level + 1):
        if l == 0:
            nodes = [Fr(j, k - 1) for j in range(k - 1)]                     # 0 : 1/(k-1) : 1 - 1/(k-1)
        else:
            nodes = [Fr(2 * j + 1, 2 * (k - 1)) for j in range(k - 1)]       # 1/(2(k-1)) : 1/(k-1) : 1 - 1/(2(k-1))
        cells = 1 << max(0, l - 1)
        for c in range(1, cells + 1):
            for node in nodes:
                x = Fr(c - 1) + node
                x = x / cells
                out.append(v_nodal(k, level, cell, mode, float(x) + 1e-16))  # eval_v_nodal_right
            x = 1.0 if l == 0 else float(Fr(2 * c - 1, 2) / cells)
            out.append(v_nodal(k, level, cell, mode, x - 1e-16))             # eval_v_nodal_left
    return np.array(out)


def threshold(M: np.ndarray, atol: float) -> np.ndarray:
    """src/additional_tools.jl:23-42 (dense result; entries with |.| < atol set to zero)"""
    M = np.array(M, dtype=np.float64)
    M[np.abs(M) < atol] = 0.0
    return M


def nodal2points_1D(k: int, max_level: int, atol: float = 1e-15) -> np.ndarray:
    cols = []
    for level, cell, mode in o.hier_index_list(k, max_level):
        cols.append(eval_points_1D(k, max_level, level, cell, mode))
    return threshold(np.array(cols).T, atol)


def points2nodal_1D(k: int, max_level: int, atol: float = 1e-15) -> np.ndarray:
    return threshold(np.linalg.inv(nodal2points_1D(k, max_level)), atol)


# ---- src/1d_nodal_basis.jl:281-362 ----------------------------------------------------------------------------
def nodal2pos_1D(k: int, max_level: int, atol: float = 1e-15) -> np.ndarray:
    ncell = 1 << max_level
    N = k * ncell
    xs, ws = np.polynomial.legendre.leggauss(max(k, 6))
    out = np.zeros((N, N))
    eps_atol = np.spacing(1e-15)                                             # eps(atol) with the default atol
    for j, (nl, nc, nm) in enumerate(o.hier_index_list(k, max_level)):
        w = 1 << max(0, nl - 1)
        nodal_min, nodal_max = (nc - 1) / w, nc / w
        for c in range(1, ncell + 1):
            pos_min, pos_med, pos_max = (c - 1) / ncell, (c - 0.5) / ncell, c / ncell
            if pos_min > nodal_max or pos_max < nodal_min:
                continue
            for m in range(1, k + 1):
                val = 0.0
                for lo, hi in ((pos_min, pos_med), (pos_med, pos_max)):
                    half, mid = 0.5 * (hi - lo), 0.5 * (hi + lo)
                    acc = 0.0
                    for xi, wi in zip(xs, ws):
                        x = mid + half * xi
                        acc += wi * v_nodal(k, nl, nc, nm, x) * o.basis(max_level, c, m, x)
                    val += half * acc
                if abs(val) > eps_atol:
                    out[k * (c - 1) + (m - 1), j] = val
    return threshold(out, atol)


def pos2nodal_1D(k: int, max_level: int, atol: float = 1e-12) -> np.ndarray:
    return threshold(np.linalg.inv(nodal2pos_1D(k, max_level)), atol)


_T1D_CACHE: dict = {}


def transform_1D(k: int, n: int, frm: str, to: str, atol: float = 1e-12) -> np.ndarray:
    """src/1d_nodal_basis.jl:383-414 (dense arrays)"""
    key = (k, n, frm, to, atol)
    if key in _T1D_CACHE:
        return _T1D_CACHE[key]
    Q = lambda: o.hier2pos(k, n).toarray()                                   # modal -> pos
    if (frm, to) == ("nodal", "pos"):
        out = nodal2pos_1D(k, n, atol=atol)
    elif (frm, to) == ("pos", "nodal"):
        out = pos2nodal_1D(k, n, atol=atol)
    elif (frm, to) == ("nodal", "points"):
        out = nodal2points_1D(k, n, atol=atol)
    elif (frm, to) == ("points", "nodal"):
        out = points2nodal_1D(k, n, atol=atol)
    elif (frm, to) == ("nodal", "modal"):
        out = threshold(Q().T @ nodal2pos_1D(k, n), atol)
    elif (frm, to) == ("modal", "nodal"):
        out = threshold(pos2nodal_1D(k, n) @ Q(), atol)
    elif (frm, to) == ("modal", "pos"):
        out = Q()
    elif (frm, to) == ("pos", "modal"):
        out = Q().T
    elif (frm, to) == ("modal", "points"):
        out = threshold(transform_1D(k, n, "nodal", "points", atol) @ transform_1D(k, n, "modal", "nodal", atol), atol)
    elif (frm, to) == ("points", "modal"):
        out = threshold(transform_1D(k, n, "nodal", "modal", atol) @ transform_1D(k, n, "points", "nodal", atol), atol)
    else:
        raise ValueError("A basis was undefined")
    _T1D_CACHE[key] = out
    return out


# ---- src/multidim_nodal_basis.jl:19-125 -----------------------------------------------------------------------
def _relevant_mask(k: int, n: int) -> np.ndarray:
    """relevant_cell_1D for every pair of 1-D indices (row = (l2, c2), column = (l1, c1)); modes ignored"""
    idx = list(o.hier_index_list(k, n))
    N = len(idx)
    lo = np.array([(c - 1) / (1 << max(0, l - 1)) for l, c, _ in idx])
    hi = np.array([c / (1 << max(0, l - 1)) for l, c, _ in idx])
    l2, r2 = lo[:, None], hi[:, None]         # rows: (level2, cell2)
    l1, r1 = lo[None, :], hi[None, :]         # columns: (level1, cell1)
    return ((l2 <= l1) & (r2 >= r1)) | ((l2 >= l1) & (r2 <= r1))


def transform(D: int, k: int, n: int, mat_1D: np.ndarray, scheme: str = "sparse", atol: float = 1e-12):
    """src/multidim_nodal_basis.jl:83-109: entry (i, j) = prod_d mat_1D[i_d, j_d] if every (cell2_d, cell1_d) pair is
    nested (relevant_cell) and |value| >= atol.  Returns scipy CSC in the vector layout."""
    import scipy.sparse as sp
    M = np.where(_relevant_mask(k, n), np.asarray(mat_1D, dtype=np.float64), 0.0)
    blocks, N = o.block_table(D, k, n, scheme)

    def rows_1d(l, ncell):            # 1-D indices of level l in (cell, mode) order, mode fastest
        base = 0 if l == 0 else (1 << (l - 1))
        return (k * (base + np.arange(ncell))[:, None] + np.arange(k)[None, :])       # (cell, mode)

    Is, Js, Vs = [], [], []
    for lv2, off2, ks2 in blocks:                 # row blocks
        r1d = [rows_1d(lv2[d], ks2[d]) for d in range(D)]
        for lv1, off1, ks1 in blocks:             # column blocks
            c1d = [rows_1d(lv1[d], ks1[d]) for d in range(D)]
            # block layout: (m_1..m_D, c_1..c_D) first index fastest; build the 2D-way tensor and reorder
            sub = [M[np.ix_(r1d[d].reshape(-1), c1d[d].reshape(-1))].reshape(ks2[d], k, ks1[d], k) for d in range(D)]
            if any(not s.any() for s in sub):
                continue
            # val[(c2_d, m2_d, c1_d, m1_d) for all d] = prod_d sub[d]; multiplication order d = 1..D from 1.0
            val = np.ones([1] * (4 * D))
            for d in range(D):
                shp = [1] * (4 * D)
                shp[4 * d:4 * d + 4] = sub[d].shape
                val = val * sub[d].reshape(shp)
            # row index = m2 (first dim fastest) then c2 ; column index likewise
            ax_c2 = [4 * d for d in range(D)]
            ax_m2 = [4 * d + 1 for d in range(D)]
            ax_c1 = [4 * d + 2 for d in range(D)]
            ax_m1 = [4 * d + 3 for d in range(D)]
            # C-order flattening has the LAST axis fastest: order axes so that m_1 is last within rows
            val = val.transpose(ax_c2[::-1] + ax_m2[::-1] + ax_c1[::-1] + ax_m1[::-1])
            nr = int(np.prod(ks2)) * k ** D
            nc = int(np.prod(ks1)) * k ** D
            blk = val.reshape(nr, nc)
            ii, jj = np.nonzero(np.abs(blk) >= atol)
            if ii.size:
                Is.append(ii + off2)
                Js.append(jj + off1)
                Vs.append(blk[ii, jj])
    A = sp.csc_matrix((np.concatenate(Vs), (np.concatenate(Is), np.concatenate(Js))), shape=(N, N))
    A.sort_indices()
    return A


def transform_literal(D: int, k: int, n: int, mat_1D: np.ndarray, scheme: str = "sparse", atol: float = 1e-12):
    """the reference's scalar loops (make_column / inner_loop), small cases only"""
    import scipy.sparse as sp
    VD = o.V2Dref(D, k, n, scheme)
    rel = _relevant_mask(k, n)
    one = {lcm: j for j, lcm in enumerate(o.hier_index_list(k, n))}
    I, J, V = [], [], []
    for j, (l1, c1, m1) in enumerate(VD):
        js = [one[(l1[d] - 1, c1[d], m1[d])] for d in range(D)]
        for i, (l2, c2, m2) in enumerate(VD):
            is_ = [one[(l2[d] - 1, c2[d], m2[d])] for d in range(D)]
            if not all(rel[is_[d], js[d]] for d in range(D)):
                continue
            val = 1.0
            for d in range(D):
                val *= mat_1D[is_[d], js[d]]
                if val == 0:
                    break
            if abs(val) < atol:
                continue
            I.append(i); J.append(j); V.append(val)
    N = len(VD)
    return sp.csc_matrix((V, (I, J)), shape=(N, N))


def make_modal2point_matrices(D: int, k: int, n: int):
    m2n = transform(D, k, n, transform_1D(k, n, "modal", "nodal"))
    n2p = transform(D, k, n, transform_1D(k, n, "nodal", "points"))
    return m2n, n2p


def make_point2modal_matrices(D: int, k: int, n: int):
    p2n = transform(D, k, n, transform_1D(k, n, "points", "nodal"))
    n2m = transform(D, k, n, transform_1D(k, n, "nodal", "modal"))
    return p2n, n2m


# ---- src/basic_function_exact_coeffs.jl + src/pdes.jl:165-192 -------------------------------------------------
def get_one_modal_1D(k: int, n: int) -> np.ndarray:
    v = np.zeros(k << n)
    v[0] = 1.0
    return v


def get_xi_modal_1D(k: int, n: int) -> np.ndarray:
    v = np.zeros(k << n)
    v[1] = 1 / np.sqrt(3)
    return v


def vlasov_steprule(D: int, k: int, n: int, Ds, m2n, n2p, p2n, n2m, F_point):
    """src/pdes.jl:165-192: returns (steprule, v_point).  Ds = grad_matrix(2D, k, n) as anything with `@`."""
    one_1D, v_1D = get_one_modal_1D(k, n), get_xi_modal_1D(k, n)
    v_modal = [o.tensor_construct(2 * D, k, n, [v_1D if j - D == i else one_1D for j in range(1, 2 * D + 1)])
               for i in range(1, D + 1)]
    v_point = [n2p @ (m2n @ v) for v in v_modal]

    def steprule(t, f):
        dfdxs = [n2p @ (m2n @ (Ds[d] @ f)) for d in range(D)]
        dfdps = [n2p @ (m2n @ (Ds[d] @ f)) for d in range(D, 2 * D)]
        contrib1 = v_point[0] * dfdxs[0]
        for d in range(1, D):
            contrib1 = contrib1 + v_point[d] * dfdxs[d]
        contrib2 = F_point[0] * dfdps[0]
        for d in range(1, D):
            contrib2 = contrib2 + F_point[d] * dfdps[d]
        return n2m @ (p2n @ (-contrib1 + contrib2))

    return steprule, v_point


def example_force_point(D: int, k: int, n: int, m2n, n2p):
    """F_point of examples/vlasov_evolve.jl:36-44: F_radial(r^2) x_i in the point basis over the 2D-dimensional phase
    space (r^2 and x_i from their exact modal coefficients, src/basic_function_exact_coeffs.jl:12-53)."""
    D2 = 2 * D
    one_1D = get_one_modal_1D(k, n)
    xi_1D = get_xi_modal_1D(k, n)
    xi2_1D = np.zeros(k << n)
    xi2_1D[2] = 2 / (3 * np.sqrt(5))
    xi2_1D = xi2_1D + 1 / 3 * one_1D

    def tc(special, i):
        return o.tensor_construct(D2, k, n, [special if d == i else one_1D for d in range(1, D2 + 1)])

    r2_modal = tc(xi2_1D, 1)
    for i in range(2, D2 + 1):
        r2_modal = r2_modal + tc(xi2_1D, i)
    r2 = n2p @ (m2n @ r2_modal)
    r2[r2 < 0] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.sqrt(r2)
        fr2 = np.where(r2 == 0, 10 / 9, 10 / 3 * (s - np.arctan(s)) / (s ** 2))
    return [(n2p @ (m2n @ tc(xi_1D, i))) * fr2 for i in range(1, D + 1)]
