"""CPU oracle for the GalerkinSparseGrids.jl hot path -- TEST INFRASTRUCTURE ONLY.

This module is a literal, op-for-op restatement (plain Python floats == IEEE
binary64, numpy only for containers) of the reference's algorithm for the path
named in BASELINE.json:north_star.  It is imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product (galerkinsparsegrids.jl_b200/) never imports it.

PARITY STATUS: "parity unpinned" at the bit level.  Julia is not installed in
this image, so the reference itself cannot be executed here and it ships no
golden vectors.  The oracle is pinned against every
known-answer / property assertion the reference's own tests hold for this path
(tests/test_oracle_pins.py): test/elementary.jl:17-27 (cell_index table),
test/vhier_DG.jl:36-65 (V2D/D2V exact round trips), test/differentiation.jl:
14-40 (L2 error bounds of D_op * coeffs), test/hier_DG.jl (reconstruction
bounds) and test/solvers.jl:54-77 (wave energy sqrt(E) ~ sqrt(2) pi).

Every function cites the reference file:line it follows (paths are relative to
/root/reference/).  Known, documented deviations from literal Julia:
  * hquadrature (HCubature 1.4.0, G7K15 adaptive) is replaced by a fixed
    Gauss-Legendre rule that is exact for the polynomial integrands of
    hier2pos (degree <= 2k-2 per cell); differs from Julia at the 1e-16 level.
  * array2poly is @fastmath in the reference (LLVM may contract to FMA); here
    it is plain multiply-then-add.
  * loops that the reference runs over pairs whose result is an exact 0.0
    (and is therefore dropped by its own thresholds) are skipped; the skipped
    set is validated against the full literal loop for small sizes in tests.
"""
from __future__ import annotations

import itertools
import math
from functools import lru_cache

import numpy as np

K_MAX = 10          # src/1d_dg_functions.jl:7
REL_TOL = 1.0e-8    # src/dg_methods.jl:10
ABS_TOL = 1.0e-12   # src/dg_methods.jl:11


# ----------------------------------------------------------------------------
# L0: index set and vector layout
# ----------------------------------------------------------------------------
def cutoff(scheme: str, level1, n: int) -> bool:
    """src/schemes.jl:21-27.  `level1` holds 1-based levels."""
    if scheme == "sparse":
        return sum(level1) > n + len(level1)
    if scheme == "full":
        return False
    raise ValueError(scheme)


def cartesian_indices(shape):
    """Julia CartesianIndices(shape) order: FIRST index fastest, 1-based."""
    D = len(shape)
    for rev in itertools.product(*[range(1, s + 1) for s in reversed(shape)]):
        yield tuple(reversed(rev))


def levels_iter(D: int, n: int, scheme: str):
    """The `for level in CartesianIndices(ls)` + cutoff loop shared by
    src/dg_vmethods.jl:35-142 (1-based level tuples, layout order)."""
    for level in cartesian_indices((n + 1,) * D):
        if cutoff(scheme, level, n):
            continue
        yield level


def cells_of(level1):
    """ks = ntuple(q -> 1<<max(0, level[q]-2), D)  (src/dg_vmethods.jl:40)."""
    return tuple(1 << max(0, l - 2) for l in level1)


def get_size(D: int, k: int, n: int, scheme: str = "sparse") -> int:
    """src/dg_vmethods.jl:35-45."""
    size = 0
    for level in levels_iter(D, n, scheme):
        size += int(np.prod(cells_of(level), dtype=np.int64)) * k ** D
    return size


def V2Dref(D: int, k: int, n: int, scheme: str = "sparse"):
    """src/dg_vmethods.jl:123-142: list index -> (level, cell, mode), all
    1-based tuples, in layout order."""
    out = []
    modes = (k,) * D
    for level in levels_iter(D, n, scheme):
        for cell in cartesian_indices(cells_of(level)):
            for mode in cartesian_indices(modes):
                out.append((level, cell, mode))
    return out


def D2Vref(D: int, k: int, n: int, scheme: str = "sparse"):
    """src/dg_vmethods.jl:102-121: dict (level, cell, mode) -> 1-based index."""
    return {lcm: j + 1 for j, lcm in enumerate(V2Dref(D, k, n, scheme))}


def V2D(D: int, k: int, n: int, vect, scheme: str = "sparse"):
    """src/dg_vmethods.jl:76-100: flat vector -> dict level -> array of shape
    cells (+) of arrays of shape (k,)*D.  Represented as dict level1 ->
    ndarray of shape cells + modes indexed [c1-1,...,cD-1,m1-1,...,mD-1]."""
    vect = np.asarray(vect)
    coeffs = {}
    j = 0
    for level in levels_iter(D, n, scheme):
        ks = cells_of(level)
        blk = np.empty(ks + (k,) * D, dtype=vect.dtype)
        for cell in cartesian_indices(ks):
            for mode in cartesian_indices((k,) * D):
                blk[tuple(c - 1 for c in cell) + tuple(m - 1 for m in mode)] = vect[j]
                j += 1
        coeffs[level] = blk
    assert j == len(vect)
    return coeffs


def D2V(D: int, k: int, n: int, coeffs, scheme: str = "sparse"):
    """src/dg_vmethods.jl:48-73."""
    size = get_size(D, k, n, scheme)
    first = next(iter(coeffs.values()))
    vect = np.empty(size, dtype=first.dtype)
    j = 0
    for level in levels_iter(D, n, scheme):
        blk = coeffs[level]
        for cell in cartesian_indices(cells_of(level)):
            for mode in cartesian_indices((k,) * D):
                vect[j] = blk[tuple(c - 1 for c in cell) + tuple(m - 1 for m in mode)]
                j += 1
    return vect


def get_index_1D(k: int, l: int, c: int, m: int) -> int:
    """src/additional_tools.jl:18-20 (l, c, m 1-based; result 1-based).
    `1<<((l-2)%UInt)`: for l == 1 the shift count wraps to 2^64-1 and the
    shift yields 0."""
    shift = (1 << (l - 2)) if l >= 2 else 0
    return k * (shift + (c - 1)) + m


def block_table(D: int, k: int, n: int, scheme: str = "sparse"):
    """Helper (derived from the D2V loop, src/dg_vmethods.jl:53-73): list of
    (level0 tuple, offset, cells tuple) per multi-level block, layout order."""
    out = []
    off = 0
    for level in levels_iter(D, n, scheme):
        ks = cells_of(level)
        out.append((tuple(l - 1 for l in level), off, ks))
        off += int(np.prod(ks, dtype=np.int64)) * k ** D
    return out, off


# ----------------------------------------------------------------------------
# L1: basis construction (src/dg_basis.jl) and evaluation (1d_dg_functions.jl)
# ----------------------------------------------------------------------------
def product_matrix(i: int, j: int, n: int) -> float:
    """src/dg_basis.jl:21-35."""
    k = int(round(n / 2))
    if i < k and j < k:
        return (1 + (-1) ** (i + j)) / (1 + i + j)
    elif i >= k and j < k:
        return (1 - (-1) ** ((i - k) + j)) / (1 + (i - k) + j)
    elif i < k and j >= k:
        return product_matrix(j, i, n)
    else:
        return (1 + (-1) ** ((i - k) + (j - k))) / (1 + (i - k) + (j - k))


def inner_product_vv(v1, v2) -> float:
    """src/dg_basis.jl:37-50."""
    value = 0.0
    n = len(v1)
    for i in range(n):
        for j in range(n):
            if v1[i] == 0 or v2[j] == 0:
                continue
            value += product_matrix(i, j, n) * v1[i] * v2[j]
    return value


def inner_product_vj(v1, j: int) -> float:
    """src/dg_basis.jl:52-60 (against x^(j-1), j 1-based)."""
    value = 0.0
    n = len(v1)
    for i in range(n):
        value += product_matrix(i, j - 1, n) * v1[i]
    return value


def _axpy(y, a, x):
    """y - a*x elementwise, multiply then subtract (Julia `y -= a * x`)."""
    return [yi - a * xi for yi, xi in zip(y, x)]


def gram_schmidt(Q_initial):
    """src/dg_basis.jl:68-85."""
    n = len(Q_initial[0])
    k = int(round(n / 2))
    Q_final = [list(q) for q in Q_initial]
    for i in range(k):
        for j in range(i):
            proj = inner_product_vv(Q_initial[i], Q_final[j]) / inner_product_vv(Q_final[j], Q_final[j])
            Q_final[i] = _axpy(Q_final[i], proj, Q_final[j])
        nrm = math.sqrt(inner_product_vv(Q_final[i], Q_final[i]))
        Q_final[i] = [q / nrm for q in Q_final[i]]
    return Q_final


def legendre(k: int):
    """src/dg_basis.jl:88-94."""
    Q = [[1.0 if i == j else 0.0 for i in range(1, 2 * (k + 1) + 1)] for j in range(1, k + 2)]
    return gram_schmidt(Q)


def orthogonalize_1(Q_initial):
    """src/dg_basis.jl:104-120."""
    n = len(Q_initial[0])
    k = int(round(n / 2))
    Q_final = [list(q) for q in Q_initial]
    legendre_polys = legendre(k - 1)
    for i in range(k):
        for j in range(k):
            proj = inner_product_vv(Q_initial[i], legendre_polys[j]) / inner_product_vv(
                legendre_polys[j], legendre_polys[j])
            Q_final[i] = _axpy(Q_final[i], proj, legendre_polys[j])
    return Q_final


def orthogonalize_2(Q_initial):
    """src/dg_basis.jl:129-146 (note: `fi` is taken from Q_initial)."""
    n = len(Q_initial[0])
    k = int(round(n / 2))
    Q_final = [list(q) for q in Q_initial]
    for i in range(1, k):                # i = 1:k-1
        fi = list(Q_initial[i - 1])
        for j in range(i + 1, k + 1):    # j = i+1:k
            a = inner_product_vj(Q_final[j - 1], k + i) / inner_product_vj(fi, k + i)
            Q_final[j - 1] = _axpy(Q_final[j - 1], a, fi)
    return Q_final


def gram_schmidt_rev(Q_initial):
    """src/dg_basis.jl:154-169."""
    n = len(Q_initial[0])
    k = int(round(n / 2))
    Q_final = [[0.0] * n for _ in range(k)]
    for i in range(k, 0, -1):
        fi = list(Q_initial[i - 1])
        Q_final[i - 1] = fi
        for j in range(i + 1, k + 1):
            proj = inner_product_vv(fi, Q_final[j - 1]) / inner_product_vv(Q_final[j - 1], Q_final[j - 1])
            Q_final[i - 1] = _axpy(Q_final[i - 1], proj, Q_final[j - 1])
        nrm = math.sqrt(inner_product_vv(Q_final[i - 1], Q_final[i - 1]))
        Q_final[i - 1] = [q / nrm for q in Q_final[i - 1]]
    return Q_final


def dg_basis(k: int):
    """src/dg_basis.jl:185-191."""
    Q = [[1.0 if j == (i - k) else 0.0 for i in range(1, 2 * k + 1)] for j in range(1, k + 1)]
    Q = orthogonalize_1(Q)
    Q = orthogonalize_2(Q)
    Q = gram_schmidt_rev(Q)
    return Q


@lru_cache(maxsize=None)
def leg_coeffs():
    """src/1d_dg_functions.jl:35: leg_coeffs = legendre(K_max)."""
    return tuple(tuple(q) for q in legendre(K_MAX))


@lru_cache(maxsize=None)
def dg_coeffs(k: int):
    """src/1d_dg_functions.jl:52-62: dg_coeffs[k][:, mode] (returned as a
    tuple over modes of 2k-vectors)."""
    return tuple(tuple(q) for q in dg_basis(k))


def _flipsign(a: float, x: float) -> float:
    """Julia flipsign: flips on the sign BIT of x (so -0.0 flips)."""
    return -a if math.copysign(1.0, x) < 0 else a


def array2poly(v, x: float) -> float:
    """src/1d_dg_functions.jl:15-28 (without @fastmath contraction)."""
    if abs(x) > 1:
        return 0.0
    n = len(v)
    k = n // 2
    s = 0.0
    for i in range(k, 0, -1):
        s *= x
        s += v[i - 1] + _flipsign(v[i - 1 + k], x)
    return s


def LegendreP(kk: int, x: float) -> float:
    """src/1d_dg_functions.jl:38-41."""
    if kk > K_MAX:
        raise ValueError("DomainError")
    return array2poly(leg_coeffs()[kk], x)


def h(k: int, mode: int, x: float) -> float:
    """src/1d_dg_functions.jl:66-69."""
    if mode > k:
        raise ValueError("DomainError")
    return array2poly(dg_coeffs(k)[mode - 1], x)


def leg(mode: int, x: float) -> float:
    """src/1d_dg_functions.jl:83-85."""
    return math.sqrt(2.0) * LegendreP(mode - 1, 2 * x - 1)


def basis(level: int, cell: int, mode: int, x: float) -> float:
    """src/1d_dg_functions.jl:91-93 (position basis, level >= 0, cell 1-based)."""
    return leg(mode, (1 << level) * x - (cell - 1)) * (2.0) ** (level / 2)


def v(k: int, level: int, cell: int, mode: int, x: float) -> float:
    """src/dg_methods.jl:27-36 (hierarchical basis, 0-based level, 1-based
    cell and mode)."""
    if level == 0:
        return LegendreP(mode - 1, 2 * x - 1) * math.sqrt(2.0)
    return h(k, mode, (1 << level) * x - (2 * cell - 1)) * math.sqrt(1.0 * (1 << level))


def cell_index(x: float, l: int) -> int:
    """src/dg_methods.jl:70-79 (0-based level l; returns 1-based cell)."""
    if l <= 1:
        return 1
    if x >= 1:
        return 2 ** (l - 1)
    return 1 + int(math.floor(2 ** (l - 1) * x))


# ----------------------------------------------------------------------------
# quadrature helper (stands in for hquadrature on polynomial integrands)
# ----------------------------------------------------------------------------
@lru_cache(maxsize=None)
def _gauss(npts: int):
    xs, ws = np.polynomial.legendre.leggauss(npts)
    return tuple(float(a) for a in xs), tuple(float(a) for a in ws)


def quad(f, a: float, b: float, npts: int = 12) -> float:
    xs, ws = _gauss(npts)
    half = 0.5 * (b - a)
    mid = 0.5 * (b + a)
    s = 0.0
    for x, w in zip(xs, ws):
        s += w * f(mid + half * x)
    return s * half


def pos_vcoeffs_DG(k: int, level: int, f, cells=None):
    """src/1d_dg_functions.jl:103-117.  `cells` (optional, 1-based iterable)
    restricts the loop to cells where f is not identically zero."""
    ncell = 1 << level
    vcoeffs = [0.0] * (ncell * k)
    rng = range(1, ncell + 1) if cells is None else cells
    for cell in rng:
        for mode in range(1, k + 1):
            left = (cell - 1) / ncell
            right = cell / ncell
            vcoeffs[(cell - 1) * k + (mode - 1)] = quad(
                lambda x: basis(level, cell, mode, x) * f(x), left, right)
    return vcoeffs


# ----------------------------------------------------------------------------
# minimal CSC container with Julia's SparseArrays semantics
# ----------------------------------------------------------------------------
class CSC:
    """Julia SparseMatrixCSC{Float64,Int64} stand-in: 0-based arrays inside,
    `.julia()` returns the 1-based Int64 fields that cross the C ABI."""

    def __init__(self, m, n, colptr, rowval, nzval):
        self.m, self.n = int(m), int(n)
        self.colptr = np.asarray(colptr, dtype=np.int64)
        self.rowval = np.asarray(rowval, dtype=np.int64)
        self.nzval = np.asarray(nzval, dtype=np.float64)

    @property
    def nnz(self):
        return int(self.colptr[-1])

    @staticmethod
    def from_coo(I, J, V, m, n):
        """sparse(I, J, V, m, n, +) with 0-based I, J: sorted rows per column,
        duplicates summed in input order."""
        cols = [dict() for _ in range(n)]
        for i, j, val in zip(I, J, V):
            c = cols[j]
            if i in c:
                c[i] += val
            else:
                c[i] = val
        colptr = [0]
        rowval, nzval = [], []
        for c in cols:
            for i in sorted(c):
                rowval.append(i)
                nzval.append(c[i])
            colptr.append(len(rowval))
        return CSC(m, n, colptr, rowval, nzval)

    def col(self, j):
        a, b = self.colptr[j], self.colptr[j + 1]
        return self.rowval[a:b], self.nzval[a:b]

    def transpose(self):
        """copy(A') : rows of the result's columns come out sorted."""
        I, J, V = [], [], []
        for j in range(self.n):
            r, v_ = self.col(j)
            I.extend([j] * len(r))
            J.extend(r.tolist())
            V.extend(v_.tolist())
        return CSC.from_coo(I, J, V, self.n, self.m)

    def toarray(self):
        out = np.zeros((self.m, self.n))
        for j in range(self.n):
            r, v_ = self.col(j)
            out[r, j] = v_
        return out

    def julia(self):
        return self.m, self.n, self.colptr + 1, self.rowval + 1, self.nzval.copy()

    def matvec(self, x):
        """SparseArrays `A*x` (Julia 1.0 mul!): column scatter
        y[rowval[p]] += nzval[p]*x[col], plain multiply-then-add."""
        y = np.zeros(self.m)
        cp, rv, nz = self.colptr, self.rowval, self.nzval
        yl = y.tolist()
        xl = np.asarray(x, dtype=np.float64).tolist()
        rvl = rv.tolist()
        nzl = nz.tolist()
        for j in range(self.n):
            xj = xl[j]
            for p in range(cp[j], cp[j + 1]):
                yl[rvl[p]] += nzl[p] * xj
        return np.array(yl)


def spmatmul(A: CSC, B: CSC) -> CSC:
    """SparseArrays.spmatmul (Julia 1.0 stdlib, Gustavson): for each column i
    of B, for each stored B[j,i] in row order, for each stored A[k,j]:
    x[k] += A[k,j]*B[j,i] (first touch assigns).  Structural fill is kept
    (numerical zeros are NOT dropped); rows are sorted afterwards."""
    assert A.n == B.m
    colptr = [0]
    rowval, nzval = [], []
    for i in range(B.n):
        acc = {}
        rB, vB = B.col(i)
        for j, nzB in zip(rB.tolist(), vB.tolist()):
            rA, vA = A.col(j)
            for kk, a in zip(rA.tolist(), vA.tolist()):
                nzC = a * nzB
                if kk in acc:
                    acc[kk] += nzC
                else:
                    acc[kk] = nzC
        for kk in sorted(acc):
            rowval.append(kk)
            nzval.append(acc[kk])
        colptr.append(len(rowval))
    return CSC(A.m, B.n, colptr, rowval, nzval)


def sp_neg_add(Dm: CSC, LF: CSC) -> CSC:
    """`-D + LF` (src/1d_derivative.jl:109).  Julia's sparse map drops entries
    whose result is exactly zero (`_map_zeropres!`)."""
    I, J, V = [], [], []
    for j in range(Dm.n):
        acc = {}
        r, v_ = Dm.col(j)
        for i, val in zip(r.tolist(), v_.tolist()):
            acc[i] = -val
        r, v_ = LF.col(j)
        for i, val in zip(r.tolist(), v_.tolist()):
            acc[i] = (acc[i] + val) if i in acc else val
        for i in sorted(acc):
            if acc[i] != 0.0:
                I.append(i); J.append(j); V.append(acc[i])
    return CSC.from_coo(I, J, V, Dm.m, Dm.n)


# ----------------------------------------------------------------------------
# L3: the 1-D operator  H = periodic_DLF_matrix(k, n)
# ----------------------------------------------------------------------------
def hier_index_list(k: int, max_level: int):
    """Column order of hier2pos (src/1d_dg_functions.jl:246-249): level 0..n,
    cell 1..2^max(0,level-1), mode 1..k."""
    out = []
    for level in range(0, max_level + 1):
        for cell in range(1, (1 << max(0, level - 1)) + 1):
            for mode in range(1, k + 1):
                out.append((level, cell, mode))
    return out


def hier2pos(k: int, max_level: int, atol: float = ABS_TOL, literal: bool = False) -> CSC:
    """src/1d_dg_functions.jl:242-263.  Column j = position-basis coefficients
    of hierarchical function j.  With literal=False the quadratures over cells
    outside the function's support (integrand identically 0.0 -> result 0.0 ->
    dropped by `abs(ans[i]) > atol`) are skipped."""
    I, J, V = [], [], []
    ncell = 1 << max_level
    for j, (level, cell, mode) in enumerate(hier_index_list(k, max_level)):
        f = lambda x, level=level, cell=cell, mode=mode: v(k, level, cell, mode, x)
        if literal:
            cells = None
        else:
            width = ncell >> max(0, level - 1)          # fine cells under the support
            cells = range((cell - 1) * width + 1, cell * width + 1)
        ans = pos_vcoeffs_DG(k, max_level, f, cells)
        for i, a in enumerate(ans):
            if abs(a) > atol:
                I.append(i); J.append(j); V.append(a)
    N = k * ncell
    return CSC.from_coo(I, J, V, N, N)


def symbolic_diff(vv):
    """src/derivative_matrix_elements.jl:80-94."""
    n = len(vv)
    k = n // 2
    ans = [0.0] * n
    for i in range(1, n + 1):
        if i < k:
            ans[i - 1] = i * vv[i]
        elif i > k and i < 2 * k:
            ans[i - 1] = (i - k) * vv[i]
        else:
            ans[i - 1] = 0.0
    return ans


@lru_cache(maxsize=None)
def legendreDlegendre(mode1: int, mode2: int) -> float:
    """src/derivative_matrix_elements.jl:101-103."""
    lc = leg_coeffs()
    return inner_product_vv(lc[mode1 - 1], symbolic_diff(lc[mode2 - 1]))


def legvDv(level, cell1, mode1, cell2, mode2) -> float:
    """src/derivative_matrix_elements.jl:106-112."""
    if cell1 == cell2:
        return (1 << (level + 1)) * legendreDlegendre(mode1, mode2)
    return 0.0


def D_matrix_1d(k: int, level: int) -> CSC:
    """src/1d_derivative.jl:21-44 (volume term).  Off-diagonal cell pairs give
    an exact 0.0 and are skipped without evaluation."""
    I, J, V = [], [], []
    for cell1 in range(1, (1 << level) + 1):
        for mode1 in range(1, k + 1):
            i = (cell1 - 1) * k + (mode1 - 1)
            for mode2 in range(1, k + 1):
                j = (cell1 - 1) * k + (mode2 - 1)
                val = legvDv(level, cell1, mode1, cell1, mode2)
                if abs(val) > 1.0e-15:
                    I.append(i); J.append(j); V.append(val)
    N = k * (1 << level)
    return CSC.from_coo(I, J, V, N, N)


def periodic_legvLFv(level, cell1, mode1, cell2, mode2, alpha=0) -> float:
    """src/1d_derivative.jl:52-74.  The `tiny = 5.0e-16` one-sided-limit
    arithmetic is kept literally (SURVEY.md headline fact 4)."""
    point1 = (cell2 - 1) / (1 << level)      # Rational -> exact dyadic Float64
    point2 = cell2 / (1 << level)
    tiny = 5.0e-16

    left1 = basis(level, cell1, mode1, point1 - tiny)
    right1 = basis(level, cell1, mode1, point1 + tiny)
    left2 = basis(level, cell1, mode1, point2 - tiny)
    right2 = basis(level, cell1, mode1, point2 + tiny)

    if cell2 == (1 << level):
        right2 = basis(level, cell1, mode1, 0.0 + tiny)
    if cell2 == 1:
        left1 = basis(level, cell1, mode1, 1.0 - tiny)

    LF1 = 0.5 * (left1 + right1) + alpha * (right1 - left1)
    LF2 = 0.5 * (left2 + right2) + alpha * (right2 - left2)

    val1 = basis(level, cell2, mode2, point1 + tiny)
    val2 = basis(level, cell2, mode2, point2 - tiny)
    return LF2 * val2 - LF1 * val1


def periodic_LF_matrix(k: int, level: int, alpha=0, literal: bool = False) -> CSC:
    """src/1d_derivative.jl:76-100.  literal=False visits only cell pairs that
    are periodic neighbours (|cell1-cell2| <= 1 mod 2^level); every other pair
    evaluates `basis` outside its support on all four one-sided points and
    yields an exact 0.0, which `abs(val) > 1.0e-15` drops."""
    I, J, V = [], [], []
    ncell = 1 << level
    for cell1 in range(1, ncell + 1):
        if literal:
            cand = range(1, ncell + 1)
        else:
            cand = sorted({(cell1 - 2) % ncell + 1, cell1, cell1 % ncell + 1})
        for mode1 in range(1, k + 1):
            i = (cell1 - 1) * k + (mode1 - 1)
            for cell2 in cand:
                for mode2 in range(1, k + 1):
                    j = (cell2 - 1) * k + (mode2 - 1)
                    val = periodic_legvLFv(level, cell1, mode1, cell2, mode2, alpha)
                    if abs(val) > 1.0e-15:
                        I.append(i); J.append(j); V.append(val)
    N = k * ncell
    return CSC.from_coo(I, J, V, N, N)


def periodic_pos_DLF_matrix(k: int, max_level: int, literal: bool = False) -> CSC:
    """src/1d_derivative.jl:108-111: A = (-D + LF)'."""
    A = sp_neg_add(D_matrix_1d(k, max_level), periodic_LF_matrix(k, max_level, literal=literal))
    return A.transpose()


_H_CACHE = {}


def periodic_DLF_matrix(k: int, max_level: int, basis_name: str = "hier", literal: bool = False) -> CSC:
    """src/1d_derivative.jl:136-148 -> :113-117: H = Q' * (A * Q)."""
    key = (k, max_level, basis_name, literal)
    if key in _H_CACHE:
        return _H_CACHE[key]
    A = periodic_pos_DLF_matrix(k, max_level, literal=literal)
    if basis_name == "pos":
        out = A
    elif basis_name == "hier":
        Q = hier2pos(k, max_level, literal=literal)
        out = spmatmul(Q.transpose(), spmatmul(A, Q))
    else:
        raise ValueError(basis_name)
    _H_CACHE[key] = out
    return out


# ----------------------------------------------------------------------------
# L3: D-dimensional assembly (src/multidim_derivative.jl)
# ----------------------------------------------------------------------------
def D_matrix_literal(D: int, d: int, k: int, n: int, scheme: str = "sparse", H: CSC | None = None) -> CSC:
    """src/multidim_derivative.jl:19-65, the literal column loop with Dict
    look-ups (d is 1-based).  Small cases only."""
    VD = V2Dref(D, k, n, scheme)
    DV = {lcm: j for j, lcm in enumerate(VD)}
    V2D_1D = V2Dref(1, k, n, "sparse")
    D2V_1D = {lcm: j for j, lcm in enumerate(V2D_1D)}
    if H is None:
        H = periodic_DLF_matrix(k, n)
    I, J, V = [], [], []
    for j, (lv, cl, md) in enumerate(VD):
        j1 = D2V_1D[((lv[d - 1],), (cl[d - 1],), (md[d - 1],))]
        rows, vals = H.col(j1)
        for i1, val in zip(rows.tolist(), vals.tolist()):
            l1, c1, m1 = V2D_1D[i1]
            level2 = lv[:d - 1] + l1 + lv[d:]
            if cutoff(scheme, level2, n):
                continue
            cell2 = cl[:d - 1] + c1 + cl[d:]
            mode2 = md[:d - 1] + m1 + md[d:]
            I.append(DV[(level2, cell2, mode2)]); J.append(j); V.append(val)
    N = len(VD)
    return CSC.from_coo(I, J, V, N, N)


def pole_tables(D: int, d: int, k: int, n: int, scheme: str = "sparse"):
    """[derived identity, SURVEY.md 8(a) a10] For sweep axis d (1-based) return
    an int64 array `poles` of shape (npoles, Nmax) holding, for every pole, the
    0-based global index of its 1-D entries in 1-D layout order (padded with -1),
    plus the pole lengths.  Built with numpy from the block table."""
    blocks, N = block_table(D, k, n, scheme)
    by_level = {lv: (off, ks) for lv, off, ks in blocks}
    kD = k ** D
    groups = {}
    for lv, off, ks in blocks:
        other = lv[:d - 1] + lv[d:]
        groups.setdefault(other, []).append(lv[d - 1])
    pole_rows, pole_len = [], []
    for other, lds in groups.items():
        lds = sorted(lds)
        assert lds == list(range(len(lds)))
        p = len(lds) - 1
        Np = k * (1 << p)
        # shape of "everything but axis d": modes of other dims, cells of other dims
        lv0 = other[:d - 1] + (0,) + other[d - 1:]
        ks0 = by_level[lv0][1]
        cells_other = ks0[:d - 1] + ks0[d:]
        npole = k ** (D - 1) * int(np.prod(cells_other, dtype=np.int64))
        idx = np.empty((npole, Np), dtype=np.int64)
        # strides inside a block: mode j -> k^j ; cell j -> kD * prod(ks[:j])
        col = 0
        for ld in lds:
            lv = other[:d - 1] + (ld,) + other[d - 1:]
            off, ks = by_level[lv]
            mstr = [k ** j for j in range(D)]
            cstr = [kD * int(np.prod(ks[:j], dtype=np.int64)) for j in range(D)]
            # enumerate other-dims (mode, cell) combos in a fixed order
            om = [np.arange(k) * mstr[j] for j in range(D) if j != d - 1]
            oc = [np.arange(ks[j]) * cstr[j] for j in range(D) if j != d - 1]
            base = np.zeros(1, dtype=np.int64)
            for arr in om + oc:
                base = (base[:, None] + arr[None, :]).reshape(-1)
            assert base.size == npole
            for c in range(ks[d - 1]):
                for m in range(k):
                    idx[:, col] = off + base + c * cstr[d - 1] + m * mstr[d - 1]
                    col += 1
        assert col == Np
        pole_rows.append(idx)
        pole_len.append(Np)
    return pole_rows, pole_len, N


def D_matrix_poles(D: int, d: int, k: int, n: int, scheme: str = "sparse", H: CSC | None = None):
    """Assemble D_d through the per-pole principal-sub-block identity; returns
    scipy CSC (0-based) -- used for mid-size cross-checks and CPU baselines.
    Equal (entry for entry, same values) to D_matrix_literal; tested."""
    import scipy.sparse as sp
    if H is None:
        H = periodic_DLF_matrix(k, n)
    Hs = sp.csc_matrix((H.nzval, H.rowval, H.colptr), shape=(H.m, H.n))
    groups, lens, N = pole_tables(D, d, k, n, scheme)
    Is, Js, Vs = [], [], []
    for idx, Np in zip(groups, lens):
        sub = Hs[:Np, :Np].tocoo()
        Is.append(idx[:, sub.row].reshape(-1))
        Js.append(idx[:, sub.col].reshape(-1))
        Vs.append(np.broadcast_to(sub.data, (idx.shape[0], sub.data.size)).reshape(-1))
    A = sp.csc_matrix((np.concatenate(Vs), (np.concatenate(Is), np.concatenate(Js))), shape=(N, N))
    A.sort_indices()
    return A


def apply_D_poles(D: int, d: int, k: int, n: int, x, scheme: str = "sparse", H: CSC | None = None,
                  tables=None):
    """y = D_d x evaluated pole by pole with each output row accumulated in
    ascending-column order, multiply-then-add -- the same per-row summation
    order as the reference's CSC column scatter (src/pdes.jl:63 `RHS*x`)."""
    if H is None:
        H = periodic_DLF_matrix(k, n)
    if tables is None:
        tables = pole_tables(D, d, k, n, scheme)
    groups, lens, N = tables
    x = np.asarray(x, dtype=np.float64)
    y = np.zeros(N)
    for idx, Np in zip(groups, lens):
        X = x[idx]                                   # (npole, Np)
        Y = np.zeros_like(X)
        for j in range(Np):                          # ascending column order
            rows, vals = H.col(j)
            keep = rows < Np
            rows, vals = rows[keep], vals[keep]
            if rows.size:
                Y[:, rows] += vals[None, :] * X[:, j:j + 1]
        y[idx] = Y
    return y


# ----------------------------------------------------------------------------
# L2: projection of 1-D functions, tensor_construct, reconstruct_DG
# ----------------------------------------------------------------------------
def coeffs_1d(k: int, n: int, f, npts: int = 20):
    """1-D hierarchical coefficients of f in VECTOR layout (what
    `vcoeffs_DG(1, k, n, f)` returns, src/dg_vmethods.jl:149-179).  The
    reference integrates each coefficient with adaptive hcubature over the
    function's support (src/dg_methods.jl:85-93, rtol 1e-8); here each half of
    the support (where the basis function is a polynomial) is integrated with
    a fixed Gauss rule.  These vectors are INPUTS to both oracle and GPU."""
    out = []
    for (level, cell, mode) in hier_index_list(k, n):
        w = 1 << max(0, level - 1)
        a, b = (cell - 1) / w, cell / w
        mid = 0.5 * (a + b)
        g = lambda x: f(x) * v(k, level, cell, mode, x)
        out.append(quad(g, a, mid, npts) + quad(g, mid, b, npts))
    return np.array(out)


def tensor_construct(D: int, k: int, n: int, vcoeff_array, scheme: str = "sparse"):
    """src/tensor_construct.jl:19-63, vector overload: coefficient of
    prod_d f_d(x_d); `val = one(T); for d in 1:D val *= coeff[d][l][c][m]`."""
    assert len(vcoeff_array) == D
    blocks, N = block_table(D, k, n, scheme)
    out = np.empty(N)
    for lv, off, ks in blocks:
        # 1-D index of (l, c, m): k*(2^(l-1)+c) + m  (0-based; l = 0 -> m)
        val = np.ones((), dtype=np.float64)
        # build with broadcasting in the block's column-major (m1..mD, c1..cD) order;
        # multiplication order d = 1..D as in the reference
        shape_m = [k] * D
        arr = np.ones([1] * (2 * D))
        for dd in range(D):
            l = lv[dd]
            base = 0 if l == 0 else (1 << (l - 1))
            idx = k * (base + np.arange(ks[dd]))[None, :] + np.arange(k)[:, None]   # (m, c)
            f1 = np.asarray(vcoeff_array[dd], dtype=np.float64)[idx]
            shp = [1] * (2 * D)
            shp[dd] = k
            shp[D + dd] = ks[dd]
            arr = arr * f1.reshape(shp)
        size = arr.size
        out[off:off + size] = arr.reshape(-1, order="F")
    return out


def reconstruct_DG(D: int, k: int, n: int, vect, xs, scheme: str = "sparse") -> float:
    """src/dg_methods.jl:150-165 on the vector layout (dict <-> vector is the
    pure permutation V2D).  Summation runs over multi-levels in layout order
    (the reference iterates Dict keys, i.e. hash order: unpinned), modes first
    dim fastest, `value += coeff[mode] * V(...)` with V's product accumulated
    from one(T) over i = 1..D (src/dg_methods.jl:47-54)."""
    blocks, N = block_table(D, k, n, scheme)
    vect = np.asarray(vect)
    value = 0.0
    kD = k ** D
    for lv, off, ks in blocks:
        cell = [cell_index(xs[i], lv[i]) for i in range(D)]
        lin = 0
        stride = 1
        for i in range(D):
            lin += (cell[i] - 1) * stride
            stride *= ks[i]
        base = off + lin * kD
        vals = [[v(k, lv[i], cell[i], m, xs[i]) for m in range(1, k + 1)] for i in range(D)]
        for e, mode in enumerate(cartesian_indices((k,) * D)):
            ans = 1.0
            for i in range(D):
                ans *= vals[i][mode[i] - 1]
            value += float(vect[base + e]) * ans
    return value


def _array2poly_vec(vv, x):
    """array2poly (src/1d_dg_functions.jl:15-28) on an array of x: same operation order as `array2poly`
    above (Horner, multiply-then-add, flipsign on the sign BIT, 0 outside [-1, 1])."""
    x = np.asarray(x, dtype=np.float64)
    half = len(vv) // 2
    neg = np.signbit(x)
    s = np.zeros_like(x)
    for i in range(half - 1, -1, -1):
        t = np.where(neg, -vv[i + half], vv[i + half])
        s = s * x + (vv[i] + t)
    return np.where(np.abs(x) > 1.0, 0.0, s)


def reconstruct_DG_batch(D: int, k: int, n: int, vect, pts, scheme: str = "sparse"):
    """reconstruct_DG (src/dg_methods.jl:150-165) for an (npts, D) array of points, vectorised over the
    points with numpy; per point the SAME operation order as `reconstruct_DG` above (multi-levels in layout
    order, modes first dim fastest, product accumulated from 1.0 over i = 1..D, value += coeff * product).
    Validated against the scalar restatement in tests/test_oracle_pins.py."""
    blocks, N = block_table(D, k, n, scheme)
    vect = np.asarray(vect, dtype=np.float64)
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, D)
    npts = pts.shape[0]
    kD = k ** D
    leg = leg_coeffs()
    dg = dg_coeffs(k)
    sqrt2 = math.sqrt(2.0)
    # 1-D tables: cell (1-based) and basis values for every (dim, level, mode)
    cell = np.empty((D, n + 1, npts), dtype=np.int64)
    val = np.empty((D, n + 1, k, npts))
    for i in range(D):
        x = pts[:, i]
        for l in range(n + 1):
            if l <= 1:
                c = np.ones(npts, dtype=np.int64)
            else:
                c = np.where(x >= 1.0, 1 << (l - 1), 1 + np.floor((1 << (l - 1)) * x).astype(np.int64))
            cell[i, l] = c
            for m in range(1, k + 1):
                if l == 0:                      # v: LegendreP(m-1, 2x-1)*sqrt(2)   (src/dg_methods.jl:27-36)
                    val[i, l, m - 1] = _array2poly_vec(leg[m - 1], 2.0 * x - 1.0) * sqrt2
                else:                           # h(k, m, 2^l x - (2c-1)) * 2^(l/2)
                    val[i, l, m - 1] = _array2poly_vec(dg[m - 1], (1 << l) * x - (2 * c - 1)) * math.sqrt(1.0 * (1 << l))
    value = np.zeros(npts)
    modes = list(cartesian_indices((k,) * D))
    for lv, off, ks in blocks:
        lin = np.zeros(npts, dtype=np.int64)
        stride = 1
        for i in range(D):
            lin += (cell[i, lv[i]] - 1) * stride
            stride *= ks[i]
        base = off + lin * kD
        for e, mode in enumerate(modes):
            ans = np.ones(npts)
            for i in range(D):
                ans = ans * val[i, lv[i], mode[i] - 1]
            value += vect[base + e] * ans
    return value


# ----------------------------------------------------------------------------
# L4: drivers (fixed-step classical RK4; ODE.jl's adaptive ode45/ode78 are
# third-party and not under /root/reference -- see SURVEY.md 8c)
# ----------------------------------------------------------------------------
def rk4(rhs, y0, dt: float, nsteps: int):
    """Classical RK4 with the stage order documented in DESIGN.md:
    k1=f(y); k2=f(y+dt/2 k1); k3=f(y+dt/2 k2); k4=f(y+dt k3);
    y += dt/6 (k1 + 2 k2 + 2 k3 + k4)."""
    y = np.array(y0, dtype=np.float64)
    for _ in range(nsteps):
        k1 = rhs(y)
        k2 = rhs(y + (0.5 * dt) * k1)
        k3 = rhs(y + (0.5 * dt) * k2)
        k4 = rhs(y + dt * k3)
        y = y + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    return y


def advect_rhs(mats, a):
    """u' = -sum_d a_d D_d u built from grad_matrix
    (src/multidim_derivative.jl:67-69; the operator vlasov_evolve uses,
    src/pdes.jl:179-180)."""
    def rhs(u):
        acc = np.zeros_like(u)
        for ad, A in zip(a, mats):
            acc += ad * (A @ u)
        return -acc
    return rhs


def wave_rhs(mats):
    """[u; v]' = [v; L u], L = sum_d D_d*D_d (src/pdes.jl:22-49,
    src/multidim_derivative.jl:71-79); L u evaluated as sum_d D_d (D_d u)."""
    def rhs(y):
        N = y.size // 2
        u, vv = y[:N], y[N:]
        lu = np.zeros(N)
        for A in mats:
            lu += A @ (A @ u)
        return np.concatenate([vv, lu])
    return rhs


def laplacian_matrix_ref(mats):
    """laplacian_matrix (src/multidim_derivative.jl:71-79) in the REFERENCE's form: the explicit sparse
    products `lap += D_op * D_op`, then `lap * x`.  `mats` are scipy CSC D_d (mid sizes); the literal
    Gustavson product `spmatmul` of this module is used by the small-size pin."""
    lap = None
    for A in mats:
        P = (A @ A).tocsc()
        lap = P if lap is None else (lap + P).tocsc()
    lap.sort_indices()
    return lap


def wave_rhs_ref(L):
    """[u; v]' = [v; L u] with the assembled L (src/pdes.jl:22-49: RHS = [[0 I];[L 0]])."""
    def rhs(y):
        N = y.size // 2
        return np.concatenate([y[N:], L @ y[:N]])
    return rhs


def energy(mats, y) -> float:
    """src/pdes.jl:258-273: E = sum_d |D_d u|^2 + |udot|^2."""
    N = y.size // 2
    u, ud = y[:N], y[N:]
    return float(sum(np.sum((A @ u) ** 2) for A in mats) + np.sum(ud ** 2))
