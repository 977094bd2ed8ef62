"""gsg_b200 -- host-side mirror of GalerkinSparseGrids.jl's operator API over libgsgb200.so.

Julia is not available in this image, so the host side above the C ABI (include/gsg_b200.h) is
written in Python and mirrors the reference's names, argument meaning and error behaviour for
the hot path: `get_size`, `V2D`/`D2V`, `cell_index`, `periodic_DLF_matrix`, `D_matrix`,
`grad_matrix`, `laplacian_matrix`, `tensor_construct`, `reconstruct_DG`, `mcerr`,
`wave_evolve`, `energy_func` (reference: src/GalerkinSparseGrids.jl:32-71).  All arithmetic on
the path runs in the CUDA library; there is no CPU fallback and nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
from typing import Sequence

import numpy as np

from ._lib import GsgError, LIB_PATH, SIGNATURES, check, lib

__all__ = [
    "GsgError", "Plan", "get_plan", "get_size", "cell_index", "basis_v", "basis_tables",
    "periodic_DLF_matrix", "coeffs_DG", "vcoeffs_DG", "tensor_construct", "V2D", "D2V", "V2Dref",
    "D2Vref", "D_matrix", "grad_matrix", "laplacian_matrix", "reconstruct_DG", "mcerr",
    "wave_evolve", "wave_evolve_1D", "advect_evolve", "energy_func", "energy_func_1D", "pos_vcoeffs_DG", "OdeIntegrator",
    "ode_solve", "VlasovRHS", "vlasov_evolve", "RHS_VLASOV", "write_operators", "write_solution", "read_dump", "RHS_ADVECT", "RHS_WAVE", "RHS_CSR", "spmv_csc", "CsrMatrix", "device_info",
    "launch_count",
]

_SCHEME = {"sparse": 0, "full": 1}


def _scheme(scheme: str) -> int:
    try:
        return _SCHEME[scheme]
    except KeyError:
        raise ValueError(f"ArgumentError: scheme={scheme!r}") from None


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _devptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor / anything with data_ptr(), or a raw int."""
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(int(t))


# ---------------------------------------------------------------------------------------------
# setup mirrors (CPU side of the library)
# ---------------------------------------------------------------------------------------------
def launch_count() -> int:
    return int(lib.gsg_launch_count())


def device_info(device: int = 0) -> dict:
    buf = C.create_string_buffer(256)
    check(lib.gsg_device_info(device, buf, 256))
    name, sms, cc, mem = buf.value.decode().split(";")
    return {"name": name, "sm_count": int(sms), "cc": cc, "global_mem": int(mem)}


def block_pattern(n: int) -> np.ndarray:
    """(2^n, 2^n) boolean structural block pattern of periodic_DLF_matrix(k, n) (include/gsg_b200.h)."""
    nq = 1 << n
    out = np.zeros((nq, nq), dtype=np.uint8)
    check(lib.gsg_block_pattern(int(n), _ptr(out)))
    return out.astype(bool)


def flat_tables(D: int, k: int, n: int, d: int, scheme: str = "sparse"):
    """CPU-side copy of the flat kernel's tables for direction d (1-based): (groups [ngroups, 20] int64 =
    {base[0..16], p, S, nitems}, cells [ncells, 3] int32 = {group, item, 1-D cell})."""
    ng, ncell = C.c_int64(0), C.c_int64(0)
    check(lib.gsg_debug_flat_tables(D, k, n, _scheme(scheme), d, None, None, C.byref(ng), C.byref(ncell)))
    groups = np.zeros((ng.value, 20), dtype=np.int64)
    cells = np.zeros((ncell.value, 3), dtype=np.int32)
    check(lib.gsg_debug_flat_tables(D, k, n, _scheme(scheme), d, groups.ctypes.data_as(C.c_void_p),
                                    cells.ctypes.data_as(C.c_void_p), C.byref(ng), C.byref(ncell)))
    return groups, cells


def rowtile_program(D: int, k: int, n: int, p: int, budget_bytes: int = 226 * 1024, nrg: int = 4):
    """CPU-side copy of the row-tile kernel's tile program for pole class p (library's own H(k, n), multi-cells of
    k^D doubles): dict with tiles [nt, 48] int32 {nx, rec_ofs, rec_bytes, grp0, rg_end[4], xq[40]}, groups [ng, 8]
    int32 {q[4], rofs, nrec, partial mask, 0} and the record blob (bytes)."""
    cnt = (C.c_int64 * 3)()
    check(lib.gsg_debug_rowtile_program(D, k, n, p, budget_bytes, nrg, None, None, None, cnt))
    tiles = np.zeros((cnt[0], 48), dtype=np.int32)
    groups = np.zeros((cnt[1], 8), dtype=np.int32)
    blob = np.zeros(cnt[2], dtype=np.uint8)
    vp_ = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib.gsg_debug_rowtile_program(D, k, n, p, budget_bytes, nrg, vp_(tiles), vp_(groups), vp_(blob), cnt))
    return {"tiles": tiles, "groups": groups, "blob": blob}


def rowtile_pole_order(k: int, A: int, PI: int, nslots: int) -> np.ndarray:
    """lane order of the row-tile kernel: in-cell pole offsets, padding lanes as ~offset"""
    out = np.zeros(nslots, dtype=np.int32)
    check(lib.gsg_debug_rowtile_pole_order(k, A, PI, nslots, out.ctypes.data_as(C.c_void_p)))
    return out


def get_size(D: int, k: int, n: int, scheme: str = "sparse") -> int:
    """get_size(Val(D), k, n, Val(scheme)) -- src/dg_vmethods.jl:35-45."""
    out = C.c_int64()
    check(lib.gsg_get_size(D, k, n, _scheme(scheme), C.byref(out)))
    return out.value


def cell_index(x: float, l: int) -> int:
    """cell_index(x, l) -- src/dg_methods.jl:70-79 (0-based level, 1-based cell)."""
    out = C.c_int64()
    check(lib.gsg_cell_index(float(x), int(l), C.byref(out)))
    return out.value


def basis_v(k: int, level: int, cell: int, mode: int, x) -> np.ndarray:
    """v(k, level, cell, mode, x) -- src/dg_methods.jl:27-36, vectorised over x."""
    x = _f64(np.atleast_1d(x))
    out = np.empty_like(x)
    check(lib.gsg_basis_v(k, level, cell, mode, _ptr(x), x.size, _ptr(out)))
    return out


def basis_tables(k: int):
    """(leg_coeffs, dg_coeffs[k]) -- src/1d_dg_functions.jl:35,52-62."""
    leg = np.empty((11, 22))
    dg = np.empty((k, 2 * k))
    check(lib.gsg_basis_tables(k, _ptr(leg), _ptr(dg)))
    return leg, dg


_H_CACHE: dict = {}


def periodic_DLF_matrix(k: int, max_level: int, basis: str = "hier"):
    """periodic_DLF_matrix(k, n; basis) -- src/1d_derivative.jl:136-148.  Returns a
    scipy.sparse.csc_matrix (0-based); the "nodal"/"point" variants are broken in the
    reference (they call the undefined modal2points_1D) and raise here too."""
    import scipy.sparse as sp
    if basis not in ("hier", "pos"):
        raise ValueError(f"ArgumentError: basis={basis!r}")
    key = (k, max_level, basis)
    if key not in _H_CACHE:
        b = 0 if basis == "hier" else 1
        nnz = C.c_int64(0)
        check(lib.gsg_dlf_matrix(k, max_level, b, C.byref(nnz), None, None, None))
        N = k << max_level
        colptr = np.empty(N + 1, dtype=np.int64)
        rowval = np.empty(nnz.value, dtype=np.int64)
        nzval = np.empty(nnz.value, dtype=np.float64)
        check(lib.gsg_dlf_matrix(k, max_level, b, C.byref(nnz), _ptr(colptr), _ptr(rowval), _ptr(nzval)))
        _H_CACHE[key] = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(N, N))
    return _H_CACHE[key]


def _hier_index_list(k: int, n: int):
    for level in range(n + 1):
        for cell in range(1, (1 << max(0, level - 1)) + 1):
            for mode in range(1, k + 1):
                yield level, cell, mode


def vcoeffs_DG(D: int, k: int, n: int, f, scheme: str = "sparse", npts: int = 20) -> np.ndarray:
    """vcoeffs_DG(1, k, n, f) -- src/dg_vmethods.jl:149-179, 1-D only: the projection of
    arbitrary D-dimensional closures by adaptive cubature stays on the host language side
    (SURVEY.md section 2); product inputs for D > 1 come from tensor_construct, as in the
    reference's own examples (examples/traveling_wave.jl:18-48)."""
    if D != 1:
        raise NotImplementedError("vcoeffs_DG: only D == 1; use tensor_construct for D > 1")
    xs, ws = np.polynomial.legendre.leggauss(npts)
    out = np.empty(k << n)
    for j, (level, cell, mode) in enumerate(_hier_index_list(k, n)):
        w = 1 << max(0, level - 1)
        a, b = (cell - 1) / w, cell / w
        mid = 0.5 * (a + b)
        tot = 0.0
        for lo, hi in ((a, mid), (mid, b)):
            half, c = 0.5 * (hi - lo), 0.5 * (hi + lo)
            x = c + half * xs
            fx = np.array([f(float(xi)) for xi in x], dtype=np.float64)
            tot += half * float(np.dot(ws, fx * basis_v(k, level, cell, mode, x)))
        out[j] = tot
    return out


def coeffs_DG(D: int, k: int, n: int, f, scheme: str = "sparse"):
    """coeffs_DG(1, k, n, f) -- src/dg_methods.jl:113-143, as a dict (see V2D)."""
    return V2D(D, k, n, vcoeffs_DG(D, k, n, f, scheme=scheme), scheme=scheme)


def tensor_construct(D: int, k: int, n: int, vcoeff_array: Sequence, scheme: str = "sparse") -> np.ndarray:
    """tensor_construct(D, k, n, [v_1..v_D]; scheme), vector overload --
    src/tensor_construct.jl:57-63."""
    if len(vcoeff_array) != D:
        raise ValueError("tensor_construct: need D coefficient vectors")
    vs = [_f64(v) for v in vcoeff_array]
    for v in vs:
        if v.size != (k << n):
            raise ValueError("tensor_construct: 1-D vectors must have length k*2^n")
    arr = (C.c_void_p * D)(*[v.ctypes.data for v in vs])
    out = np.empty(get_size(D, k, n, scheme))
    check(lib.gsg_tensor_construct(D, k, n, _scheme(scheme), arr, _ptr(out)))
    return out


# ---------------------------------------------------------------------------------------------
# vector <-> dict layout (pure permutations; src/dg_vmethods.jl:48-142)
# ---------------------------------------------------------------------------------------------
def _levels(D: int, n: int, scheme: str):
    full = _scheme(scheme) == 1
    for rev in itertools.product(range(1, n + 2), repeat=D):
        level = rev[::-1]                      # first index fastest
        if not full and sum(level) > n + D:    # src/schemes.jl:21-23
            continue
        yield level


def V2D(D: int, k: int, n: int, vect, scheme: str = "sparse") -> dict:
    """V2D -- src/dg_vmethods.jl:76-100.  dict: 1-based level tuple -> ndarray indexed
    [m_1..m_D, c_1..c_D] (0-based), i.e. the block exactly as it sits in the vector."""
    vect = np.asarray(vect)
    out, j = {}, 0
    for level in _levels(D, n, scheme):
        cells = tuple(1 << max(0, l - 2) for l in level)
        size = int(np.prod(cells)) * k ** D
        out[level] = vect[j:j + size].reshape((k,) * D + cells, order="F").copy()
        j += size
    if j != vect.size:
        raise ValueError("V2D: vector length does not match get_size")
    return out


def D2V(D: int, k: int, n: int, coeffs: dict, scheme: str = "sparse") -> np.ndarray:
    """D2V -- src/dg_vmethods.jl:48-73."""
    parts = [np.asarray(coeffs[level]).reshape(-1, order="F") for level in _levels(D, n, scheme)]
    return np.concatenate(parts) if parts else np.empty(0)


def V2Dref(D: int, k: int, n: int, scheme: str = "sparse"):
    """V2Dref -- src/dg_vmethods.jl:123-142: list of 1-based (level, cell, mode) tuples."""
    out = []
    for level in _levels(D, n, scheme):
        cells = tuple(1 << max(0, l - 2) for l in level)
        for crev in itertools.product(*[range(1, c + 1) for c in cells[::-1]]):
            for mrev in itertools.product(range(1, k + 1), repeat=D):
                out.append((level, crev[::-1], mrev[::-1]))
    return out


def D2Vref(D: int, k: int, n: int, scheme: str = "sparse") -> dict:
    """D2Vref -- src/dg_vmethods.jl:102-121 (1-based indices)."""
    return {lcm: j + 1 for j, lcm in enumerate(V2Dref(D, k, n, scheme))}


# ---------------------------------------------------------------------------------------------
# plan + operators
# ---------------------------------------------------------------------------------------------
class Plan:
    """Device-resident operator plan for (D, k, n, scheme).  Replaces the assembled
    SparseMatrixCSC of grad_matrix / laplacian_matrix (src/multidim_derivative.jl:61-79)."""

    def __init__(self, D: int, k: int, n: int, scheme: str = "sparse", H=None, device: int = 0):
        import scipy.sparse as sp
        self.D, self.k, self.n, self.scheme = D, k, n, scheme
        if H is None:
            H = periodic_DLF_matrix(k, n)
        H = sp.csc_matrix(H)
        H.sort_indices()
        colptr = (H.indptr.astype(np.int64) + 1)
        rowval = (H.indices.astype(np.int64) + 1)
        nzval = _f64(H.data)
        handle = C.c_void_p()
        check(lib.gsg_plan_create(D, k, n, _scheme(scheme), H.shape[0], _ptr(colptr), _ptr(rowval),
                                  _ptr(nzval), device, C.byref(handle)))
        self._h = handle
        size = C.c_int64()
        check(lib.gsg_plan_size(self._h, C.byref(size)))
        self.size = size.value
        check(lib.gsg_plan_dev_size(self._h, C.byref(size)))
        self.dev_size = size.value      # padded length of device-layout vectors

    def close(self):
        if getattr(self, "_h", None):
            lib.gsg_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host vectors ----------------------------------------------------------------------
    def _vec(self, x) -> np.ndarray:
        x = _f64(x)
        if x.shape != (self.size,):
            raise ValueError(f"DimensionMismatch: expected length {self.size}, got {x.shape}")
        return x

    def apply_D(self, d: int, x) -> np.ndarray:
        x = self._vec(x)
        y = np.empty_like(x)
        check(lib.gsg_apply_D(self._h, d, _ptr(x), _ptr(y)))
        return y

    def apply_grad(self, a, x) -> np.ndarray:
        x = self._vec(x)
        a = _f64(a)
        if a.size != self.D:
            raise ValueError("apply_grad: need D coefficients")
        y = np.empty_like(x)
        check(lib.gsg_apply_grad(self._h, _ptr(a), _ptr(x), _ptr(y)))
        return y

    def apply_laplacian(self, x) -> np.ndarray:
        x = self._vec(x)
        y = np.empty_like(x)
        check(lib.gsg_apply_laplacian(self._h, _ptr(x), _ptr(y)))
        return y

    def rk4_advect(self, a, y, dt: float, nsteps: int) -> np.ndarray:
        y = self._vec(y).copy()
        a = _f64(a)
        check(lib.gsg_rk4_advect(self._h, _ptr(a), _ptr(y), float(dt), int(nsteps)))
        return y

    def rk4_wave(self, u, v, dt: float, nsteps: int):
        u = self._vec(u).copy()
        v = self._vec(v).copy()
        check(lib.gsg_rk4_wave(self._h, _ptr(u), _ptr(v), float(dt), int(nsteps)))
        return u, v

    def energy(self, u, udot) -> float:
        u, udot = self._vec(u), self._vec(udot)
        out = C.c_double()
        check(lib.gsg_energy(self._h, _ptr(u), _ptr(udot), C.byref(out)))
        return out.value

    def reconstruct(self, vcoeffs, points) -> np.ndarray:
        """points: (npts, D) array (row i = point i; same memory as Julia's D x npts)."""
        vcoeffs = self._vec(vcoeffs)
        pts = _f64(points).reshape(-1, self.D)
        out = np.empty(pts.shape[0])
        check(lib.gsg_reconstruct(self._h, _ptr(vcoeffs), _ptr(pts), pts.shape[0], _ptr(out)))
        return out

    # -- device vectors (torch tensors or raw pointers) ------------------------------------
    def set_stream(self, stream) -> None:
        s = getattr(stream, "cuda_stream", stream)
        check(lib.gsg_plan_set_stream(self._h, C.c_void_p(int(s) if s else 0)))

    def sync(self) -> None:
        check(lib.gsg_plan_sync(self._h))

    def set_rk4_mode(self, mode: int) -> None:
        """0 = automatic (Taylor form for the linear right-hand sides), 1 = always the staged form."""
        check(lib.gsg_plan_set_rk4_mode(self._h, int(mode)))

    def set_flat(self, mode: int) -> None:
        """0 = tiled class kernels, 1 = flat kernel (one launch per right-hand side), 2 = automatic (default)."""
        check(lib.gsg_plan_set_flat(self._h, int(mode)))

    def describe(self) -> str:
        """which kernel serves which pole classes, per direction"""
        buf = C.create_string_buffer(8192)
        check(lib.gsg_plan_describe(self._h, buf, 8192))
        return buf.value.decode()

    @property
    def flat_active(self) -> bool:
        out = C.c_int(0)
        check(lib.gsg_plan_flat_active(self._h, C.byref(out)))
        return bool(out.value)

    def set_partition(self, rank: int, nranks: int) -> None:
        """Block partition over nranks = 2^b ranks (include/gsg_b200.h): afterwards the plan's sweeps
        cover only the pole groups this rank computes."""
        check(lib.gsg_plan_set_partition(self._h, int(rank), int(nranks)))

    @property
    def cell_stride(self) -> int:
        """doubles per multi-cell in the device layout (k^D rounded up to even)"""
        kd = self.k ** self.D
        return kd + (kd & 1)

    def partition_blocks(self, kind: int, d: int = 0):
        """(offsets, sizes, partner): kind 0 = blocks this rank owns, kind 1 = level_d == 0 blocks exchanged
        with `partner` along partition dimension d (empty if d is not one)."""
        cnt, partner = C.c_int64(), C.c_int(-1)
        check(lib.gsg_plan_partition_blocks(self._h, kind, d, None, None, C.byref(cnt), C.byref(partner)))
        offs = np.zeros(max(cnt.value, 1), dtype=np.int64)
        sizes = np.zeros(max(cnt.value, 1), dtype=np.int64)
        check(lib.gsg_plan_partition_blocks(self._h, kind, d, _ptr(offs), _ptr(sizes), C.byref(cnt), C.byref(partner)))
        return offs[:cnt.value], sizes[:cnt.value], partner.value

    def rk4_taylor_cells_dev(self, cells, u, v1, v2, v3, v4, c1, c2, c3, c4) -> None:
        check(lib.gsg_rk4_taylor_cells_dev(self._h, _devptr(cells), int(cells.numel()), _devptr(u), _devptr(v1),
                                           _devptr(v2), _devptr(v3), _devptr(v4), float(c1), float(c2), float(c3),
                                           float(c4)))

    def rk_stage_dev(self, length: int, u, k, acc, w, cw: float, ca: float, first: bool) -> None:
        check(lib.gsg_rk_stage_dev(self._h, int(length), _devptr(u), _devptr(k), _devptr(acc), _devptr(w),
                                   float(cw), float(ca), 1 if first else 0))

    def rk_final_dev(self, length: int, u, k, acc, ca: float) -> None:
        check(lib.gsg_rk_final_dev(self._h, int(length), _devptr(u), _devptr(k), _devptr(acc), float(ca)))

    def profile_enable(self, on=True) -> None:
        """True: time every streaming-kernel launch; an int > 1: only the first `on` launches; False: off."""
        check(lib.gsg_profile_enable(self._h, int(on)))

    def profile_read(self):
        """(launches, total_ms, dofs) of the timed streaming-kernel launches."""
        n, ms, dofs = C.c_int64(), C.c_double(), C.c_double()
        check(lib.gsg_profile_read(self._h, C.byref(n), C.byref(ms), C.byref(dofs)))
        return n.value, ms.value, dofs.value

    def pack_dev(self, ref_vec, dev_vec) -> None:
        """reference-layout device vector (length size) -> device layout (length dev_size)."""
        check(lib.gsg_pack_dev(self._h, _devptr(ref_vec), _devptr(dev_vec)))

    def unpack_dev(self, dev_vec, ref_vec) -> None:
        check(lib.gsg_unpack_dev(self._h, _devptr(dev_vec), _devptr(ref_vec)))

    def tensor_construct_dev(self, vcoeff_array, device=None):
        """tensor_construct(D, k, n, [v_1..v_D]) expanded on the device: returns a zero-padded torch tensor in
        device layout (src/tensor_construct.jl:19-63)."""
        import torch
        if len(vcoeff_array) != self.D:
            raise ValueError("tensor_construct: need D coefficient vectors")
        vs = [_f64(v) for v in vcoeff_array]
        for v in vs:
            if v.size != (self.k << self.n):
                raise ValueError("tensor_construct: 1-D vectors must have length k*2^n")
        arr = (C.c_void_p * self.D)(*[v.ctypes.data for v in vs])
        out = torch.zeros(self.dev_size, dtype=torch.float64, device=device if device is not None else "cuda")
        torch.cuda.current_stream(out.device).synchronize()      # the fill runs on torch's stream, the expansion on the plan's
        check(lib.gsg_tensor_construct_dev(self._h, arr, _devptr(out)))
        return out

    def to_device(self, host_vec, device="cuda:0"):
        """numpy reference-layout vector -> zero-padded torch tensor in device layout."""
        import torch
        ref = torch.from_numpy(self._vec(host_vec)).to(device)
        dev = torch.zeros(self.dev_size, dtype=torch.float64, device=device)
        torch.cuda.current_stream(dev.device).synchronize()      # copy and fill run on torch's stream, the pack on the plan's
        self.pack_dev(ref, dev)
        self.sync()
        return dev

    def to_host(self, dev_vec) -> np.ndarray:
        import torch
        ref = torch.empty(self.size, dtype=torch.float64, device=dev_vec.device)
        torch.cuda.current_stream(dev_vec.device).synchronize()  # work torch may still have queued on dev_vec
        self.unpack_dev(dev_vec, ref)
        self.sync()
        return ref.cpu().numpy()

    def apply_D_dev(self, d: int, x, y, alpha: float = 1.0, beta: float = 0.0) -> None:
        check(lib.gsg_apply_D_dev(self._h, d, alpha, _devptr(x), beta, _devptr(y)))

    def apply_grad_dev(self, a, x, y) -> None:
        a = _f64(a)
        check(lib.gsg_apply_grad_dev(self._h, _ptr(a), _devptr(x), _devptr(y)))

    def apply_dirs_dev(self, c, dirs, x, y, beta: float = 0.0) -> None:
        """y = beta*y + sum_{d in dirs} c[d-1] D_d x (1-based directions; pairs inside `dirs` are swept fused)."""
        c = _f64(c)
        mask = 0
        for d in dirs:
            mask |= 1 << (d - 1)
        check(lib.gsg_apply_dirs_dev(self._h, _ptr(c), mask, float(beta), _devptr(x), _devptr(y)))

    def apply_laplacian_dev(self, x, y, tmp) -> None:
        check(lib.gsg_apply_laplacian_dev(self._h, _devptr(x), _devptr(y), _devptr(tmp)))

    def rk4_advect_dev(self, a, y, dt: float, nsteps: int) -> None:
        a = _f64(a)
        check(lib.gsg_rk4_advect_dev(self._h, _ptr(a), _devptr(y), float(dt), int(nsteps)))

    def rk4_wave_dev(self, u, v, dt: float, nsteps: int) -> None:
        check(lib.gsg_rk4_wave_dev(self._h, _devptr(u), _devptr(v), float(dt), int(nsteps)))

    def reconstruct_dev(self, vcoeffs, points, npts: int, out) -> None:
        check(lib.gsg_reconstruct_dev(self._h, _devptr(vcoeffs), _devptr(points), int(npts), _devptr(out)))


_PLANS: dict = {}


def get_plan(D: int, k: int, n: int, scheme: str = "sparse") -> Plan:
    key = (D, k, n, scheme)
    if key not in _PLANS:
        _PLANS[key] = Plan(D, k, n, scheme)
    return _PLANS[key]


class _Operator:
    """What `D_matrix(...)` returns: supports `A * x` and `A @ x` like a SparseMatrixCSC."""

    def __init__(self, plan: Plan):
        self.plan = plan
        self.shape = (plan.size, plan.size)

    def __mul__(self, x):
        return self.matvec(x)

    __matmul__ = __mul__


class DOperator(_Operator):
    def __init__(self, plan: Plan, d: int):
        super().__init__(plan)
        if not 1 <= d <= plan.D:
            raise ValueError("BoundsError: axis d out of range")
        self.d = d

    def matvec(self, x):
        return self.plan.apply_D(self.d, x)


class LaplacianOperator(_Operator):
    def matvec(self, x):
        return self.plan.apply_laplacian(x)


def D_matrix(D: int, d: int, k: int, n: int, scheme: str = "sparse") -> DOperator:
    """D_matrix(D, d, k, n; scheme) -- src/multidim_derivative.jl:61-65 (matrix-free)."""
    return DOperator(get_plan(D, k, n, scheme), d)


def grad_matrix(D: int, k: int, n: int, scheme: str = "sparse"):
    """grad_matrix -- src/multidim_derivative.jl:67-69."""
    return [D_matrix(D, d, k, n, scheme) for d in range(1, D + 1)]


def laplacian_matrix(D: int, k: int, n: int, scheme: str = "sparse") -> LaplacianOperator:
    """laplacian_matrix -- src/multidim_derivative.jl:71-79 (applied as sum_d D_d(D_d x))."""
    return LaplacianOperator(get_plan(D, k, n, scheme))


# ---------------------------------------------------------------------------------------------
# reconstruction and error measurement
# ---------------------------------------------------------------------------------------------
def reconstruct_DG(coeffs, xs, D: int | None = None, k: int | None = None, n: int | None = None,
                   scheme: str = "sparse"):
    """reconstruct_DG(coeffs, xs) -- src/dg_methods.jl:150-165, batched.  `coeffs` is either
    the dict V2D returns (D, k, n inferred) or a vector with D, k, n given; `xs` is one
    point (length D) or an (npts, D) array.  Returns a float or an array."""
    if isinstance(coeffs, dict):
        first_level = next(iter(coeffs))
        D = len(first_level)
        k = np.asarray(coeffs[first_level]).shape[0]
        nlev = len(coeffs)
        n = max(max(l) for l in coeffs) - 1
        if scheme == "sparse" and nlev != math.comb(n + D, D):
            scheme = "full"
        vect = D2V(D, k, n, coeffs, scheme=scheme)
    else:
        if D is None or k is None or n is None:
            raise ValueError("reconstruct_DG: D, k, n required with a coefficient vector")
        vect = coeffs
    pts = _f64(xs)
    single = pts.ndim == 1 and pts.size == D
    out = get_plan(D, k, n, scheme).reconstruct(vect, pts.reshape(-1, D))
    return float(out[0]) if single else out


def mcerr(coeffs, g, D: int, k: int, n: int, count: int = 1000, scheme: str = "sparse", rng=None) -> float:
    """mcerr(x -> reconstruct_DG(dict, x), g, D; count) -- src/error_measure.jl:12-19,39-41:
    sqrt(mean over `count` uniform points of (f - g)^2), with f evaluated in one batch."""
    rng = np.random.default_rng() if rng is None else rng
    pts = rng.random((count, D))
    f = reconstruct_DG(coeffs, pts, D, k, n, scheme) if not isinstance(coeffs, dict) else reconstruct_DG(coeffs, pts)
    gv = np.array([g(p) for p in pts])
    return float(np.sqrt(np.mean((f - gv) ** 2)))


# ---------------------------------------------------------------------------------------------
# PDE drivers (src/pdes.jl)
# ---------------------------------------------------------------------------------------------
def _steps(time0: float, time1: float, dt: float):
    nsteps = max(1, int(math.ceil((time1 - time0) / dt - 1e-12)))
    return nsteps, (time1 - time0) / nsteps


RHS_ADVECT, RHS_WAVE, RHS_CSR, RHS_VLASOV = 0, 1, 2, 3


class OdeIntegrator:
    """Device-resident adaptive Runge-Kutta stepper (gsg_ode_*): the drop-in for ODE.jl's ode45 (Dormand-Prince
    5(4)) / ode78 (Fehlberg 7(8)) calls of src/pdes.jl:62-68, 113-119, 206-213."""

    def __init__(self, plan: "Plan", rhs_kind: int, y0, t0: float, t1: float, order: str = "45", a=None, A=None,
                 reltol: float = 0.0, abstol: float = 0.0, vlasov=None):
        if order not in ("45", "78"):
            raise ValueError("ArgumentError(:order)")
        self.plan, self.A, self.vlasov = plan, A, vlasov     # keep the matrix / right-hand side alive
        y0 = _f64(y0)
        self.n = y0.size
        a_arr = _f64(a) if a is not None else None
        h = C.c_void_p()
        if rhs_kind == RHS_VLASOV:
            check(lib.gsg_ode_create_vlasov(vlasov._h, int(order), float(reltol), float(abstol), _ptr(y0), float(t0),
                                            float(t1), C.byref(h)))
        else:
            check(lib.gsg_ode_create(plan._h, rhs_kind, _ptr(a_arr) if a_arr is not None else None,
                                     A._h if A is not None else None, int(order), float(reltol), float(abstol), _ptr(y0),
                                     float(t0), float(t1), C.byref(h)))
        self._h = h
        self.t, self.done = float(t0), False

    def step(self):
        """advance to the next accepted step; returns (t, dt, done)"""
        t, dt, done = C.c_double(), C.c_double(), C.c_int()
        check(lib.gsg_ode_step(self._h, C.byref(t), C.byref(dt), C.byref(done)))
        self.t, self.done = t.value, bool(done.value)
        return t.value, dt.value, self.done

    def state(self) -> np.ndarray:
        y = np.empty(self.n)
        check(lib.gsg_ode_state(self._h, _ptr(y)))
        return y

    def interp(self, tquery: float) -> np.ndarray:
        y = np.empty(self.n)
        check(lib.gsg_ode_interp(self._h, float(tquery), _ptr(y)))
        return y

    def stats(self) -> dict:
        a, r, f = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.gsg_ode_stats(self._h, C.byref(a), C.byref(r), C.byref(f)))
        return {"accepted": a.value, "rejected": r.value, "rhs_evals": f.value}

    def close(self):
        if getattr(self, "_h", None):
            lib.gsg_ode_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ode_solve(plan: "Plan", rhs_kind: int, y0, tspan, order: str = "45", points: str = "all", a=None, A=None,
              reltol: float = 0.0, abstol: float = 0.0, stats: dict | None = None, vlasov=None):
    """`ode45(F, y0, tspan; points)` / `ode78(...)` of ODE.jl on the device: returns (tout, yout) as ODE.jl does --
    points="all": every accepted step plus the requested times strictly inside a step (Hermite); points="specified":
    the requested times only (src/pdes.jl:207-213 uses this with range(t0, t1, length=nout))."""
    tspan = [float(t) for t in tspan]
    ode = OdeIntegrator(plan, rhs_kind, y0, tspan[0], tspan[-1], order, a=a, A=A, reltol=reltol, abstol=abstol, vlasov=vlasov)
    tdir = 1.0 if tspan[-1] > tspan[0] else -1.0
    tout, yout = [tspan[0]], [_f64(y0).copy()]
    if points == "specified":
        tout = list(tspan)
        yout = yout + [None] * (len(tspan) - 1)
    it = 1
    t_prev = tspan[0]
    while not ode.done:
        t, dt, done = ode.step()
        if points == "specified":
            while it < len(tspan) and (tdir * tspan[it] < tdir * t or done):
                yout[it] = ode.state() if tspan[it] == t else ode.interp(tspan[it])
                it += 1
        else:
            while it < len(tspan) and tdir * t_prev < tdir * tspan[it] < tdir * t:
                yout.append(ode.interp(tspan[it]))
                tout.append(tspan[it])
                it += 1
            yout.append(ode.state())
            tout.append(t)
        t_prev = t
    if stats is not None:
        stats.update(ode.stats())
    ode.close()
    return tout, yout


def wave_evolve(D: int, k: int, n: int, f0coeffs, v0coeffs, time0: float, time1: float,
                order: str = "45", scheme: str = "sparse", dt: float | None = None, nout: int = 2, **kwargs):
    """wave_evolve(D, k, n, f0coeffs, v0coeffs, t0, t1; order, scheme) -- src/pdes.jl:54-70.
    order "45" / "78": ODE.jl's adaptive ode45 / ode78 on the device (the reference's branches, (tout, yout) with
    every accepted step); order "4": the added fixed-step RK4 branch (resident, `nout` equally spaced outputs).
    Returns (times, [state_i]) with state = [u; v] like ODE.jl's (tout, yout)."""
    plan = get_plan(D, k, n, scheme)
    u, v = _f64(f0coeffs).copy(), _f64(v0coeffs).copy()
    if order in ("45", "78"):
        tout, yout = ode_solve(plan, RHS_WAVE, np.concatenate([u, v]), [time0, time1], order=order, **kwargs)
        return np.array(tout), yout
    if order != "4":
        raise ValueError("ArgumentError(:order)")
    if dt is None:
        dt = 0.25 * 2.785 / (math.sqrt(D) * 8.081 * (1 << n))   # RK4 stability, rho(H) = 8.081 2^n
    times = np.linspace(time0, time1, nout)
    states = [np.concatenate([u, v])]
    for t0, t1 in zip(times[:-1], times[1:]):
        ns, h = _steps(t0, t1, dt)
        u, v = plan.rk4_wave(u, v, h, ns)
        states.append(np.concatenate([u, v]))
    return times, states


class VlasovRHS:
    """`steprule` of vlasov_evolve (src/pdes.jl:174-192) resident on the device (gsg_vlasov_*)."""

    def __init__(self, plan: "Plan", m2n, n2p, p2n, n2m, F_point):
        self.plan = plan
        self.mats = [A if isinstance(A, CsrMatrix) else CsrMatrix(A) for A in (m2n, n2p, p2n, n2m)]
        Fs = [_f64(F) for F in F_point]
        if len(Fs) != plan.D // 2 or any(F.shape != (plan.size,) for F in Fs):
            raise ValueError("DimensionMismatch: F_point must hold D vectors of length N")
        arr = (C.c_void_p * len(Fs))(*[F.ctypes.data for F in Fs])
        h = C.c_void_p()
        check(lib.gsg_vlasov_create(plan._h, *[A._h for A in self.mats], arr, C.byref(h)))
        self._h = h

    def __call__(self, t, f_modal) -> np.ndarray:
        f = self.plan._vec(f_modal)
        out = np.empty_like(f)
        check(lib.gsg_vlasov_rhs(self._h, _ptr(f), _ptr(out)))
        return out

    def v_point(self, i: int) -> np.ndarray:
        out = np.empty(self.plan.size)
        check(lib.gsg_vlasov_v_point(self._h, int(i), _ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib.gsg_vlasov_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def vlasov_evolve(D: int, k: int, n: int, m2n, n2p, p2n, n2m, f0_modal, F_point, time0: float, time1: float,
                  nout: int = 2, order: str = "45", scheme: str = "sparse", dump: str | None = None, **kwargs):
    """vlasov_evolve(D, k, n, m2n, n2p, p2n, n2m, f0_modal, F_point, t0, t1, nout; order, scheme) -- src/pdes.jl:131-227.
    The four transform matrices are inputs, as in the reference (scipy sparse, or resident CsrMatrix handles); the
    right-hand side and the adaptive integrator run on the device; outputs at range(t0, t1, length=nout) when
    points="specified" (what examples/vlasov_evolve.jl asks for).  `dump`: write the reference's operator / solution
    file layout (write_operators / write_solution) to this path."""
    if order not in ("45", "78"):
        raise ValueError("ArgumentError(:order)")
    plan = get_plan(2 * D, k, n, scheme)
    rhs = VlasovRHS(plan, m2n, n2p, p2n, n2m, F_point)
    tspan = np.linspace(time0, time1, nout)
    soln = ode_solve(plan, RHS_VLASOV, f0_modal, tspan, order=order, vlasov=rhs, **kwargs)
    if dump:
        ops = {"m2n": m2n, "n2p": n2p, "p2n": p2n, "n2m": n2m}
        write_operators(dump, D, k, n, {name: A for name, A in ops.items() if not isinstance(A, CsrMatrix)})
        write_solution(dump, soln)
    rhs.close()
    return soln


def pos_vcoeffs_DG(k: int, max_level: int, f, npts: int = 20) -> np.ndarray:
    """pos_vcoeffs_DG(k, level, f) -- src/1d_dg_functions.jl:103-117: coefficients in the position basis
    (cell-wise scaled Legendre functions), one Gauss rule per cell instead of hquadrature."""
    xs, ws = np.polynomial.legendre.leggauss(npts)
    ncell = 1 << max_level
    out = np.empty(k * ncell)
    sc = 2.0 ** (max_level / 2)
    leg, _ = basis_tables(k)
    for c in range(ncell):
        a, b = c / ncell, (c + 1) / ncell
        half, mid = 0.5 * (b - a), 0.5 * (a + b)
        x = mid + half * xs
        fx = np.array([f(float(xi)) for xi in x])
        for m in range(k):
            # basis(level, cell, mode, x) = sqrt(2) P_{m}(2 (2^l x - c) - 1) 2^(l/2), P in monomial coefficients
            xi = 2.0 * (ncell * x - c) - 1.0
            pv = np.zeros_like(xi)
            for co in leg[m, :11][::-1]:
                pv = pv * xi + co
            out[k * c + m] = half * float(np.dot(ws, fx * pv * math.sqrt(2.0) * sc))
    return out


def wave_evolve_1D(k: int, max_level: int, f0, v0, time0: float, time1: float, basis: str = "hier", order: str = "45",
                   **kwargs):
    """wave_evolve_1D -- src/pdes.jl:89-121 (BASELINE config 1): D_op = periodic_DLF_matrix(k, n; basis),
    laplac = D_op * D_op as an explicit sparse product, RHS = [[0 I]; [laplac 0]] (wave_data), integrated with
    ode45 / ode78; the RHS matrix lives on the device as a resident CSR matrix (GSG_RHS_CSR)."""
    import scipy.sparse as sp
    if basis == "pos":
        f0c, v0c = pos_vcoeffs_DG(k, max_level, f0), pos_vcoeffs_DG(k, max_level, v0)
    elif basis == "hier":
        f0c, v0c = vcoeffs_DG(1, k, max_level, f0), vcoeffs_DG(1, k, max_level, v0)
    else:
        raise ValueError(f"ArgumentError(:basis) {basis!r}")       # "nodal"/"point" are broken in the reference too
    D_op = periodic_DLF_matrix(k, max_level, basis=basis)
    laplac = (D_op @ D_op).tocsc()
    N = f0c.size
    RHS = sp.bmat([[None, sp.identity(N, format="csc")], [laplac, None]], format="csc")
    A = CsrMatrix(RHS)
    plan = get_plan(1, k, max_level, "sparse")
    tout, yout = ode_solve(plan, RHS_CSR, np.concatenate([f0c, v0c]), [time0, time1], order=order, A=A, **kwargs)
    return np.array(tout), yout


def energy_func_1D(k: int, level: int, soln, basis: str = "hier"):
    """energy_func_1D -- src/pdes.jl:239-255: |D_op u|^2 + |udot|^2 per saved state (generic SpMV for D_op)."""
    times, states = soln
    D_op = CsrMatrix(periodic_DLF_matrix(k, level, basis=basis))
    N = states[0].size // 2
    return np.array(times), np.array([float(np.sum((D_op @ s[:N]) ** 2) + np.sum(s[N:] ** 2)) for s in states])


def advect_evolve(D: int, k: int, n: int, a, u0coeffs, time0: float, time1: float,
                  scheme: str = "sparse", dt: float | None = None):
    """u' = -sum_d a_d D_d u from time0 to time1 with fixed-step RK4 (BASELINE config 4; the
    operator is the one vlasov_evolve applies, src/pdes.jl:179-180)."""
    plan = get_plan(D, k, n, scheme)
    if dt is None:
        dt = 0.5 * 2.785 / (float(np.sum(np.abs(a))) * 8.081 * (1 << n))
    ns, h = _steps(time0, time1, dt)
    return plan.rk4_advect(a, u0coeffs, h, ns)


def energy_func(D: int, k: int, n: int, soln, scheme: str = "sparse"):
    """energy_func(D, k, n, soln; scheme) -- src/pdes.jl:258-273."""
    plan = get_plan(D, k, n, scheme)
    times, states = soln
    N = plan.size
    return np.array(times), np.array([plan.energy(s[:N], s[N:]) for s in states])


# ---------------------------------------------------------------------------------------------
# on-disk format (src/pdes.jl:145-163, 217-223)
# ---------------------------------------------------------------------------------------------
# The reference dumps its operators and solution snapshots into "vlasov.h5" through JLD2 with the dataset names
#   dimensions, order, levels, "<name>.m", "<name>.n", "<name>.colptr", "<name>.rowval", "<name>.nzval"
#   (name in m2n, n2p, p2n, n2m, "Ds[d]"; colptr / rowval 1-BASED Int64), times, "f_modal.%06d".
# No HDF5 library is available in this image, so the container here is numpy's .npz with EXACTLY those dataset names
# and conventions (1-based Int64 CSC fields, Float64 vectors): a post-processing script only swaps the file opener.
def _csc_fields(name: str, A) -> dict:
    import scipy.sparse as sp
    A = sp.csc_matrix(A)
    A.sort_indices()
    return {f"{name}.m": np.int64(A.shape[0]), f"{name}.n": np.int64(A.shape[1]),
            f"{name}.colptr": A.indptr.astype(np.int64) + 1, f"{name}.rowval": A.indices.astype(np.int64) + 1,
            f"{name}.nzval": _f64(A.data)}


def write_operators(path: str, D: int, k: int, n: int, operators: dict) -> None:
    """jldopen(path, "w"): dimensions / order / levels + every matrix as the five SparseMatrixCSC fields."""
    fields = {"dimensions": np.int64(D), "order": np.int64(k), "levels": np.int64(n)}
    for name, A in operators.items():
        fields.update(_csc_fields(name, A))
    np.savez(path, **fields)


def write_solution(path: str, soln) -> None:
    """jldopen(path, "r+"): adds `times` and `f_modal.%06d` (1-based) to the file written by write_operators."""
    times, states = soln
    fields = dict(np.load(path)) if __import__("os").path.exists(path) else {}
    fields["times"] = _f64(times)
    for i, sol in enumerate(states, start=1):
        fields["f_modal.%06d" % i] = _f64(sol)
    np.savez(path, **fields)


def read_dump(path: str):
    """-> (meta, operators as scipy CSC, times, states) from a file written by the two functions above."""
    import scipy.sparse as sp
    z = np.load(path)
    meta = {key: int(z[key]) for key in ("dimensions", "order", "levels") if key in z}
    ops = {}
    for key in z.files:
        if key.endswith(".colptr"):
            name = key[:-7]
            ops[name] = sp.csc_matrix((z[f"{name}.nzval"], z[f"{name}.rowval"] - 1, z[f"{name}.colptr"] - 1),
                                      shape=(int(z[f"{name}.m"]), int(z[f"{name}.n"])))
    times = z["times"] if "times" in z else None
    states = [z[key] for key in sorted(f for f in z.files if f.startswith("f_modal."))]
    return meta, ops, times, states


# ---------------------------------------------------------------------------------------------
# generic SpMV cross-check
# ---------------------------------------------------------------------------------------------
class CsrMatrix:
    """Resident copy of a reference-style assembled SparseMatrixCSC for the cross-check SpMV."""

    def __init__(self, A, device: int = 0):
        import scipy.sparse as sp
        A = sp.csc_matrix(A)
        A.sort_indices()
        self.shape = A.shape
        colptr = A.indptr.astype(np.int64) + 1
        rowval = A.indices.astype(np.int64) + 1
        nzval = _f64(A.data)
        h = C.c_void_p()
        check(lib.gsg_csr_create(A.shape[0], A.shape[1], _ptr(colptr), _ptr(rowval), _ptr(nzval), device, C.byref(h)))
        self._h = h

    def __matmul__(self, x):
        x = _f64(x)
        y = np.empty(self.shape[0])
        check(lib.gsg_csr_apply(self._h, _ptr(x), _ptr(y)))
        return y

    __mul__ = __matmul__

    def apply_dev(self, x, y, stream=None) -> None:
        s = getattr(stream, "cuda_stream", stream)
        check(lib.gsg_csr_apply_dev(self._h, _devptr(x), _devptr(y), C.c_void_p(int(s) if s else 0)))

    def close(self):
        if getattr(self, "_h", None):
            lib.gsg_csr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spmv_csc(A, x) -> np.ndarray:
    """y = A * x on the GPU for a SparseMatrixCSC-style matrix (`*(RHS, x)`, src/pdes.jl:63)."""
    import scipy.sparse as sp
    A = sp.csc_matrix(A)
    A.sort_indices()
    colptr = A.indptr.astype(np.int64) + 1
    rowval = A.indices.astype(np.int64) + 1
    nzval = _f64(A.data)
    x = _f64(x)
    y = np.empty(A.shape[0])
    check(lib.gsg_spmv_csc(A.shape[0], A.shape[1], _ptr(colptr), _ptr(rowval), _ptr(nzval), _ptr(x), _ptr(y)))
    return y
