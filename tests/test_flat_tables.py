"""CPU check of the flat path's host tables (sweep_flat_kernel, kernels.cuh): the kernel's gather loop is replayed
in numpy from the library's own tables (gsg_debug_flat_tables: pole groups + per-cell {group, item, 1-D cell}) and
the oracle's H, and must reproduce the oracle's operator apply.  No compute entry point is called."""
import numpy as np
import pytest

from helpers import random_state, relerr


def _block_rows(H, k, nq):
    """dense k x k blocks of H by block row: list of (qc, block) with ascending qc, as build_matrix stores them"""
    A = H.toarray()
    rows = []
    for q in range(nq):
        r = []
        for qc in range(nq):
            blk = A[q * k:(q + 1) * k, qc * k:(qc + 1) * k]
            if np.any(blk != 0.0):
                r.append((qc, blk))
        rows.append(r)
    return rows


def _q_decode(q):
    ld = 0 if q == 0 else int(q).bit_length()
    cd = 0 if q == 0 else q - (1 << (ld - 1))
    Cd = 1 if ld <= 1 else 1 << (ld - 1)
    return ld, cd, Cd


def _emulate(D, k, n, d, scheme, groups, cells, rows, xdev, KD, KDp):
    """the kernel's loop nest for one direction (0-based d), device layout in and out"""
    A = k ** d
    PI = KD // k
    y = np.zeros_like(xdev)
    for ci in range(cells.shape[0]):
        g, r, q = (int(v) for v in cells[ci])
        base, p, S = groups[g, :17], int(groups[g, 17]), int(groups[g, 18])
        NQ = 1 << p
        lo, hi = r % S, r // S
        for j in range(PI):
            b, a = divmod(j, A)
            po = a + k * A * b
            acc = np.zeros(k)
            for qc, blk in rows[q]:
                if qc >= NQ:
                    break
                ld, cd, Cd = _q_decode(qc)
                off = int(base[ld]) + KDp * (lo + S * (cd + Cd * hi)) + po
                acc += blk @ xdev[off:off + A * k:A]
            y[ci * KDp + po: ci * KDp + po + A * k: A] += acc
    return y


def _to_dev(x, KD, KDp):
    return np.pad(x.reshape(-1, KD), ((0, 0), (0, KDp - KD))).ravel()


@pytest.mark.parametrize("D,k,n,scheme", [(2, 3, 4, "sparse"), (3, 2, 3, "sparse"), (2, 2, 3, "full"), (1, 3, 4, "sparse"),
                                           (4, 2, 2, "sparse"), (3, 3, 2, "full")])
def test_flat_tables_reproduce_operator(gsg, oracle, D, k, n, scheme):
    H = oracle.periodic_DLF_matrix(k, n)
    rows = _block_rows(H, k, 1 << n)
    N = oracle.get_size(D, k, n, scheme=scheme)
    KD = k ** D
    KDp = (KD + 1) & ~1
    x = random_state(N, seed=D * 100 + n)
    xdev = _to_dev(x, KD, KDp)
    for d in range(1, D + 1):
        groups, cells = gsg.flat_tables(D, k, n, d, scheme=scheme)
        assert cells.shape[0] * KD == N
        assert (cells[:, 0] >= 0).all()
        ydev = _emulate(D, k, n, d - 1, scheme, groups, cells, rows, xdev, KD, KDp)
        y = ydev.reshape(-1, KDp)[:, :KD].ravel()
        ref = oracle.apply_D_poles(D, d, k, n, x, H=H, scheme=scheme)
        assert relerr(y, ref) <= 1e-13, (d, relerr(y, ref))
