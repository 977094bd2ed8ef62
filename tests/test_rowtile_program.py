"""CPU check of the row-tile kernel's tile program (csrc/rowtile.inl): the program of a pole class is replayed in
numpy for one pole -- load the tile's x cells, walk the records of every row group (one x cell + one k x k block per
row of the record's mask), store complete rows, add partial ones -- and must reproduce H_p x; every complete row is
produced exactly once and never also receives a partial sum.  Also the bank-permuted lane order of the poles.
No compute entry point is called."""
import struct

import numpy as np
import pytest


def _replay(prog, k, p, x, KDp):
    NQ = 1 << p
    y = np.zeros(k * NQ)
    complete = np.zeros(NQ, dtype=int)
    partial = np.zeros(NQ, dtype=int)
    xc = x.reshape(NQ, k)
    blob = prog["blob"].tobytes()
    stats = {"records": 0, "blocks": 0}
    for T in prog["tiles"]:
        nx, rec_ofs, rec_bytes, grp0 = (int(v) for v in T[:4])
        rg_end = [int(v) for v in T[4:8]]
        xs = [xc[int(q)] for q in T[8:8 + nx]]
        assert len(set(int(q) for q in T[8:8 + nx])) == nx
        assert rg_end[-1] >= 1 and all(a <= b for a, b in zip(rg_end, rg_end[1:]))
        assert rec_ofs % 16 == 0 and rec_bytes % 16 == 0
        used = 0
        for gi in range(grp0, grp0 + rg_end[-1]):
            G = [int(v) for v in prog["groups"][gi]]
            q, rofs, nrec, pmask = G[:4], G[4], G[5], G[6]
            assert rofs % 8 == 0 and nrec >= 1
            acc = np.zeros((4, k))
            seen = set()
            ptr = rec_ofs + rofs
            for _ in range(nrec):
                xofs, mask = struct.unpack_from("<ii", blob, ptr)
                ptr += 8
                assert xofs % (KDp * 8) == 0 and 0 <= xofs // (KDp * 8) < nx
                assert mask >> 8 == (xofs // (KDp * 8) * 4) // nx     # the barrier the x cell arrives on
                mask &= 0xFF
                assert 0 < mask < 16
                assert xofs not in seen          # one record per distinct x cell of a group
                seen.add(xofs)
                stats["records"] += 1
                for r in range(4):
                    if mask >> r & 1:
                        assert q[r] >= 0
                        h = np.frombuffer(blob, dtype=np.float64, count=k * k, offset=ptr).reshape(k, k)
                        ptr += k * k * 8
                        acc[r] += h @ xs[xofs // (KDp * 8)]
                        stats["blocks"] += 1
            used += ptr - (rec_ofs + rofs)
            for r in range(4):
                if q[r] < 0:
                    assert not (pmask >> r & 1)
                    continue
                y[q[r] * k:(q[r] + 1) * k] += acc[r]
                if pmask >> r & 1:
                    partial[q[r]] += 1
                else:
                    complete[q[r]] += 1
        assert used <= rec_bytes < used + 16
    return y, complete, partial, stats


@pytest.mark.parametrize("D,k,n,p,budget,nrg", [
    (6, 3, 8, 4, 226 * 1024, 4), (6, 3, 8, 5, 226 * 1024, 4), (6, 3, 8, 6, 226 * 1024, 4), (6, 3, 8, 7, 226 * 1024, 4),
    (6, 3, 8, 8, 226 * 1024, 4), (6, 3, 8, 8, 140 * 1024, 2), (6, 3, 8, 5, 140 * 1024, 4), (6, 3, 6, 6, 112 * 1024, 2),
    (4, 4, 6, 4, 112 * 1024, 4), (4, 4, 6, 6, 112 * 1024, 4), (5, 3, 5, 5, 60 * 1024, 2), (4, 5, 5, 5, 112 * 1024, 2),
    (5, 2, 7, 7, 20 * 1024, 2),
])
def test_rowtile_program_reproduces_subblock(gsg, D, k, n, p, budget, nrg):
    prog = gsg.rowtile_program(D, k, n, p, budget, nrg)
    H = gsg.periodic_DLF_matrix(k, n).toarray()
    Np = k << p
    KDp = (k ** D + 1) & ~1
    x = np.random.default_rng(p * 10 + k).standard_normal(Np)
    y, complete, partial, stats = _replay(prog, k, p, x, KDp)
    ref = H[:Np, :Np] @ x
    assert np.linalg.norm(y - ref) <= 1e-13 * np.linalg.norm(ref)
    # a row is either produced completely by exactly one tile, or only ever receives partial sums
    assert np.all((complete == 1) & (partial == 0) | (complete == 0) & (partial >= 1))
    for T in prog["tiles"]:
        assert 64 + (2 * nrg + int(T[0])) * KDp * 8 + int(T[2]) <= budget
    # every stored block of H_p is used exactly once
    Hp = H[:Np, :Np]
    nblk = sum(1 for q in range(1 << p) for r in range(1 << p) if np.any(Hp[q * k:(q + 1) * k, r * k:(r + 1) * k] != 0))
    assert stats["blocks"] == nblk
    print(f"class p={p}: {len(prog['tiles'])} tiles, {int(prog['tiles'][:, 0].sum())} cell loads for {1 << p} cells, "
          f"{stats['blocks']} blocks in {stats['records']} records ({stats['blocks'] / stats['records']:.2f} rows per x load), "
          f"{int((partial > 0).sum())} partial rows")


@pytest.mark.parametrize("k,D,C,PW", [(3, 6, 4, 2), (3, 6, 2, 4), (3, 5, 1, 3), (5, 4, 2, 2), (2, 7, 1, 2)])
def test_rowtile_pole_order_is_a_conflict_free_permutation(gsg, k, D, C, PW):
    PI = k ** (D - 1)
    nslots = 32 * C * PW
    for d in range(D):
        A = k ** d
        tab = gsg.rowtile_pole_order(k, A, PI, nslots)
        valid = tab[tab >= 0]
        want = sorted(a + k * A * b for b in range(PI // A) for a in range(A))
        assert sorted(valid.tolist()) == want                      # every pole exactly once
        # 16 different 8-byte banks per half-warp, except for what the pigeonhole principle forces (a bank that holds
        # more poles than there are half-warps)
        nhw = nslots // 16
        forced = sum(max(0, int(c) - nhw) for c in np.bincount(valid % 16, minlength=16))
        extra = 0
        for h in range(nhw):
            hw = tab[h * 16:(h + 1) * 16]
            v = hw[hw >= 0]
            extra += len(v) - len(set((v % 16).tolist()))
            # padding lanes repeat an address of their own half-warp (a broadcast)
            for raw in hw[hw < 0]:
                assert len(v) == 0 or ~raw in v.tolist()
        assert extra == forced, (d, extra, forced)
        if k == 3:
            assert forced <= 8      # A = 27 at D = 6: the banks (a + b) mod 16 hold 13 .. 18 poles for 16 half-warps


def _plan_shape(k, D):
    """poles per lane / pole warps / row groups as build_rowtile_program chooses them (csrc/gsg_b200.cu)"""
    PI = k ** (D - 1)
    C = 4 if (PI >= 192 and k <= 3) else 2 if (PI > 96 and k <= 4) else 1
    PW = (PI + 32 * C - 1) // (32 * C)
    return PI, C, PW, max(1, min(4, 8 // PW))


@pytest.mark.parametrize("k,D", [(2, 7), (2, 8), (2, 9), (2, 10), (3, 5), (3, 6), (4, 4), (4, 5), (5, 4)])
def test_rowtile_program_every_plan_shape(gsg, k, D):
    """every (k, D) whose plans take the row-tile kernel (k <= 5, 64 <= k^(D-1) <= 512), default shared-memory budget,
    every long class up to n = 7 (n <= 10 was swept once by hand): the program reproduces H_p x and fits"""
    PI, C, PW, RG = _plan_shape(k, D)
    assert 64 <= PI <= 512 and PW <= 8
    pmin = 0
    while (k << pmin) <= 32 and pmin <= 3:            # first class the streaming kernel does not serve
        pmin += 1
    n = 7
    H = gsg.periodic_DLF_matrix(k, n).toarray()
    KDp = (k ** D + 1) & ~1
    for p in range(pmin, n + 1):
        prog = gsg.rowtile_program(D, k, n, p, 226 * 1024, RG)
        Np = k << p
        x = np.random.default_rng(p).standard_normal(Np)
        y, complete, partial, _ = _replay(prog, k, p, x, KDp)
        ref = H[:Np, :Np] @ x
        assert np.linalg.norm(y - ref) <= 1e-13 * np.linalg.norm(ref), (k, D, p)
        assert np.all((complete == 1) & (partial == 0) | (complete == 0) & (partial >= 1))
        assert max(64 + (2 * RG + int(T[0])) * KDp * 8 + int(T[2]) for T in prog["tiles"]) <= 226 * 1024
