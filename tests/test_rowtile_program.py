"""CPU check of the row-tile kernel's tile program (csrc/rowtile.inl): the program of a pole class is replayed in
numpy for one pole -- load the tile's x cells, run the records of every row, store complete rows, add partial ones --
and must reproduce H_p x; every complete row is produced exactly once and never also receives a partial sum.
No compute entry point is called."""
import numpy as np
import pytest


def _replay(prog, k, p, x):
    NQ = 1 << p
    y = np.zeros(k * NQ)
    complete = np.zeros(NQ, dtype=int)
    partial = np.zeros(NQ, dtype=int)
    xc = x.reshape(NQ, k)
    for T in prog["tiles"]:
        nx, rec0, nrec, row0 = (int(v) for v in T[:4])
        rg_end = [int(v) for v in T[4:8]]
        xs = [xc[int(q)] for q in T[8:8 + nx]]
        assert len(set(int(q) for q in T[8:8 + nx])) == nx
        assert rg_end[-1] >= 1 and all(a <= b for a, b in zip(rg_end, rg_end[1:]))
        used = 0
        for ri in range(row0, row0 + rg_end[-1]):
            q, rb, re, part = (int(v) for v in prog["rows"][ri])
            assert 0 <= rb < re <= nrec
            acc = np.zeros(k)
            for r in range(rec0 + rb, rec0 + re):
                s = int(prog["rec_slot"][r])
                assert 0 <= s < nx
                acc += prog["rec_h"][r] @ xs[s]
            used += re - rb
            y[q * k:(q + 1) * k] += acc
            if part:
                partial[q] += 1
            else:
                complete[q] += 1
        assert used == nrec
    return y, complete, partial


@pytest.mark.parametrize("D,k,n,p,budget,nrg", [
    (6, 3, 8, 4, 112 * 1024, 2), (6, 3, 8, 5, 112 * 1024, 2), (6, 3, 8, 6, 112 * 1024, 2), (6, 3, 8, 7, 112 * 1024, 2),
    (6, 3, 8, 8, 112 * 1024, 2), (6, 3, 8, 8, 225 * 1024, 4), (6, 3, 8, 5, 225 * 1024, 4), (6, 3, 6, 6, 80 * 1024, 2),
    (4, 4, 6, 4, 112 * 1024, 4), (4, 4, 6, 6, 112 * 1024, 4), (5, 3, 5, 5, 60 * 1024, 2), (4, 5, 5, 5, 112 * 1024, 2),
    (5, 2, 7, 7, 20 * 1024, 2),
])
def test_rowtile_program_reproduces_subblock(gsg, D, k, n, p, budget, nrg):
    prog = gsg.rowtile_program(D, k, n, p, budget, nrg)
    H = gsg.periodic_DLF_matrix(k, n).toarray()
    Np = k << p
    x = np.random.default_rng(p * 10 + k).standard_normal(Np)
    y, complete, partial = _replay(prog, k, p, x)
    ref = H[:Np, :Np] @ x
    assert np.linalg.norm(y - ref) <= 1e-13 * np.linalg.norm(ref)
    # a row is either produced completely by exactly one tile, or only ever receives partial sums
    assert np.all((complete == 1) & (partial == 0) | (complete == 0) & (partial >= 1))
    KDp = (k ** D + 1) & ~1
    rec_bytes = (k * k * 8 + 8 + 15) & ~15
    for T in prog["tiles"]:
        assert 64 + int(T[0]) * KDp * 8 + int(T[2]) * rec_bytes <= budget
    print(f"class p={p}: {len(prog['tiles'])} tiles, {int(prog['tiles'][:, 0].sum())} cell loads for {1 << p} cells, "
          f"{len(prog['rec_slot'])} records, {int((partial > 0).sum())} partial rows")
