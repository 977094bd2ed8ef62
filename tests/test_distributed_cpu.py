"""world_size-2/4 gloo tests of the multi-GPU control flow (galerkinsparsegrids.jl_b200/distributed.py) on CPU:
the block-partitioned RK4 driver with the oracle standing in for the local CUDA sweeps."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


# ------------------------------------------------------------------------------------------------
# block-partitioned scheme (PartitionedRK4): same driver as bench.py --gpus N, gloo, oracle sweeps
# ------------------------------------------------------------------------------------------------
class OraclePartPlan:
    """CPU stand-in for the partitioned gsg_b200.Plan: reference layout (cell stride k^D), ownership rule
    restated independently of the C++ (level == 0 of dimension D-j <-> rank bit j)."""

    def __init__(self, oracle, D, k, n):
        self.o, self.D, self.k, self.n = oracle, D, k, n
        self.H = oracle.periodic_DLF_matrix(k, n)
        self.blocks, self.N = oracle.block_table(D, k, n)
        self.cell_stride = k ** D
        self.dev_size = self.N
        self.rank, self.bits = 0, 0

    def set_partition(self, rank, world):
        self.rank, self.bits = rank, world.bit_length() - 1
        self.tables = []
        for d in range(1, self.D + 1):
            rows, lens, N = self.o.pole_tables(self.D, d, self.k, self.n)
            keys = list(dict.fromkeys(lv[:d - 1] + lv[d:] for lv, _, _ in self.blocks))     # same order as pole_tables
            assert len(keys) == len(rows)
            mine_rows, mine_lens = [], []
            for other, idx, Np in zip(keys, rows, lens):
                p = (Np // self.k).bit_length() - 1
                mine = True
                for j in range(self.bits):
                    e, mybit = self.D - 1 - j, (rank >> j) & 1
                    if e == d - 1:
                        mine &= (mybit == 1) if p == 0 else (mybit == 0)
                    else:
                        lvl = other[e] if e < d - 1 else other[e - 1]
                        mine &= (lvl == 0) == (mybit == 1)
                if mine:
                    mine_rows.append(idx)
                    mine_lens.append(Np)
            self.tables.append((mine_rows, mine_lens, N))

    def _owner(self, lv, skip=-1):
        o = 0
        for j in range(self.bits):
            e = self.D - 1 - j
            if e != skip and lv[e] == 0:
                o |= 1 << j
        return o

    def partition_blocks(self, kind, d=0):
        offs, sizes, partner = [], [], -1
        if kind == 1:
            j = self.D - d
            if j >= self.bits:
                return np.zeros(0, np.int64), np.zeros(0, np.int64), -1
            partner = self.rank ^ (1 << j)
        for lv, off, ks in self.blocks:
            size = int(np.prod(ks, dtype=np.int64)) * self.cell_stride
            if kind == 0:
                take = self._owner(lv) == self.rank
            else:
                mask = ((1 << self.bits) - 1) & ~(1 << (self.D - d))
                take = lv[d - 1] == 0 and sum(lv) < self.n and ((self._owner(lv, d - 1) ^ self.rank) & mask) == 0
            if take:
                offs.append(off)
                sizes.append(size)
        return np.array(offs, np.int64), np.array(sizes, np.int64), partner

    def apply_D_dev(self, d, w, k, alpha=1.0, beta=0.0):
        rows, lens, N = self.tables[d - 1]
        if not rows:
            return
        y = self.o.apply_D_poles(self.D, d, self.k, self.n, w.numpy(), H=self.H, tables=self.tables[d - 1])
        idx = np.concatenate([r.reshape(-1) for r in rows])
        kn = k.numpy()
        kn[idx] = alpha * y[idx] + (beta * kn[idx] if beta != 0.0 else 0.0)

    def apply_dirs_dev(self, c, dirs, w, k, beta=0.0):
        first = True
        for d in dirs:
            self.apply_D_dev(d, w, k, alpha=c[d - 1], beta=beta if first else 1.0)
            first = False

    def rk4_taylor_cells_dev(self, cells, u, v1, v2, v3, v4, c1, c2, c3, c4):
        cs = self.cell_stride
        idx = (cells.numpy().astype(np.int64)[:, None] * cs + np.arange(cs)[None, :]).reshape(-1)
        un = u.numpy()
        un[idx] += c1 * v1.numpy()[idx] + c2 * v2.numpy()[idx] + c3 * v3.numpy()[idx] + c4 * v4.numpy()[idx]


def _part_worker(rank, world, port, D, k, n, nsteps, dt, outdir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gsg_oracle as oracle
    import gsg_b200  # noqa: F401
    from gsg_b200.distributed import DistComm, PartitionedRK4
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    a = [1.0, -0.5, 0.25][:D]
    plan = OraclePartPlan(oracle, D, k, n)
    drv = PartitionedRK4(plan, a, rank, world, torch.device("cpu"), DistComm())
    import math
    v1 = oracle.coeffs_1d(k, n, lambda x: math.sin(2 * math.pi * x))
    u0 = oracle.tensor_construct(D, k, n, [v1] * D)
    drv.set_state(torch.from_numpy(u0.copy()))
    drv.step(dt, nsteps)
    np.save(os.path.join(outdir, f"part{rank}.npy"), drv.owned_state().numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_rk4_matches_serial(tmp_path, oracle, world):
    D, k, n, nsteps, dt = 3, 2, 3, 3, 1e-3
    mp.spawn(_part_worker, args=(world, _free_port(), D, k, n, nsteps, dt, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"part{r}.npy") for r in range(world)]
    got = sum(parts)
    # every DOF is owned by exactly one rank
    assert all(np.count_nonzero(pr) > 0 for pr in parts)
    import math
    v1 = oracle.coeffs_1d(k, n, lambda x: math.sin(2 * math.pi * x))
    u0 = oracle.tensor_construct(D, k, n, [v1] * D)
    mats = [oracle.D_matrix_poles(D, d, k, n) for d in range(1, D + 1)]
    ref = oracle.rk4(oracle.advect_rhs(mats, [1.0, -0.5, 0.25][:D]), u0, dt, nsteps)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
