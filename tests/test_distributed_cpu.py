"""world_size-2 gloo test of the multi-GPU control flow (galerkinsparsegrids.jl_b200/distributed.py)
on CPU: the same ShardedRK4 driver bench.py runs over NCCL, with the oracle standing in for the
local CUDA sweeps (each rank applies its share of the pole tiles)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleOps:
    """CPU stand-in for GpuOps: rank r owns every world-th pole of every direction."""

    def __init__(self, oracle, D, k, n, a, rank, world):
        self.o, self.D, self.k, self.n, self.a = oracle, D, k, n, a
        self.rank, self.world = rank, world
        self.H = oracle.periodic_DLF_matrix(k, n)
        self.tables = [oracle.pole_tables(D, d, k, n) for d in range(1, D + 1)]
        self.full_len = oracle.get_size(D, k, n)

    def zeros(self, m):
        return torch.zeros(m, dtype=torch.float64)

    def apply_partial(self, w, k):
        x = w[: self.full_len].numpy()
        out = k[: self.full_len].numpy()
        for d, ad in enumerate(self.a):
            groups, lens, N = self.tables[d]
            share = ([g[self.rank::self.world] for g in groups], lens, N)      # this rank's poles
            out -= ad * self.o.apply_D_poles(self.D, d + 1, self.k, self.n, x, H=self.H, tables=share)

    def rk_stage(self, u, k, acc, w, cw, ca, first):
        base = u if first else acc
        new_acc = base + ca * k
        w.copy_(u + cw * k)
        acc.copy_(new_acc)

    def rk_final(self, u, k, acc, ca):
        u.copy_(acc + ca * k)


def _worker(rank, world, port, D, k, n, nsteps, dt, out):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gsg_oracle as oracle
    import gsg_b200
    from gsg_b200.distributed import ShardedRK4
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    a = [1.0] * D
    ops = OracleOps(oracle, D, k, n, a, rank, world)
    drv = ShardedRK4(ops, rank, world)
    import math
    v1 = oracle.coeffs_1d(k, n, lambda x: math.sin(2 * math.pi * x))
    u0 = oracle.tensor_construct(D, k, n, [v1] * D)
    drv.set_state(torch.from_numpy(u0.copy()))
    drv.step(dt, nsteps)
    if rank == 0:
        np.save(out, drv.get_state().numpy())
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(300)
def test_sharded_rk4_matches_serial(tmp_path, oracle):
    D, k, n, nsteps, dt = 2, 3, 3, 4, 1e-3
    out = str(tmp_path / "state.npy")
    mp.spawn(_worker, args=(2, _free_port(), D, k, n, nsteps, dt, out), nprocs=2, join=True)
    got = np.load(out)
    import math
    v1 = oracle.coeffs_1d(k, n, lambda x: math.sin(2 * math.pi * x))
    u0 = oracle.tensor_construct(D, k, n, [v1] * D)
    mats = [oracle.D_matrix_poles(D, d, k, n) for d in range(1, D + 1)]
    ref = oracle.rk4(oracle.advect_rhs(mats, [1.0] * D), u0, dt, nsteps)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13


def test_shard_ranges_cover_everything():
    """The tile-range rule of gsg_plan_set_shard: contiguous, disjoint, exhaustive."""
    for ntiles in (0, 1, 7, 148, 12345):
        for world in (1, 2, 3, 8):
            ranges = [(ntiles * r // world, ntiles * (r + 1) // world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == ntiles
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
