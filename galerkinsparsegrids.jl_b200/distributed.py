"""Multi-GPU RK4 for u' = -sum_d a_d D_d u: one process per GPU, torch.distributed for the
plumbing (baseline collective).

Round-1 scheme ("work-shared sweeps + reduce-scatter / all-gather"):
  * every rank holds the full stage input w (device layout, padded to a multiple of world_size);
  * a sweep along d couples multi-levels that differ in level_d, so the pole tiles of every
    direction are split evenly over the ranks (gsg_plan_set_shard); each rank accumulates its
    partial k_r = -sum_d a_d D_d w restricted to its tiles;
  * k = sum_r k_r is formed with ONE reduce-scatter per right-hand side, so each rank owns a
    contiguous slice of k; the RK stage update runs on that slice only;
  * the next stage input is re-assembled with ONE all-gather.
The exchange volume per RHS is ~2 N (W-1)/W doubles per rank, which bounds the speed-up well
below linear; the slab partition of SURVEY.md 8e that keeps most directions local is the
next step.

The driver is backend-agnostic: `ops` supplies the local operator and the stage kernels, so the
same control flow is exercised on CPU (gloo, oracle operator) by tests/test_distributed_cpu.py
and on GPUs (nccl, libgsgb200 kernels) by bench.py.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GpuOps:
    """Local pieces on one GPU through the C ABI (device-layout vectors)."""

    def __init__(self, plan, a, rank: int, world: int, device):
        self.plan, self.a = plan, [float(x) for x in a]
        self.device = device
        plan.set_shard(rank, world)
        self.full_len = plan.dev_size

    def zeros(self, n):
        return torch.zeros(n, dtype=torch.float64, device=self.device)

    def apply_partial(self, w, k):
        """k += this rank's share of -sum_d a_d D_d w (k must be zeroed by the caller)."""
        for d, ad in enumerate(self.a, start=1):
            if ad != 0.0:
                self.plan.apply_D_dev(d, w, k, alpha=-ad, beta=1.0)

    def rk_stage(self, u, k, acc, w, cw, ca, first):
        self.plan.rk_stage_dev(u.numel(), u, k, acc, w, cw, ca, first)

    def rk_final(self, u, k, acc, ca):
        self.plan.rk_final_dev(u.numel(), u, k, acc, ca)


class ShardedRK4:
    def __init__(self, ops, rank: int, world: int, group=None):
        self.ops, self.rank, self.world, self.group = ops, rank, world, group
        n = ops.full_len
        self.shard = (n + world - 1) // world
        self.padded = self.shard * world
        self.w = ops.zeros(self.padded)        # full stage input
        self.k = ops.zeros(self.padded)        # partial RHS of this rank
        self.k_sh = ops.zeros(self.shard)
        self.u_sh = ops.zeros(self.shard)
        self.acc_sh = ops.zeros(self.shard)
        self.w_sh = ops.zeros(self.shard)
        self._use_rs = world > 1 and dist.get_backend(group) == "nccl"

    # -- state in / out (full vector in the ops' layout) ---------------------------------------
    def set_state(self, full):
        self.w.zero_()
        self.w[: full.numel()].copy_(full)
        lo = self.rank * self.shard
        self.u_sh.copy_(self.w[lo: lo + self.shard])

    def get_state(self):
        return self.w[: self.ops.full_len]

    # -- collectives -----------------------------------------------------------------------------
    def _reduce_scatter(self):
        if self.world == 1:
            self.k_sh.copy_(self.k[: self.shard])
        elif self._use_rs:
            dist.reduce_scatter_tensor(self.k_sh, self.k, group=self.group)
        else:                                   # gloo has no reduce_scatter: all-reduce + slice
            dist.all_reduce(self.k, group=self.group)
            lo = self.rank * self.shard
            self.k_sh.copy_(self.k[lo: lo + self.shard])

    def _all_gather(self, shard):
        if self.world == 1:
            self.w[: self.shard].copy_(shard)
        else:
            dist.all_gather_into_tensor(self.w, shard, group=self.group)

    def _rhs(self):
        self.k.zero_()
        self.ops.apply_partial(self.w, self.k)
        self._reduce_scatter()

    def step(self, dt: float, nsteps: int = 1):
        """Classical RK4 (same stage order as gsg_rk4_advect / the oracle)."""
        o = self.ops
        for _ in range(nsteps):
            self._rhs()
            o.rk_stage(self.u_sh, self.k_sh, self.acc_sh, self.w_sh, 0.5 * dt, dt / 6.0, True)
            self._all_gather(self.w_sh)
            self._rhs()
            o.rk_stage(self.u_sh, self.k_sh, self.acc_sh, self.w_sh, 0.5 * dt, dt / 3.0, False)
            self._all_gather(self.w_sh)
            self._rhs()
            o.rk_stage(self.u_sh, self.k_sh, self.acc_sh, self.w_sh, dt, dt / 3.0, False)
            self._all_gather(self.w_sh)
            self._rhs()
            o.rk_final(self.u_sh, self.k_sh, self.acc_sh, dt / 6.0)
            self._all_gather(self.u_sh)
