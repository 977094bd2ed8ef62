"""Multi-GPU RK4 for u' = -sum_d a_d D_d u.

Two drivers of the same block partition:
  * `MultiGpuRK4` -- the shipped path.  The whole partitioned RK4 runs inside libgsgb200 (gsg_mg_*): peer-mapped
    state slabs, pull / pull-add kernels over NVLink, flag counters for rank synchronisation, the step replayed
    from a CUDA graph.  Python only exchanges the 64-byte CUDA IPC handles once (torch.distributed all_gather,
    any backend) -- nothing per right-hand side.
  * `PartitionedRK4` -- the portable path: the same partition driven from Python with NCCL / gloo point-to-point
    messages (kept for CPU tests of the control flow and as a fallback where peer mapping is unavailable).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import check, lib


class MultiGpuRK4:
    """One rank of the in-library multi-GPU RK4 (include/gsg_b200.h, gsg_mg_*)."""

    def __init__(self, plan, rank: int, world: int):
        self.plan, self.rank, self.world = plan, rank, world
        h = C.c_void_p()
        check(lib.gsg_mg_create(plan._h, rank, world, C.byref(h)))
        self._h = h

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib.gsg_mg_ipc_handle(self._h, buf))
        return buf.raw

    def connect_ipc(self, handles) -> None:
        """handles: the 64-byte handles of all ranks in rank order"""
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        check(lib.gsg_mg_connect_ipc(self._h, C.c_char_p(blob)))

    def connect_torch(self, group=None) -> None:
        """Exchange the IPC handles through torch.distributed (one all_gather of 64 bytes per rank)."""
        if self.world == 1:
            self.connect_ipc([self.ipc_handle()])
            return
        mine = torch.frombuffer(bytearray(self.ipc_handle()), dtype=torch.uint8).clone()
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine = mine.to(dev)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=group)
        self.connect_ipc([bytes(t.cpu().numpy().tobytes()) for t in out])

    @staticmethod
    def connect_local(ranks) -> None:
        """ranks: every MultiGpuRK4 of the partition, living in this process (rank order)."""
        arr = (C.c_void_p * len(ranks))(*[r._h for r in ranks])
        check(lib.gsg_mg_connect_local(arr, len(ranks)))

    def set_state(self, u_host: np.ndarray) -> None:
        u_host = np.ascontiguousarray(u_host, dtype=np.float64)
        assert u_host.shape == (self.plan.size,)
        check(lib.gsg_mg_set_state(self._h, u_host.ctypes.data_as(C.c_void_p)))

    def get_state(self, out_host: np.ndarray) -> np.ndarray:
        assert out_host.dtype == np.float64 and out_host.flags.c_contiguous and out_host.shape == (self.plan.size,)
        check(lib.gsg_mg_get_state(self._h, out_host.ctypes.data_as(C.c_void_p)))
        return out_host

    def owned_fraction(self):
        f, b = C.c_double(), C.c_int64()
        check(lib.gsg_mg_owned_fraction(self._h, C.byref(f), C.byref(b)))
        return f.value, b.value

    def step(self, a, dt: float, nsteps: int) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64)
        check(lib.gsg_mg_rk4_advect(self._h, a.ctypes.data_as(C.c_void_p), float(dt), int(nsteps)))

    @staticmethod
    def step_all(ranks, a, dt: float, nsteps: int) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64)
        arr = (C.c_void_p * len(ranks))(*[r._h for r in ranks])
        check(lib.gsg_mg_rk4_advect_all(arr, len(ranks), a.ctypes.data_as(C.c_void_p), float(dt), int(nsteps)))

    def sync(self) -> None:
        check(lib.gsg_mg_sync(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.gsg_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# nranks = 2^b.  Dimension D-j (j < b) splits the multi-level blocks into {level == 0} (rank bit j = 1)
# and {level >= 1} (bit 0): every block has exactly one owner, the shares are balanced to ~10 % at
# D=6, n=8 (50.5/49.5 at 2 ranks), and ownership depends on levels only, so
#   * a sweep along a non-partition dimension is entirely local (every pole of an owned block is owned);
#   * along a partition dimension d the poles with p >= 1 straddle exactly one rank pair.  The bit-0 rank
#     (which holds all their level >= 1 cells) sweeps them: its partner sends the level_d == 0 blocks of
#     the stage input before the sweep and receives the sweep's contribution to those blocks after it.
# So one right-hand side costs 2 point-to-point messages per partition dimension per rank pair -- no
# all-gather of the state -- and D-b of the D directions need no communication at all.
# Vectors are full length on every rank (the device layout is global); a rank only reads and writes
# the cells it owns plus the exchanged level-0 blocks.

class ThreadComm:
    """In-process stand-in for torch.distributed point-to-point (tests: one thread per virtual rank)."""

    def __init__(self, world):
        import queue
        self.q = {(s, d): queue.Queue() for s in range(world) for d in range(world)}

    def endpoint(self, rank):
        comm = self

        class _EP:
            def send(self, t, dst):
                comm.q[(rank, dst)].put(t.clone())

            def recv(self, t, src):
                t.copy_(comm.q[(src, rank)].get(timeout=120))

            def exchange(self, sends, recvs):
                for t, dst in sends:
                    self.send(t, dst)
                for t, src in recvs:
                    self.recv(t, src)

            def start(self, sends, recvs):
                for t, dst in sends:
                    self.send(t, dst)
                return lambda: [self.recv(t, src) for t, src in recvs]

        return _EP()


class DistComm:
    """torch.distributed point-to-point (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        self.group = group

    def exchange(self, sends, recvs):
        ops = [dist.P2POp(dist.isend, t, dst, self.group) for t, dst in sends]
        ops += [dist.P2POp(dist.irecv, t, src, self.group) for t, src in recvs]
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def start(self, sends, recvs):
        """post the messages now, return a callable that waits for them (NCCL: the transfer runs on the
        communicator's stream beside whatever is launched in between)"""
        ops = [dist.P2POp(dist.isend, t, dst, self.group) for t, dst in sends]
        ops += [dist.P2POp(dist.irecv, t, src, self.group) for t, src in recvs]
        reqs = dist.batch_isend_irecv(ops) if ops else []
        return lambda: [r.wait() for r in reqs]


class PartitionedRK4:
    """RK4 of u' = -sum_d a_d D_d u on a block-partitioned plan (Taylor form)."""

    def __init__(self, plan, a, rank: int, world: int, device, comm):
        self.plan, self.a, self.rank, self.world, self.comm = plan, [float(x) for x in a], rank, world, comm
        self.device = device
        plan.set_partition(rank, world)
        self.D = plan.D
        self.cs = plan.cell_stride
        self.bits = max(world.bit_length() - 1, 0)

        def cells_of(offs, sizes):
            parts = [torch.arange(o // self.cs, (o + s) // self.cs, dtype=torch.int64) for o, s in zip(offs, sizes)]
            return torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64)

        offs, sizes, _ = plan.partition_blocks(0)
        self.owned = cells_of(offs, sizes).to(device)
        self.owned_i32 = self.owned.to(torch.int32)
        self.owned_doubles = int(sizes.sum())
        self.ex = {}                      # partition dimension d -> (cells, partner, my bit)
        for j in range(self.bits):
            d = self.D - j
            offs, sizes, partner = plan.partition_blocks(1, d)
            self.ex[d] = (cells_of(offs, sizes).to(device), partner, (rank >> j) & 1)
        n = plan.dev_size
        self.v = [torch.zeros(n, dtype=torch.float64, device=device) for _ in range(5)]   # u, v1..v4
        self.timing = [] if __import__("os").environ.get("GSG_PART_TIMING") else None
        self.use_graphs = bool(__import__("os").environ.get("GSG_PART_GRAPH")) and torch.device(device).type == "cuda"
        self.graphs, self.graph_seen, self.recv_bufs = {}, {}, {}
        self.exchange_bytes_per_rhs = sum(2 * 8 * self.cs * c.numel() for c, _, _ in self.ex.values())

    def _2d(self, t):
        return t.view(-1, self.cs)

    def set_state(self, full):
        """full: the whole state in device layout (every rank passes the same vector)."""
        self.v[0].zero_()
        self._2d(self.v[0]).index_copy_(0, self.owned, self._2d(full).index_select(0, self.owned))

    def owned_state(self):
        """this rank's part of the state, zero elsewhere (sum over ranks = the full state)."""
        out = torch.zeros_like(self.v[0])
        self._2d(out).index_copy_(0, self.owned, self._2d(self.v[0]).index_select(0, self.owned))
        return out

    def _run(self, tag, w, k, fn, staged=None):
        """Run a group of sweeps: eagerly, or (GSG_PART_GRAPH=1, CUDA only) replayed from a CUDA graph captured
        per (phase, input vector, output vector) -- a rank's sweeps are launch-latency-bound, a replay costs one
        launch.  Receive buffers of a captured phase are persistent (the graph holds their addresses)."""
        if not self.use_graphs:
            fn()
            return
        key = (tag, w.data_ptr(), k.data_ptr())
        ent = self.graphs.get(key)
        if ent is None:
            self.graph_seen[key] = self.graph_seen.get(key, 0) + 1
            if self.graph_seen[key] < 2:          # first visit: eager (kernel attributes get configured)
                fn()
                return
            g = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream()
            with torch.cuda.graph(g):
                self.plan.set_stream(torch.cuda.current_stream())
                fn()
            self.plan.set_stream(cur)
            self.graphs[key] = g
            ent = g
        ent.replay()

    def _mark(self, name):
        if self.timing is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timing.append((name, ev, __import__("time").perf_counter()))

    def timing_report(self):
        """average device / host milliseconds per phase of rhs() (development aid)"""
        torch.cuda.synchronize()
        import statistics
        dev, host = {}, {}
        for (n0, e0, t0), (n1, e1, t1) in zip(self.timing[:-1], self.timing[1:]):
            if n1 == "begin":
                continue
            dev.setdefault(n1, []).append(e0.elapsed_time(e1))
            host.setdefault(n1, []).append(1e3 * (t1 - t0))
        half = {n: dev[n][len(dev[n]) // 2:] for n in dev}          # second half: past warm-up / connection setup
        return {n: (statistics.median(dev[n]), statistics.median(host[n]), statistics.mean(half[n]), max(half[n]))
                for n in dev}

    def rhs(self, w, k):
        """k[owned] = -sum_d a_d D_d w   (w valid on the owned cells)."""
        plan = self.plan
        self._mark("begin")
        # 1. level-0 blocks of the stage input travel to the rank that sweeps the straddling poles
        sends, recvs, staged = [], [], {}
        for d, (cells, partner, bit) in self.ex.items():
            if cells.numel() == 0 or self.a[d - 1] == 0.0:
                continue
            if bit == 1:
                sends.append((self._2d(w).index_select(0, cells), partner))
            else:
                bkey = (d, w.data_ptr())
                if bkey not in self.recv_bufs:
                    self.recv_bufs[bkey] = torch.empty(cells.numel(), self.cs, dtype=torch.float64, device=self.device)
                staged[d] = self.recv_bufs[bkey]
                recvs.append((staged[d], partner))
        wait = self.comm.start(sends, recvs)
        self._mark("pack+post")
        # 2. sweeps: the local directions run while the messages are in flight (the first one initialises k
        #    on the owned cells); then the partition dimensions
        local = [d for d in range(1, self.D + 1) if d not in self.ex]
        any_local = any(self.a[d - 1] != 0.0 for d in local) or not self.ex

        neg_a = [-x for x in self.a]

        def local_sweeps():
            dirs = [d for d in local if self.a[d - 1] != 0.0] or local[:1]
            plan.apply_dirs_dev(neg_a, dirs, w, k, 0.0)       # direction pairs inside `dirs` are swept fused

        def partition_sweeps():
            first = not any_local
            for d, buf in staged.items():
                cells = self.ex[d][0]
                self._2d(w).index_copy_(0, cells, buf)
                self._2d(k).index_fill_(0, cells, 0.0)      # scratch for the partner's contribution
            for d in sorted(self.ex):
                ad = self.a[d - 1]
                if ad == 0.0 and not first:
                    continue
                plan.apply_D_dev(d, w, k, alpha=-ad, beta=0.0 if first else 1.0)
                first = False

        self._run("local", w, k, local_sweeps)
        self._mark("local sweeps")
        wait()
        self._mark("wait recv")
        self._staged = staged
        self._run("part", w, k, partition_sweeps, staged)
        self._mark("unpack+partition sweeps")
        # 3. contributions to the partner's level-0 blocks travel back and are accumulated there
        sends, recvs, back = [], [], {}
        for d, (cells, partner, bit) in self.ex.items():
            if cells.numel() == 0 or self.a[d - 1] == 0.0:
                continue
            if bit == 0:
                sends.append((self._2d(k).index_select(0, cells), partner))
            else:
                back[d] = torch.empty(cells.numel(), self.cs, dtype=torch.float64, device=self.device)
                recvs.append((back[d], partner))
        self.comm.exchange(sends, recvs)
        for d, buf in back.items():
            self._2d(k).index_add_(0, self.ex[d][0], buf)
        self._mark("return exchange")

    def step(self, dt: float, nsteps: int = 1):
        u, v1, v2, v3, v4 = self.v
        for _ in range(nsteps):
            self.rhs(u, v1)
            self.rhs(v1, v2)
            self.rhs(v2, v3)
            self.rhs(v3, v4)
            self.plan.rk4_taylor_cells_dev(self.owned_i32, u, v1, v2, v3, v4, dt, dt * dt / 2.0, dt ** 3 / 6.0,
                                           dt ** 4 / 24.0)
