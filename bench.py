#!/usr/bin/env python
"""bench.py -- RK4 DOF-updates/s of the sparse-grid DG advection u' = -sum_d a_d D_d u
(BASELINE.json metric; D=6 sparse, k=3, n=8) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU)
  python bench.py --impl reference ...                      the reference's CPU path (oracle port)

One "step" = one classical RK4 step of the whole state = N_dof DOF-updates: 4 right-hand sides, each D
directional operator applies.  The right-hand side is linear and time-independent, so the library advances
the step in the Taylor form of RK4 (u += dt L u + dt^2/2 L^2 u + dt^3/6 L^3 u + dt^4/24 L^4 u: the same four
operator applications, one combine pass instead of four stage updates); the staged
form is timed beside it and reported as `staged_ms_per_step`.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (D, k, n)  -- BASELINE.json configs
    "d6k3n8": (6, 3, 8),     # config 4, the one the metric is quoted on (fits one B200: 276 MB state)
    "d4k3n7": (4, 3, 7),     # config 3 operator (latency regime)
    "d2k3n8": (2, 3, 8),     # config 2 operator (latency regime)
    "recon": (4, 4, 8),      # config 5: reconstruct_DG of a D=4 sparse k=4 n=8 interpolant (points/s)
    "wave2d": (2, 3, 8),     # config 2 as the reference runs it: wave equation [u; v]' = [v; lap u], classical RK4
}
FP64_PEAK_TFLOPS = 37.0      # nominal B200 fp64 (MEASURED_PEAKS.json holds only the copy bandwidth and the bf16 GEMM)
DT = 1.0e-4                  # stable for RK4 at D=6, n=8 (dt_max ~ 2.2e-4)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region, in-process through NVML (a thread
    polling every 100 ms).  An external `nvidia-smi -lms 50` loop was measured to stall the GPU for ~11 ms
    per query (4.7 vs 2.7 ms per step at 4 GPUs), so nvidia-smi is only the fallback, at -lms 500."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.thread, self.stop_flag = gpu_index, None, None, False
        self.sm, self.mx, self.reasons, self.how = [], [], set(), None

    def _nvml_loop(self, pynvml, handle):
        bits = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown if hasattr(pynvml, "nvmlClocksEventReasonHwSlowdown")
                else pynvml.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown",
                                               getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0)),
                "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown",
                                               getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0)),
                "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap",
                                        getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0))}
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons",
                              getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
        while not self.stop_flag:
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    r = int(get_reasons(handle))
                    for name, bit in bits.items():
                        if bit and (r & bit):
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        if os.environ.get("GSG_NO_SAMPLER"):
            return
        try:
            import threading

            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: match by PCI bus id when torch is available
            handle = None
            try:
                import torch
                bus = torch.cuda.get_device_properties(self.gpu).pci_bus_id
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        handle = h
                        break
            except Exception:
                handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.how = "nvml thread, 100 ms"
            self.thread = threading.Thread(target=self._nvml_loop, args=(pynvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.how = "nvidia-smi -lms 500"
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "500", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "samples": len(self.sm),
                    "reasons": sorted(self.reasons), "how": self.how}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "how": self.how}


def synthetic_state(g, D, k, n):
    """sin-product initial condition prod_d sin(2 pi x_d) via tensor_construct."""
    v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
    return g.tensor_construct(D, k, n, [v1] * D)


def _ref_common(D, k, n):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np

    import cbaseline
    import gsg_oracle as o
    H = o.periodic_DLF_matrix(k, n)
    N = cbaseline.get_size(D, k, n)
    x = np.random.default_rng(0).standard_normal(N)
    return cbaseline, H, N, x


def _rk4_step_estimate(D, t_dir, t_axpy):
    """seconds of one RK4 step of the reference's CPU path from MEASURED complete products: 4 right-hand sides x
    (one `Ds[d]*f` per direction, src/pdes.jl:179-180) + the integrator's vector arithmetic (4 D + 10 axpy-like passes)"""
    dirs = [sum(v) / len(v) for v in t_dir.values() if v]
    per_rhs = sum(dirs) * (D / len(dirs))              # directions not sampled count as the mean of those that were
    return 4.0 * per_rhs + (4 * D + 10) * t_axpy


def cpu_baseline(D, k, n, ndirs=2):
    """cpu_baseline leg of our own arm (rank 0, N = 1): the reference's stock CPU path -- SERIAL CSC column-scatter
    SpMV on the assembled D_d (Julia SparseArrays) -- timed on `ndirs` COMPLETE directions (every column of D_d)."""
    cb, H, N, x = _ref_common(D, k, n)
    t_dir, nnz = {}, 0
    for d in ([1, D] if ndirs == 2 else list(range(1, ndirs + 1))):
        ts, nnz, _ = cb.time_direction(D, d, k, n, H, x, threads=1)
        t_dir.setdefault(d, []).append(ts)
    t_axpy = cb.time_axpy(N)
    t_step = _rk4_step_estimate(D, t_dir, t_axpy)
    return {
        "value": N / t_step, "unit": "DOF-updates/s", "cores": 1, "kind": "port",
        "sample": (f"serial CSC column-scatter SpMV (Julia SparseArrays `A*x`) on the COMPLETE assembled D_d of directions "
                   f"{sorted(t_dir)} ({nnz} nnz each, assembled slab by slab, assembly untimed); RK4 step = 4 RHS x {D} products "
                   f"(other directions counted at the measured mean) + {4 * D + 10} vector passes"),
        "seconds_per_product": {str(d): v[0] for d, v in t_dir.items()}, "t_axpy_s": t_axpy, "t_step_est_s": t_step,
        "nnz_per_s": nnz / (sum(v[0] for v in t_dir.values()) / len(t_dir)),
    }


def run_reference(args, D, k, n, rank, world):
    """The reference's own CPU implementation of the path (oracle port; Julia is not installed): one timed "step" is
    a bounded sample of the RK4 workload = ONE COMPLETE product D_d x (every column of the assembled matrix),
    cycling d = 1..D over the steps, measured for both CPU variants of the reference:
      serial CSC column scatter (stock Julia SparseArrays)  and  OpenMP CSR on all host threads (MKLSparse hook)."""
    if rank != 0:
        return
    t_wall0 = time.perf_counter()
    cb, H, N, x = _ref_common(D, k, n)
    threads = os.cpu_count() or 1
    cb.lib().gsgo_set_threads(threads)          # torchrun exports OMP_NUM_THREADS=1: set the count explicitly
    W, K = max(args.warmup, 1), max(args.steps, 1)
    t_ser, t_omp, step_s, nnz = {}, {}, [], 0
    for i in range(W + K):
        d = i % D + 1
        ts, nnz, _ = cb.time_direction(D, d, k, n, H, x, threads=1)
        tp = cb.time_direction(D, d, k, n, H, x, threads=threads)[0] if threads > 1 else ts
        if i >= W:
            t_ser.setdefault(d, []).append(ts)
            t_omp.setdefault(d, []).append(tp)
            step_s.append(ts + tp)
    t_axpy = cb.time_axpy(N)
    variants = [{"cores": 1, "kind": "serial CSC column scatter (stock Julia SparseArrays)",
                 "t_step_est_s": _rk4_step_estimate(D, t_ser, t_axpy)}]
    if threads > 1:
        variants.append({"cores": threads, "kind": f"OpenMP row-parallel CSR, {threads} threads (MKLSparse hook)",
                         "t_step_est_s": _rk4_step_estimate(D, t_omp, t_axpy)})
    for v in variants:
        v["value"] = N / v["t_step_est_s"]
    best = max(variants, key=lambda v: v["value"])
    sample = (f"each step = one COMPLETE product D_d x on the assembled matrix ({nnz} nnz, every column; d cycles 1..{D}), "
              f"timed for both variants ({K} timed steps cover directions {sorted(t_ser)}); value = N / (4 RHS x {D} measured "
              f"products + {4 * D + 10} vector passes) of the faster variant; assembly is untimed (the reference assembles once)")
    line = {
        "impl": "reference", "metric": "RK4 DOF-updates/sec (D=%d sparse, k=%d, n=%d)" % (D, k, n),
        "value": best["value"], "unit": "DOF-updates/s", "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * sum(step_s) / len(step_s), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(D, k, n, N),
                   "note": "reference = serial Julia SparseMatrixCSC*Vector (MKLSparse optional); Julia absent, oracle port timed; "
                           "ms_per_step is the measured SpMV time of one bounded-sample step (one complete product, both variants)"},
        "cpu_baseline": {"value": best["value"], "unit": "DOF-updates/s", "cores": best["cores"], "kind": "port",
                         "sample": sample, "rk4_step_est_s": best["t_step_est_s"]},
        "cpu_variants": variants,
        "e2e": {"value": best["value"], "unit": "DOF-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_wall0,
    }
    print(json.dumps(line), flush=True)


def workload_name(D, k, n, N):
    return (f"{D}-D sparse-grid advection u'=-sum_d D_d u, k={k} n={n} sparse, classical RK4, N={N} DOFs "
            f"(BASELINE config {4 if (D, k, n) == (6, 3, 8) else '-'})")


def run_wave(args, rank, local_rank, world):
    """--config wave2d: the 2-D wave evolution of BASELINE config 2 (src/pdes.jl:54-68 with order = "4"): one step =
    one classical RK4 step of [u; v]' = [v; laplacian u] through gsg_rk4_wave_dev (the Laplacian takes the pre-squared
    blocks S_p = (H[:N',:N'])^2, one sweep per direction, src/multidim_derivative.jl:71-79).  Replicas at N > 1."""
    import numpy as np
    import torch

    import gsg_b200 as g
    D, k, n = CONFIGS["wave2d"]
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    W, K = max(args.warmup, 3), args.steps
    plan = g.Plan(D, k, n, device=local_rank)
    N = plan.size
    v1d = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
    u0 = g.tensor_construct(D, k, n, [v1d] * D)
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    plan.set_stream(stream)
    u = plan.tensor_construct_dev([v1d] * D, device=device)
    v = torch.zeros_like(u)
    plan.rk4_wave_dev(u, v, DT, W)
    torch.cuda.synchronize(device)
    l0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    plan.rk4_wave_dev(u, v, DT, K)
    e1.record(stream)
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    launches = g.launch_count() - l0
    hu, hv = torch.from_numpy(u0.copy()).pin_memory(), torch.zeros(N, dtype=torch.float64).pin_memory()
    plan.rk4_wave(hu.numpy(), hv.numpy(), DT, 1)                 # warm the host-buffer path
    t0 = time.perf_counter()
    rc = g.lib.gsg_rk4_wave(plan._h, g._ptr(hu.numpy()), g._ptr(hv.numpy()), DT, K)
    t1 = time.perf_counter()
    if rc != 0:
        raise SystemExit("gsg_rk4_wave failed")
    energy = plan.energy(hu.numpy(), hv.numpy())
    if rank == 0:
        print(json.dumps({
            "metric": "RK4 DOF-updates/sec, 2-D wave (D=%d sparse, k=%d, n=%d)" % (D, k, n), "value": 2 * N * K / (ms * 1e-3),
            "unit": "DOF-updates/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"2-D wave equation on the sparse grid, k={k} n={n}, state [u; v] = 2 x {N} DOFs, classical RK4 "
                                   "(BASELINE config 2 with order 4), u0 = sin(2 pi x) sin(2 pi y), v0 = 0", "dt": DT,
                       "sweep_path": "flat kernel, pre-squared Laplacian blocks" if plan.flat_active else "tiled class kernels"},
            "roofline": None, "cpu_baseline": None,
            "e2e": {"value": 2 * N * K / (t1 - t0), "unit": "DOF-updates/s", "h2d_bytes_per_step": 16.0 * N / K,
                    "d2h_bytes_per_step": 16.0 * N / K, "call": f"gsg_rk4_wave(plan, u_host, v_host, dt, nsteps={K})"},
            "gpu_launches": launches, "energy_sqrt": math.sqrt(max(energy, 0.0)),
        }), flush=True)


def run_recon(args, rank, local_rank, world):
    """--config recon: batched reconstruct_DG (BASELINE config 5).  One step = one batch of `npts` uniform points
    (counter-based generator, seed 20240); points are sharded over the ranks, coefficients replicated."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import gsg_b200 as g
    D, k, n = CONFIGS["recon"]
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    npts_total = int(float(os.environ.get("GSG_RECON_NPTS", "2e7")))
    npts = npts_total // world
    W, K = max(args.warmup, 3), args.steps
    plan = g.Plan(D, k, n, device=local_rank)
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    plan.set_stream(stream)
    v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
    coef = plan.tensor_construct_dev([v1] * D, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(20240 + rank)
    pts = torch.rand(npts, D, dtype=torch.float64, device=device, generator=gen)
    out = torch.empty(npts, dtype=torch.float64, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(W):
        plan.reconstruct_dev(coef, pts, npts, out)
    barrier()
    sampler = ClockSampler(local_rank)
    l0 = g.launch_count()
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        plan.reconstruct_dev(coef, pts, npts, out)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = g.launch_count() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = npts * world * K / (ms * 1e-3)
    exact = torch.prod(torch.sin(2 * math.pi * pts), dim=1)
    rms = float(torch.sqrt(torch.mean((out - exact) ** 2)))
    # e2e: host points in, host values out, through the host-pointer entry point
    hp = pts[: min(npts, 4_000_000)].cpu().pin_memory()
    hv = g.tensor_construct(D, k, n, [v1] * D)
    plan.reconstruct(hv, hp.numpy()[:1000])
    barrier()
    t0 = time.perf_counter()
    vals = plan.reconstruct(hv, hp.numpy())
    t1 = time.perf_counter()
    e2e_rate = hp.shape[0] / (t1 - t0)
    if world > 1:
        tt = torch.tensor([e2e_rate], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        e2e_rate = float(tt.item())
    nlev = math.comb(n + D, D)
    flop_pt = 2.0 * nlev * sum(k ** j for j in range(1, D + 1))            # separable contraction
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import gsg_oracle as o
        sample = hp.numpy()[:1024]
        o.reconstruct_DG_batch(D, k, n, hv, sample[:8])
        tc0 = time.perf_counter()
        ref = o.reconstruct_DG_batch(D, k, n, hv, sample)
        tc1 = time.perf_counter()
        cb = {"value": sample.shape[0] / (tc1 - tc0), "unit": "points/s", "cores": 1, "kind": "port",
              "sample": f"oracle.reconstruct_DG_batch (numpy restatement of src/dg_methods.jl:150-165) on the first "
                        f"{sample.shape[0]} points", "max_abs_diff_vs_gpu": float(np.abs(ref - vals[:sample.shape[0]]).max())}
    if rank == 0:
        line = {
            "metric": "reconstruct_DG points/sec (D=%d sparse, k=%d, n=%d)" % (D, k, n), "value": value, "unit": "points/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"reconstruct_DG batch, D={D} sparse k={k} n={n} interpolant of prod sin(2 pi x_d), "
                                   f"{npts * world} uniform points per step (BASELINE config 5 evaluates 1e8; GSG_RECON_NPTS sets it)",
                       "l2": "coefficients (21.5 MB) are L2-resident by design; points stream from HBM (40 B per point)",
                       "parallelism": "single GPU" if world == 1 else f"points sharded over {world} GPUs, coefficients replicated"},
            "roofline": {"bound": "fp64", "kernel": "reconstruct3_kernel", "achieved": value * flop_pt / 1e12,
                         "peak": FP64_PEAK_TFLOPS * world, "unit": "TFLOP/s", "frac": value * flop_pt / 1e12 / (FP64_PEAK_TFLOPS * world),
                         "traffic": None, "peak_source": "nominal B200 fp64 FMA rate (no measured fp64 peak in MEASURED_PEAKS.json)",
                         "flop_per_point": flop_pt,
                         "note": "flops = 2 |Lambda| sum_{j<=D} k^j per point (separable contraction, SURVEY 8d); the timed step "
                                 "includes the Morton-key radix sort of the points (cub) and the contraction kernel"},
            "cpu_baseline": cb,
            "e2e": {"value": e2e_rate, "unit": "points/s", "h2d_bytes_per_step": 8.0 * D * hp.shape[0], "d2h_bytes_per_step": 8.0 * hp.shape[0],
                    "call": f"gsg_reconstruct(plan, vcoeffs_host, points_host[{hp.shape[0]}], out_host): coefficients + points H2D, kernel, values D2H"},
            "gpu_launches": int(launches), "clocks": clocks, "rms_error_vs_exact": rms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="d6k3n8", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    D, k, n = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.config == "wave2d":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the RK4 advection path; --config wave2d is a GPU-only line"}))
            return
        run_wave(args, rank, local_rank, world)
        return
    if args.config == "recon":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the RK4 path; --config recon carries its own cpu_baseline"}))
            return
        run_recon(args, rank, local_rank, world)
        return
    if args.impl == "reference":
        run_reference(args, D, k, n, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import gsg_b200 as g

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libgsgb200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    W = max(args.warmup, 3)
    K = args.steps
    a = np.ones(D)

    plan = g.Plan(D, k, n, device=local_rank)
    N = plan.size
    if rank == 0 and os.environ.get("GSG_DESCRIBE"):
        print(plan.describe(), file=sys.stderr, flush=True)
    v1d = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
    u0 = g.tensor_construct(D, k, n, [v1d] * D)                  # host copy for the end-to-end (host buffer) leg
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    plan.set_stream(stream)
    y = plan.tensor_construct_dev([v1d] * D, device=device)      # initial data expanded on the device (gsg_tensor_construct_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    mg = drv = None
    if world == 1:
        def run(nsteps):
            plan.rk4_advect_dev(a, y, DT, nsteps)
    elif os.environ.get("GSG_MG_NCCL"):
        # portable path: the same block partition driven from Python with NCCL point-to-point messages
        from gsg_b200.distributed import DistComm, PartitionedRK4
        drv = PartitionedRK4(plan, a, rank, world, device, DistComm())
        drv.set_state(y)

        def run(nsteps):
            drv.step(DT, nsteps)
    else:
        # shipped path: the partitioned RK4 runs inside libgsgb200 over peer-mapped slabs (CUDA IPC); torch.distributed
        # only carries the 64-byte handles once and the barriers around the timed region
        from gsg_b200.distributed import MultiGpuRK4
        del y
        torch.cuda.empty_cache()
        mg = MultiGpuRK4(plan, rank, world)
        mg.connect_torch()
        dist.barrier()
        mg.set_state(u0)
        mg.sync()
        dist.barrier()

        def run(nsteps):
            mg.step(a, DT, nsteps)

    run(W)
    barrier()
    sampler = ClockSampler(local_rank)
    # CUDA events around the dominant kernel for a bounded sample of the timed region's launches (the first
    # 72 = two steps' worth: timing all of them costs ~8 % of the step, this sample ~0.5 %)
    # (the in-library multi-GPU driver replays the step from a CUDA graph: its sample is taken after the timed region)
    # (the flat path of the small configurations replays its steps from a CUDA graph as well: no event sample there)
    plan.profile_enable(0 if (os.environ.get("GSG_NO_PROFILE_EVENTS") or mg is not None or plan.flat_active) else 72)
    l0 = g.launch_count()
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run(K)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = g.launch_count() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if drv is not None and os.environ.get("GSG_PART_TIMING"):
        rep = drv.timing_report()
        print(f"[rank {rank}] owned {100.0 * drv.owned_doubles / plan.dev_size:.1f}% phases (device ms, host ms): "
              + "; ".join(f"{k_}: {v_[0]:.3f}/{v_[1]:.3f} mean {v_[2]:.3f} max {v_[3]:.3f}" for k_, v_ in rep.items()), file=sys.stderr, flush=True)
    if mg is not None and not os.environ.get("GSG_NO_PROFILE_EVENTS"):
        barrier()
        plan.profile_enable(72)
        mg.step(a, DT, 2)                       # eager (not replayed) steps with events around the dominant kernel
        mg.sync()
        barrier()
    if mg is None and plan.flat_active and not os.environ.get("GSG_NO_PROFILE_EVENTS"):
        barrier()
        plan.profile_enable(8)
        run(2)                                  # eager (not replayed) steps with events around the flat launches
        torch.cuda.synchronize(device)
        barrier()
    n_prof, prof_ms, prof_dofs = plan.profile_read()
    plan.profile_enable(False)
    value = N * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (TMA streaming sweep over the register-resident classes)
    peak, peak_src = load_peaks()
    roofline = None
    if n_prof > 0 and prof_ms > 0:
        algo_bytes = 16.0 * prof_dofs / n_prof                 # 16 B per DOF per directional apply
        avg_ms = prof_ms / n_prof
        achieved = algo_bytes / (avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")   # from the committed ncu --set full capture
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("sweep_stream_kernel_bytes_per_launch")
            except Exception:
                traffic = None
        flat = plan.flat_active
        roofline = {
            "bound": "hbm", "kernel": "sweep_flat_kernel" if flat else "sweep_stream_kernel", "achieved": achieved,
            "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None if flat else traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg_ms, "timed_launches": n_prof,
            # per step: 4 RHS x (D/2 PAIR launches + D reduced launches) of the streaming family; flat: 4 launches
            "kernel_share_of_step": (prof_ms / n_prof) * (4 if flat else 4 * (D + D // 2)) / (ms / K),
            "note": ("sweep_flat_kernel: ONE launch per right-hand side covers every direction (small index sets: the "
                     "state sits in L2, a step is launch-latency bound, so the HBM roofline fraction is not the figure of "
                     "merit here -- ms_per_step is); algorithmic = 16 B per DOF per directional apply" if flat else
                     "sweep_stream_kernel<K, PAIR> family: per RHS D/2 direction-pair launches (2-D sub-planes with "
                     "n' <= 2, two directional applies from one load / one store) + D launches over the remaining "
                     "short-pole groups; algorithmic = 16 B per DOF per directional apply (SURVEY 8d), so a PAIR "
                     "launch counts 32 B/DOF while moving 16-24; 'traffic' = ncu DRAM bytes per launch averaged over "
                     "one RHS; timed: the first launches of the timed region (bounded sample)"),
            # whole-step effective bandwidth under three byte models: SURVEY 8(d)'s contract model
            # (4 x 16 D + 144 B/DOF: staged stage updates), what the timed Taylor form needs by the same accounting
            # (4 x 16 D + 48), and the staged form measured beside it (filled in below)
            "step_model": {"bytes_per_dof_model": 64 * D + 144,
                           "achieved_gbs": (64 * D + 144) * N * K / (ms * 1e-3) / 1e9,
                           "frac": (64 * D + 144) * N * K / (ms * 1e-3) / 1e9 / peak,
                           "taylor": {"bytes_per_dof": 64 * D + 48,
                                      "frac": (64 * D + 48) * N * K / (ms * 1e-3) / 1e9 / peak}},
        }

    # ---- e2e: the C-ABI evolve call on HOST buffers (H2D + K steps + D2H inside the timed region)
    if world == 1:
        host = torch.from_numpy(u0.copy()).pin_memory()
        hv = host.numpy()
        plan.rk4_advect(a, hv, DT, 1)                            # warm the workspaces
        torch.cuda.synchronize(device)
        hv[:] = u0
        t0 = time.perf_counter()
        check = g.lib.gsg_rk4_advect(plan._h, g._ptr(a), g._ptr(hv), DT, K)
        t1 = time.perf_counter()
        if check != 0:
            raise RuntimeError("gsg_rk4_advect failed")
        e2e = {"value": N * K / (t1 - t0), "unit": "DOF-updates/s", "h2d_bytes_per_step": 8.0 * N / K,
               "d2h_bytes_per_step": 8.0 * N / K,
               "call": f"gsg_rk4_advect(plan, a, y_host(pinned), dt, nsteps={K}): one H2D + {K} RK4 steps + one D2H",
               "seconds": t1 - t0}
    elif mg is not None:
        # every rank: owned blocks of the pinned host state -> device, K steps, owned blocks -> pinned host
        host = torch.from_numpy(u0.copy()).pin_memory()
        out_host = torch.zeros(N, dtype=torch.float64).pin_memory()
        barrier()
        t0 = time.perf_counter()
        mg.set_state(host.numpy())
        mg.step(a, DT, K)
        mg.get_state(out_host.numpy())          # synchronises this rank
        barrier()
        t1 = time.perf_counter()
        tt = torch.tensor([t1 - t0], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt.item())
        frac, xbytes = mg.owned_fraction()
        e2e = {"value": N * K / sec, "unit": "DOF-updates/s", "h2d_bytes_per_step": 8.0 * N * frac / K,
               "d2h_bytes_per_step": 8.0 * N * frac / K, "seconds": sec,
               "call": f"per rank: gsg_mg_set_state(pinned host state; owned blocks H2D) + gsg_mg_rk4_advect(dt, {K}) + "
                       "gsg_mg_get_state(owned blocks D2H); wall clock, max over ranks"}
    else:
        host = torch.from_numpy(u0.copy()).pin_memory()
        ref_dev = torch.empty(N, dtype=torch.float64, device=device)
        full_dev = torch.zeros(plan.dev_size, dtype=torch.float64, device=device)
        out_host = torch.empty(plan.dev_size, dtype=torch.float64).pin_memory()
        barrier()
        t0 = time.perf_counter()
        ref_dev.copy_(host, non_blocking=True)
        plan.pack_dev(ref_dev, full_dev)
        drv.set_state(full_dev)
        drv.step(DT, K)
        out_host.copy_(drv.owned_state(), non_blocking=True)
        barrier()
        t1 = time.perf_counter()
        tt = torch.tensor([t1 - t0], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt.item())
        e2e = {"value": N * K / sec, "unit": "DOF-updates/s", "h2d_bytes_per_step": 8.0 * N / K,
               "d2h_bytes_per_step": 8.0 * plan.dev_size / K, "seconds": sec,
               "call": f"per rank: pinned host state -> device, PartitionedRK4.step(dt, {K}), owned part -> pinned host "
                       "(wall clock, max over ranks)"}

    staged_ms = None
    if world == 1:
        plan.set_rk4_mode(1)
        plan.rk4_advect_dev(a, y, DT, 2)
        torch.cuda.synchronize(device)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks = max(4, K // 4)
        s0.record(stream)
        plan.rk4_advect_dev(a, y, DT, ks)
        s1.record(stream)
        torch.cuda.synchronize(device)
        staged_ms = s0.elapsed_time(s1) / ks
        plan.set_rk4_mode(0)

    if roofline is not None and staged_ms:
        roofline["step_model"]["staged"] = {"bytes_per_dof": 64 * D + 144, "ms_per_step": staged_ms,
                                            "frac": (64 * D + 144) * N / (staged_ms * 1e-3) / 1e9 / peak}
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(D, k, n)

    if rank == 0:
        line = {
            "metric": "RK4 DOF-updates/sec (D=%d sparse, k=%d, n=%d)" % (D, k, n),
            "value": value, "unit": "DOF-updates/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "staged_ms_per_step": staged_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(D, k, n, N),
                       "rk4_form": "Taylor form for the linear RHS: 4 operator applies + 1 combine pass (staged form timed beside it)",
                       "initial_condition": "prod_d sin(2 pi x_d) via tensor_construct", "dt": DT,
                       "l2": ("inputs larger than L2 (4 state-sized vectors x %.0f MB vs 126 MB L2)" % (8e-6 * N) if 40e-6 * N > 126 else
                              "state smaller than L2 (%.2f MB): latency regime, no flush between steps (a flush would time the flush)" % (8e-6 * N)),
                       "sweep_path": "flat kernel (one launch per right-hand side)" if plan.flat_active else "tiled class kernels",
                       "parallelism": ("single GPU" if world == 1 else
                                       f"multi-level blocks partitioned over {world} GPUs by level==0 of the last "
                                       f"{world.bit_length() - 1} dimension(s); per RHS 2 point-to-point messages per "
                                       f"partition dimension per rank pair, "
                                       + (f"{drv.exchange_bytes_per_rhs / 1e6:.1f} MB exchanged per RHS on rank 0, "
                                          f"rank 0 owns {100.0 * drv.owned_doubles / plan.dev_size:.1f}% of the state"
                                          if mg is None else
                                          f"in-library driver (gsg_mg_*): peer-mapped slabs over CUDA IPC, pull / pull-add "
                                          f"kernels over NVLink, flag counters, step replayed from a CUDA graph; rank 0 pulls "
                                          f"{mg.owned_fraction()[1] / 1e6:.1f} MB per RHS and owns "
                                          f"{100.0 * mg.owned_fraction()[0]:.1f}% of the state"))},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
