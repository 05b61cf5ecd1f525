#!/usr/bin/env python
"""bench.py -- throughput of the stable-fluids step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 512|1024|256|128] [--impl reference]

A "step" is one full Simulate() (velocity step + density step + obstacle pass) over the whole grid,
preceded -- as in the reference's Update() -- by the smoke-plume source injection.
Workload (default): BASELINE.json configs[3], the configuration the metric ("... at 1/2/4/8 GPUs") is
quoted on: 512^3, K_d = 20, K_p = 80 Jacobi, dt = 0.1*128/N, obstacle sphere r = 0.1N, z-slabs across
the N GPUs (strong scaling: the grid is fixed).  The state (6 GB) is far larger than L2, so no
explicit L2 flush is needed between timed steps.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel = the 3D Jacobi sweep (relax_vec4): algorithmic 13 B/voxel (12 without
                obstacles: no flag stream) * voxels per launch / average launch duration (CUDA events on the
                solver's stream, 20 back-to-back launches on the live fields in this process right after the
                timed steps), against MEASURED_PEAKS.json hbm_gbs; traffic = ncu dram bytes (profiles/).
  cpu_baseline  the CPU oracle (C restatement of FluidSim.cs, OpenMP) on a bounded sample (rank 0, N=1)
  e2e           same metric through the C ABI with HOST buffers: per step the source cells go host->
                device and density + pressure (what UpdateVisualization reads, FluidSim.cs:761-768)
                come back into pinned host memory.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, K_d, K_p, solver_kind, BASELINE.json config)
    "128": (128, 20, 40, 0, "configs[1]: 128^3, K_p=40"),
    "256": (256, 20, 20, 0, "configs[2]: 256^3 via the C ABI"),
    "512": (512, 20, 80, 0, "configs[3]: 512^3, K_p=80, z-slabs at 1/2/4/8 GPUs"),
    "1024": (1024, 20, 100, 0, "configs[4] grid with Jacobi: 1024^3, K_p=100"),
    "1024rb": (1024, 20, 100, 1, "configs[4]: 1024^3, K_p=100 red-black"),
}


def step_bytes_per_voxel(kd, kp):
    return 88 * kd + 26 * kp + 182  # SURVEY.md section 8(d) / BASELINE.md section 3


def plume(n, nz):
    """Smoke-plume source cells: ball at (0.5N, 0.2N, 0.5N), r = max(1.5, N/16),
    density +100*falloff, upward (+y) velocity v_src*falloff with CFL = dt*(N-2)*v_src ~ 2.5."""
    sx, sy, sz, rad = 0.5 * n, 0.2 * n, 0.5 * nz, max(1.5, n / 16)
    r = int(np.ceil(rad)) + 1
    ii, jj, kk = np.meshgrid(np.arange(int(sx) - r, int(sx) + r + 1), np.arange(int(sy) - r, int(sy) + r + 1),
                             np.arange(int(sz) - r, int(sz) + r + 1), indexing="ij")
    d = np.sqrt((ii - sx) ** 2 + (jj - sy) ** 2 + (kk - sz) ** 2)
    keep = d <= rad
    fall = (1.0 - d[keep] / rad).astype(np.float32)
    return (ii[keep].astype(np.float32), jj[keep].astype(np.float32), kk[keep].astype(np.float32), fall)


def sphere_mask(n, nz):
    z, y, x = np.ogrid[:nz, :n, :n]
    return (((x - 0.5 * n) ** 2 + (y - 0.5 * n) ** 2 + (z - 0.5 * nz) ** 2) < (0.1 * n) ** 2).astype(np.uint8)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region (NVML, 100 ms period)."""

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.dev, self.samples, self.reasons, self.maxclk = device_index, [], set(), None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.maxclk = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxclk, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.maxclk, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(n, kd, kp, steps, warmup=1):
    """Times the CPU oracle (OpenMP over all host cores) on an n^3 sample of the workload."""
    import oracle

    oracle.set_threads(os.cpu_count())  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    o = oracle.OracleSolver(n, n, n, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=True, cell_size=1.0 / n)
    o.obstacles[...] = sphere_mask(n, n)
    x, y, z, fall = plume(n, n)
    dt = 0.1 * 128 / n
    vsrc = 2.5 / (dt * (n - 2))

    def one():
        idx = (z.astype(np.int64) * n + y.astype(np.int64)) * n + x.astype(np.int64)
        o.f["density"].reshape(-1)[idx] += np.float32(100) * fall
        o.f["vy"].reshape(-1)[idx] += np.float32(vsrc) * fall
        o.step(dt, 1e-4, 1e-4)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dtm = time.perf_counter() - t0
    return n ** 3 * steps / dtm / 1e9, dtm / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, kd, kp, kind, cfg = WORKLOADS[args.workload]
    sample = 128 if n >= 128 else n
    cores = os.cpu_count()
    val, ms = cpu_oracle_rate(sample, kd, kp, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "Gvoxel-updates/s per full fluid step", "value": val, "unit": "Gvoxel-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg, "grid": [n, n, n], "iters_diffuse": kd, "iters_pressure": kp, "solver": "jacobi",
                   "note": "reference C# cannot run here (no dotnet/mono/Unity); this arm times the CPU oracle, a C restatement of FluidSim.cs"},
        "cpu_baseline": {"value": val, "unit": "Gvoxel-updates/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}^3 sub-grid of the workload, same K_d/K_p, {args.steps} steps after {max(args.warmup, 1)} warm-up"},
        "e2e": {"value": val, "unit": "Gvoxel-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="512", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-obstacle", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg = importlib.import_module("3dfluidsimulation_b200")
    bld = importlib.import_module("3dfluidsimulation_b200.build")
    lib = bld.build()
    n, kd, kp, kind, cfg = WORKLOADS[args.workload]
    warmup = max(args.warmup, 3)
    dt = 0.1 * 128 / n
    vsrc = 2.5 / (dt * (n - 2))
    obstacle = not args.no_obstacle

    s = pkg.NativeSolver(n, n, n, iters_diffuse=kd, iters_pressure=kp, solver_kind=kind, enable_obstacle=obstacle,
                         cell_size=1.0 / n, device_id=local, slab_rank=rank, slab_count=world,
                         use_cuda_graph=not args.no_graph, lib_path=lib)
    if world > 1:
        blobs = [None] * world
        dist.all_gather_object(blobs, s.halo_export())
        s.halo_connect(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
        dist.barrier()
    if obstacle:
        s.set_obstacles(sphere_mask(n, n))
    px, py, pz, fall = plume(n, n)
    dens_amt, vy_amt = (np.float32(100) * fall), (np.float32(vsrc) * fall)
    h2d_bytes = px.size * (8 + 4 * 2)  # what fs_add_source_cells uploads: int64 index + 2 amounts per cell

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def one_step():
        s.add_source_cells(px, py, pz, density=dens_amt, ay=vy_amt)
        s.step(dt, 1e-4, 1e-4)

    # ---- device-resident throughput ----------------------------------------------------------------------
    for _ in range(warmup):
        one_step()
    s.sync()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = s.launch_count()
    s.timer_start()
    for _ in range(args.steps):
        one_step()
    ms = s.timer_stop()
    launches = s.launch_count() - l0
    barrier()
    clocks = sampler.result()
    ms = max_over_ranks(ms)
    value = n ** 3 * args.steps / (ms * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    # Every step: source cells host->device, fs_step, then density + pressure device->host into pinned buffers.
    # The readback is the pipelined form of the ABI (fs_get_field_async): step t's fields travel over PCIe while
    # step t+1 computes; two host buffer sets alternate so that the "application" can still read set t while
    # set t+1 is in flight.  The timed region ends only when the last transfer has landed (fs_wait_transfers).
    host = [[torch.empty(s.shape, dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)] for _ in range(2)]
    one_step(); s.get_field_async("density", host[0][0]); s.get_field_async("pressure", host[0][1]); s.wait_transfers()
    barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        one_step()
        s.get_field_async("density", host[it & 1][0])
        s.get_field_async("pressure", host[it & 1][1])
    s.wait_transfers()
    s.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n ** 3 * args.steps / e2e_s / 1e9
    d2h_bytes = int(host[0][0].nbytes + host[0][1].nbytes)
    # the same loop with the blocking fs_get_field, for comparison
    t0 = time.perf_counter()
    for it in range(args.steps):
        one_step()
        s.get_field("density", host[0][0])
        s.get_field("pressure", host[0][1])
    e2e_blocking_s = max_over_ranks(time.perf_counter() - t0)

    # ---- roofline of the dominant kernel (3D Jacobi sweep), live, CUDA events on the solver stream ------------
    peak, peak_src = measured_peak()
    sweep_ms, sweep_bytes = s.bench_sweep(1, 0, 20)
    sweep_ms = max_over_ranks(sweep_ms)
    smooth_ms, smooth_bytes = s.bench_sweep(0, 0, 20)
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    own = s.owned_voxels
    # the ncu capture is of one launch at 512^3 on one GPU: only quote it for that case
    traffic = ncu_traffic("relax_vec4_jacobi_3d") if (world == 1 and n == 512) else None
    roofline = {
        "bound": "hbm", "kernel": "relax_vec4<JACOBI,3D>", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": sweep_bytes, "avg_launch_ms": sweep_ms,
        "frac_of_8000_nominal": achieved / 8000.0,
        "smoother_GBps": smooth_bytes / (smooth_ms * 1e-3) / 1e9,
        "step_level": {"bytes_per_voxel": step_bytes_per_voxel(kd, kp),
                       "achieved_GBps": step_bytes_per_voxel(kd, kp) * own * world / (ms / args.steps * 1e-3) / 1e9 / world,
                       "note": "per-GPU, reference-structure algorithmic bytes / measured step time"},
    }

    line = {
        "metric": "Gvoxel-updates/s per full fluid step", "value": value, "unit": "Gvoxel-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg, "grid": [n, n, n], "iters_diffuse": kd, "iters_pressure": kp,
                   "solver": "red-black" if kind else "jacobi", "dt": dt, "obstacle": "sphere r=0.1N" if obstacle else "none",
                   "parallelism": f"z-slabs x{world}", "cuda_graph": not args.no_graph,
                   "l2": "state (45 B/voxel) is larger than L2; no flush needed"},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "Gvoxel-updates/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_s / args.steps * 1e3,
                "blocking_readback_value": n ** 3 * args.steps / e2e_blocking_s / 1e9,
                "note": "per step: fs_add_source_cells (host->device) + fs_step + fs_get_field_async(density, pressure) into "
                        "pinned host memory, fs_wait_transfers before the clock stops; blocking_readback_value = same loop with fs_get_field"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = 128 if n >= 128 else n
        csteps = 8
        val, cms = cpu_oracle_rate(sample, kd, kp, csteps, 1)
        line["cpu_baseline"] = {"value": val, "unit": "Gvoxel-updates/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{sample}^3 sub-grid, same K_d/K_p, {csteps} steps after 1 warm-up ({cms:.0f} ms/step)"}
    else:
        line["cpu_baseline"] = None
    s.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
