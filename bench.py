#!/usr/bin/env python
"""bench.py -- throughput of the stable-fluids step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 512|1024|1024rb|256|128|32]
                    [--scaling strong|weak] [--impl reference]

A "step" is one full Simulate() (velocity step + density step + obstacle pass) over the whole grid,
preceded -- as in the reference's Update() -- by the smoke-plume source injection.
Workload (default): BASELINE.json configs[3], the configuration the metric ("... at 1/2/4/8 GPUs") is
quoted on: 512^3, K_d = 20, K_p = 80 Jacobi, dt = 0.1*128/N, obstacle sphere r = 0.1N, z-slabs across
the N GPUs (strong scaling: the grid is fixed; --scaling weak keeps 512 x 512 x 512 per GPU instead).
The state (6 GB) is far larger than L2, so no explicit L2 flush is needed between timed steps.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      the kernel with the largest share of the step (the Jacobi sweep relax_vec4; the fused single-pass sweep
                when the solver is red-black): ALGORITHMIC bytes per launch (in 4 + rhs 4 + flags 1 + out 4 =
                13 B/voxel) / average launch duration (CUDA events on the solver's stream, back-to-back
                launches on the live fields right after the timed steps) against MEASURED_PEAKS.json hbm_gbs; `kernels`
                lists the same figure for every kernel (in_step: launched by the default policy); traffic = ncu dram
                bytes per launch (profiles/).
  cpu_baseline  the CPU oracle (C restatement of FluidSim.cs, OpenMP, all host cores) on a bounded sample: a
                z-slab of the SAME nx x ny grid (same rows, same coefficients) with 256^3 voxels' worth of planes;
                "port-tidy" = oracle/fluid_oracle.c, "port-faithful" = oracle/ref_faithful3d.c (static-64 job
                batches, serial BoundaryJob scans, per-call allocate-and-copy, as FluidSim.cs executes).
  e2e           same metric through the C ABI with HOST buffers, wall clock.  value = the drop-in's frame (Update(),
                FluidSim.cs:390-450): source cells host->device, fs_step, fs_render_rgba of the viewed plane and
                fs_get_metrics device->host, a host sync every step (use the on-device consumers
                in the e2e loop).  full_fields_value = round 1's definition, kept beside it: density + pressure of the
                whole grid (what UpdateVisualization reads in the 2D reference, :761-768; 1 GB at 512^3) come back into
                pinned host memory every step through the pipelined readback.
  parity_check  (N > 1) before timing, a small slab case over the same N ranks is compared bit for bit with a
                1-GPU handle on rank 0.
  extra         short 1024^3 K_p = 100 runs (BASELINE configs[4]: Jacobi and red-black, strong scaling) and the
                weak-scaling variant 1024 x 1024 x 128 per GPU.
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
    "32": (32, 20, 20, 0, "configs[0]: 32^3, K_d=K_p=20"),
    "128": (128, 20, 40, 0, "configs[1]: 128^3, K_p=40"),
    "256": (256, 20, 20, 0, "configs[2]: 256^3 via the C ABI"),
    "512": (512, 20, 80, 0, "configs[3]: 512^3, K_p=80, z-slabs at 1/2/4/8 GPUs"),
    "1024": (1024, 20, 100, 0, "configs[4] grid with Jacobi: 1024^3, K_p=100"),
    "1024rb": (1024, 20, 100, 1, "configs[4]: 1024^3, K_p=100 red-black"),
}
METRIC = "Gvoxel-updates/s per full fluid step"


def step_bytes_per_voxel(kd, kp):
    return 88 * kd + 26 * kp + 182  # SURVEY.md section 8(d) / BASELINE.md section 3


def plume(nx, ny, nz):
    """Smoke-plume source cells: ball at (0.5 nx, 0.2 ny, 0.5 nz), r = max(1.5, nx/16),
    density +100*falloff, upward (+y) velocity v_src*falloff with CFL = dt*(N-2)*v_src ~ 2.5."""
    sx, sy, sz, rad = 0.5 * nx, 0.2 * ny, 0.5 * nz, max(1.5, nx / 16)
    r = int(np.ceil(rad)) + 1
    ks = np.arange(max(int(sz) - r, 0), min(int(sz) + r, nz - 1) + 1)
    js = np.arange(max(int(sy) - r, 0), min(int(sy) + r, ny - 1) + 1)
    is_ = np.arange(max(int(sx) - r, 0), min(int(sx) + r, nx - 1) + 1)
    kk, jj, ii = np.meshgrid(ks, js, is_, indexing="ij")
    d = np.sqrt((ii - sx) ** 2 + (jj - sy) ** 2 + (kk - sz) ** 2)
    keep = d <= rad
    fall = (1.0 - d[keep] / rad).astype(np.float32)
    return (ii[keep].astype(np.float32), jj[keep].astype(np.float32), kk[keep].astype(np.float32), fall)


def sphere_mask(nx, ny, nz):
    """Obstacle sphere r = 0.1 nx at the centre, built plane by plane (no nx*ny*nz float temporaries at 1024^3)."""
    y, x = np.ogrid[:ny, :nx]
    d2 = (x - 0.5 * nx) ** 2 + (y - 0.5 * ny) ** 2
    m = np.zeros((nz, ny, nx), np.uint8)
    r2 = (0.1 * nx) ** 2
    for k in range(nz):
        dz2 = (k - 0.5 * nz) ** 2
        if dz2 < r2:
            m[k] = d2 + dz2 < r2
    return m


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


# ---- CPU arm ---------------------------------------------------------------------------------------------
def cpu_sample_dims(n):
    """Bounded sample of the n^3 workload for the CPU arm: the whole grid up to 256^3, else a z-slab of the same
    nx x ny grid holding 256^3 voxels (512 -> 512x512x64, 1024 -> 1024x1024x16): same row length, same plane size,
    same coefficients (N = nx), the per-voxel work of the full grid."""
    if n <= 256:
        return n, n, n
    return n, n, max(256 ** 3 // (n * n), 8)


def cpu_oracle_rate(dims, kd, kp, steps, warmup=1, faithful=False):
    """Times the CPU oracle (OpenMP over all host cores) on an nx x ny x nz sample of the workload."""
    import oracle

    oracle.set_threads(os.cpu_count())  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    nx, ny, nz = dims
    o = oracle.OracleSolver(nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=True, cell_size=1.0 / nx)
    o.obstacles[...] = sphere_mask(nx, ny, nz)
    x, y, z, fall = plume(nx, ny, nz)
    dt = 0.1 * 128 / nx
    vsrc = 2.5 / (dt * (nx - 2))
    idx = (z.astype(np.int64) * ny + y.astype(np.int64)) * nx + x.astype(np.int64)

    def one():
        o.f["density"].reshape(-1)[idx] += np.float32(100) * fall
        o.f["vy"].reshape(-1)[idx] += np.float32(vsrc) * fall
        o.step(dt, 1e-4, 1e-4, faithful=faithful)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dtm = time.perf_counter() - t0
    return nx * ny * nz * steps / dtm / 1e9, dtm / steps * 1e3


def cpu_baseline_block(n, kd, kp, steps, warmup, faithful_steps=2):
    dims = cpu_sample_dims(n)
    val, ms = cpu_oracle_rate(dims, kd, kp, steps, warmup)
    fval, fms = cpu_oracle_rate(dims, kd, kp, faithful_steps, 1, faithful=True)
    sample = f"{dims[0]}x{dims[1]}x{dims[2]} voxels" + ("" if dims[2] == n else f" (z-slab of the {n}^3 grid, same rows/planes/coefficients)")
    return dims, {
        "value": val, "unit": "Gvoxel-updates/s", "cores": os.cpu_count(), "kind": "port-tidy",
        "sample": f"{sample}, same K_d/K_p, {steps} steps after {warmup} warm-up ({ms:.0f} ms/step); oracle/fluid_oracle.c, OpenMP over planes",
        "faithful": {"value": fval, "unit": "Gvoxel-updates/s", "cores": os.cpu_count(), "kind": "port-faithful",
                     "sample": f"{sample}, {faithful_steps} steps after 1 warm-up ({fms:.0f} ms/step); oracle/ref_faithful3d.c: "
                               "static-64 job batches, single-threaded BoundaryJob scan after every sweep, per-call allocate-and-copy "
                               "(FluidSim.cs:1299-1301, :1324, :1645)"},
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, kd, kp, kind, cfg = WORKLOADS[args.workload]
    warm = max(args.warmup, 1)
    dims, block = cpu_baseline_block(n, kd, kp, args.steps, warm)
    val = block["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gvoxel-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": dims[0] * dims[1] * dims[2] / val / 1e6,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg, "grid": list(dims), "workload_grid": [n, n, n], "iters_diffuse": kd, "iters_pressure": kp,
                   "solver": "jacobi", "dt": 0.1 * 128 / n, "obstacle": "sphere r=0.1N",
                   "note": "reference C# cannot run here (no dotnet/mono/Unity); this arm times the CPU oracle, a C restatement of "
                           "FluidSim.cs, on `grid` (a bounded sample of `workload_grid`); per-voxel rate"},
        "cpu_baseline": block,
        "e2e": {"value": val, "unit": "Gvoxel-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm -----------------------------------------------------------------------------------------------
class Job:
    """Process-group plumbing shared by the timed runs (one process per GPU)."""

    def __init__(self, torch, dist, rank, world, local):
        self.torch, self.dist, self.rank, self.world, self.local = torch, dist, rank, world, local

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def connect(self, s):
        if self.world > 1:
            blobs = [None] * self.world
            self.dist.all_gather_object(blobs, s.halo_export())
            s.halo_connect(blobs[self.rank - 1] if self.rank > 0 else None,
                           blobs[self.rank + 1] if self.rank < self.world - 1 else None)
            self.dist.barrier()


def make_plume_solver(pkg, job, lib, dims, kd, kp, kind, obstacle, graph):
    nx, ny, nz = dims
    s = pkg.NativeSolver(nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, solver_kind=kind, enable_obstacle=obstacle,
                         cell_size=1.0 / nx, device_id=job.local, slab_rank=job.rank, slab_count=job.world,
                         use_cuda_graph=graph, lib_path=lib)
    job.connect(s)
    if obstacle:
        s.set_obstacles(sphere_mask(nx, ny, nz))
    px, py, pz, fall = plume(nx, ny, nz)
    dt = 0.1 * 128 / nx
    vsrc = 2.5 / (dt * (nx - 2))
    dens_amt, vy_amt = (np.float32(100) * fall), (np.float32(vsrc) * fall)

    def one_step():
        s.add_source_cells(px, py, pz, density=dens_amt, ay=vy_amt)
        s.step(dt, 1e-4, 1e-4)

    return s, one_step, px.size


def timed_run(job, s, one_step, steps, warmup, sample_clocks=False, settle_s=0.5):
    """W warm-up steps (more when W steps are shorter than settle_s: at 8 GPUs a 512^3 step is 17 ms and three of them
    do not get the graph, the peer mappings and the clocks into steady state -- the same build measured 17.7 and 20.8
    ms/step there back to back, while the later frame loop sat at 16.6), then exactly `steps` timed steps."""
    times = []
    for _ in range(warmup):
        t0 = time.perf_counter()
        one_step()
        s.sync()
        times.append(time.perf_counter() - t0)
    per_step = max(min(times), 1e-4) if times else settle_s  # (the first step also captures the graph: take the fastest)
    extra = int(job.max_over_ranks(float(max(0, int(np.ceil(settle_s / per_step)) - warmup))))
    extra = min(extra, 200)
    for _ in range(extra):
        one_step()
    s.sync()
    timed_run.warmup_steps_run = warmup + extra
    sampler = ClockSampler(job.local) if sample_clocks else None
    job.barrier()
    if sampler:
        sampler.start()
    l0 = s.launch_count()
    s.timer_start()
    for _ in range(steps):
        one_step()
    ms = s.timer_stop()
    launches = s.launch_count() - l0
    job.barrier()
    clocks = sampler.result() if sampler else None
    return job.max_over_ranks(ms), launches, clocks


def parity_check(pkg, job, lib):
    """Slab correctness inside the benchmark job: 64 x 40 x 48, obstacles, K_d = 6, K_p = 8,
    3 steps with the CUDA graph, on the same `world` ranks / GPUs as the timed run; rank 0 repeats it on one GPU with
    a single handle and compares every owned plane of every rank bit for bit."""
    nx, ny, nz, steps = 64, 40, 48, 3
    rng = np.random.default_rng(5)
    shape = (nz, ny, nx)
    mask = (rng.random(shape) < 0.04).astype(np.uint8)
    fields = {n: ((rng.random(shape, dtype=np.float32) * 2 - 1) * np.float32(2.0)).astype(np.float32) for n in ("density", "vx", "vy", "vz")}
    kw = dict(iters_diffuse=6, iters_pressure=8, enable_obstacle=True, cell_size=1.0 / nx, use_cuda_graph=True, lib_path=lib)
    out = {"ranks": job.world, "grid": [nx, ny, nz], "steps": steps, "solvers": {}}
    ok_all = True
    for kind, name in ((0, "jacobi"), (1, "red-black")):
        s = pkg.NativeSolver(nx, ny, nz, device_id=job.local, slab_rank=job.rank, slab_count=job.world, solver_kind=kind, **kw)
        job.connect(s)
        s.set_obstacles(mask)
        for n, a in fields.items():
            s.set_field(n, a[s.z_begin:s.z_end])
        for _ in range(steps):
            s.step(0.05, 3e-3, 2e-3)
        mine = {n: s.get_field(n) for n in ("density", "vx", "vy", "vz", "pressure")}
        s.close()
        parts = [None] * job.world
        job.dist.all_gather_object(parts, mine)
        if job.rank == 0:
            ref = pkg.NativeSolver(nx, ny, nz, device_id=job.local, solver_kind=kind, **kw)
            ref.set_obstacles(mask)
            for n, a in fields.items():
                ref.set_field(n, a)
            for _ in range(steps):
                ref.step(0.05, 3e-3, 2e-3)
            exact = True
            for n in mine:
                got = np.concatenate([p[n] for p in parts], axis=0)
                exact = exact and bool(np.array_equal(got, ref.get_field(n)))
            ref.close()
            out["solvers"][name] = exact
            ok_all = ok_all and exact
    out["bit_exact"] = ok_all
    return out


def short_run(pkg, job, lib, dims, kd, kp, kind, steps, warmup, graph=True):
    s, one_step, _ = make_plume_solver(pkg, job, lib, dims, kd, kp, kind, True, graph)
    try:
        ms, launches, _ = timed_run(job, s, one_step, steps, warmup)
        vox = dims[0] * dims[1] * dims[2]
        return {"grid": list(dims), "iters_diffuse": kd, "iters_pressure": kp, "solver": "red-black" if kind else "jacobi",
                "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "value": vox * steps / (ms * 1e-3) / 1e9,
                "unit": "Gvoxel-updates/s", "n_gpus": job.world, "gpu_launches": int(launches)}
    finally:
        s.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="512", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-obstacle", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the 1024^3 / weak-scaling side measurements")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline list")
    ap.add_argument("--grid", default=None, help="nx,ny,nz override of the workload's grid (experiments: e.g. 512,512,128 on 2 GPUs "
                                                 "has the slab thickness of 512^3 on 8)")
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
    job = Job(torch, dist, rank, world, local)

    pkg = importlib.import_module("3dfluidsimulation_b200")
    bld = importlib.import_module("3dfluidsimulation_b200.build")
    if local == 0:
        lib = bld.build()          # one builder per node (atomic rename inside); the others wait and load
    if world > 1:
        dist.barrier()
    lib = bld.LIB
    n, kd, kp, kind, cfg = WORKLOADS[args.workload]
    warmup = max(args.warmup, 3)
    obstacle = not args.no_obstacle
    graph = not args.no_graph
    dims = (n, n, n * world) if args.scaling == "weak" else (n, n, n)
    if args.grid:
        dims = tuple(int(v) for v in args.grid.split(","))
        cfg += f" [grid overridden: {args.grid}]"
    voxels = dims[0] * dims[1] * dims[2]

    check = parity_check(pkg, job, lib) if world > 1 else None

    s, one_step, ncells = make_plume_solver(pkg, job, lib, dims, kd, kp, kind, obstacle, graph)
    h2d_bytes = ncells * (8 + 4 * 2)  # what fs_add_source_cells uploads: int64 index + 2 amounts per cell
    dt = 0.1 * 128 / n

    # ---- device-resident throughput ----------------------------------------------------------------------
    ms, launches, clocks = timed_run(job, s, one_step, args.steps, warmup, sample_clocks=True)
    warmup_steps_run = timed_run.warmup_steps_run
    value = voxels * args.steps / (ms * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    # Every step: source cells host->device, fs_step, then density + pressure device->host into pinned buffers.
    # The readback is the pipelined form of the ABI (fs_get_field_async): step t's fields travel over PCIe while
    # step t+1 computes; two host buffer sets alternate so that the "application" can still read set t while
    # set t+1 is in flight.  The timed region ends only when the last transfer has landed (fs_wait_transfers).
    host = [[torch.empty(s.shape, dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)] for _ in range(2)]
    one_step(); s.get_field_async("density", host[0][0]); s.get_field_async("pressure", host[0][1]); s.wait_transfers()
    job.barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        one_step()
        s.get_field_async("density", host[it & 1][0])
        s.get_field_async("pressure", host[it & 1][1])
    s.wait_transfers()
    s.sync()
    e2e_s = job.max_over_ranks(time.perf_counter() - t0)
    e2e_value = voxels * args.steps / e2e_s / 1e9
    d2h_bytes = int(host[0][0].nbytes + host[0][1].nbytes)
    # the drop-in's frame: Update() = sources, Simulate(), UpdateVisualization() (colour mapping on the device, one
    # RGBA plane back) and LogCurrentMetrics() (two reductions back) -- FluidSim.cs:390-450, :755-866, :578-607
    vis = pkg.native.FsVisParams.reference_defaults(n, 2)
    vis.z_slice = dims[2] // 2
    renders = s.z_begin <= vis.z_slice < s.z_end
    rgba = np.empty((dims[1], dims[0], 4), np.float32)
    job.barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        one_step()
        if renders:
            s.render_rgba(vis, rgba)
        s.metrics()
    s.sync()
    frame_s = job.max_over_ranks(time.perf_counter() - t0)
    frame_value = voxels * args.steps / frame_s / 1e9
    del host

    # ---- roofline of the step's kernels, live, CUDA events on the solver stream --------------------------------
    peak, peak_src = measured_peak()
    own = s.owned_voxels

    def sweep(kind_id, reps):
        try:
            ms_k, by = s.bench_sweep(kind_id, 0, reps)
        except pkg.FluidSolverError:
            return None
        return job.max_over_ranks(ms_k), by

    fl = 1 if obstacle else 0
    kernel_table = [
        # (bench kind, name, reference job, compulsory B/voxel of THIS kernel, sweeps of the reference it replaces)
        (7, "relax_pair_kernel<JACOBI>", "2 x (LinearSolveIterationJob + BoundaryJob)", 12 + fl, 2),
        (8, "relax_pair_kernel<SMOOTH>", "2 x (DiffuseJob + BoundaryJob)", 8 + fl, 2),
        (9, "relax_pair_kernel<RED_BLACK>", "red-black full sweep (both colours + set_bnd)", 12 + fl, 1),
        (1, "relax_vec4<JACOBI>", "LinearSolveIterationJob + BoundaryJob", 12 + fl, 1),
        (0, "relax_vec4<SMOOTH>", "DiffuseJob + BoundaryJob", 8 + fl, 1),
        (2, "rb_vec4 x2", "red-black full sweep as two colour launches", 24 + 2 * fl, 1),
        (3, "advect (scalar)", "AdvectJob + BoundaryJob", 20 + fl, 1),
        (4, "advect_velocity (3 components fused)", "3 x (AdvectJob + BoundaryJob)", 24 + fl, 3),
        (5, "divergence_vec4", "ProjectDivergenceJob + 2 BoundaryJob", 16, 1),
        (6, "gradient_vec4", "ProjectVelocityAdjustJob + 3 BoundaryJob", 28 + fl, 1),
    ]
    kernels = []
    measure = [1, 0] + ([9] if kind else []) if args.no_kernels else [k[0] for k in kernel_table]
    if world > 1:
        measure = [k for k in measure if k in (0, 1, 2, 7, 8, 9)]   # the once-per-step kernels peer-read neighbours' live fields
    order = [k for k in (3, 4, 5, 7, 8, 9, 1, 0, 2, 6) if k in measure]   # gradient last: it overwrites the velocities
    results = {}
    for kid in order:
        results[kid] = sweep(kid, 5 if kid in (3, 4, 5, 6) else 20)
    for kid, name, job_name, bpv, nsweeps in kernel_table:
        r = results.get(kid)
        if not r:
            continue
        ms_k, _ = r
        gbs = bpv * own / (ms_k * 1e-3) / 1e9
        kernels.append({"kernel": name, "replaces": job_name, "bytes_per_voxel": bpv, "avg_launch_ms": ms_k,
                        "achieved_GBps": gbs, "frac": gbs / peak, "reference_sweeps_per_launch": nsweeps})
    by_name = {k["kernel"]: k for k in kernels}
    # which kernels the step launches under the default policy (csrc/fluidsolver.cu pair_supported): single sweeps for
    # Jacobi / smoother, the fused kernel for red-black; the others are listed for comparison
    fused_env = os.environ.get("FS_PAIR", "")
    for k in kernels:
        name = k["kernel"]
        if name.startswith("relax_pair_kernel"):
            k["in_step"] = fused_env == "1" or (name.endswith("<RED_BLACK>") and kind == 1 and fused_env != "0")
        elif name == "rb_vec4 x2":
            k["in_step"] = kind == 1 and fused_env == "0"
        else:
            k["in_step"] = True
    # the dominant kernel of the workload: the pressure sweep (160 of the 328 sweeps of the 512^3 step)
    dom = (by_name.get("relax_pair_kernel<RED_BLACK>") if kind else None) or by_name.get("relax_vec4<JACOBI>")
    traffic_key = {"relax_pair_kernel<JACOBI>": "relax_pair_jacobi_3d", "relax_pair_kernel<RED_BLACK>": "relax_pair_rb_3d",
                   "relax_vec4<JACOBI>": "relax_vec4_jacobi_3d"}[dom["kernel"]]
    traffic = ncu_traffic(traffic_key) if (world == 1 and n == 512 and args.scaling == "strong") else None
    step_bpv = step_bytes_per_voxel(kd, kp)
    roofline = {
        "bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_GBps"], "peak": peak, "unit": "GB/s",
        "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src,
        "bytes_per_voxel": dom["bytes_per_voxel"], "algorithmic_bytes_per_launch": dom["bytes_per_voxel"] * own,
        "avg_launch_ms": dom["avg_launch_ms"], "frac_of_8000_nominal": dom["achieved_GBps"] / 8000.0,
        "reference_sweeps_per_launch": dom["reference_sweeps_per_launch"],
        "reference_structure_GBps": dom["achieved_GBps"] * dom["reference_sweeps_per_launch"],
        "note": "achieved = this kernel's compulsory bytes (in + rhs + flags + out) / its launch time; a fused pair does two "
                "sweeps of the reference per launch, so against SURVEY.md 8(d)'s per-sweep figure it counts twice "
                "(reference_structure_GBps)",
        "kernels": kernels,
        "step_level": {"bytes_per_voxel": step_bpv,
                       "achieved_GBps": step_bpv * own / (ms / args.steps * 1e-3) / 1e9,
                       "frac": step_bpv * own / (ms / args.steps * 1e-3) / 1e9 / peak,
                       "note": "per-GPU, reference-structure algorithmic bytes (88 K_d + 26 K_p + 182 B/voxel) / measured step time; "
                               "fused sweeps make this exceed 1"},
    }
    s.close()

    line = {
        "metric": METRIC, "value": value, "unit": "Gvoxel-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg, "grid": list(dims), "iters_diffuse": kd, "iters_pressure": kp,
                   "solver": "red-black" if kind else "jacobi", "dt": dt, "obstacle": "sphere r=0.1N" if obstacle else "none",
                   "parallelism": f"z-slabs x{world}", "cuda_graph": graph, "warmup_steps_run": warmup_steps_run,
                   "l2": "state (45 B/voxel) is larger than L2; no flush needed"},
        "roofline": roofline,
        "e2e": {"value": frame_value, "unit": "Gvoxel-updates/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(rgba.nbytes + 12), "ms_per_step": frame_s / args.steps * 1e3,
                "frame_value": frame_value,
                "full_fields_value": e2e_value, "full_fields_d2h_bytes_per_step": d2h_bytes,
                "full_fields_ms_per_step": e2e_s / args.steps * 1e3,
                "note": "value = the drop-in's Update() through the C ABI, wall clock, a host sync every step: "
                        "fs_add_source_cells (host->device through pinned staging) + fs_step + fs_render_rgba (the viewed "
                        "plane, device->host) + fs_get_metrics (device->host) -- what FluidSim.cs:390-450 does per frame; "
                        "full_fields_value = the same with density + pressure of the WHOLE grid read back every step "
                        "(fs_get_field_async into pinned host memory, fs_wait_transfers before the clock stops): this was "
                        "`value` in round 1 and is PCIe / host-ingest bound at 8 GPUs (1 GB per step)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if check is not None:
        line["parity_check"] = check
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, line["cpu_baseline"] = cpu_baseline_block(n, kd, kp, 6, 1)
    else:
        line["cpu_baseline"] = None

    # ---- side measurements: BASELINE configs[4] (1024^3, K_p = 100) and weak scaling ------------------------------
    if not args.no_extra and args.workload == "512" and args.scaling == "strong":
        extra = {}
        for key, edims, ekind in (("1024_jacobi", (1024, 1024, 1024), 0), ("1024_red_black", (1024, 1024, 1024), 1),
                                  ("weak_1024x1024x128_per_gpu_red_black", (1024, 1024, 128 * world), 1)):
            try:
                extra[key] = short_run(pkg, job, lib, edims, 20, 100, ekind, steps=3, warmup=2, graph=graph)
                extra[key]["scaling"] = "weak" if key.startswith("weak") else "strong"
            except Exception as e:  # keep the headline line even if a side run fails (e.g. out of memory)
                extra[key] = {"error": repr(e)[:300]}
        line["extra"] = extra
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
