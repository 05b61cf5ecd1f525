"""z-slab decomposition.

CPU tier: (1) the solver core's halo placement -- P slab handles of the host-emulated core, one thread
each, same lock-step protocol as the CUDA executor -- must reproduce the single-grid oracle bit for
bit (K10); (2) a world_size-2 gloo run in which two PROCESSES exchange one-plane halos per Jacobi
iteration around oracle sweeps; (3) the partition helper agrees with fs_slab_range.
GPU tier (needs >= 2 devices, skipped otherwise): the same bit-exactness with real P2P stores between
two B200s, in-process (peer access) and across processes (CUDA IPC under torchrun)."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

import parity_cases as P
from conftest import ROOT


def slab_mod():
    return importlib.import_module("3dfluidsimulation_b200.slab")


def run_group_vs_oracle(lib, O, nx, ny, nz, count, obstacles, vscale, steps=2, kd=4, kp=6, devices=None, graph=False,
                        local_obstacle=False, red_black=False):
    rng = np.random.default_rng(3)
    shape = (nz, ny, nx)
    mask = P.random_mask(shape, rng, 0.06) if obstacles else np.zeros(shape, np.uint8)
    if local_obstacle:   # obstacle cells in ONE slab only: every slab must still run the same op sequence
        mask = np.zeros(shape, np.uint8)
        mask[nz // 2 - 1:nz // 2 + 1, 3:6, 4:8] = 1
    g = slab_mod().SlabGroup(nx, ny, nz, count, lib_path=lib, devices=devices, iters_diffuse=kd, iters_pressure=kp,
                             enable_obstacle=obstacles, cell_size=1.0 / nx, use_cuda_graph=graph,
                             solver_kind=1 if red_black else 0)
    o = O.OracleSolver(nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=obstacles, cell_size=1.0 / nx,
                       red_black=red_black)
    try:
        g.set_obstacles(mask); o.obstacles[...] = mask
        for n in ("density", "vx", "vy", "vz"):
            a = P.rnd(shape, rng, vscale)
            g.set_field(n, a); o.f[n][...] = a
        for _ in range(steps):
            g.step(0.05, 3e-3, 2e-3); o.step(0.05, 3e-3, 2e-3)
        for n in ("density", "vx", "vy", "vz", "pressure"):
            if obstacles:
                P.assert_close(g.get_field(n), o.f[n], 2e-6, f"P={count} {n}")   # drag uses exp()
            else:
                P.assert_exact(g.get_field(n), o.f[n], f"P={count} {n}")
        mean, mx = g.metrics()
        omean, omx = o.metrics()
        assert abs(mean - omean) <= 1e-5 * abs(omean) + 1e-12 and abs(mx - omx) <= 1e-5 * omx
    finally:
        g.close()


@pytest.mark.parametrize("count", [2, 3, 4])
@pytest.mark.parametrize("obstacles", [False, True])
def test_emulated_slabs_match_single_grid(emul_lib, oracle, count, obstacles):
    run_group_vs_oracle(emul_lib, oracle, 12, 10, 16, count, obstacles, vscale=1.5)


@pytest.mark.parametrize("count", [2, 3, 4])
@pytest.mark.parametrize("obstacles", [False, True])
@pytest.mark.parametrize("kd,kp", [(4, 6), (5, 7), (1, 2)])
def test_emulated_slabs_extended_sweeps(emul_lib, oracle, count, obstacles, kd, kp, monkeypatch):
    """FS_EXTEND=1: two sweeps per halo operation -- the first of each couple also computes the first ghost plane from the
    two-deep ghost zone and exchanges nothing, the second exchanges; readers of ghost planes outside an operation
    acknowledge.  Even and odd iteration counts; still the single-grid oracle bit for bit."""
    monkeypatch.setenv("FS_EXTEND", "1")
    run_group_vs_oracle(emul_lib, oracle, 12, 10, 16, count, obstacles, vscale=1.5, kd=kd, kp=kp)


@pytest.mark.parametrize("count", [2, 3])
@pytest.mark.parametrize("kp", [5, 6])
def test_emulated_slabs_red_black(emul_lib, oracle, count, kp):
    """Red-black pressure solve (BASELINE config 5) across slabs: colour parity uses the GLOBAL z, halos travel
    between the colour passes; the slab run equals the single-grid oracle."""
    run_group_vs_oracle(emul_lib, oracle, 12, 10, 16, count, True, vscale=1.5, kp=kp, red_black=True)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("count", [3, 4])
def test_emulated_slabs_obstacle_in_one_slab_only(emul_lib, oracle, count):
    """Regression: with the obstacle confined to the middle slab(s) the outer ranks used to skip the mirror /
    drag halo operations, the ranks' sequence numbers diverged and the exchange deadlocked (seen on 4 GPUs)."""
    run_group_vs_oracle(emul_lib, oracle, 12, 10, 18, count, True, vscale=1.5, local_obstacle=True)


def test_emulated_slabs_cross_slab_backtrace(emul_lib, oracle):
    """dt0*|v| up to 3 cells: the trilinear gather reads the neighbour slab's planes (peer view)."""
    run_group_vs_oracle(emul_lib, oracle, 12, 10, 16, 2, False, vscale=6.0)


def test_partition_helper_matches_core(emul_lib, pkg):
    for nz, count in ((16, 2), (17, 3), (64, 8), (512, 8), (30, 4)):
        want = slab_mod().slab_bounds(nz, count)
        for r in range(count):
            s = pkg.NativeSolver(8, 8, nz, slab_rank=r, slab_count=count, lib_path=emul_lib)
            assert (s.z_begin, s.z_end) == want[r]
            s.close()
    with pytest.raises(pkg.FluidSolverError):
        pkg.NativeSolver(8, 8, 1, slab_rank=0, slab_count=2, lib_path=emul_lib)   # 2D grids do not slab


GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)
shape = (14, 9, 10)
x0 = rng.random(shape, dtype=np.float32); guess = rng.random(shape, dtype=np.float32)
mask = (rng.random(shape) < 0.05).astype(np.uint8)
a, c, iters = np.float32(0.3), np.float32(2.8), 6
zb, ze = shape[0] * rank // world, shape[0] * (rank + 1) // world
lo, hi = max(zb - 1, 0), min(ze + 1, shape[0])
cur = guess[lo:hi].copy()
for it in range(iters):
    nxt = O.lin_solve(0, cur, x0[lo:hi], a, c, mask[lo:hi], 1)      # one sweep on the slab (ghosts included)
    # the slab-local set_bnd treated the ghost planes as z faces: restore them, then exchange real halos
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(nxt[zb - lo].copy()), rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(nxt[ze - 1 - lo].copy()), rank + 1))
    if rank > 0:
        t = torch.empty(shape[1:], dtype=torch.float32); dist.recv(t, rank - 1); nxt[0] = t.numpy()
    if rank < world - 1:
        t = torch.empty(shape[1:], dtype=torch.float32); dist.recv(t, rank + 1); nxt[-1] = t.numpy()
    for r in reqs: r.wait()
    cur = nxt
parts = [None] * world
dist.all_gather_object(parts, cur[zb - lo:ze - lo])
if rank == 0:
    full = O.lin_solve(0, guess, x0, a, c, mask, iters)
    got = np.concatenate(parts, axis=0)
    # interior planes of every slab must match the single-grid run bit for bit
    assert np.array_equal(got[1:-1], full[1:-1]), float(np.abs(got - full).max())
    print("GLOO_SLAB_OK")
dist.destroy_process_group()
'''


def test_two_process_gloo_halo_exchange(oracle, tmp_path):
    """world_size 2, gloo, CPU: one-plane halo exchange per Jacobi iteration between two processes."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "GLOO_SLAB_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


# ---- GPU tier -----------------------------------------------------------------------------------------
def _gpu_count():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("obstacles", [False, True])
def test_two_gpu_slabs_extended_sweeps(cuda_lib, oracle, obstacles, monkeypatch):
    """FS_EXTEND=1 on two real GPUs (the executor reads the variable when the handle is created)."""
    monkeypatch.setenv("FS_EXTEND", "1")
    test_two_gpu_slabs_in_process(cuda_lib, oracle, obstacles, True)


@pytest.mark.gpu
@pytest.mark.parametrize("obstacles", [False, True])
@pytest.mark.parametrize("graph", [False, True])
def test_two_gpu_slabs_in_process(cuda_lib, oracle, obstacles, graph):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    run_group_vs_oracle(cuda_lib, oracle, 64, 40, 48, 2, obstacles, vscale=3.0, steps=3, kd=6, kp=8, devices=[0, 1], graph=graph)


@pytest.mark.gpu
@pytest.mark.parametrize("dims,kp", [((64, 40, 48), 8), ((128, 128, 128), 100)])
def test_two_gpu_slabs_red_black(cuda_lib, oracle, dims, kp):
    """BASELINE config 5's solver over 2 slabs on 2 GPUs (P2P halos between the colour passes / fused sweeps):
    equal to the single-grid oracle, 128^3 with K_p = 100 included."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    oracle.set_threads(os.cpu_count())
    run_group_vs_oracle(cuda_lib, oracle, *dims, 2, True, vscale=1.0, steps=1, kd=4, kp=kp, devices=[0, 1], graph=True,
                        red_black=True)


IPC_WORKER = r'''
import os, sys, importlib, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
import parity_cases as P
pkg = importlib.import_module("3dfluidsimulation_b200"); slab = importlib.import_module("3dfluidsimulation_b200.slab")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx, ny, nz = 64, 40, 48
rng = np.random.default_rng(5)
shape = (nz, ny, nx)
mask = P.random_mask(shape, rng, 0.04)
fields = {n: P.rnd(shape, rng, 3.0) for n in ("density", "vx", "vy", "vz")}
kw = dict(iters_diffuse=6, iters_pressure=8, enable_obstacle=True, cell_size=1.0 / nx, use_cuda_graph=True)
s = pkg.NativeSolver(nx, ny, nz, device_id=local, slab_rank=rank, slab_count=world, **kw)
slab.connect_distributed(s, dist, rank, world)
s.set_obstacles(mask)
for n, a in fields.items(): s.set_field(n, a[s.z_begin:s.z_end])
for _ in range(4): s.step(0.05, 3e-3, 2e-3)
mine = {n: s.get_field(n) for n in ("density", "vx", "vy", "vz", "pressure")}
parts = [None] * world
dist.all_gather_object(parts, mine)
s.close()
if rank == 0:
    ref = pkg.NativeSolver(nx, ny, nz, device_id=0, **kw)
    ref.set_obstacles(mask)
    for n, a in fields.items(): ref.set_field(n, a)
    for _ in range(4): ref.step(0.05, 3e-3, 2e-3)
    for n in mine:
        got = np.concatenate([p[n] for p in parts], axis=0)
        assert np.array_equal(got, ref.get_field(n)), (n, float(np.abs(got - ref.get_field(n)).max()))
    ref.close()
    print("IPC_SLAB_OK")
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_two_gpu_slabs_across_processes(cuda_lib, tmp_path):
    """torchrun, one process per GPU, CUDA IPC peer mapping: the slab run equals the 1-GPU run bit for bit."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "ipc_worker.py"
    script.write_text(IPC_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29612", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "IPC_SLAB_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
