"""SURVEY.md section 8 rows a12 (AddForceToArea / UpdateCustomSource) and N4 (obstacle mask builder), on both tiers:
CPU (host-emulated core behind the same C ABI) and GPU (libfluidsolver.so).

* fs_build_obstacles (device-side SetupObstacles + RecursiveFloodFill + IsInsideShape, FluidSim.cs:302-388) against the
  host restatement of the same lines (solver.reference_mask) for the reference's three shapes, in 2D (the reference's own
  case: the two scene configurations included) and in 3D; per slab as well: every slab builds only its own planes.
* fs_set_obstacles_slab: the slab-local upload gives the same solver state as the global one.
* AddForceToArea (:452-483) through the host mirror against a cell-by-cell restatement applied to the oracle.
"""
import importlib
import os
import math

import numpy as np
import pytest

import parity_cases as P

f32 = np.float32

SHAPES_2D = [
    dict(size=128, obstacleShape="Circle"),                                             # SampleScene.unity:529-612
    dict(size=64, resolutionMultiplier=3.0, obstacleShape="Airfoil"),                   # SampleScene.unity:260-343
    dict(size=96, obstacleShape="Rectangle", obstacleWidth=0.3, obstacleHeight=0.12),
    dict(size=64, obstacleShape="Airfoil", obstacleWidth=0.4, obstaclePositionX=0.3, obstaclePositionY=0.6),
    dict(size=48, obstacleShape="Circle", obstacleRadius=0.01),                         # smaller than a cell around the seed
    dict(size=64, obstacleShape="Rectangle", obstacleWidth=0.01, obstacleHeight=0.01),  # seed outside the strict box
    dict(size=40, obstacleShape="Circle", obstacleRadius=0.45, obstaclePositionX=0.9),  # clipped by the grid
]
SHAPES_3D = [
    dict(size=32, depth=24, obstacleShape="Circle", obstacleRadius=0.2),
    dict(size=32, depth=20, obstacleShape="Rectangle", obstacleWidth=0.3, obstacleHeight=0.2, obstaclePositionZ=0.4),
    dict(size=48, depth=16, obstacleShape="Airfoil", obstacleWidth=0.35),
    dict(size=24, depth=12, obstacleShape="Circle", obstacleRadius=0.3, obstaclePositionZ=0.1),
]


def sim_for(pkg, lib, **kw):
    return pkg.FluidSimulation(lib_path=lib, use_cuda_graph=False, **kw)


def check_builder(pkg, lib, kw):
    sim = sim_for(pkg, lib, **kw)
    try:
        want = sim.reference_mask()
        P.assert_exact(sim.obstacles, want, f"device mask builder {kw}")
        assert sim.obstacleCells == int(want.sum())
        # the solver state behind it equals the one the uploaded mask gives: one step, every field bit exact
        other = sim_for(pkg, lib, **kw)
        other.native.set_obstacles(want)
        rng = np.random.default_rng(1)
        for s in (sim, other):
            s.enableCustomSource = True; s.sourceEmitsVelocity = True; s.sourceDirection = 30.0; s.sourceRadius = 2.0
            s.sourcePositionX = 0.25
        for _ in range(2):
            sim.Update(); other.Update()
        for name in ("density", "vx", "vy", "pressure"):
            P.assert_exact(sim.field(name), other.field(name), f"built vs uploaded mask: {name} {kw}")
        other.close()
    finally:
        sim.close()


def check_slab_builder(pkg, lib, slab_mod, devices):
    """Each slab handle builds its own planes; together they equal the single-grid mask, and the global presence
    bits agree (a step over the slabs equals the single-grid step)."""
    kw = dict(size=32, depth=24, obstacleShape="Airfoil", obstacleWidth=0.35)
    sim = sim_for(pkg, lib, **kw)
    want = sim.obstacles.copy()
    shape = sim.obstacle_shape()
    n, nz = sim.currentSize, sim.currentDepth
    g = slab_mod.SlabGroup(n, n, nz, len(devices), lib_path=lib, devices=devices, iters_diffuse=4, iters_pressure=6,
                           cell_size=1.0 / n)
    try:
        counts = g.call("build_obstacles", shape)
        assert set(counts) == {int(want.sum())}
        got = np.concatenate(g.call("get_obstacles"), axis=0)
        P.assert_exact(got, want, "slab-built mask")
        rng = np.random.default_rng(2)
        ref = pkg.NativeSolver(n, n, nz, iters_diffuse=4, iters_pressure=6, cell_size=1.0 / n, lib_path=lib)
        ref.set_obstacles(want)
        for name in ("density", "vx", "vy", "vz"):
            a = P.rnd((nz, n, n), rng, 1.0)
            g.set_field(name, a); ref.set_field(name, a)
        g.step(0.05, 2e-3, 2e-3); ref.step(0.05, 2e-3, 2e-3)
        for name in ("density", "vx", "vy", "vz", "pressure"):
            P.assert_exact(g.get_field(name), ref.get_field(name), f"slab-built obstacles: {name}")
        # the slab-local upload path: same state again
        g2 = slab_mod.SlabGroup(n, n, nz, len(devices), lib_path=lib, devices=devices, iters_diffuse=4, iters_pressure=6,
                                cell_size=1.0 / n)
        interior = bool(want[1:-1, 1:-1, 1:-1].any())
        g2.each(lambda r, s: s.set_obstacles_slab(want[slice(*s.halo_range())], bool(want.any()), interior))
        rng = np.random.default_rng(2)
        for name in ("density", "vx", "vy", "vz"):
            g2.set_field(name, P.rnd((nz, n, n), rng, 1.0))
        g2.step(0.05, 2e-3, 2e-3)
        for name in ("density", "vx", "vy", "vz", "pressure"):
            P.assert_exact(g2.get_field(name), ref.get_field(name), f"slab-uploaded obstacles: {name}")
        g2.close(); ref.close()
    finally:
        g.close(); sim.close()


def check_add_force_to_area(pkg, lib, O, size=48):
    """AddForceToArea :452-483: velocity with linear fall-off inside `radius`, density inside 0.3 radius; applied before
    Simulate as in Update() :414-442."""
    sim = sim_for(pkg, lib, size=size, obstacleShape="Circle")
    o = O.OracleSolver(size, size, 1, cell_size=float(sim.cellSize), raw_viscosity=1e-4)
    o.obstacles[...] = sim.obstacles
    try:
        for center, force, radius in (((20.3, 17.8), (3.0, -1.5), 5.0), ((2.2, 46.9), (-2.0, 0.5), 4.0), ((30.0, 30.0), (0.7, 0.7), 2.0)):
            sim.AddForceToArea(center, force, radius)
            cx, cy = f32(center[0]), f32(center[1])
            clamp = lambda v: min(max(v, 0), size - 1)
            for x in range(clamp(int(cx - f32(radius))), clamp(int(cx + f32(radius))) + 1):
                for y in range(clamp(int(cy - f32(radius))), clamp(int(cy + f32(radius))) + 1):
                    dist = f32(math.sqrt(f32(f32(x - cx) * f32(x - cx) + f32(y - cy) * f32(y - cy))))   # Vector2.Distance
                    if dist <= f32(radius):
                        fall = f32(1) - dist / f32(radius)
                        o.add_velocity(x, y, 0, f32(force[0]) * fall, f32(force[1]) * fall)
                        if dist < f32(radius) * f32(0.3):
                            o.add_density(x, y, 0, f32(sim.sourceStrength) * fall)
            for name in ("density", "vx", "vy"):
                P.assert_exact(sim.field(name), o.f[name], f"AddForceToArea {name} at {center}")
            sim.Simulate(); o.step(*sim.effective_parameters())
        for name in ("density", "vx", "vy", "pressure"):
            P.assert_close(sim.field(name), o.f[name], 3e-6, f"after AddForceToArea + Simulate: {name}")
    finally:
        sim.close()


# ---- CPU tier ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", SHAPES_2D + SHAPES_3D)
def test_mask_builder_emulated(pkg, emul_lib, kw):
    check_builder(pkg, emul_lib, kw)


@pytest.mark.parametrize("count", [2, 3])
def test_slab_mask_builder_emulated(pkg, emul_lib, count):
    check_slab_builder(pkg, emul_lib, importlib.import_module("3dfluidsimulation_b200.slab"), [0] * count)


def test_add_force_to_area_emulated(pkg, emul_lib, oracle):
    check_add_force_to_area(pkg, emul_lib, oracle)


def test_source_ring_many_calls(pkg, emul_lib, oracle):
    """More add-source calls than staging slots between two steps: every call lands exactly once."""
    with P.make_solver(emul_lib, 16, 12, 9) as s:
        o = oracle.OracleSolver(16, 12, 9)
        for n in range(11):
            xs = np.array([1 + n % 5, 3, 7], f32); ys = np.array([2, 4 + n % 3, 6], f32); zs = np.array([1, 2, 3 + n % 4], f32)
            s.add_source_cells(xs, ys, zs, density=np.array([1, 2, 3], f32) * f32(n + 1), ax=np.array([4, 5, 6], f32))
            for x, y, z, d, a in zip(xs, ys, zs, np.array([1, 2, 3], f32) * f32(n + 1), (4, 5, 6)):
                o.add_density(x, y, z, d); o.add_velocity(x, y, z, a, 0.0, 0.0)
        for name in ("density", "vx"):
            P.assert_exact(s.get_field(name), o.f[name], f"source ring {name}")


# ---- GPU tier ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kw", SHAPES_2D + SHAPES_3D + [dict(size=256, depth=192, obstacleShape="Airfoil", obstacleWidth=0.3),
                                                          dict(size=512, obstacleShape="Airfoil", obstacleWidth=0.4)])
def test_mask_builder_gpu(pkg, cuda_lib, kw):
    check_builder(pkg, cuda_lib, kw)


@pytest.mark.gpu
def test_slab_mask_builder_two_gpus(pkg, cuda_lib):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    check_slab_builder(pkg, cuda_lib, importlib.import_module("3dfluidsimulation_b200.slab"), [0, 1])


@pytest.mark.gpu
def test_add_force_to_area_gpu(pkg, cuda_lib, oracle):
    check_add_force_to_area(pkg, cuda_lib, oracle)


@pytest.mark.gpu
def test_source_ring_many_calls_gpu(pkg, cuda_lib, oracle):
    with P.make_solver(cuda_lib, 16, 12, 9) as s:
        o = oracle.OracleSolver(16, 12, 9)
        for n in range(23):
            xs = np.array([1 + n % 5, 3, 7], f32); ys = np.array([2, 4 + n % 3, 6], f32); zs = np.array([1, 2, 3 + n % 4], f32)
            s.add_source_cells(xs, ys, zs, density=np.array([1, 2, 3], f32) * f32(n + 1), ax=np.array([4, 5, 6], f32))
            for x, y, z, d, a in zip(xs, ys, zs, np.array([1, 2, 3], f32) * f32(n + 1), (4, 5, 6)):
                o.add_density(x, y, z, d); o.add_velocity(x, y, z, a, 0.0, 0.0)
        for name in ("density", "vx"):
            P.assert_exact(s.get_field(name), o.f[name], f"source ring {name}")


# ---- persistence sink (N4, second half) -------------------------------------------------------------------------
@pytest.mark.parametrize("jsonl", [False, True])
def test_run_log_replaces_sql_cs(pkg, emul_lib, tmp_path, jsonl):
    """SQL.SaveSimRunParams / LogRuntimeMetrics (SQL.cs:46-127) through the portable sink: one SimulationRuns row per
    attached run with the reference's columns, one RuntimeMetrics row per step whose mean density / max speed are the
    device reductions, and the reference's timeStep == 0.1f quirk on request."""
    log = pkg.RunLog(str(tmp_path / ("runs.jsonl" if jsonl else "runs.db")), jsonl=jsonl)
    sim = pkg.FluidSimulation(size=32, timeStep=0.05, lib_path=emul_lib, use_cuda_graph=False)
    sim.enableCustomSource = True; sim.sourceEmitsVelocity = True; sim.sourcePositionY = 0.2   # outside the obstacle
    run = sim.AttachRunLog(log)
    assert run == 1
    for _ in range(3):
        sim.Update()
    runs, metrics = log.rows("SimulationRuns"), log.rows("RuntimeMetrics")
    assert len(runs) == 1 and runs[0]["Size"] == 32 and runs[0]["ObstacleType"] == "Circle" and abs(runs[0]["TimeStep"] - 0.05) < 1e-9
    assert len(metrics) == 3 and all(m["RunID"] == 1 for m in metrics)
    mean, mx = sim.metrics()
    assert metrics[-1]["AverageDensity"] == pytest.approx(mean) and metrics[-1]["MaxVelocityMagnitude"] == pytest.approx(mx)
    sim.close()
    quirky = pkg.RunLog(str(tmp_path / "q.jsonl"), jsonl=True, reference_quirks=True)
    sim = pkg.FluidSimulation(size=32, lib_path=emul_lib, use_cuda_graph=False)   # timeStep = 0.1f: SQL.cs:53-56 returns -1
    assert sim.AttachRunLog(quirky) == -1
    sim.Update()
    assert not os.path.exists(str(tmp_path / "q.jsonl"))
    sim.close(); log.close()
