"""CPU tier, randomized: the solver core (host-emulated executor, same per-cell functions and orchestration as the
CUDA library) against the oracle on random tiny grids -- dimensions down to 3, random obstacle masks (ring cells
included), random iteration counts (odd counts rotate the ping-pong buffers), random physical parameters.
Bit exact without obstacles; with obstacles the drag's exp() allows 2e-6 relative per step."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import parity_cases as P


@st.composite
def grids(draw):
    three_d = draw(st.booleans())
    nx = draw(st.integers(3, 11))
    ny = draw(st.integers(3, 9))
    nz = draw(st.integers(3, 8)) if three_d else 1
    return nx, ny, nz


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(dims=grids(), seed=st.integers(0, 2 ** 31 - 1), kd=st.integers(0, 5), kp=st.integers(0, 6),
       fill=st.sampled_from([0.0, 0.05, 0.3]), dt=st.sampled_from([0.01, 0.1, 0.4]),
       visc=st.sampled_from([0.0, 1e-4, 0.05]), vscale=st.sampled_from([0.0, 0.5, 20.0]))
def test_random_steps_match_oracle(emul_lib, oracle, dims, seed, kd, kp, fill, dt, visc, vscale):
    nx, ny, nz = dims
    rng = np.random.default_rng(seed)
    shape = P.shape_of(nx, ny, nz)
    mask = (rng.random(shape) < fill).astype(np.uint8)
    obstacles = bool(mask.any())
    with P.make_solver(emul_lib, nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=True,
                       cell_size=1.0 / nx, raw_viscosity=visc) as s:
        o = oracle.OracleSolver(nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=True,
                                cell_size=1.0 / nx, raw_viscosity=visc)
        s.set_obstacles(mask); o.obstacles[...] = mask
        for name in ("density", "vx", "vy") + (("vz",) if nz > 1 else ()):
            a = P.rnd(shape, rng, vscale if name != "density" else 5.0)
            s.set_field(name, a); o.f[name][...] = a
        for _ in range(2):
            s.step(dt, visc, 2e-3); o.step(dt, visc, 2e-3)
        for name in ("density", "vx", "vy", "pressure") + (("vz",) if nz > 1 else ()):
            got, want = s.get_field(name), o.f[name]
            if obstacles:
                P.assert_close(got, want, 4e-6, f"{name} {dims} seed={seed}")
            else:
                P.assert_exact(got, want, f"{name} {dims} seed={seed}")


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(dims=grids(), seed=st.integers(0, 2 ** 31 - 1), b=st.integers(0, 3), iters=st.integers(1, 5),
       a=st.sampled_from([0.0, 0.009, 0.65, 40.0]))
def test_random_relaxation_matches_oracle(emul_lib, oracle, dims, seed, b, iters, a):
    nx, ny, nz = dims
    if b == 3 and nz == 1:
        b = 0
    rng = np.random.default_rng(seed)
    shape = P.shape_of(nx, ny, nz)
    mask = (rng.random(shape) < 0.2).astype(np.uint8)
    x0, guess = P.rnd(shape, rng), P.rnd(shape, rng)
    c = np.float32(1 + 6 * a)
    with P.make_solver(emul_lib, nx, ny, nz) as s:
        s.set_obstacles(mask)
        s.set_field("vx", x0); s.set_field("vx0", P.rnd(shape, rng))
        s.op_smooth("vx0", "vx", b, a, c, iters)
        P.assert_exact(s.get_field("vx0"), oracle.diffuse_smooth(b, x0, a, c, mask, iters), f"smooth {dims} b={b}")
        s.set_field("vy0", guess)
        s.op_lin_solve("vy0", "vx", b, a, c, iters)
        P.assert_exact(s.get_field("vy0"), oracle.lin_solve(b, guess, x0, a, c, mask, iters), f"jacobi {dims} b={b}")
        s.set_field("vy0", guess)
        s.op_lin_solve("vy0", "vx", b, a, c, iters, solver_kind=1)
        P.assert_exact(s.get_field("vy0"), oracle.lin_solve(b, guess, x0, a, c, mask, iters, red_black=True), f"red-black {dims} b={b}")


# ---- the same randomized checks against the real CUDA library (float4 kernels when nx % 4 == 0) ------------------
@st.composite
def gpu_grids(draw):
    three_d = draw(st.booleans())
    nx = draw(st.sampled_from([4, 8, 12, 16, 20, 36, 132, 5, 7, 30]))
    ny = draw(st.integers(3, 12))
    nz = draw(st.integers(3, 20)) if three_d else 1
    return nx, ny, nz


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(dims=gpu_grids(), seed=st.integers(0, 2 ** 31 - 1), kd=st.integers(0, 5), kp=st.integers(0, 6),
       fill=st.sampled_from([0.0, 0.05, 0.3]), dt=st.sampled_from([0.01, 0.1, 0.4]),
       visc=st.sampled_from([0.0, 1e-4, 0.05]), vscale=st.sampled_from([0.0, 0.5, 20.0]), graph=st.booleans())
def test_random_steps_match_oracle_gpu(cuda_lib, oracle, dims, seed, kd, kp, fill, dt, visc, vscale, graph):
    nx, ny, nz = dims
    rng = np.random.default_rng(seed)
    shape = P.shape_of(nx, ny, nz)
    mask = (rng.random(shape) < fill).astype(np.uint8)
    obstacles = bool(mask.any())
    with P.make_solver(cuda_lib, nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=True,
                       cell_size=1.0 / nx, raw_viscosity=visc, use_cuda_graph=graph) as s:
        o = oracle.OracleSolver(nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=True,
                                cell_size=1.0 / nx, raw_viscosity=visc)
        s.set_obstacles(mask); o.obstacles[...] = mask
        for name in ("density", "vx", "vy") + (("vz",) if nz > 1 else ()):
            a = P.rnd(shape, rng, vscale if name != "density" else 5.0)
            s.set_field(name, a); o.f[name][...] = a
        for _ in range(3):
            s.step(dt, visc, 2e-3); o.step(dt, visc, 2e-3)
        for name in ("density", "vx", "vy", "pressure") + (("vz",) if nz > 1 else ()):
            got, want = s.get_field(name), o.f[name]
            if obstacles:
                P.assert_close(got, want, 6e-6, f"{name} {dims} seed={seed}")
            else:
                P.assert_exact(got, want, f"{name} {dims} seed={seed}")


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(dims=gpu_grids(), seed=st.integers(0, 2 ** 31 - 1), b=st.integers(0, 3), iters=st.integers(1, 5),
       a=st.sampled_from([0.0, 0.009, 0.65, 40.0]))
def test_random_relaxation_matches_oracle_gpu(cuda_lib, oracle, dims, seed, b, iters, a):
    nx, ny, nz = dims
    if b == 3 and nz == 1:
        b = 0
    rng = np.random.default_rng(seed)
    shape = P.shape_of(nx, ny, nz)
    mask = (rng.random(shape) < 0.2).astype(np.uint8)
    x0, guess = P.rnd(shape, rng), P.rnd(shape, rng)
    c = np.float32(1 + 6 * a)
    with P.make_solver(cuda_lib, nx, ny, nz) as s:
        s.set_obstacles(mask)
        s.set_field("vx", x0); s.set_field("vx0", P.rnd(shape, rng))
        s.op_smooth("vx0", "vx", b, a, c, iters)
        P.assert_exact(s.get_field("vx0"), oracle.diffuse_smooth(b, x0, a, c, mask, iters), f"smooth {dims} b={b}")
        s.set_field("vy0", guess)
        s.op_lin_solve("vy0", "vx", b, a, c, iters)
        P.assert_exact(s.get_field("vy0"), oracle.lin_solve(b, guess, x0, a, c, mask, iters), f"jacobi {dims} b={b}")
        s.set_field("vy0", guess)
        s.op_lin_solve("vy0", "vx", b, a, c, iters, solver_kind=1)
        P.assert_exact(s.get_field("vy0"), oracle.lin_solve(b, guess, x0, a, c, mask, iters, red_black=True), f"red-black {dims} b={b}")
