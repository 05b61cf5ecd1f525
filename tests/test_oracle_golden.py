"""CPU tier: the C oracle (literal 2D restatement AND the 3D oracle in nz == 1 mode) against the
golden vectors produced by the independent numpy restatement (oracle/make_golden.py).  Bit exact,
except the drag (exp), which is compared at 2e-7 relative."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

KERNEL_FILES = ["kernels2d_n24.npz", "kernels2d_n30.npz"]
TRAJ_FILES = ["traj2d_n32_obst.npz", "traj2d_n32_free.npz"]


@pytest.mark.parametrize("name", KERNEL_FILES)
def test_kernels_against_golden(oracle, name):
    O, R2 = oracle, oracle.Ref2D
    g = np.load(os.path.join(GOLDEN, name))
    obs, x, n = g["obs"], g["field"], int(g["n"])
    eq = np.testing.assert_array_equal
    for b in (0, 1, 2):
        eq(R2.boundary(b, x.copy(), obs), g[f"boundary_b{b}"])
        eq(O.set_bnd(b, x.copy(), obs), g[f"boundary_b{b}"])
    for tag in ("small", "large"):
        diff, dt = g[f"diff_{tag}"]
        a, c = O.diffuse_coeffs(n, diff, dt)
        for b in (0, 1, 2):
            eq(R2.diffuse_with_jobs(b, x, diff, dt, obs), g[f"smooth_{tag}_b{b}"])
            eq(O.diffuse_smooth(b, x, a, c, obs, 20), g[f"smooth_{tag}_b{b}"])
            eq(R2.diffuse(b, x, diff, dt, obs), g[f"diffuse_{tag}_b{b}"])
            eq(O.diffuse(b, x, diff, dt, obs, 20), g[f"diffuse_{tag}_b{b}"])
    for b in (0, 1, 2):
        eq(R2.linear_solve_with_jobs(b, g["ls_guess"], g["ls_rhs"], 0.37, 1 + 6 * 0.37, obs, 7), g[f"linsolve_b{b}"])
        eq(O.lin_solve(b, g["ls_guess"], g["ls_rhs"], 0.37, 1 + 6 * 0.37, obs, 7), g[f"linsolve_b{b}"])
    vx, vy = g["vx"], g["vy"]
    for got in (R2.project_with_jobs(vx, vy, obs), tuple(O.project(vx, vy, None, obs, 20)[i] for i in (0, 1, 3))):
        eq(got[0], g["proj_vx"]); eq(got[1], g["proj_vy"]); eq(got[2], g["proj_p"])
    dt = float(g["adv_dt"])
    for b in (0, 1, 2):
        eq(R2.advect_with_jobs(b, x, vx, vy, dt, obs), g[f"advect_b{b}"])
        eq(O.advect(b, x, vx, vy, None, dt, obs), g[f"advect_b{b}"])
    eq(R2.advect_with_jobs(0, x, g["vbig"], vy, dt, obs), g["advect_clamped"])
    eq(O.advect(0, x, g["vbig"], vy, None, dt, obs), g["advect_clamped"])
    ex, ey = R2.enforce_obstacles(vx, vy, obs, 1.0 / n, 1e-4)
    np.testing.assert_allclose(ex, g["enf_vx"], rtol=0, atol=2e-7 * np.abs(g["enf_vx"]).max())
    np.testing.assert_allclose(ey, g["enf_vy"], rtol=0, atol=2e-7 * np.abs(g["enf_vy"]).max())
    ex, ey, _ = O.enforce_obstacles(vx, vy, None, obs, 1.0 / n, 1e-4)
    eq(ex, R2.enforce_obstacles(vx, vy, obs, 1.0 / n, 1e-4)[0])


@pytest.mark.parametrize("name", TRAJ_FILES)
def test_trajectory_against_golden(oracle, name):
    """Update() order (sources, then Simulate) for 6 steps: literal 2D C restatement and the 3D oracle
    (nz == 1) both reproduce the numpy restatement's trajectory."""
    O = oracle
    g = np.load(os.path.join(GOLDEN, name))
    n, obs = int(g["n"]), g["obs"]
    dt, visc, diff, cell, rawv = (float(v) for v in g["params"])
    has_obst = bool(obs.any())
    st = {k: np.zeros((n, n), np.float32) for k in ("density", "vx", "vy", "vx0", "vy0", "pressure")}
    st["vx"][...] = g["init_vx"]; st["vy"][...] = g["init_vy"]
    o3 = O.OracleSolver(n, n, 1, enable_obstacle=has_obst, cell_size=cell, raw_viscosity=rawv)
    o3.obstacles[...] = obs
    o3.f["vx"][...] = g["init_vx"]; o3.f["vy"][...] = g["init_vy"]
    for step in range(1, int(g["steps"]) + 1):
        for x, y, d, ax, ay in g["sources"]:
            O.lib().r2_add_density(n, st["density"].ctypes.data_as(O._F), O._cf(x), O._cf(y), O._cf(d))
            O.lib().r2_add_velocity(n, st["vx"].ctypes.data_as(O._F), st["vy"].ctypes.data_as(O._F), O._cf(x), O._cf(y), O._cf(ax), O._cf(ay))
            o3.add_density(x, y, 0.0, d); o3.add_velocity(x, y, 0.0, ax, ay)
        O.Ref2D.simulate(st, obs, dt, visc, diff, has_obst, cell, rawv)
        o3.step(dt, visc, diff)
        for k in ("density", "vx", "vy", "pressure"):
            np.testing.assert_array_equal(st[k], o3.f[k], err_msg=f"K9: 3D oracle (nz=1) != literal 2D at step {step} {k}")
            if f"step{step}_{k}" in g:
                want = g[f"step{step}_{k}"]
                tol = 1e-6 * step * np.abs(want).max() if has_obst else 0.0
                np.testing.assert_allclose(st[k], want, rtol=0, atol=tol, err_msg=f"step {step} {k}")
