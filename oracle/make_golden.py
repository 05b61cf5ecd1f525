"""Generates tests/golden/*.npz from the independent numpy restatement (oracle/np_restatement.py).

Run from the repo root:  python -m oracle.make_golden
The fixtures are committed; tests compare the C oracle (CPU suite) and the CUDA path (GPU suite)
against them.  Inputs are seeded (numpy default_rng) and stored alongside the outputs so that no
test needs this script or /root/reference at run time.
"""
from __future__ import annotations

import os

import numpy as np

from . import np_restatement as R

f32 = np.float32
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_mask(n, rng, ring_cells=True):
    m = R.circle_mask(n, 0.5, 0.5, 0.1)
    extra = rng.random((n, n)) < 0.04          # isolated cells, pairs, cells touching the ring
    m |= extra.astype(np.uint8)
    m[1, 3] = 1                                # interior cell adjacent to the ring (face reads it pre-mirror)
    m[n - 2, n - 3] = 1
    m[5, 1] = 1
    if ring_cells:                             # the flood fill may mark ring cells (FluidSim.cs:332)
        m[0, 4] = 1
        m[6, n - 1] = 1
    return m


def kernels_case(n, seed):
    rng = np.random.default_rng(seed)
    obs = make_mask(n, rng)
    rnd = lambda s=1.0: (rng.random((n, n), dtype=f32) * 2 - 1) * f32(s)
    out = {"obs": obs, "n": np.int32(n)}
    x = rnd()
    out["field"] = x
    for b in (0, 1, 2):
        y = x.copy()
        R.boundary(y, obs, b)
        out[f"boundary_b{b}"] = y
    # diffusion: small a (reference defaults) and a large a where pass 1 is visible
    for tag, diff, dt in (("small", 1e-4, 0.1), ("large", 0.05, 0.4)):
        out[f"diff_{tag}"] = np.array([diff, dt], f32)
        for b in (0, 1, 2):
            out[f"smooth_{tag}_b{b}"] = R.diffuse_with_jobs(b, x, diff, dt, obs, 20)
            out[f"diffuse_{tag}_b{b}"] = R.diffuse(b, x, diff, dt, obs, 20)
    guess, rhs = rnd(), rnd()
    out["ls_guess"], out["ls_rhs"] = guess, rhs
    for b in (0, 1, 2):
        out[f"linsolve_b{b}"] = R.linsolve_iterations(b, guess, rhs, 0.37, 1 + 6 * 0.37, obs, 7)
    vx, vy = rnd(2.0), rnd(2.0)
    out["vx"], out["vy"] = vx, vy
    pvx, pvy, p = R.project_with_jobs(vx, vy, obs, 20)
    out["proj_vx"], out["proj_vy"], out["proj_p"] = pvx, pvy, p
    dt = 0.05
    out["adv_dt"] = f32(dt)
    for b in (0, 1, 2):
        out[f"advect_b{b}"] = R.advect_with_jobs(b, x, vx, vy, dt, obs)
    big = rnd(40.0)                            # forces both clamps
    out["vbig"] = big
    out["advect_clamped"] = R.advect_with_jobs(0, x, big, vy, dt, obs)
    evx, evy = R.enforce_obstacles(vx, vy, obs, 1.0 / n, 1e-4)
    out["enf_vx"], out["enf_vy"] = evx, evy
    return out


def trajectory_case(n, steps, with_obstacle, seed):
    """Smoke plume in the reference's Update() order: sources, then Simulate (FluidSim.cs:405-442)."""
    rng = np.random.default_rng(seed)
    obs = R.circle_mask(n, 0.5, 0.5, 0.1) if with_obstacle else np.zeros((n, n), np.uint8)
    st = {k: np.zeros((n, n), f32) for k in ("density", "vx", "vy", "vx0", "vy0", "pressure")}
    st["vx"] = (rng.random((n, n), dtype=f32) - f32(0.5)) * f32(0.02)   # tiny seed flow so every term is live
    st["vy"] = (rng.random((n, n), dtype=f32) - f32(0.5)) * f32(0.02)
    dt, visc, diff = 0.1, 1e-4, 1e-4
    out = {"obs": obs, "n": np.int32(n), "params": np.array([dt, visc, diff, 1.0 / n, 1e-4], f32),
           "init_vx": st["vx"].copy(), "init_vy": st["vy"].copy(), "steps": np.int32(steps)}
    sx, sy, rad = 0.5 * n, 0.2 * n, max(1.5, n / 16)
    src = []
    for j in range(n):
        for i in range(n):
            d = np.sqrt(f32((i - sx) ** 2 + (j - sy) ** 2))
            if d <= rad:
                fall = f32(1.0) - f32(d / rad)
                src.append((i, j, f32(100.0) * fall, f32(0.0), f32(1.0) * fall))
    out["sources"] = np.array(src, f32)        # rows: x, y, density, vx, vy  (AddDensity/AddVelocity calls)
    for s in range(1, steps + 1):
        for i, j, dd, ax, ay in src:
            st["density"][int(j), int(i)] += dd
            st["vx"][int(j), int(i)] += ax
            st["vy"][int(j), int(i)] += ay
        R.simulate(st, obs, dt, visc, diff, enable_obstacle=with_obstacle, cell=1.0 / n, rawvisc=1e-4, iters=20)
        if s in (1, 2, steps):
            for k in ("density", "vx", "vy", "pressure"):
                out[f"step{s}_{k}"] = st[k].copy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "kernels2d_n24.npz"), **kernels_case(24, 1234))
    np.savez_compressed(os.path.join(OUT, "kernels2d_n30.npz"), **kernels_case(30, 4321))   # nx % 4 != 0
    np.savez_compressed(os.path.join(OUT, "traj2d_n32_obst.npz"), **trajectory_case(32, 6, True, 7))
    np.savez_compressed(os.path.join(OUT, "traj2d_n32_free.npz"), **trajectory_case(32, 6, False, 8))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
