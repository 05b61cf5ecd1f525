"""Independent numpy restatement of the reference's 2D solver -- TEST INFRASTRUCTURE.

Written directly from Assets/Scripts/FluidSim.cs (not from oracle/*.c), vectorised with array
slices instead of per-index loops, so that an indexing or ordering mistake in one restatement
shows up as a mismatch with the other.  ``oracle/make_golden.py`` runs it to produce the fixtures
in tests/golden/; the C oracle and the CUDA path are both checked against those fixtures.

PARITY UNPINNED by the reference itself (no C# toolchain, no reference tests): two restatements
by the same reader agreeing is weaker evidence than the reference's own output.

Arrays are float32, shape (size, size), indexed [y, x] so that flat index = x + y*size
(FluidSim.cs:749-752).  All arithmetic stays in float32; sums keep the reference's left-to-right
association.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def boundary(x: np.ndarray, obs: np.ndarray, b: int) -> None:
    """BoundaryJob.Execute, FluidSim.cs:1243-1288 (in place)."""
    n = x.shape[0]
    sx = f32(-1.0) if b == 1 else f32(1.0)
    sy = f32(-1.0) if b == 2 else f32(1.0)
    r = slice(1, n - 1)
    x[r, 0] = sx * x[r, 1]              # :1248   x[IX(0,i)]      <- x[IX(1,i)]
    x[r, n - 1] = sx * x[r, n - 2]      # :1249
    x[0, r] = sy * x[1, r]              # :1250   x[IX(i,0)]      <- x[IX(i,1)]
    x[n - 1, r] = sy * x[n - 2, r]      # :1251
    x[0, 0] = f32(0.5) * (x[0, 1] + x[1, 0])                           # :1255
    x[n - 1, 0] = f32(0.5) * (x[n - 1, 1] + x[n - 2, 0])               # :1256  IX(0,size-1)
    x[0, n - 1] = f32(0.5) * (x[0, n - 2] + x[1, n - 1])               # :1257  IX(size-1,0)
    x[n - 1, n - 1] = f32(0.5) * (x[n - 1, n - 2] + x[n - 2, n - 1])   # :1258
    if b not in (1, 2):
        return
    o = obs.astype(bool)
    inner = np.zeros_like(o)
    inner[r, r] = o[r, r]
    if b == 1:   # neighbours along x: (i-1, j) then (i+1, j)   :1269-1276
        lo_free = ~np.roll(o, 1, axis=1)
        hi_free = ~np.roll(o, -1, axis=1)
        lo_val = np.roll(x, 1, axis=1)
        hi_val = np.roll(x, -1, axis=1)
    else:        # neighbours along y: (i, j-1) then (i, j+1)   :1277-1284
        lo_free = ~np.roll(o, 1, axis=0)
        hi_free = ~np.roll(o, -1, axis=0)
        lo_val = np.roll(x, 1, axis=0)
        hi_val = np.roll(x, -1, axis=0)
    m = np.zeros_like(x)
    m = np.where(lo_free, m + (-lo_val), m)
    m = np.where(hi_free, m + (-hi_val), m)
    cnt = lo_free.astype(np.int32) + hi_free.astype(np.int32)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = np.where(cnt > 0, m / cnt.astype(f32), f32(0))
    x[inner] = val.astype(f32)[inner]


def _nb_sum(a: np.ndarray) -> np.ndarray:
    """((right + left) + top) + bottom over the interior, FluidSim.cs:1063-1066 / :1228-1229."""
    return ((a[1:-1, 2:] + a[1:-1, :-2]) + a[2:, 1:-1]) + a[:-2, 1:-1]


def diffuse_with_jobs(b, x0, diff, dt, obs, iters=20):
    """DiffuseWithJobs, FluidSim.cs:1292-1357."""
    n = x0.shape[0]
    a = f32(f32(f32(f32(dt) * f32(diff)) * f32(n - 2)) * f32(n - 2))   # :1295
    c = f32(f32(1) + f32(f32(6) * a))                                   # :1296
    fluid = ~obs.astype(bool)[1:-1, 1:-1]
    src, dst = x0.copy(), x0.copy()                                     # :1299-1300
    for _ in range(iters):
        new = (src[1:-1, 1:-1] + a * _nb_sum(src)) / c                  # :1062-1067
        inner = dst[1:-1, 1:-1]
        inner[fluid] = new[fluid]                                       # obstacles / ring not written
        boundary(dst, obs, b)
        src, dst = dst, src
    return src.copy()                                                   # :1348


def linsolve_iterations(b, x, x0, a, c, obs, iters=20):
    """LinearSolveWithJobs / PressureSolveWithJobs loops, FluidSim.cs:1378-1405, :1594-1625."""
    a, c = f32(a), f32(c)
    fluid = ~obs.astype(bool)[1:-1, 1:-1]
    rd = x.copy()
    for _ in range(iters):
        wr = rd.copy()                                                  # ring / obstacle: copy (:1209, :1216)
        new = (x0[1:-1, 1:-1] + a * _nb_sum(rd)) / c                    # :1227-1230
        inner = wr[1:-1, 1:-1]
        inner[fluid] = new[fluid]
        boundary(wr, obs, b)
        rd = wr
    return rd


def diffuse(b, x0, diff, dt, obs, iters=20):
    """Diffuse, FluidSim.cs:740-745."""
    n = x0.shape[0]
    x = diffuse_with_jobs(b, x0, diff, dt, obs, iters)
    a = f32(f32(f32(f32(dt) * f32(diff)) * f32(n - 2)) * f32(n - 2))
    return linsolve_iterations(b, x, x0, a, f32(f32(1) + f32(f32(6) * a)), obs, iters)


def project_with_jobs(vx, vy, obs, iters=20):
    """ProjectWithJobs, FluidSim.cs:1417-1521. Returns (vx, vy, p)."""
    n = vx.shape[0]
    nf = f32(n)
    div = np.zeros_like(vx)
    div[1:-1, 1:-1] = f32(-0.5) * (((vx[1:-1, 2:] - vx[1:-1, :-2]) + vy[2:, 1:-1]) - vy[:-2, 1:-1]) / nf  # :1089-1092
    p = np.zeros_like(vx)
    boundary(div, obs, 0)
    boundary(p, obs, 0)
    p = linsolve_iterations(0, p, div, 1.0, 6.0, obs, iters)           # :1581-1582
    fluid = ~obs.astype(bool)[1:-1, 1:-1]
    vx, vy = vx.copy(), vy.copy()
    gx = vx[1:-1, 1:-1] - f32(0.5) * (p[1:-1, 2:] - p[1:-1, :-2]) * nf  # :1120
    gy = vy[1:-1, 1:-1] - f32(0.5) * (p[2:, 1:-1] - p[:-2, 1:-1]) * nf  # :1121
    vx[1:-1, 1:-1][fluid] = gx[fluid]
    vy[1:-1, 1:-1][fluid] = gy[fluid]
    boundary(vx, obs, 1)
    boundary(vy, obs, 2)
    return vx, vy, p


def advect_with_jobs(b, d0, vx, vy, dt, obs):
    """AdvectWithJobs, FluidSim.cs:1523-1576 (+ AdvectJob :1138-1185)."""
    n = d0.shape[0]
    dt0 = f32(f32(dt) * f32(n - 2))                                     # :1526
    jj, ii = np.meshgrid(np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32), indexing="ij")
    x = ii.astype(f32) - dt0 * vx                                       # :1158
    y = jj.astype(f32) - dt0 * vy                                       # :1159
    hi = f32(n) - f32(1.5)
    x = np.where(x < f32(0.5), f32(0.5), x)
    x = np.where(x > hi, hi, x).astype(f32)
    y = np.where(y < f32(0.5), f32(0.5), y)
    y = np.where(y > hi, hi, y).astype(f32)
    i0 = x.astype(np.int32)
    j0 = y.astype(np.int32)
    i1, j1 = i0 + 1, j0 + 1
    s1 = x - i0.astype(f32)
    s0 = f32(1) - s1
    t1 = y - j0.astype(f32)
    t0 = f32(1) - t1
    val = s0 * (t0 * d0[j0, i0] + t1 * d0[j1, i0]) + s1 * (t0 * d0[j0, i1] + t1 * d0[j1, i1])  # :1183-1184
    out = np.zeros_like(d0)                                             # :1529
    fluid = ~obs.astype(bool)
    inner = np.zeros_like(fluid)
    inner[1:-1, 1:-1] = fluid[1:-1, 1:-1]
    out[inner] = val.astype(f32)[inner]
    boundary(out, obs, b)
    return out


def enforce_obstacles(vx, vy, obs, cell, visc):
    """EnforceObstacleBoundaries + ApplyDragNearObstacle, FluidSim.cs:617-673 (sequential, as written)."""
    import math

    n = vx.shape[0]
    vx, vy = vx.copy(), vy.copy()
    o = obs.astype(bool)
    for i in range(1, n - 1):
        for j in range(1, n - 1):
            if not o[j, i]:
                continue
            vx[j, i] = 0
            vy[j, i] = 0
            for di, dj in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                ni, nj = i + di, j + dj
                if ni < 1 or ni >= n - 1 or nj < 1 or nj >= n - 1 or o[nj, ni]:
                    continue
                U = f32(math.sqrt(float(f32(vx[nj, ni] * vx[nj, ni]) + f32(vy[nj, ni] * vy[nj, ni]))))
                Re = f32(U * f32(cell)) / f32(max(f32(visc), f32(1e-5)))
                t = f32(1.0) - f32(math.exp(float(f32(-Re) * f32(0.01))))
                t = min(max(t, f32(0)), f32(1))
                drag = f32(0.8) + f32(f32(0.98) - f32(0.8)) * f32(t)
                vx[nj, ni] = f32(vx[nj, ni] * drag)
                vy[nj, ni] = f32(vy[nj, ni] * drag)
    return vx, vy


def simulate(st: dict, obs, dt, visc, diff, enable_obstacle=True, cell=1.0 / 128, rawvisc=1e-4, iters=20):
    """Simulate / VelocityStep / DensityStep, FluidSim.cs:551-570, :703-721. st is updated in place."""
    st["vx0"] = diffuse(1, st["vx"], visc, dt, obs, iters)              # :705
    st["vy0"] = diffuse(2, st["vy"], visc, dt, obs, iters)              # :706
    st["vx0"], st["vy0"], p = project_with_jobs(st["vx0"], st["vy0"], obs, iters)   # :708
    st["vx"] = p                                                        # p aliases velocityX (:708, :1508)
    st["pressure"] = p.copy()                                           # :1509
    nvx = advect_with_jobs(1, st["vx0"], st["vx0"], st["vy0"], dt, obs)  # :710
    nvy = advect_with_jobs(2, st["vy0"], st["vx0"], st["vy0"], dt, obs)  # :711
    st["vx"], st["vy"] = nvx, nvy
    st["vx"], st["vy"], p = project_with_jobs(st["vx"], st["vy"], obs, iters)       # :713
    st["vx0"] = p                                                       # p aliases velocityX0
    st["pressure"] = p.copy()
    tmp = diffuse(0, st["density"], diff, dt, obs, iters)               # :718-719
    st["density"] = advect_with_jobs(0, tmp, st["vx"], st["vy"], dt, obs)  # :720
    if enable_obstacle:                                                 # :567-570
        st["vx"], st["vy"] = enforce_obstacles(st["vx"], st["vy"], obs, cell, rawvisc)
    return st


# ---- scene helpers (host-side, not on the timed path) ------------------------------------------
def circle_mask(size: int, cx: float, cy: float, radius: float) -> np.ndarray:
    """SetupObstacles/IsInsideShape for ObstacleShape.Circle, FluidSim.cs:302-361.  A disc is
    4-connected, so the recursive flood fill from the centre marks exactly the cells inside."""
    m = np.zeros((size, size), np.uint8)
    centre_x, centre_y = f32(cx * size), f32(cy * size)
    r = f32(radius * size)
    for y in range(size):
        for x in range(size):
            if f32((x - centre_x) * (x - centre_x)) + f32((y - centre_y) * (y - centre_y)) < r * r:
                m[y, x] = 1
    return m
