"""Parity cases shared by the CPU tier (host emulation of the solver core) and the GPU tier (the real
libfluidsolver.so).  Every case drives the C ABI exactly as the reference drives its own jobs and
compares with the oracle on identical seeded inputs.

Tolerances (stated here once):
  * every kernel except the obstacle drag: BIT-EXACT (np.array_equal).  The library is built with
    -fmad=false / IEEE division, the oracle with -ffp-contract=off, both keep the reference's
    association order, so fp32 results must agree to the last bit.
  * obstacle drag (exp in double, FluidSim.cs:667): <= 2e-7 relative to max|V| -- CUDA's exp() is
    within 1 ulp (double) of glibc's, which can flip the float rounding in rare cases.
  * whole steps that include the drag inherit that tolerance: 1e-6 * max|field| per step.
"""
from __future__ import annotations

import importlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32


def pkg():
    return importlib.import_module("3dfluidsimulation_b200")


def shape_of(nx, ny, nz):
    return (ny, nx) if nz == 1 else (nz, ny, nx)


def random_mask(shape, rng, fill=0.05):
    m = (rng.random(shape) < fill).astype(np.uint8)
    # make sure interior cells next to the ring and ring cells themselves are covered
    def put(*idx):
        if all(0 <= i < n for i, n in zip(idx, shape)):
            m[idx] = 1
    if len(shape) == 2:
        put(1, 2); put(0, 3); put(shape[0] - 2, shape[1] - 2)
    else:
        put(1, 1, 2); put(0, 2, 2); put(shape[0] - 2, shape[1] - 2, shape[2] - 2); put(2, 1, 1)
    return m


def rnd(shape, rng, scale=1.0):
    return ((rng.random(shape, dtype=f32) * 2 - 1) * f32(scale)).astype(f32)


def make_solver(lib, nx, ny, nz, **kw):
    return pkg().NativeSolver(nx, ny, nz, lib_path=lib, **kw)


def assert_exact(got, want, what):
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first {bad[0]}, max abs diff "
                             f"{np.abs(got - want).max():.3g} (max |want| {np.abs(want).max():.3g})")


def assert_close(got, want, rel, what):
    scale = max(float(np.abs(want).max()), 1e-30)
    err = float(np.abs(got - want).max())
    assert err <= rel * scale + 1e-12, f"{what}: max abs err {err:.3g} > {rel:g} * {scale:.3g}"


# ---- per-operator cases ---------------------------------------------------------------------------------
def case_set_bnd(lib, O, nx, ny, nz, seed=1):
    rng = np.random.default_rng(seed)
    shape = shape_of(nx, ny, nz)
    mask = random_mask(shape, rng)
    x = rnd(shape, rng)
    with make_solver(lib, nx, ny, nz) as s:
        s.set_obstacles(mask)
        for b in range(0, 4 if nz > 1 else 3):
            s.set_field("density", x)
            s.op_set_bnd("density", b)
            assert_exact(s.get_field("density"), O.set_bnd(b, x.copy(), mask), f"set_bnd b={b} {shape}")


def case_smooth_and_linsolve(lib, O, nx, ny, nz, seed=2, obstacles=True, iters=(1, 2, 3, 6)):
    rng = np.random.default_rng(seed)
    shape = shape_of(nx, ny, nz)
    mask = random_mask(shape, rng) if obstacles else np.zeros(shape, np.uint8)
    x0, guess = rnd(shape, rng), rnd(shape, rng)
    a, c = f32(0.37), f32(1 + 6 * 0.37)
    with make_solver(lib, nx, ny, nz) as s:
        s.set_obstacles(mask)
        for b in range(0, 4 if nz > 1 else 3):
            for it in iters:
                s.set_field("vx", x0)
                s.set_field("vx0", rnd(shape, rng))   # garbage in the destination must not matter
                s.op_smooth("vx0", "vx", b, a, c, it)
                assert_exact(s.get_field("vx0"), O.diffuse_smooth(b, x0, a, c, mask, it), f"smooth b={b} it={it} {shape}")
                s.set_field("vy0", guess)
                s.op_lin_solve("vy0", "vx", b, a, c, it)
                assert_exact(s.get_field("vy0"), O.lin_solve(b, guess, x0, a, c, mask, it), f"lin_solve b={b} it={it} {shape}")
            s.set_field("vx", x0)
            s.op_diffuse("vx0", "vx", b, 2e-3, 0.3)
            assert_exact(s.get_field("vx0"), O.diffuse(b, x0, 2e-3, 0.3, mask, 20), f"diffuse b={b} {shape}")


def case_project(lib, O, nx, ny, nz, seed=3, obstacles=True, iters=7):
    rng = np.random.default_rng(seed)
    shape = shape_of(nx, ny, nz)
    mask = random_mask(shape, rng) if obstacles else np.zeros(shape, np.uint8)
    v = [rnd(shape, rng, 2.0) for _ in range(3)]
    with make_solver(lib, nx, ny, nz, iters_pressure=iters) as s:
        s.set_obstacles(mask)
        names = ("vx", "vy", "vz") if nz > 1 else ("vx", "vy")
        for n, a in zip(names, v):
            s.set_field(n, a)
        s.op_project(False)
        want = O.project(v[0], v[1], v[2] if nz > 1 else None, mask, iters)
        assert_exact(s.get_field("divergence"), O.divergence(v[0], v[1], v[2] if nz > 1 else None, mask), f"divergence {shape}")
        assert_exact(s.get_field("pressure"), want[3], f"pressure {shape}")
        for n, w in zip(names, want[:3]):
            assert_exact(s.get_field(n), w, f"project {n} {shape}")


def case_advect(lib, O, nx, ny, nz, seed=4, obstacles=True, vscale=3.0, coherent=False):
    """coherent=True: sub-cell displacements whose sign is constant over large blocks (plus a block of exact zeros), the
    regime of a real flow away from the plume core -- what the float4 kernel's shared-displacement path handles."""
    rng = np.random.default_rng(seed)
    shape = shape_of(nx, ny, nz)
    mask = random_mask(shape, rng) if obstacles else np.zeros(shape, np.uint8)
    d0 = rnd(shape, rng)
    v = [rnd(shape, rng, vscale) for _ in range(3)]
    if coherent:
        for n, a in enumerate(v):
            a[...] = np.abs(a) * f32(0.2 / (0.05 * max(nx - 2, 1)) / max(vscale, 1e-6))   # |displacement| < 0.2 cells
            sl = [slice(None)] * a.ndim
            sl[n % a.ndim] = slice(a.shape[n % a.ndim] // 2, None)
            a[tuple(sl)] *= f32(-1)                                                          # sign flips half way along one axis
            a[tuple(slice(0, max(1, d // 3)) for d in a.shape)] = 0                          # a corner block at rest
    vz = v[2] if nz > 1 else None
    dt = 0.05
    with make_solver(lib, nx, ny, nz) as s:
        s.set_obstacles(mask)
        s.set_field("vx", v[0]); s.set_field("vy", v[1])
        if nz > 1:
            s.set_field("vz", v[2])
        for b in range(0, 4 if nz > 1 else 3):
            s.set_field("vx0", d0)
            s.op_advect("density", "vx0", b, dt, use_v0_fields=False)
            assert_exact(s.get_field("density"), O.advect(b, d0, v[0], v[1], vz, dt, mask), f"advect b={b} {shape}")
        # fused self-advection of the velocity, FluidSim.cs:710-711
        s.set_field("vx0", v[0]); s.set_field("vy0", v[1])
        if nz > 1:
            s.set_field("vz0", v[2])
        s.op_advect_velocity(dt)
        assert_exact(s.get_field("vx"), O.advect(1, v[0], v[0], v[1], vz, dt, mask), f"advect_velocity x {shape}")
        assert_exact(s.get_field("vy"), O.advect(2, v[1], v[0], v[1], vz, dt, mask), f"advect_velocity y {shape}")
        if nz > 1:
            assert_exact(s.get_field("vz"), O.advect(3, v[2], v[0], v[1], vz, dt, mask), f"advect_velocity z {shape}")


def case_enforce(lib, O, nx, ny, nz, seed=5):
    rng = np.random.default_rng(seed)
    shape = shape_of(nx, ny, nz)
    mask = random_mask(shape, rng, 0.1)
    v = [rnd(shape, rng, 2.0) for _ in range(3)]
    vz = v[2] if nz > 1 else None
    with make_solver(lib, nx, ny, nz, cell_size=1.0 / nx, raw_viscosity=1e-4) as s:
        s.set_obstacles(mask)
        s.set_field("vx", v[0]); s.set_field("vy", v[1])
        if nz > 1:
            s.set_field("vz", v[2])
        s.op_enforce_obstacles()
        want = O.enforce_obstacles(v[0], v[1], vz, mask, 1.0 / nx, 1e-4)
        for n, w in zip(("vx", "vy", "vz"), want):
            if w is not None:
                assert_close(s.get_field(n), w, 2e-7, f"enforce {n} {shape}")


def case_sources(lib, O, nx, ny, nz):
    """AddDensity/AddVelocity: (int) truncation toward zero, clamp to the grid (FluidSim.cs:723-738)."""
    with make_solver(lib, nx, ny, nz) as s:
        o = O.OracleSolver(nx, ny, nz)
        pts = [(2.9, 3.1, 1.5), (-4.0, 0.2, -1.0), (nx + 7.0, ny - 0.01, nz + 3.0), (0.999, ny * 0.5, nz * 0.5), (-0.5, -0.5, -0.5)]
        for n, (x, y, z) in enumerate(pts):
            s.add_density(x, y, z, 10.0 + n); o.add_density(x, y, z, 10.0 + n)
            s.add_velocity(x, y, z, 1.0 + n, -2.0, 0.5); o.add_velocity(x, y, z, 1.0 + n, -2.0, 0.5)
        # batched form with a duplicate cell
        xs = np.array([3.2, 3.7, 1.0], f32); ys = np.array([2.0, 2.0, 1.0], f32); zs = np.array([1.0, 1.0, 1.0], f32)
        s.add_source_cells(xs, ys, zs, density=np.array([1, 2, 3], f32), ax=np.array([4, 5, 6], f32))
        for x, y, z, d, a in zip(xs, ys, zs, (1, 2, 3), (4, 5, 6)):
            o.add_density(x, y, z, d); o.add_velocity(x, y, z, a, 0.0, 0.0)
        for name in ("density", "vx", "vy") + (("vz",) if nz > 1 else ()):
            assert_exact(s.get_field(name), o.f[name], f"sources {name}")
        dense = rnd(shape_of(nx, ny, nz), np.random.default_rng(0))
        s.add_sources(density=dense); o.add_sources(d=dense)
        assert_exact(s.get_field("density"), o.f["density"], "dense add")


def plume_cells(nx, ny, nz):
    """Smoke-plume source cells: ball at (0.5 nx, 0.2 ny, 0.5 nz), r = max(1.5, nx/16), linear
    fall-off 1 - dist/r as in UpdateCustomSource (FluidSim.cs:503-512).  Integer cell coordinates as float32."""
    sx, sy, sz, rad = 0.5 * nx, 0.2 * ny, 0.5 * nz, max(1.5, nx / 16)
    r = int(np.ceil(rad)) + 1
    ks = np.arange(max(int(sz) - r, 0), min(int(sz) + r, nz - 1) + 1) if nz > 1 else np.array([0])
    js = np.arange(max(int(sy) - r, 0), min(int(sy) + r, ny - 1) + 1)
    is_ = np.arange(max(int(sx) - r, 0), min(int(sx) + r, nx - 1) + 1)
    kk, jj, ii = np.meshgrid(ks, js, is_, indexing="ij")
    d = np.sqrt((ii - sx) ** 2 + (jj - sy) ** 2 + ((kk - sz) ** 2 if nz > 1 else 0.0))
    keep = d <= rad
    return (ii[keep].astype(f32), jj[keep].astype(f32), kk[keep].astype(f32), (1.0 - d[keep] / rad).astype(f32))


def run_steps(lib, O, nx, ny, nz, steps, *, kd=20, kp=20, obstacles=True, use_graph=False, seed=6, dt=0.1,
              visc=1e-4, diff=1e-4, vsrc=1.0, solver_kwargs=None, oracle_kwargs=None):
    """Plume in Update() order (sources, then the step) on solver and oracle; returns both."""
    rng = np.random.default_rng(seed)
    shape = shape_of(nx, ny, nz)
    mask = np.zeros(shape, np.uint8)
    if obstacles:
        if nz == 1:
            yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
            mask = (((xx - 0.5 * nx) ** 2 + (yy - 0.5 * ny) ** 2) < (0.1 * nx) ** 2).astype(np.uint8)
        else:
            zz, yy, xx = np.ogrid[:nz, :ny, :nx]   # broadcast: no nx*ny*nz index arrays (512^3 would need 3 GB)
            mask = (((xx - 0.5 * nx) ** 2 + (yy - 0.5 * ny) ** 2 + (zz - 0.5 * nz) ** 2) < (0.1 * nx) ** 2).astype(np.uint8)
    s = make_solver(lib, nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=obstacles,
                    cell_size=1.0 / nx, use_cuda_graph=use_graph, **(solver_kwargs or {}))
    o = O.OracleSolver(nx, ny, nz, iters_diffuse=kd, iters_pressure=kp, enable_obstacle=obstacles, cell_size=1.0 / nx,
                       **(oracle_kwargs or {}))
    s.set_obstacles(mask); o.obstacles[...] = mask
    for name in ("vx", "vy") + (("vz",) if nz > 1 else ()):
        a = rnd(shape, rng, 0.01)
        s.set_field(name, a); o.f[name][...] = a
    cx, cy, cz, fall = plume_cells(nx, ny, nz)
    dens, vyamt = (f32(100) * fall).astype(f32), (f32(vsrc) * fall).astype(f32)
    flat = (cz.astype(np.int64) * ny + cy.astype(np.int64)) * nx + cx.astype(np.int64)   # unique cells: a fancy += is exact
    for _ in range(steps):
        s.add_source_cells(cx, cy, cz, density=dens, ay=vyamt)
        o.f["density"].reshape(-1)[flat] += dens
        o.f["vy"].reshape(-1)[flat] += vyamt
        s.step(dt, visc, diff); o.step(dt, visc, diff)
    return s, o


def case_steps(lib, O, nx, ny, nz, steps, rel=1e-6, **kw):
    s, o = run_steps(lib, O, nx, ny, nz, steps, **kw)
    try:
        for name in ("density", "vx", "vy", "pressure") + (("vz",) if nz > 1 else ()):
            got, want = s.get_field(name), o.f[name]
            if kw.get("obstacles", True):
                assert_close(got, want, rel * steps, f"{steps}-step {name} {(nx, ny, nz)}")
            else:
                assert_exact(got, want, f"{steps}-step {name} {(nx, ny, nz)} (no drag => bit exact)")
        mean, mx, _ = s.metrics()
        omean, omx = o.metrics()
        assert abs(mean - omean) <= 1e-5 * max(abs(omean), 1e-12) and abs(mx - omx) <= 1e-5 * max(omx, 1e-12)
    finally:
        s.close()


def case_golden_2d(lib, name):
    """The CUDA path (or the emulated core) against the committed numpy-restatement fixtures."""
    g = np.load(os.path.join(GOLDEN, name))
    n, obs, x = int(g["n"]), g["obs"], g["field"]
    with make_solver(lib, n, n, 1, iters_diffuse=20, iters_pressure=20, cell_size=1.0 / n, raw_viscosity=1e-4) as s:
        s.set_obstacles(obs)
        for b in (0, 1, 2):
            s.set_field("density", x); s.op_set_bnd("density", b)
            assert_exact(s.get_field("density"), g[f"boundary_b{b}"], f"golden boundary b{b}")
        for tag in ("small", "large"):
            diff, dt = (float(v) for v in g[f"diff_{tag}"])
            a = f32(f32(f32(f32(dt) * f32(diff)) * f32(n - 2)) * f32(n - 2)); c = f32(f32(1) + f32(6) * a)
            for b in (0, 1, 2):
                s.set_field("vx", x)
                s.op_smooth("vx0", "vx", b, a, c, 20)
                assert_exact(s.get_field("vx0"), g[f"smooth_{tag}_b{b}"], f"golden smooth {tag} b{b}")
                s.op_diffuse("vx0", "vx", b, diff, dt)
                assert_exact(s.get_field("vx0"), g[f"diffuse_{tag}_b{b}"], f"golden diffuse {tag} b{b}")
        for b in (0, 1, 2):
            s.set_field("vx", g["ls_rhs"]); s.set_field("vy0", g["ls_guess"])
            s.op_lin_solve("vy0", "vx", b, 0.37, 1 + 6 * 0.37, 7)
            assert_exact(s.get_field("vy0"), g[f"linsolve_b{b}"], f"golden linsolve b{b}")
        s.set_field("vx", g["vx"]); s.set_field("vy", g["vy"])
        s.op_project(False)
        assert_exact(s.get_field("vx"), g["proj_vx"], "golden project vx")
        assert_exact(s.get_field("vy"), g["proj_vy"], "golden project vy")
        assert_exact(s.get_field("pressure"), g["proj_p"], "golden project p")
        dt = float(g["adv_dt"])
        s.set_field("vx", g["vx"]); s.set_field("vy", g["vy"]); s.set_field("vx0", x)
        for b in (0, 1, 2):
            s.op_advect("density", "vx0", b, dt)
            assert_exact(s.get_field("density"), g[f"advect_b{b}"], f"golden advect b{b}")
        s.set_field("vx", g["vbig"])
        s.op_advect("density", "vx0", 0, dt)
        assert_exact(s.get_field("density"), g["advect_clamped"], "golden advect clamped")
        s.set_field("vx", g["vx"]); s.set_field("vy", g["vy"])
        s.op_enforce_obstacles()
        assert_close(s.get_field("vx"), g["enf_vx"], 2e-7, "golden enforce vx")
        assert_close(s.get_field("vy"), g["enf_vy"], 2e-7, "golden enforce vy")


def case_golden_trajectory(lib, name, use_graph=False):
    g = np.load(os.path.join(GOLDEN, name))
    n, obs = int(g["n"]), g["obs"]
    dt, visc, diff, cell, rawv = (float(v) for v in g["params"])
    has_obst = bool(obs.any())
    with make_solver(lib, n, n, 1, iters_diffuse=20, iters_pressure=20, enable_obstacle=has_obst, cell_size=cell,
                     raw_viscosity=rawv, use_cuda_graph=use_graph) as s:
        s.set_obstacles(obs)
        s.set_field("vx", g["init_vx"]); s.set_field("vy", g["init_vy"])
        src = g["sources"]
        for step in range(1, int(g["steps"]) + 1):
            s.add_source_cells(src[:, 0].copy(), src[:, 1].copy(), None, density=src[:, 2].copy(), ax=src[:, 3].copy(), ay=src[:, 4].copy())
            s.step(dt, visc, diff)
            if f"step{step}_density" in g:
                for k in ("density", "vx", "vy", "pressure"):
                    if has_obst:
                        assert_close(s.get_field(k), g[f"step{step}_{k}"], 1e-6 * step, f"golden trajectory step {step} {k}")
                    else:
                        assert_exact(s.get_field(k), g[f"step{step}_{k}"], f"golden trajectory step {step} {k}")


def case_async_readback(lib, nx=16, ny=12, nz=9):
    """fs_get_field_async + fs_wait_transfers returns what fs_get_field returns, also when a step is enqueued
    between the request and the wait (the snapshot is taken in stream order)."""
    rng = np.random.default_rng(21)
    shape = shape_of(nx, ny, nz)
    with make_solver(lib, nx, ny, nz, iters_diffuse=3, iters_pressure=4, enable_obstacle=False) as s:
        for n in ("density", "vx", "vy") + (("vz",) if nz > 1 else ()):
            s.set_field(n, rnd(shape, rng))
        s.step(0.05, 1e-3, 1e-3)
        want_d, want_p = s.get_field("density"), s.get_field("pressure")
        out_d, out_p = np.empty(shape, f32), np.empty(shape, f32)
        s.get_field_async("density", out_d)
        s.get_field_async("pressure", out_p)
        s.step(0.05, 1e-3, 1e-3)            # must not disturb the snapshots
        s.wait_transfers()
        assert_exact(out_d, want_d, "async density")
        assert_exact(out_p, want_p, "async pressure")
        s.get_field_async("density", out_d)  # second use of the same slot
        s.wait_transfers()
        assert_exact(out_d, s.get_field("density"), "async density, second transfer")


def case_visualize(lib, O, nx=24, ny=20, nz=1):
    """fs_render_rgba (UpdateVisualizationJob on the device) vs the oracle's restatement, every colour mode,
    obstacles, source marker, 2D and a 3D slice.  Bit exact (mul/add/div only)."""
    rng = np.random.default_rng(31)
    shape = shape_of(nx, ny, nz)
    V = pkg().native.FsVisParams
    with make_solver(lib, nx, ny, nz) as s:
        d = (rng.random(shape, dtype=f32) * f32(300)).astype(f32)
        p = ((rng.random(shape, dtype=f32) - f32(0.5)) * f32(250)).astype(f32)
        m = (rng.random(shape) < 0.1).astype(np.uint8)
        s.set_obstacles(m); s.set_field("density", d); s.set_field("pressure", p)
        for mode in range(5):
            for marker in (0, 1):
                v = V.reference_defaults(nx, mode)
                v.enable_custom_source = marker
                v.source_x, v.source_y = 0.3 * nx, 0.6 * ny
                v.colour_intensity = 0.004
                if mode == 1:   # three gradient keys, unevenly spaced
                    v.gradient_key_count = 3
                    v.gradient_colors[1][:] = (0.2, 0.9, 0.1, 0.5)
                    v.gradient_colors[2][:] = (1.0, 0.0, 0.0, 1.0)
                    v.gradient_times[1], v.gradient_times[2] = 0.35, 1.0
                k = 0
                if nz > 1:
                    k = nz // 2
                    v.z_slice = k
                got = s.render_rgba(v)
                want = O.visualize(d[k] if nz > 1 else d, p[k] if nz > 1 else p, m[k] if nz > 1 else m, v)
                assert_exact(got, want, f"visualize mode={mode} marker={marker} nz={nz}")
        if nz > 1:
            v = V.reference_defaults(nx, 0)
            v.z_slice = nz + 3
            try:
                s.render_rgba(v)
                raise AssertionError("z_slice outside the grid must be rejected")
            except pkg().FluidSolverError:
                pass


def case_streamlines(lib, O, nx=40, ny=32, nz=1):
    """fs_streamlines (StreamlineCalculationJob + StreamlineDrawJob on the device) vs the oracle.  The glyph validity
    pattern and start points are exact; end points involve atan2/cos/sin and agree to 1e-4 pixels."""
    rng = np.random.default_rng(41)
    shape = shape_of(nx, ny, nz)
    with make_solver(lib, nx, ny, nz) as s:
        ux, uy = rnd(shape, rng, 3.0), rnd(shape, rng, 3.0)
        ux[np.abs(ux) < 0.3] = 0; uy[np.abs(uy) < 0.3] = 0      # some glyphs fall under the 0.01 magnitude cut
        m = (rng.random(shape) < 0.15).astype(np.uint8)
        s.set_obstacles(m); s.set_field("vx", ux); s.set_field("vy", uy)
        k = nz // 2 if nz > 1 else 0
        pl = (lambda a: a[k]) if nz > 1 else (lambda a: a)
        for skip, scale in ((1, 1.0), (3, 1.0), (4, 2.5), (7, 10.0)):
            got = s.streamlines(skip, scale, k)
            want = O.streamlines(pl(ux), pl(uy), pl(m), skip, scale)
            assert got.shape == want.shape
            assert_exact(got[:, :2], want[:, :2], f"streamline starts skip={skip}")
            assert np.array_equal(got[:, 2] < 0, want[:, 2] < 0), f"streamline validity pattern skip={skip}"
            err = float(np.abs(got[:, 2:] - want[:, 2:]).max()) if got.size else 0.0
            assert err <= 1e-4, f"streamline ends skip={skip}: {err}"
