"""GPU tier (`-m gpu`): the parity tests proper.  Everything goes through the C ABI of the real
libfluidsolver.so (CUDA kernels on cuda:0) and is compared with the oracle on the same seeded inputs,
with the committed golden fixtures, and -- at BASELINE.json's full sizes -- through size-independent
properties."""
import os

import numpy as np
import pytest

import parity_cases as P

pytestmark = pytest.mark.gpu

# nx % 4 == 0 takes the float4 z-marching kernel (relax_vec4); other nx take the per-cell kernels
GRIDS = [(16, 12, 1), (64, 48, 1), (30, 30, 1), (4, 4, 1), (16, 12, 9), (64, 40, 35), (30, 17, 11), (4, 3, 3),
         (8, 3, 3), (132, 20, 70)]


@pytest.fixture(scope="module")
def lib(cuda_lib):
    import torch

    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    return cuda_lib


@pytest.mark.parametrize("dims", GRIDS)
def test_set_bnd(lib, oracle, dims):
    P.case_set_bnd(lib, oracle, *dims)


@pytest.mark.parametrize("dims", GRIDS)
@pytest.mark.parametrize("obstacles", [True, False])
def test_smooth_linsolve_diffuse(lib, oracle, dims, obstacles):
    P.case_smooth_and_linsolve(lib, oracle, *dims, obstacles=obstacles)


@pytest.mark.parametrize("dims", GRIDS)
def test_project(lib, oracle, dims):
    P.case_project(lib, oracle, *dims)


@pytest.mark.parametrize("dims", GRIDS)
@pytest.mark.parametrize("coherent", [False, True])
def test_advect(lib, oracle, dims, coherent):
    P.case_advect(lib, oracle, *dims, coherent=coherent)


def test_advect_vec4_matches_per_cell_kernel_256(lib, monkeypatch):
    """The float4 advect kernel (shared-displacement fast path + per-cell gathers) against the per-cell kernel on a 256^3
    plume-like state: bit identical."""
    rng = np.random.default_rng(31)
    n = 256
    shape = (n, n, n)
    f = {nme: P.rnd(shape, rng, 0.02) for nme in ("vx0", "vy0", "vz0")}
    f["vy0"][96:160, 40:120, 96:160] += np.float32(3.0)          # a fast core, sub-cell motion elsewhere
    f["vx0"][:, :, : n // 2] = np.abs(f["vx0"][:, :, : n // 2])
    f["vz0"][: n // 3] = 0
    mask = P.random_mask(shape, rng, 0.01)
    out = {}
    for tag, env in (("vec4", "0"), ("cell", "1")):
        monkeypatch.setenv("FS_NO_ADVECT_VEC4", env)
        with P.make_solver(lib, n, n, n) as s:
            s.set_obstacles(mask)
            for nme, a in f.items():
                s.set_field(nme, a)
            s.op_advect_velocity(0.025)
            s.set_field("vx0", f["vx0"])
            s.op_advect("density", "vx0", 0, 0.025)
            out[tag] = {nme: s.get_field(nme) for nme in ("vx", "vy", "vz", "density")}
    for nme in out["vec4"]:
        P.assert_exact(out["vec4"][nme], out["cell"][nme], f"advect vec4 vs per-cell {nme}")


@pytest.mark.parametrize("dims", [(16, 12, 1), (64, 40, 35)])
def test_enforce(lib, oracle, dims):
    P.case_enforce(lib, oracle, *dims)


@pytest.mark.parametrize("dims", [(16, 12, 1), (16, 12, 9)])
def test_sources(lib, oracle, dims):
    P.case_sources(lib, oracle, *dims)


@pytest.mark.parametrize("name", ["kernels2d_n24.npz", "kernels2d_n30.npz"])
def test_golden_kernels(lib, name):
    P.case_golden_2d(lib, name)


@pytest.mark.parametrize("name", ["traj2d_n32_obst.npz", "traj2d_n32_free.npz"])
@pytest.mark.parametrize("graph", [False, True])
def test_golden_trajectory(lib, name, graph):
    P.case_golden_trajectory(lib, name, use_graph=graph)


@pytest.mark.parametrize("kd,kp", [(3, 5), (1, 1), (2, 3)])
def test_graph_replay_with_rotating_buffer_roles(lib, oracle, kd, kp):
    """Odd iteration counts leave the ping-pong buffers swapped after a step, so consecutive steps need different
    captured graphs (keyed by the role configuration); 6 steps must still follow the oracle."""
    P.case_steps(lib, oracle, 16, 12, 10, 6, kd=kd, kp=kp, obstacles=False, use_graph=True)
    P.case_steps(lib, oracle, 16, 12, 1, 6, kd=kd, kp=kp, obstacles=True, use_graph=True)


@pytest.mark.parametrize("obstacles", [True, False])
def test_config1_32cube_trajectory(lib, oracle, obstacles):
    """BASELINE config 1 AS WRITTEN: 32^3 smoke plume, 100 steps, K_d = K_p = 20, dt = 0.1 (source velocity 1 keeps
    CFL <~ 3).  Without obstacles every field is bit exact after all 100 steps; with the
    obstacle sphere the drag's exp() allows 1e-6 * max|field| per step."""
    P.case_steps(lib, oracle, 32, 32, 32, 100, kd=20, kp=20, obstacles=obstacles, use_graph=True)


def test_config3_256cube_step_vs_oracle(lib, oracle):
    """BASELINE config 3 size: 256^3, K_d = K_p = 20, ONE step through the ABI in Update() order (sources, then
    fs_step), obstacle sphere, compared field by field with the CPU oracle (~2 s of host time)."""
    n = 256
    P.case_steps(lib, oracle, n, n, n, 1, kd=20, kp=20, obstacles=True, use_graph=True, dt=0.1 * 128 / n,
                 vsrc=2.5 / (0.1 * 128 / n * (n - 2)))


def test_config4_512cube_step_vs_oracle(lib, oracle):
    """BASELINE config 4 size: 512^3, K_d = 20, K_p = 80, obstacle sphere, ONE step vs the CPU oracle (the oracle
    needs ~20-40 s and ~6 GB on the GPU box's host)."""
    n = 512
    oracle.set_threads(os.cpu_count())
    P.case_steps(lib, oracle, n, n, n, 1, kd=20, kp=80, obstacles=True, use_graph=False, dt=0.1 * 128 / n,
                 vsrc=2.5 / (0.1 * 128 / n * (n - 2)))


def _residual_inf(p, div, mask):
    """max |6 p - sum_nb p - div| over interior fluid cells, relative to max |div| (the fixed point of the pressure
    iteration, FluidSim.cs:1581-1582 with the six 3D neighbours)."""
    c = p[1:-1, 1:-1, 1:-1].astype(np.float64)
    nb = (p[1:-1, 1:-1, 2:].astype(np.float64) + p[1:-1, 1:-1, :-2] + p[1:-1, 2:, 1:-1] + p[1:-1, :-2, 1:-1]
          + p[2:, 1:-1, 1:-1] + p[:-2, 1:-1, 1:-1])
    r = np.abs(6.0 * c - nb - div[1:-1, 1:-1, 1:-1])
    r[mask[1:-1, 1:-1, 1:-1] != 0] = 0.0
    return float(r.max()) / max(float(np.abs(div).max()), 1e-30), float(np.sqrt((r * r).mean()))


@pytest.mark.parametrize("n,iters", [(128, 100), (160, 40)])
def test_red_black_at_scale_on_cuda_output(lib, oracle, n, iters):
    """BASELINE config 5's solver at sizes where the float4 / fused kernels run (K = 100 red-black pressure iterations).
    (1) the CUDA pressure equals the oracle's red-black bit for bit;
    (2) the documented Jacobi <-> red-black equivalence, evaluated ON THE CUDA OUTPUT of
        both solver kinds after the same iteration count:
          * RMS residual of 6p - sum_nb p = div: red-black <= Jacobi;
          * distance to the converged solution (3000 oracle red-black iterations), max norm and RMS: red-black <= Jacobi;
          * max-norm residual: right after the black half sweep the whole residual sits on the red cells, so it is
            only required to stay within 2x of Jacobi's (measured ~1.4x at 128^3)."""
    rng = np.random.default_rng(17)
    shape = (n, n, n)
    zz, yy, xx = np.ogrid[:n, :n, :n]
    mask = (((xx - 0.5 * n) ** 2 + (yy - 0.5 * n) ** 2 + (zz - 0.5 * n) ** 2) < (0.1 * n) ** 2).astype(np.uint8)
    v = [P.rnd(shape, rng, 1.0) for _ in range(3)]
    out = {}
    for kind in (0, 1):
        with P.make_solver(lib, n, n, n, iters_pressure=iters, solver_kind=kind) as s:
            s.set_obstacles(mask)
            for name, a in zip(("vx", "vy", "vz"), v):
                s.set_field(name, a)
            s.op_project(False)
            out[kind] = (s.get_field("pressure"), s.get_field("divergence"))
    want = oracle.project(v[0], v[1], v[2], mask, iters, red_black=True)[3]
    P.assert_exact(out[1][0], want, f"red-black pressure {n}^3 K={iters}")
    if n != 128:
        return
    rb_inf, rb_rms = _residual_inf(out[1][0], out[1][1], mask)
    ja_inf, ja_rms = _residual_inf(out[0][0], out[0][1], mask)
    assert rb_rms <= ja_rms and rb_inf <= 2.0 * ja_inf, (rb_inf, ja_inf, rb_rms, ja_rms)
    oracle.set_threads(os.cpu_count())
    star = oracle.lin_solve(0, np.zeros(shape, np.float32), out[1][1], 1.0, 6.0, mask, 3000, red_black=True).astype(np.float64)
    err = {k: (out[k][0] - star)[1:-1, 1:-1, 1:-1] for k in out}
    assert np.abs(err[1]).max() <= np.abs(err[0]).max()
    assert np.sqrt((err[1] ** 2).mean()) <= np.sqrt((err[0] ** 2).mean())


def test_reset_with_obstacles_and_graph(lib):
    """fs_reset after a captured step with interior obstacles: the next step must not replay the old graph (it
    references the freed obstacle list and the mirror / drag decisions): compare with a fresh handle."""
    rng = np.random.default_rng(23)
    shape = (12, 16, 16)
    mask = P.random_mask(shape, rng, 0.08)
    f = {nme: P.rnd(shape, rng, 0.5) for nme in ("density", "vx", "vy", "vz")}
    kw = dict(iters_diffuse=4, iters_pressure=6, enable_obstacle=True, use_cuda_graph=True)

    def prime(s, with_mask):
        if with_mask:
            s.set_obstacles(mask)
        for nme, a in f.items():
            s.set_field(nme, a)
        for _ in range(2):
            s.step(0.05, 1e-3, 1e-3)

    with P.make_solver(lib, 16, 16, 12, **kw) as s, P.make_solver(lib, 16, 16, 12, **kw) as fresh:
        prime(s, True)
        s.reset()
        prime(s, False)          # same (dt, visc, diff) and buffer roles as the captured graph
        prime(fresh, False)
        for nme in ("density", "vx", "vy", "vz", "pressure"):
            P.assert_exact(s.get_field(nme), fresh.get_field(nme), f"after reset: {nme}")
        s.reset()
        prime(s, True)           # and obstacles again after a reset
        fresh.reset()
        prime(fresh, True)
        for nme in ("density", "vx", "vy", "vz", "pressure"):
            P.assert_exact(s.get_field(nme), fresh.get_field(nme), f"after second reset: {nme}")


def test_config2_128cube_single_step(lib, oracle):
    """BASELINE config 2: 128^3, 40 pressure iterations, field-by-field check after 1 and 2 steps."""
    P.case_steps(lib, oracle, 128, 128, 128, 2, kd=20, kp=40, obstacles=True, use_graph=True, dt=0.1 * 128 / 128)


def test_scene_configs_2d(lib, oracle):
    """The two scene parameter sets (SampleScene.unity): 192^2 and 128^2, 3 steps each."""
    P.case_steps(lib, oracle, 192, 192, 1, 3, obstacles=True, dt=0.1 * 128 / 192, visc=1e-4 / 3, diff=1e-4 / 3)
    P.case_steps(lib, oracle, 128, 128, 1, 3, obstacles=True)


def test_async_readback(lib):
    P.case_async_readback(lib)
    P.case_async_readback(lib, 64, 40, 35)


def test_red_black_matches_oracle(lib, oracle):
    rng = np.random.default_rng(9)
    shape = (10, 9, 12)
    mask = P.random_mask(shape, rng)
    rhs, guess = P.rnd(shape, rng), P.rnd(shape, rng)
    with P.make_solver(lib, 12, 9, 10) as s:
        s.set_obstacles(mask)
        for b in (0, 1, 3):
            s.set_field("vx", rhs); s.set_field("vy0", guess)
            s.op_lin_solve("vy0", "vx", b, 0.3, 2.8, 4, solver_kind=1)
            P.assert_exact(s.get_field("vy0"), oracle.lin_solve(b, guess, rhs, 0.3, 2.8, mask, 4, red_black=True), f"rb b={b}")


def test_vector_and_scalar_kernels_agree(lib, oracle):
    """relax_vec4 against the per-cell kernels (FS_FORCE_GENERIC=1) on a grid the oracle would take
    minutes for: 256^3, full step, bit-identical fields."""
    def run(force, no_pair=False):
        os.environ["FS_FORCE_GENERIC"] = "1" if force else "0"
        os.environ["FS_PAIR"] = "0" if no_pair else "1"      # 1: fuse wherever the kernel applies (the default policy is auto)
        try:
            s, _ = None, None
            pk = P.pkg()
            s = pk.NativeSolver(256, 256, 256, iters_diffuse=4, iters_pressure=6, enable_obstacle=True, lib_path=lib)
            rng = np.random.default_rng(11)
            s.set_obstacles(P.random_mask((256, 256, 256), rng, 0.02))
            for n in ("density", "vx", "vy", "vz"):
                s.set_field(n, P.rnd((256, 256, 256), rng, 0.5))
            s.step(0.02, 1e-3, 1e-3)
            out = {n: s.get_field(n) for n in ("density", "vx", "vy", "vz", "pressure")}
            s.close()
            return out
        finally:
            os.environ.pop("FS_FORCE_GENERIC", None)
            os.environ.pop("FS_PAIR", None)
    a, b, c = run(False), run(True), run(False, no_pair=True)
    for n in a:
        P.assert_exact(a[n], b[n], f"fused/vec4 vs per-cell {n}")
        P.assert_exact(a[n], c[n], f"fused two-stage sweeps vs single vec4 sweeps {n}")


@pytest.mark.parametrize("kind", ["jacobi", "smooth", "rb"])
@pytest.mark.parametrize("dims", [(64, 40, 35), (132, 20, 70), (8, 3, 3), (4, 3, 3), (256, 100, 9), (16, 12, 9)])
def test_fused_pair_kernel_vs_oracle(lib, oracle, kind, dims, monkeypatch):
    """The fused two-stage sweep (two Jacobi / smoother iterations or both red-black colours per pass) against the
    oracle for every field kind b it is used for (b = 0 with obstacles; b = 1, 2, 3 without), odd iteration counts
    (pair + single mix) and tile shapes that leave partial tiles in x, y and z."""
    nx, ny, nz = dims
    rng = np.random.default_rng(5)
    shape = (nz, ny, nx)
    x0, guess = P.rnd(shape, rng), P.rnd(shape, rng)
    a, c = np.float32(0.37), np.float32(1 + 6 * 0.37)
    monkeypatch.setenv("FS_PAIR", "1")   # read at fs_create: fuse wherever the kernel applies (default: auto policy)
    for obstacles, bs in ((True, (0,)), (False, (0, 1, 2, 3))):
        mask = P.random_mask(shape, rng) if obstacles else np.zeros(shape, np.uint8)
        with P.make_solver(lib, nx, ny, nz) as s:
            s.set_obstacles(mask)
            for b in bs:
                for it in (2, 3, 6, 7):
                    s.set_field("vx", x0); s.set_field("vy0", guess); s.set_field("vx0", P.rnd(shape, rng))
                    if kind == "smooth":
                        s.op_smooth("vx0", "vx", b, a, c, it)
                        P.assert_exact(s.get_field("vx0"), oracle.diffuse_smooth(b, x0, a, c, mask, it), f"pair smooth b={b} it={it}")
                    else:
                        rb = kind == "rb"
                        s.op_lin_solve("vy0", "vx", b, a, c, it, solver_kind=1 if rb else 0)
                        P.assert_exact(s.get_field("vy0"), oracle.lin_solve(b, guess, x0, a, c, mask, it, red_black=rb),
                                       f"pair {kind} b={b} it={it}")


def test_full_size_properties_512(lib):
    """BASELINE config 4 size (512^3, K_p = 80): properties that need no oracle run.
    (1) a uniform density with zero velocity stays exactly uniform through a full step (3D stencil
    preserves constants; advect with V = 0 is the identity); (2) with zero velocity the pressure and
    divergence are identically zero; (3) mirror symmetry: a y-symmetric plume stays y... x-symmetric."""
    pk = P.pkg()
    n = 512
    s = pk.NativeSolver(n, n, n, iters_diffuse=20, iters_pressure=80, enable_obstacle=False, lib_path=lib, use_cuda_graph=True)
    d = np.full((n, n, n), 2.5, np.float32)
    s.set_field("density", d)
    s.step(0.025, 1e-4, 1e-4)
    got = s.get_field("density")
    assert float(np.abs(got - 2.5).max()) <= 2.5 * 3e-6          # (1+6a)/(1+6a) rounding only
    assert not s.get_field("pressure").any() and not s.get_field("divergence").any()
    # x-mirror symmetry of a centred plume: density(x) == density(n-1-x), vx antisymmetric
    s.reset()
    zz, yy, xx = np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij")
    blob = np.zeros((n, n, n), np.float32)
    c = n // 2
    blob[c - 4:c + 4, 100:108, c - 4:c + 4] = 50.0
    vy = np.zeros_like(blob); vy[c - 4:c + 4, 100:108, c - 4:c + 4] = 2.0
    s.set_field("density", blob); s.set_field("vy", vy)
    s.step(0.025, 1e-4, 1e-4)
    dd, vx = s.get_field("density"), s.get_field("vx")
    assert dd.any()
    # not bit exact: the back-trace x = i - dt0*v rounds differently at i and n-1-i (ulp(512) = 6e-5 in the
    # interpolation weights), so the mirror images agree to ~1e-4 of the field maximum
    np.testing.assert_allclose(dd, dd[:, :, ::-1], rtol=0, atol=5e-4 * float(dd.max()))
    np.testing.assert_allclose(vx, -vx[:, :, ::-1], rtol=0, atol=5e-4 * max(float(np.abs(vx).max()), 1e-20))
    s.close()


@pytest.mark.parametrize("divisor", [6.0, 1.0 + 6 * 0.009, 1.0 + 6 * 0.37, 1.0001, 3.0, 1023.9999, 1.9999999, 0.7])
def test_constant_divisor_division_is_ieee_exact(lib, divisor):
    """fs_div (reciprocal + two FMA corrections, no slow path) against __fdiv_rn for ALL 2^32 numerator bit
    patterns -- zeros, denormals, infinities and NaNs included -- for the divisors the step uses (6, 1+6a)
    and a few awkward ones."""
    with P.make_solver(lib, 8, 8, 1) as s:
        assert s.selftest_division(np.float32(divisor), 0, 1 << 32) == 0


def test_denormal_and_zero_fields_are_exact(lib, oracle):
    """Fields that decay through the denormal range (what a plume's far field does) stay bit exact."""
    rng = np.random.default_rng(12)
    shape = (20, 24, 32)
    x0 = (P.rnd(shape, rng) * np.float32(1e-36)).astype(np.float32)
    x0[:, :, 16:] = 0
    x0[3, 3, 3] = -0.0
    mask = np.zeros(shape, np.uint8)
    with P.make_solver(lib, 32, 24, 20) as s:
        s.set_obstacles(mask)
        s.set_field("vx", x0)
        s.op_diffuse("vx0", "vx", 0, 1e-4, 0.1)
        P.assert_exact(s.get_field("vx0"), oracle.diffuse(0, x0, 1e-4, 0.1, mask, 20), "denormal diffuse")


@pytest.mark.parametrize("dims", [(24, 20, 1), (192, 192, 1), (64, 40, 35)])
def test_visualize(lib, oracle, dims):
    """next row N2: UpdateVisualizationJob on the device."""
    P.case_visualize(lib, oracle, *dims)


@pytest.mark.parametrize("dims", [(40, 32, 1), (192, 192, 1), (64, 40, 35)])
def test_streamlines(lib, oracle, dims):
    """next row N3: streamline glyph segments on the device."""
    P.case_streamlines(lib, oracle, *dims)


def test_cpp_host_on_gpu(lib, tmp_path):
    """The compiled-language host mirror (include/fluid_simulation.hpp) linked against the product library gives
    the same fields as the Python mirror on the same GPU."""
    import subprocess

    from conftest import ROOT

    exe = str(tmp_path / "host_demo_cuda")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "host_demo.cpp"), "-o", exe, "-L", os.path.dirname(lib), "-lfluidsolver",
                    f"-Wl,-rpath,{os.path.dirname(lib)}", "-Wl,-rpath-link,/usr/local/cuda/lib64"], check=True)
    for size, frames, depth, shape in ((64, 3, 1, 0), (48, 2, 16, 2)):
        out = subprocess.run([exe, str(size), str(frames), str(depth), str(shape)], capture_output=True, text=True, check=True).stdout
        got = {l.split()[0]: [float(v) for v in l.split()[1:]] for l in out.strip().splitlines()}
        sim = P.pkg().FluidSimulation(size=size, depth=depth, obstacleShape=["Circle", "Rectangle", "Airfoil"][shape], lib_path=lib,
                                      use_cuda_graph=False)
        sim.enableCustomSource = True; sim.sourceEmitsVelocity = True
        sim.sourceDirection = 90.0; sim.sourceRadius = 2.0; sim.sourcePositionY = 0.2
        for f in range(frames):
            sim.AddForceToArea((np.float32(0.3) * np.float32(size) + np.float32(f), np.float32(0.6) * np.float32(size)), (2.5, -1.25), 3.0)
            sim.Update()
        assert int(sim.obstacles.sum()) == int(got["obstacle_cells"][0])
        for name in ("density", "vx", "vy", "pressure"):
            a = sim.field(name).astype(np.float64)
            np.testing.assert_allclose(got[name], [a.sum(), (a * a).sum()], rtol=2e-5, atol=1e-12, err_msg=name)
        vis = P.pkg().native.FsVisParams.reference_defaults(size, 2)
        rgba = sim.UpdateVisualization(vis).astype(np.float64)
        np.testing.assert_allclose(got["rgba"], [rgba.sum(), (rgba * rgba).sum()], rtol=2e-5, err_msg="rgba")
        assert int(sim.DrawStreamlines().sum()) == int(got["streamline_pixels"][0])
        sim.close()
