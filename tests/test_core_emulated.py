"""CPU tier: the solver core (csrc/fs_core.h orchestration + csrc/fs_cellops.cuh per-cell functions +
the C ABI) compiled with a plain-loop executor (tests/host_emul) against the oracle and the golden
fixtures.  Same cases as the GPU tier (tests/parity_cases.py); what this tier cannot see is the CUDA
launch geometry and relax_vec4 -- that is what `-m gpu` is for."""
import pytest

import parity_cases as P

GRIDS = [(12, 10, 1), (16, 16, 1), (3, 3, 1), (12, 10, 9), (8, 8, 8), (3, 3, 3), (5, 4, 3)]


@pytest.mark.parametrize("dims", GRIDS)
def test_set_bnd(emul_lib, oracle, dims):
    P.case_set_bnd(emul_lib, oracle, *dims)


@pytest.mark.parametrize("dims", GRIDS)
@pytest.mark.parametrize("obstacles", [True, False])
def test_smooth_linsolve_diffuse(emul_lib, oracle, dims, obstacles):
    P.case_smooth_and_linsolve(emul_lib, oracle, *dims, obstacles=obstacles)


@pytest.mark.parametrize("dims", GRIDS)
def test_project(emul_lib, oracle, dims):
    P.case_project(emul_lib, oracle, *dims)


@pytest.mark.parametrize("dims", GRIDS)
def test_advect(emul_lib, oracle, dims):
    P.case_advect(emul_lib, oracle, *dims)


@pytest.mark.parametrize("dims", [(12, 10, 1), (12, 10, 9)])
def test_enforce(emul_lib, oracle, dims):
    P.case_enforce(emul_lib, oracle, *dims)


@pytest.mark.parametrize("dims", [(12, 10, 1), (12, 10, 9)])
def test_sources(emul_lib, oracle, dims):
    P.case_sources(emul_lib, oracle, *dims)


@pytest.mark.parametrize("dims,steps,kd,kp", [((16, 16, 1), 3, 20, 20), ((12, 12, 12), 2, 5, 7), ((12, 12, 12), 2, 4, 6)])
@pytest.mark.parametrize("obstacles", [True, False])
def test_steps(emul_lib, oracle, dims, steps, kd, kp, obstacles):
    P.case_steps(emul_lib, oracle, *dims, steps, kd=kd, kp=kp, obstacles=obstacles)


@pytest.mark.parametrize("kd,kp", [(3, 5), (1, 1), (0, 0)])
def test_odd_iteration_counts_rotate_buffer_roles(emul_lib, oracle, kd, kp):
    P.case_steps(emul_lib, oracle, 12, 10, 9, 4, kd=kd, kp=kp, obstacles=False)
    P.case_steps(emul_lib, oracle, 12, 10, 1, 4, kd=kd, kp=kp, obstacles=True)


@pytest.mark.parametrize("name", ["kernels2d_n24.npz", "kernels2d_n30.npz"])
def test_golden_kernels(emul_lib, name):
    P.case_golden_2d(emul_lib, name)


@pytest.mark.parametrize("name", ["traj2d_n32_obst.npz", "traj2d_n32_free.npz"])
def test_golden_trajectory(emul_lib, name):
    P.case_golden_trajectory(emul_lib, name)


def test_red_black_matches_oracle(emul_lib, oracle):
    import numpy as np

    rng = np.random.default_rng(9)
    shape = (10, 9, 12)
    mask = P.random_mask(shape, rng)
    rhs, guess = P.rnd(shape, rng), P.rnd(shape, rng)
    with P.make_solver(emul_lib, 12, 9, 10) as s:
        s.set_obstacles(mask)
        for b in (0, 1, 3):
            s.set_field("vx", rhs); s.set_field("vy0", guess)
            s.op_lin_solve("vy0", "vx", b, 0.3, 2.8, 4, solver_kind=1)
            P.assert_exact(s.get_field("vy0"), oracle.lin_solve(b, guess, rhs, 0.3, 2.8, mask, 4, red_black=True), f"rb b={b}")


def test_async_readback(emul_lib):
    P.case_async_readback(emul_lib)


def test_abi_errors(emul_lib, pkg):
    """Error behaviour of the boundary: negative status + message, never an exception across the ABI."""
    import numpy as np

    with pytest.raises(pkg.FluidSolverError) as e:
        pkg.NativeSolver(2, 8, 1, lib_path=emul_lib)
    assert e.value.code == -1
    with pytest.raises(pkg.FluidSolverError):
        pkg.NativeSolver(8, 8, 2, lib_path=emul_lib)            # nz must be 1 or >= 3
    with pkg.NativeSolver(8, 8, 1, lib_path=emul_lib) as s:
        with pytest.raises(pkg.FluidSolverError):
            s.set_obstacles(np.zeros(5, np.uint8))              # wrong mask size
        with pytest.raises(pkg.FluidSolverError):
            s.get_field("vz")                                   # not allocated in 2D
        with pytest.raises(pkg.FluidSolverError):
            s.op_smooth("vx", "vx", 0, 0.1, 1.6, 2)             # dst == src
        with pytest.raises(pkg.FluidSolverError):
            s.op_set_bnd("density", 3)                          # b = 3 needs a z axis
        assert b"" != s.lib.fs_last_error(s.h)


@pytest.mark.parametrize("dims", [(24, 20, 1), (16, 12, 9)])
def test_visualize(emul_lib, oracle, dims):
    P.case_visualize(emul_lib, oracle, *dims)


@pytest.mark.parametrize("dims", [(40, 32, 1), (24, 20, 9)])
def test_streamlines(emul_lib, oracle, dims):
    P.case_streamlines(emul_lib, oracle, *dims)
