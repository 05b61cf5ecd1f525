"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package (see oracle/fluid_oracle.c header).
The product package ``3dfluidsimulation_b200`` never imports it.

PARITY UNPINNED by the reference: no C# toolchain here and the reference has no tests or golden
vectors; the oracle is pinned by hand-derived known answers and an independent numpy restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfluid_oracle.so")

_F = C.POINTER(C.c_float)
_U8 = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile oracle/*.c with gcc (see oracle/Makefile). Returns the .so path."""
    srcs = [os.path.join(_HERE, f) for f in ("fluid_oracle.c", "ref2d.c", "ref_faithful3d.c", "fluid_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "-s"], check=True)
    return _SO


class _State(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("iters_diffuse", C.c_int), ("iters_pressure", C.c_int), ("red_black", C.c_int),
        ("enable_obstacle", C.c_int), ("cell_size", C.c_float), ("raw_viscosity", C.c_float),
        ("density", _F), ("vx", _F), ("vy", _F), ("vz", _F), ("vx0", _F), ("vy0", _F), ("vz0", _F),
        ("pressure", _F), ("obstacles", _U8),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.fo_cell_index.restype = C.c_longlong
        _lib.fo_cell_index.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    return _lib


def set_threads(n: int | None = None) -> int:
    """OpenMP thread count for the oracle's sweeps (torchrun exports OMP_NUM_THREADS=1, which would make the CPU
    baseline single threaded).  Default: every host core."""
    n = int(n or os.cpu_count() or 1)
    lib()
    C.CDLL("libgomp.so.1").omp_set_num_threads(n)
    return n


def _f(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_F)


def _u(a):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(_U8)


def _dims(shape):
    """numpy shape is (nz, ny, nx) or (ny, nx); returns nx, ny, nz."""
    if len(shape) == 2:
        return shape[1], shape[0], 1
    return shape[2], shape[1], shape[0]


def _cf(v):
    return C.c_float(float(v))


# ---- 3D oracle (nz == 1 reproduces the 2D reference) ------------------------------------------
def set_bnd(b, x, obs):
    nx, ny, nz = _dims(x.shape)
    lib().fo_set_bnd(nx, ny, nz, int(b), _f(x), _u(obs))
    return x


def diffuse_coeffs(n, diff, dt):
    a, c = C.c_float(), C.c_float()
    lib().fo_diffuse_coeffs(int(n), _cf(diff), _cf(dt), C.byref(a), C.byref(c))
    return np.float32(a.value), np.float32(c.value)


def diffuse_smooth(b, x0, a, c, obs, iters):
    nx, ny, nz = _dims(x0.shape)
    x = np.empty_like(x0)
    lib().fo_diffuse_smooth(nx, ny, nz, int(b), _f(x), _f(x0), _cf(a), _cf(c), _u(obs), int(iters))
    return x


def lin_solve(b, x, x0, a, c, obs, iters, red_black=False):
    nx, ny, nz = _dims(x0.shape)
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    fn = lib().fo_lin_solve_rb if red_black else lib().fo_lin_solve
    fn(nx, ny, nz, int(b), _f(x), _f(x0), _cf(a), _cf(c), _u(obs), int(iters))
    return x


def diffuse(b, x0, diff, dt, obs, iters):
    nx, ny, nz = _dims(x0.shape)
    x = np.zeros_like(x0)
    lib().fo_diffuse(nx, ny, nz, int(b), _f(x), _f(x0), _cf(diff), _cf(dt), _u(obs), int(iters))
    return x


def divergence(vx, vy, vz, obs):
    nx, ny, nz = _dims(vx.shape)
    div = np.empty_like(vx)
    lib().fo_divergence(nx, ny, nz, _f(div), _f(vx), _f(vy), _f(vz), _u(obs))
    return div


def project(vx, vy, vz, obs, iters, red_black=False):
    """Returns (vx, vy, vz, p) -- new arrays."""
    nx, ny, nz = _dims(vx.shape)
    vx, vy = vx.copy(), vy.copy()
    vz = vz.copy() if vz is not None else None
    p = np.zeros_like(vx)
    lib().fo_project(nx, ny, nz, _f(vx), _f(vy), _f(vz), _f(p), _u(obs), int(iters), int(bool(red_black)))
    return vx, vy, vz, p


def advect(b, d0, vx, vy, vz, dt, obs):
    nx, ny, nz = _dims(d0.shape)
    d = np.empty_like(d0)
    lib().fo_advect(nx, ny, nz, int(b), _f(d), _f(d0), _f(vx), _f(vy), _f(vz), _cf(dt), _u(obs))
    return d


def enforce_obstacles(vx, vy, vz, obs, cell, rawvisc):
    nx, ny, nz = _dims(vx.shape)
    vx, vy = vx.copy(), vy.copy()
    vz = vz.copy() if vz is not None else None
    lib().fo_enforce_obstacles(nx, ny, nz, _f(vx), _f(vy), _f(vz), _u(obs), _cf(cell), _cf(rawvisc))
    return vx, vy, vz


def visualize(density, pressure, obstacles, vis):
    """fo_visualize on one (ny, nx) plane; `vis` is any ctypes struct with fs_vis_params' layout."""
    ny, nx = density.shape
    out = np.empty((ny, nx, 4), np.float32)
    lib().fo_visualize(nx, ny, _f(density), _f(pressure), _u(obstacles), C.byref(vis), out.ctypes.data_as(_F))
    return out


def streamlines(vx, vy, obstacles, skip, scale):
    ny, nx = vx.shape
    out = np.empty(((nx // skip) * (ny // skip), 4), np.float32)
    lib().fo_streamlines(nx, ny, int(skip), _cf(scale), _f(vx), _f(vy), _u(obstacles), out.ctypes.data_as(_F))
    return out


def cell_index(shape, x, y, z=0.0):
    nx, ny, nz = _dims(shape)
    return int(lib().fo_cell_index(nx, ny, nz, _cf(x), _cf(y), _cf(z)))


@dataclass
class OracleSolver:
    """State holder mirroring FluidSimulation's fields (FluidSim.cs:112-117, :132)."""

    nx: int
    ny: int
    nz: int = 1
    iters_diffuse: int = 20
    iters_pressure: int = 20
    red_black: bool = False
    enable_obstacle: bool = True
    cell_size: float = 1.0 / 128
    raw_viscosity: float = 1e-4
    f: dict = field(default_factory=dict)

    def __post_init__(self):
        shape = (self.ny, self.nx) if self.nz == 1 else (self.nz, self.ny, self.nx)
        for n in ("density", "vx", "vy", "vz", "vx0", "vy0", "vz0", "pressure"):
            self.f[n] = np.zeros(shape, np.float32)
        self.obstacles = np.zeros(shape, np.uint8)

    @property
    def shape(self):
        return self.obstacles.shape

    def _state(self):
        s = _State(self.nx, self.ny, self.nz, self.iters_diffuse, self.iters_pressure, int(self.red_black),
                   int(self.enable_obstacle), self.cell_size, self.raw_viscosity)
        for n in ("density", "vx", "vy", "vz", "vx0", "vy0", "vz0", "pressure"):
            setattr(s, n, _f(self.f[n]))
        s.obstacles = _u(self.obstacles)
        return s

    def add_density(self, x, y, z, amount):
        self.f["density"].reshape(-1)[cell_index(self.shape, x, y, z)] += np.float32(amount)

    def add_velocity(self, x, y, z, ax, ay, az=0.0):
        i = cell_index(self.shape, x, y, z)
        self.f["vx"].reshape(-1)[i] += np.float32(ax)
        self.f["vy"].reshape(-1)[i] += np.float32(ay)
        if self.nz > 1:
            self.f["vz"].reshape(-1)[i] += np.float32(az)

    def add_sources(self, d=None, vx=None, vy=None, vz=None):
        for n, a in (("density", d), ("vx", vx), ("vy", vy), ("vz", vz)):
            if a is not None:
                self.f[n] += a.astype(np.float32).reshape(self.shape)

    def step(self, dt, visc, diff, faithful=False):
        """One Simulate().  faithful=True runs ref_faithful3d.c: same bits, the reference's execution structure
        (static-64 job batches, single-threaded BoundaryJob scans, per-call allocate-and-copy; Jacobi only)."""
        s = self._state()
        if faithful:
            assert not self.red_black, "the reference has no red-black solver"
            lib().rf_step(C.byref(s), _cf(dt), _cf(visc), _cf(diff))
        else:
            lib().fo_step(C.byref(s), _cf(dt), _cf(visc), _cf(diff))

    def metrics(self):
        s = self._state()
        a, b = C.c_float(), C.c_float()
        lib().fo_metrics(C.byref(s), C.byref(a), C.byref(b))
        return a.value, b.value


# ---- literal 2D restatement (ref2d.c) ----------------------------------------------------------
class Ref2D:
    """Thin wrappers over ref2d.c; arrays are (size, size) float32, index [y, x]."""

    @staticmethod
    def boundary(b, x, obs):
        lib().r2_boundary(x.shape[0], int(b), _f(x), _u(obs))
        return x

    @staticmethod
    def diffuse_with_jobs(b, x0, diff, dt, obs, iters=20):
        x = np.zeros_like(x0)
        lib().r2_diffuse_with_jobs(x0.shape[0], int(b), _f(x), _f(x0), _cf(diff), _cf(dt), _u(obs), int(iters))
        return x

    @staticmethod
    def linear_solve_with_jobs(b, x, x0, a, c, obs, iters=20):
        x = x.copy()
        lib().r2_linear_solve_with_jobs(x0.shape[0], int(b), _f(x), _f(x0), _cf(a), _cf(c), _u(obs), int(iters))
        return x

    @staticmethod
    def diffuse(b, x0, diff, dt, obs, iters=20):
        x = np.zeros_like(x0)
        lib().r2_diffuse(x0.shape[0], int(b), _f(x), _f(x0), _cf(diff), _cf(dt), _u(obs), int(iters))
        return x

    @staticmethod
    def project_with_jobs(vx, vy, obs, iters=20):
        vx, vy = vx.copy(), vy.copy()
        p = np.zeros_like(vx)
        lib().r2_project_with_jobs(vx.shape[0], _f(vx), _f(vy), _f(p), _u(obs), int(iters))
        return vx, vy, p

    @staticmethod
    def advect_with_jobs(b, d0, vx, vy, dt, obs):
        d = np.zeros_like(d0)
        lib().r2_advect_with_jobs(d0.shape[0], int(b), _f(d), _f(d0), _f(vx), _f(vy), _cf(dt), _u(obs))
        return d

    @staticmethod
    def enforce_obstacles(vx, vy, obs, cell, visc):
        vx, vy = vx.copy(), vy.copy()
        lib().r2_enforce_obstacles(vx.shape[0], _f(vx), _f(vy), _u(obs), _cf(cell), _cf(visc))
        return vx, vy

    @staticmethod
    def simulate(state: dict, obs, dt, visc, diff, enable_obstacle, cell, rawvisc, iters=20):
        """state: dict with density, vx, vy, vx0, vy0, pressure (modified in place)."""
        n = obs.shape[0]
        lib().r2_simulate(n, _f(state["density"]), _f(state["vx"]), _f(state["vy"]), _f(state["vx0"]),
                          _f(state["vy0"]), _f(state["pressure"]), _u(obs), _cf(dt), _cf(visc), _cf(diff),
                          int(bool(enable_obstacle)), _cf(cell), _cf(rawvisc), int(iters))
