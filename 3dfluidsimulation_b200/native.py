"""ctypes binding of include/fluidsolver.h -- the same entry points the C# P/Invoke layer
(Assets/Plugin/NativeFluidSolver.cs) binds.

There is no CPU fallback: if libfluidsolver.so is missing, cannot be loaded, or no CUDA device is
present, every constructor raises.  The library itself links only libcudart.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libfluidsolver.so")

FS_ABI_VERSION = 1
FS_IPC_BLOB_BYTES = 1024

# fs_field
DENSITY, VX, VY, VZ, VX0, VY0, VZ0, PRESSURE, DIVERGENCE = range(9)
FIELD_IDS = {"density": DENSITY, "vx": VX, "vy": VY, "vz": VZ, "vx0": VX0, "vy0": VY0, "vz0": VZ0,
             "pressure": PRESSURE, "divergence": DIVERGENCE}
JACOBI, RED_BLACK = 0, 1

EXPORTS = [
    "fs_abi_version", "fs_create", "fs_destroy", "fs_reset", "fs_last_error", "fs_slab_range",
    "fs_set_obstacles", "fs_slab_halo_range", "fs_set_obstacles_slab", "fs_build_obstacles", "fs_get_obstacles", "fs_add_density", "fs_add_velocity", "fs_add_source_cells", "fs_add_sources",
    "fs_step", "fs_sync", "fs_get_field", "fs_set_field", "fs_get_field_async", "fs_wait_transfers", "fs_get_metrics",
    "fs_render_rgba", "fs_streamlines",
    "fs_op_set_bnd", "fs_op_diffuse", "fs_op_smooth", "fs_op_lin_solve", "fs_op_project", "fs_op_advect",
    "fs_op_advect_velocity", "fs_op_enforce_obstacles",
    "fs_timer_start", "fs_timer_stop", "fs_launch_count", "fs_bench_sweep",
    "fs_selftest_division", "fs_halo_export", "fs_halo_connect",
]


class FsParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("iters_diffuse", C.c_int32), ("iters_pressure", C.c_int32), ("solver_kind", C.c_int32),
        ("enable_obstacle", C.c_int32), ("cell_size", C.c_float), ("raw_viscosity", C.c_float),
        ("device_id", C.c_int32), ("slab_rank", C.c_int32), ("slab_count", C.c_int32),
        ("use_cuda_graph", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class FsVisParams(C.Structure):
    """fs_vis_params (include/fluidsolver.h): parameters of the reference's UpdateVisualizationJob."""
    _fields_ = [
        ("color_mode", C.c_int32), ("visualize_source_position", C.c_int32), ("enable_custom_source", C.c_int32),
        ("gradient_key_count", C.c_int32), ("z_slice", C.c_int32),
        ("source_x", C.c_float), ("source_y", C.c_float), ("visual_marker_radius", C.c_float),
        ("colour_intensity", C.c_float),
        ("medium_density_threshold", C.c_float), ("high_density_threshold", C.c_float),
        ("low_pressure_threshold", C.c_float), ("high_pressure_threshold", C.c_float),
        ("fluid_color", C.c_float * 4), ("obstacle_color", C.c_float * 4), ("source_position_color", C.c_float * 4),
        ("low_density_color", C.c_float * 4), ("medium_density_color", C.c_float * 4), ("high_density_color", C.c_float * 4),
        ("low_pressure_color", C.c_float * 4), ("neutral_pressure_color", C.c_float * 4), ("high_pressure_color", C.c_float * 4),
        ("gradient_colors", (C.c_float * 4) * 8), ("gradient_times", C.c_float * 8),
    ]

    @classmethod
    def reference_defaults(cls, size, mode=0):
        """The inspector defaults of FluidSim.cs:57-83 (Unity's Color constants spelled out)."""
        v = cls()
        v.color_mode, v.visualize_source_position, v.enable_custom_source = mode, 1, 0
        v.source_x, v.source_y, v.visual_marker_radius = 0.5 * size, 0.5 * size, 3.0
        v.colour_intensity = 1.0
        v.medium_density_threshold, v.high_density_threshold = 50.0, 200.0
        v.low_pressure_threshold, v.high_pressure_threshold = -50.0, 50.0
        white, blue, green, red, gray, yellow = (1, 1, 1, 1), (0, 0, 1, 1), (0, 1, 0, 1), (1, 0, 0, 1), (0.5, 0.5, 0.5, 1), (1, 0.92156863, 0.01568628, 1)
        v.fluid_color[:], v.obstacle_color[:], v.source_position_color[:] = white, gray, yellow
        v.low_density_color[:], v.medium_density_color[:], v.high_density_color[:] = blue, green, red
        v.low_pressure_color[:], v.neutral_pressure_color[:], v.high_pressure_color[:] = blue, white, red
        v.gradient_key_count = 2                      # Start(): blue at 0, red at 1 (:188-203)
        v.gradient_colors[0][:], v.gradient_colors[1][:] = blue, red
        v.gradient_times[0], v.gradient_times[1] = 0.0, 1.0
        return v


class FsObstacleShape(C.Structure):
    """fs_obstacle_shape (include/fluidsolver.h): parameters of the device-side SetupObstacles (FluidSim.cs:302-388)."""
    _fields_ = [
        ("kind", C.c_int32), ("center_x", C.c_float), ("center_y", C.c_float), ("center_z", C.c_float),
        ("radius", C.c_float), ("width", C.c_float), ("height", C.c_float), ("depth", C.c_float),
        ("seed_x", C.c_int32), ("seed_y", C.c_int32), ("seed_z", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


SHAPE_KINDS = {"Circle": 0, "Rectangle": 1, "Airfoil": 2}


class FluidSolverError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fluidsolver error {code}: {message}")
        self.code = code


_F = C.POINTER(C.c_float)
_libs: dict[str, C.CDLL] = {}


def load(path: str | None = None) -> C.CDLL:
    """Load the native library.  Raises if it is absent -- the product has no other path."""
    path = os.path.abspath(path or os.environ.get("FLUIDSOLVER_LIB", DEFAULT_LIB))
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python 3dfluidsimulation_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    sig = {
        "fs_abi_version": (C.c_int, []),
        "fs_create": (C.c_int, [C.POINTER(FsParams), C.POINTER(vp)]),
        "fs_destroy": (None, [vp]),
        "fs_reset": (C.c_int, [vp]),
        "fs_last_error": (C.c_char_p, [vp]),
        "fs_slab_range": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
        "fs_set_obstacles": (C.c_int, [vp, vp, i64]),
        "fs_slab_halo_range": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "fs_set_obstacles_slab": (C.c_int, [vp, vp, i64, i32, i32]),
        "fs_build_obstacles": (C.c_int, [vp, C.POINTER(FsObstacleShape), C.POINTER(i64)]),
        "fs_get_obstacles": (C.c_int, [vp, vp, i64]),
        "fs_add_density": (C.c_int, [vp, f32, f32, f32, f32]),
        "fs_add_velocity": (C.c_int, [vp, f32, f32, f32, f32, f32, f32]),
        "fs_add_source_cells": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
        "fs_add_sources": (C.c_int, [vp, vp, vp, vp, vp]),
        "fs_step": (C.c_int, [vp, f32, f32, f32]),
        "fs_sync": (C.c_int, [vp]),
        "fs_get_field": (C.c_int, [vp, i32, vp, i64]),
        "fs_set_field": (C.c_int, [vp, i32, vp, i64]),
        "fs_get_field_async": (C.c_int, [vp, i32, vp, i64]),
        "fs_wait_transfers": (C.c_int, [vp]),
        "fs_get_metrics": (C.c_int, [vp, _F, _F, C.POINTER(C.c_double)]),
        "fs_render_rgba": (C.c_int, [vp, C.POINTER(FsVisParams), vp, i64]),
        "fs_streamlines": (C.c_int, [vp, i32, f32, i32, vp, i64]),
        "fs_op_set_bnd": (C.c_int, [vp, i32, i32]),
        "fs_op_diffuse": (C.c_int, [vp, i32, i32, i32, f32, f32]),
        "fs_op_smooth": (C.c_int, [vp, i32, i32, i32, f32, f32, i32]),
        "fs_op_lin_solve": (C.c_int, [vp, i32, i32, i32, f32, f32, i32, i32]),
        "fs_op_project": (C.c_int, [vp, i32]),
        "fs_op_advect": (C.c_int, [vp, i32, i32, i32, i32, f32]),
        "fs_op_advect_velocity": (C.c_int, [vp, f32]),
        "fs_op_enforce_obstacles": (C.c_int, [vp]),
        "fs_timer_start": (C.c_int, [vp]),
        "fs_timer_stop": (C.c_int, [vp, _F]),
        "fs_launch_count": (i64, [vp]),
        "fs_bench_sweep": (C.c_int, [vp, i32, i32, i32, _F, C.POINTER(C.c_double)]),
        "fs_selftest_division": (C.c_int, [vp, f32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
        "fs_halo_export": (C.c_int, [vp, vp, i64]),
        "fs_halo_connect": (C.c_int, [vp, vp, vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    if lib.fs_abi_version() != FS_ABI_VERSION:
        raise RuntimeError("libfluidsolver ABI version mismatch")
    _libs[path] = lib
    return lib


def _ptr(a: np.ndarray | None, dtype):
    if a is None:
        return None
    if a.dtype != dtype or not a.flags.c_contiguous:
        raise TypeError(f"expected C-contiguous {np.dtype(dtype).name} array")
    return a.ctypes.data


class NativeSolver:
    """One fs_solver handle (one GPU, one z-slab)."""

    def __init__(self, nx, ny, nz=1, *, iters_diffuse=20, iters_pressure=20, solver_kind=JACOBI,
                 enable_obstacle=True, cell_size=None, raw_viscosity=1e-4, device_id=0, slab_rank=0,
                 slab_count=1, use_cuda_graph=False, lib_path=None):
        self.lib = load(lib_path)
        p = FsParams()
        p.abi_version = FS_ABI_VERSION
        p.nx, p.ny, p.nz = int(nx), int(ny), int(nz)
        p.iters_diffuse, p.iters_pressure = int(iters_diffuse), int(iters_pressure)
        p.solver_kind = int(solver_kind)
        p.enable_obstacle = int(bool(enable_obstacle))
        p.cell_size = float(cell_size if cell_size is not None else 1.0 / nx)
        p.raw_viscosity = float(raw_viscosity)
        p.device_id, p.slab_rank, p.slab_count = int(device_id), int(slab_rank), int(slab_count)
        p.use_cuda_graph = int(bool(use_cuda_graph))
        self.params = p
        h = C.c_void_p()
        rc = self.lib.fs_create(C.byref(p), C.byref(h))
        if rc != 0:
            raise FluidSolverError(rc, (self.lib.fs_last_error(None) or b"").decode())
        self.h = h
        zb, ze, n = C.c_int32(), C.c_int32(), C.c_int64()
        self.lib.fs_slab_range(self.h, C.byref(zb), C.byref(ze), C.byref(n))
        self.z_begin, self.z_end, self.owned_voxels = zb.value, ze.value, n.value
        self.nx, self.ny, self.nz = p.nx, p.ny, p.nz
        self.shape = (p.ny, p.nx) if p.nz == 1 else (self.z_end - self.z_begin, p.ny, p.nx)

    # -- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise FluidSolverError(rc, (self.lib.fs_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.fs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- reference surface
    def reset(self):
        self._ck(self.lib.fs_reset(self.h))

    def set_obstacles(self, mask: np.ndarray):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self._ck(self.lib.fs_set_obstacles(self.h, mask.ctypes.data, mask.size))

    def halo_range(self):
        """Global planes [z0, z1) this handle holds (owned planes plus ghost planes)."""
        z0, z1 = C.c_int32(), C.c_int32()
        self._ck(self.lib.fs_slab_halo_range(self.h, C.byref(z0), C.byref(z1)))
        return z0.value, z1.value

    def set_obstacles_slab(self, mask: np.ndarray, global_any: bool, global_interior: bool):
        """Mask of the planes halo_range() reports only (no rank needs the global mask)."""
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self._ck(self.lib.fs_set_obstacles_slab(self.h, mask.ctypes.data, mask.size, int(global_any), int(global_interior)))

    def build_obstacles(self, shape: "FsObstacleShape") -> int:
        """Device-side SetupObstacles: builds this handle's planes of the mask; returns the global obstacle cell count."""
        n = C.c_int64()
        self._ck(self.lib.fs_build_obstacles(self.h, C.byref(shape), C.byref(n)))
        return int(n.value)

    def get_obstacles(self) -> np.ndarray:
        out = np.empty(self.shape, np.uint8)
        self._ck(self.lib.fs_get_obstacles(self.h, out.ctypes.data, out.size))
        return out

    def add_density(self, x, y, z, amount):
        self._ck(self.lib.fs_add_density(self.h, x, y, z, amount))

    def add_velocity(self, x, y, z, ax, ay, az=0.0):
        self._ck(self.lib.fs_add_velocity(self.h, x, y, z, ax, ay, az))

    def add_source_cells(self, x, y, z=None, density=None, ax=None, ay=None, az=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z, density, ax, ay, az)]
        self._ck(self.lib.fs_add_source_cells(self.h, arrs[0].size, *[_ptr(a, np.float32) for a in arrs]))

    def add_sources(self, density=None, vx=None, vy=None, vz=None):
        self._ck(self.lib.fs_add_sources(self.h, *[_ptr(a, np.float32) for a in (density, vx, vy, vz)]))

    def step(self, dt, visc, diff):
        self._ck(self.lib.fs_step(self.h, dt, visc, diff))

    def sync(self):
        self._ck(self.lib.fs_sync(self.h))

    def get_field(self, field, out: np.ndarray | None = None) -> np.ndarray:
        fid = FIELD_IDS[field] if isinstance(field, str) else int(field)
        if out is None:
            out = np.empty(self.shape, np.float32)
        self._ck(self.lib.fs_get_field(self.h, fid, _ptr(out, np.float32), out.size))
        return out

    def get_field_async(self, field, out: np.ndarray):
        """Pipelined readback into `out` (pinned for full speed); valid after wait_transfers()."""
        fid = FIELD_IDS[field] if isinstance(field, str) else int(field)
        self._ck(self.lib.fs_get_field_async(self.h, fid, _ptr(out, np.float32), out.size))

    def wait_transfers(self):
        self._ck(self.lib.fs_wait_transfers(self.h))

    def set_field(self, field, a: np.ndarray):
        fid = FIELD_IDS[field] if isinstance(field, str) else int(field)
        a = np.ascontiguousarray(a, dtype=np.float32)
        self._ck(self.lib.fs_set_field(self.h, fid, a.ctypes.data, a.size))

    def render_rgba(self, vis: "FsVisParams", out: np.ndarray | None = None) -> np.ndarray:
        """UpdateVisualizationJob on the device: (ny, nx, 4) float32 RGBA of one xy plane."""
        if out is None:
            out = np.empty((self.ny, self.nx, 4), np.float32)
        self._ck(self.lib.fs_render_rgba(self.h, C.byref(vis), _ptr(out, np.float32), out.size))
        return out

    def streamlines(self, skip, scale, z_slice=0) -> np.ndarray:
        """Streamline glyph segments of one plane: (count, 4) float32 rows (startX, startY, endX, endY), -1 = invalid."""
        count = (self.nx // skip) * (self.ny // skip)
        out = np.empty((count, 4), np.float32)
        self._ck(self.lib.fs_streamlines(self.h, skip, scale, z_slice, _ptr(out, np.float32), count))
        return out

    def metrics(self):
        m, s, t = C.c_float(), C.c_float(), C.c_double()
        self._ck(self.lib.fs_get_metrics(self.h, C.byref(m), C.byref(s), C.byref(t)))
        return m.value, s.value, t.value

    # -- operators
    def op_set_bnd(self, field, b):
        self._ck(self.lib.fs_op_set_bnd(self.h, FIELD_IDS[field], b))

    def op_diffuse(self, dst, src, b, diff, dt):
        self._ck(self.lib.fs_op_diffuse(self.h, FIELD_IDS[dst], FIELD_IDS[src], b, diff, dt))

    def op_smooth(self, dst, src, b, a, c, iters):
        self._ck(self.lib.fs_op_smooth(self.h, FIELD_IDS[dst], FIELD_IDS[src], b, a, c, iters))

    def op_lin_solve(self, dst, rhs, b, a, c, iters, solver_kind=JACOBI):
        self._ck(self.lib.fs_op_lin_solve(self.h, FIELD_IDS[dst], FIELD_IDS[rhs], b, a, c, iters, solver_kind))

    def op_project(self, use_v0_fields=False):
        self._ck(self.lib.fs_op_project(self.h, int(use_v0_fields)))

    def op_advect(self, dst, src, b, dt, use_v0_fields=False):
        self._ck(self.lib.fs_op_advect(self.h, FIELD_IDS[dst], FIELD_IDS[src], b, int(use_v0_fields), dt))

    def op_advect_velocity(self, dt):
        self._ck(self.lib.fs_op_advect_velocity(self.h, dt))

    def op_enforce_obstacles(self):
        self._ck(self.lib.fs_op_enforce_obstacles(self.h))

    # -- measurement
    def timer_start(self):
        self._ck(self.lib.fs_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.fs_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(self.lib.fs_launch_count(self.h))

    def bench_sweep(self, kind, b, reps):
        ms, by = C.c_float(), C.c_double()
        self._ck(self.lib.fs_bench_sweep(self.h, kind, b, reps, C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def selftest_division(self, divisor, first_bits=0, count=1 << 32) -> int:
        bad = C.c_uint64()
        self._ck(self.lib.fs_selftest_division(self.h, divisor, first_bits, count, C.byref(bad)))
        return int(bad.value)

    # -- multi-GPU wiring
    def halo_export(self) -> bytes:
        buf = C.create_string_buffer(FS_IPC_BLOB_BYTES)
        self._ck(self.lib.fs_halo_export(self.h, buf, FS_IPC_BLOB_BYTES))
        return buf.raw

    def halo_connect(self, lower: bytes | None, upper: bytes | None, same_process=False):
        lo = C.create_string_buffer(lower, FS_IPC_BLOB_BYTES) if lower else None
        up = C.create_string_buffer(upper, FS_IPC_BLOB_BYTES) if upper else None
        self._ck(self.lib.fs_halo_connect(self.h, lo, up, int(same_process)))
