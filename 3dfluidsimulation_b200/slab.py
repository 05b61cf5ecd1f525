"""z-slab decomposition helpers.

The C ABI gives one handle per GPU / slab (``fs_params.slab_rank / slab_count``); neighbours are wired
with ``fs_halo_export`` / ``fs_halo_connect``.  Two drivers:

* ``connect_distributed``: one process per GPU under torchrun; the IPC blobs travel through
  ``torch.distributed.all_gather_object`` (plumbing only -- the data path is P2P stores between the
  slabs' own kernels, there is no collective on it).
* ``SlabGroup``: all slabs in ONE process (what a single-process host such as the Unity player would
  do with several GPUs), one host thread per handle because every handle's calls block on its own
  stream while its neighbours must keep being fed.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import native


def slab_bounds(nz: int, count: int):
    """The partition fs_create uses (csrc/fs_core.h): balanced, contiguous."""
    return [(nz * r // count, nz * (r + 1) // count) for r in range(count)]


def connect_distributed(solver: native.NativeSolver, dist, rank: int, world: int):
    blobs = [None] * world
    dist.all_gather_object(blobs, solver.halo_export())
    solver.halo_connect(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None,
                        same_process=False)
    dist.barrier()


class SlabGroup:
    """P slab handles in one process; every method fans out to all handles on their own threads."""

    def __init__(self, nx, ny, nz, count, *, devices=None, lib_path=None, **kw):
        self.count, self.nx, self.ny, self.nz = count, nx, ny, nz
        lib = native.load(lib_path)
        emulated = hasattr(lib, "fs_host_emulation")   # tests/host_emul: host threads, no devices involved
        devices = list(devices) if devices is not None else ([0] * count if emulated else list(range(count)))
        if len(devices) != count:
            raise ValueError("one device per slab is required")
        if not emulated and len(set(devices)) != count:
            # every slab's halo kernel spins on flags its neighbours' kernels write; kernels of different handles on
            # ONE GPU are not guaranteed to be co-resident, so two slabs sharing a device can deadlock
            raise ValueError("z-slabs need distinct devices (slabs that wait on one another cannot share a GPU)")
        self.solvers = [native.NativeSolver(nx, ny, nz, slab_rank=r, slab_count=count, device_id=devices[r],
                                            lib_path=lib_path, **kw) for r in range(count)]
        self.pool = ThreadPoolExecutor(max_workers=count)
        if count > 1:
            blobs = [s.halo_export() for s in self.solvers]
            for r, s in enumerate(self.solvers):
                s.halo_connect(blobs[r - 1] if r > 0 else None, blobs[r + 1] if r < count - 1 else None, same_process=True)

    def each(self, fn):
        """fn(rank, solver) on every slab concurrently; re-raises the first failure."""
        futs = [self.pool.submit(fn, r, s) for r, s in enumerate(self.solvers)]
        return [f.result() for f in futs]

    def set_obstacles(self, mask):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self.each(lambda r, s: s.set_obstacles(mask))

    def set_field(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self.each(lambda r, s: s.set_field(name, a[s.z_begin:s.z_end]))

    def get_field(self, name):
        parts = self.each(lambda r, s: s.get_field(name))
        return np.concatenate(parts, axis=0)

    def add_source_cells(self, *a, **kw):
        self.each(lambda r, s: s.add_source_cells(*a, **kw))

    def step(self, dt, visc, diff):
        self.each(lambda r, s: s.step(dt, visc, diff))

    def call(self, method, *a, **kw):
        return self.each(lambda r, s: getattr(s, method)(*a, **kw))

    def metrics(self):
        res = self.each(lambda r, s: s.metrics())
        total = sum(t for _, _, t in res)
        return total / (self.nx * self.ny * self.nz), max(m for _, m, _ in res)

    def close(self):
        self.each(lambda r, s: s.sync())
        for s in self.solvers:
            s.close()
        self.pool.shutdown()
