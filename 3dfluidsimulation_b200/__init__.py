"""3dfluidsimulation_b200 -- B200-native (sm_100a CUDA) implementation of the stable-fluids solver hot
path of ChrisWangstpauls/3DFluidSimulation, behind a C ABI (include/fluidsolver.h).

The package name starts with a digit, so import it with
``importlib.import_module("3dfluidsimulation_b200")``.

Contents: ``csrc/`` (kernels + C ABI, built into ``libfluidsolver.so`` by ``build.py``),
``native`` (ctypes binding, no CPU fallback), ``solver`` (host-side mirror of the reference's
``FluidSimulation`` surface), ``slab`` (z-slab partitioning for multi-GPU runs), ``runlog`` (portable
run-parameter / metrics sink replacing the reference's SQL.cs).
"""
from . import native  # noqa: F401
from .native import FluidSolverError, NativeSolver  # noqa: F401
from .runlog import RunLog  # noqa: F401
from .solver import FluidSimulation  # noqa: F401

__all__ = ["native", "NativeSolver", "FluidSimulation", "FluidSolverError", "RunLog"]
