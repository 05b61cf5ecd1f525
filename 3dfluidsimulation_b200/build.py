"""In-tree build of libfluidsolver.so (nvcc, sm_100a only).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only container (``__graft_entry__.build()``).
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfluidsolver.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # reference arithmetic is unfused fp32 (Burst strict float mode)
    "-prec-div=true", "-prec-sqrt=true",
    "--extended-lambda",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
    "-shared", "-cudart", "shared",
    "-diag-suppress", "20011,20014",   # HD templates instantiated with device-only samplers
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfluidsolver.so cannot be built (there is no CPU fallback)")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(os.path.dirname(HERE), "include", "fluidsolver.h")
    ]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    # build into a private file and rename it into place: a concurrent process (another torchrun rank, a test
    # worker) either sees the old complete library or the new complete one, never a half-written file
    tmp = f"{LIB}.{os.getpid()}.tmp"
    cmd = [nvcc_path(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", tmp, os.path.join(CSRC, "fluidsolver.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force=True, verbose="-v" in sys.argv))
