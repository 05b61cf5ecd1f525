"""CPU tier: the product library loads and exports every symbol include/fluidsolver.h declares (no
compute call is made without a GPU), creating a handle without a GPU fails loudly, and the
host-side mirror's managed-side logic (parameter scaling, obstacle mask) follows the reference."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "fluidsolver.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fs_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(pkg):
    assert declared_functions() == sorted(pkg.native.EXPORTS)


def test_library_exports_every_declared_symbol(cuda_lib):
    lib = ctypes.CDLL(cuda_lib)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/fluidsolver.h but not exported"
    assert lib.fs_abi_version() == 1


def test_no_cpu_fallback_without_gpu(pkg, cuda_lib):
    """On a box without CUDA the product must refuse to run rather than fall back to anything."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU-only container")
    with pytest.raises(pkg.FluidSolverError) as e:
        pkg.NativeSolver(16, 16, 1, lib_path=cuda_lib)
    assert "CUDA" in str(e.value) or "device" in str(e.value)


def test_missing_library_is_an_error(pkg, tmp_path):
    with pytest.raises(FileNotFoundError):
        pkg.native.load(str(tmp_path / "nope.so"))


def test_product_package_never_imports_the_oracle():
    code = ("import importlib,sys; importlib.import_module('3dfluidsimulation_b200'); "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    import subprocess

    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for root, _, files in os.walk(os.path.join(ROOT, "3dfluidsimulation_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "fluid_oracle" not in text, f"{f} references the oracle"


def test_parameter_scaling_and_masks(pkg, emul_lib):
    """FluidSim.cs:216-222, :554-556 and SetupObstacles :302-361 with the two scene configurations
    (SampleScene.unity:260-343: size 64 x multiplier 3, airfoil; :529-612: 128, circle)."""
    sim = pkg.FluidSimulation(size=64, resolutionMultiplier=3.0, obstacleShape="Airfoil", lib_path=emul_lib, use_cuda_graph=False)
    assert sim.currentSize == 192
    dt, visc, diff = sim.effective_parameters()
    assert np.float32(dt) == np.float32(0.1) * (np.float32(128.0) / np.float32(192))
    assert np.float32(visc) == np.float32(1e-4) / np.float32(3.0)
    assert sim.obstacles.any() and sim.obstacles.shape == (192, 192)
    sim.close()
    sim = pkg.FluidSimulation(size=128, lib_path=emul_lib, use_cuda_graph=False)
    n = 128
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    want = ((xx - 64.0) ** 2 + (yy - 64.0) ** 2 < (0.1 * n) ** 2).astype(np.uint8)
    np.testing.assert_array_equal(sim.obstacles, want)
    assert sim.effective_parameters()[0] == pytest.approx(0.1)
    sim.close()


def test_update_order_matches_reference(pkg, emul_lib, oracle):
    """Update(): custom source first, then Simulate (FluidSim.cs:405-442), driven through the mirror."""
    sim = pkg.FluidSimulation(size=32, lib_path=emul_lib, use_cuda_graph=False)
    sim.enableCustomSource = True
    sim.sourceEmitsVelocity = True
    sim.sourceDirection = 90.0
    sim.sourceRadius = 2.0
    sim.sourcePositionY = 0.2
    o = oracle.OracleSolver(32, 32, 1, cell_size=float(sim.cellSize), raw_viscosity=1e-4)
    o.obstacles[...] = sim.obstacles
    import math
    for _ in range(3):
        sim.Update()
        # the same disc, applied to the oracle cell by cell (UpdateCustomSource :503-531)
        sx, sy, rad = 0.5 * 32, 0.2 * 32, 2.0
        for i in range(max(0, math.floor(sx - rad)), min(31, math.ceil(sx + rad)) + 1):
            for j in range(max(0, math.floor(np.float32(sy) - rad)), min(31, math.ceil(np.float32(sy) + rad)) + 1):
                dist = np.float32(math.sqrt(np.float32((i - np.float32(sx)) ** 2 + (j - np.float32(sy)) ** 2)))
                if dist <= rad:
                    fall = np.float32(1) - dist / np.float32(rad)
                    ang = np.float32(90.0) * np.float32(math.pi / 180.0)
                    o.add_density(i, j, 0, np.float32(100.0) * fall)
                    o.add_velocity(i, j, 0, np.float32(math.cos(ang)) * np.float32(10.0) * fall, np.float32(math.sin(ang)) * np.float32(10.0) * fall)
        o.step(*sim.effective_parameters())
    for k in ("density", "vx", "vy", "pressure"):
        got, want = sim.field(k), o.f[k]
        assert np.abs(got - want).max() <= 3e-6 * max(np.abs(want).max(), 1e-30), k
    sim.close()


# ---- C# binding: struct layouts ---------------------------------------------------------------------------------
def _c_struct_fields(name):
    """[(field, array_suffix)] of a struct in include/fluidsolver.h, in declaration order."""
    src = open(os.path.join(ROOT, "include", "fluidsolver.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        for n in names.split(","):
            fields.append(re.match(r"\s*([a-z_0-9]+)", n.strip()).group(1))
    return fields


def _csharp_struct_layout(name):
    """Sequential layout of a [StructLayout(Sequential)] struct of NativeFluidSolver.cs: [(field, offset)], size."""
    src = open(os.path.join(ROOT, "Assets", "Plugin", "NativeFluidSolver.cs")).read()
    body = re.search(r"public struct %s\s*\{(.*?)\n    \}" % name, src, re.S).group(1)
    sizes = {"int": 4, "float": 4, "Color": 16}
    out, off = [], 0
    for line in body.splitlines():
        line = line.split("//")[0].strip()
        if "const" in line or not line.startswith(("public", "[MarshalAs")):
            continue
        m = re.match(r"(?:\[MarshalAs\(UnmanagedType\.ByValArray, SizeConst = (\d+)\)\]\s*)?public (\w+)(\[\])? ([\w, ]+);", line)
        assert m, line
        count, typ, names = int(m.group(1) or 1), m.group(2), [n.strip() for n in m.group(4).split(",")]
        for n in names:
            out.append((n, off))
            off += sizes[typ] * count
    declared = int(re.search(r"public const int NativeSize = (\d+);", body).group(1))
    return out, off, declared


@pytest.mark.parametrize("cname,csname", [("fs_params", "FsParams"), ("fs_vis_params", "FsVisParams"), ("fs_obstacle_shape", "FsObstacleShape")])
def test_csharp_struct_layouts(tmp_path, cname, csname):
    """The P/Invoke structs of Assets/Plugin/NativeFluidSolver.cs have the size and the field offsets of the C structs
    they mirror (measured with a compiled C probe: sizeof / offsetof of every field of include/fluidsolver.h)."""
    import subprocess

    fields = _c_struct_fields(cname)
    probe = tmp_path / "probe.c"
    probe.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fluidsolver.h"\nint main(void){printf("%zu", sizeof(' + cname + '));'
                     + "".join('printf(" %%zu", offsetof(%s, %s));' % (cname, f) for f in fields) + "return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    nums = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    csize, coffs = nums[0], nums[1:]
    cs, size, declared = _csharp_struct_layout(csname)
    cs_offs = [o for _, o in cs]
    if cname in ("fs_params", "fs_obstacle_shape"):   # `reserved[n]` is spelled as n scalar fields on the C# side
        cs_offs = cs_offs[:len(coffs)]
    assert size == csize == declared, (size, csize, declared)
    assert cs_offs == coffs, list(zip(fields, coffs, cs))


def test_csharp_binds_every_entry_point():
    """One [DllImport] per function of include/fluidsolver.h in Assets/Plugin/NativeFluidSolver.cs (the P/Invoke layer the
    Unity scene loads), and the component uses the entry points of the frame path."""
    cs = open(os.path.join(ROOT, "Assets", "Plugin", "NativeFluidSolver.cs")).read()
    bound = set(re.findall(r"public static extern \w+ (fs_[a-z_0-9]+)\(", cs))
    assert sorted(bound) == declared_functions()
    comp = open(os.path.join(ROOT, "Assets", "Plugin", "FluidSimulationNative.cs")).read()
    for name in ("fs_create", "fs_build_obstacles", "fs_add_source_cells", "fs_step", "fs_render_rgba", "fs_streamlines",
                 "fs_get_metrics", "fs_add_density", "fs_add_velocity"):
        assert "Native." + name in comp, name
    for method in ("AddForceToArea", "UpdateVisualization", "DrawStreamlines", "SetupObstacles", "SetPaused", "GetSourcePosition",
                   "SetSourcePosition", "SaveCurrentConfiguration"):
        assert re.search(r"\b%s\(" % method, comp), method
