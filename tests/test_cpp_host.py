"""CPU tier: the compiled-language host mirror (include/fluid_simulation.hpp) drives the C ABI from C++ and lands on
the same numbers as the Python mirror -- parameter scaling, custom-source disc, obstacle flood fill and call order
are the managed-side logic of the reference (FluidSim.cs:213-235, :302-388, :390-450, :485-533)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, EMUL_DIR


@pytest.fixture(scope="module")
def demo(emul_lib, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "host_demo")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "host_demo.cpp"), "-o", exe,
                    "-L", EMUL_DIR, "-lfluidsolver_hostemul", f"-Wl,-rpath,{EMUL_DIR}"], check=True)
    return exe


SHAPES = ["Circle", "Rectangle", "Airfoil"]


@pytest.mark.parametrize("size,frames,depth,shape", [(48, 3, 1, 0), (48, 2, 1, 2), (24, 2, 12, 0), (32, 2, 10, 1)])
def test_cpp_host_matches_python_mirror(demo, emul_lib, pkg, size, frames, depth, shape):
    """Same frames through both hosts: custom source, AddForceToArea, Simulate, then UpdateVisualization (DensityBased)
    and DrawStreamlines; obstacle masks come from the native builder in both, 2D and 3D."""
    out = subprocess.run([demo, str(size), str(frames), str(depth), str(shape)], capture_output=True, text=True, check=True).stdout
    got = {l.split()[0]: [float(v) for v in l.split()[1:]] for l in out.strip().splitlines()}
    sim = pkg.FluidSimulation(size=size, depth=depth, obstacleShape=SHAPES[shape], lib_path=emul_lib, use_cuda_graph=False)
    sim.enableCustomSource = True
    sim.sourceEmitsVelocity = True
    sim.sourceDirection = 90.0
    sim.sourceRadius = 2.0
    sim.sourcePositionY = 0.2
    for f in range(frames):
        sim.AddForceToArea((np.float32(0.3) * np.float32(size) + np.float32(f), np.float32(0.6) * np.float32(size)), (2.5, -1.25), 3.0)
        sim.Update()
    assert int(sim.obstacles.sum()) == int(got["obstacle_cells"][0])
    for name in ("density", "vx", "vy", "pressure"):
        a = sim.field(name).astype(np.float64)
        np.testing.assert_allclose(got[name], [a.sum(), (a * a).sum()], rtol=2e-5, atol=1e-12, err_msg=name)
    vis = pkg.native.FsVisParams.reference_defaults(size, 2)
    vis.visualize_source_position = 1
    rgba = sim.UpdateVisualization(vis).astype(np.float64)
    np.testing.assert_allclose(got["rgba"], [rgba.sum(), (rgba * rgba).sum()], rtol=2e-5, err_msg="rgba")
    assert int(sim.DrawStreamlines().sum()) == int(got["streamline_pixels"][0])
    sim.close()
