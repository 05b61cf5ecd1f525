"""Host-side mirror of the reference's ``FluidSimulation`` MonoBehaviour for the solver hot path.

The reference's host language is C# (Unity); no C# toolchain exists in the build image, so the
compiled-language host is ``Assets/Plugin/FluidSimulationNative.cs`` (written against the same C
ABI, not compilable here) and THIS module is the host side that the tests and the benchmark drive.
It keeps the reference's field names, argument meanings and call order
(Assets/Scripts/FluidSim.cs: public fields :12-110, Update :390-450, Simulate :551-576,
AddDensity/AddVelocity :723-738, UpdateCustomSource :485-533, AddForceToArea :452-483,
SetupObstacles :302-388, ResetSimulation :213-235).

Everything numerical happens in libfluidsolver.so (CUDA); this file only does what the reference
does on the managed side: parameter scaling, source lists, the obstacle mask.
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np

from . import native

f32 = np.float32


def _round_half_even(v: float) -> int:
    """Mathf.RoundToInt: banker's rounding on the float value."""
    return int(np.rint(f32(v)))


class FluidSimulation:
    """Drop-in for the reference component's solver surface.  ``depth`` (nz) is the 3D extension:
    ``depth=1`` reproduces the reference's 2D solver."""

    # FluidSim.cs:19-31 defaults
    def __init__(self, size=128, resolutionMultiplier=1.0, *, depth=1, physicalSize=1.0, diffusion=1e-4,
                 viscosity=1e-4, timeStep=0.1, autoAdjustParameters=True, enableObstacle=True,
                 obstacleShape="Circle", obstaclePositionX=0.5, obstaclePositionY=0.5, obstaclePositionZ=0.5,
                 obstacleRadius=0.1, obstacleWidth=0.2, obstacleHeight=0.2, itersDiffuse=20, itersPressure=20,
                 solverKind=native.JACOBI, device_id=0, use_cuda_graph=True, lib_path=None):
        self.paused = False
        self.size, self.resolutionMultiplier, self.depth = size, resolutionMultiplier, depth
        self.physicalSize, self.diffusion, self.viscosity, self.timeStep = physicalSize, diffusion, viscosity, timeStep
        self.autoAdjustParameters = autoAdjustParameters
        # customizable source, :34-55
        self.enableCustomSource = False
        self.sourceStrength, self.sourceEmitsVelocity, self.sourceDirection = 100.0, False, 0.0
        self.sourceVelocity, self.sourceRadius, self.sourcePulseRate, self.sourcePulsing = 10.0, 1.0, 1.0, False
        self.sourcePositionX, self.sourcePositionY, self.sourcePositionZ = 0.5, 0.5, 0.5
        # obstacle, :96-110
        self.enableObstacle, self.obstacleShape = enableObstacle, obstacleShape
        self.obstaclePositionX, self.obstaclePositionY, self.obstaclePositionZ = obstaclePositionX, obstaclePositionY, obstaclePositionZ
        self.obstacleRadius, self.obstacleWidth, self.obstacleHeight = obstacleRadius, obstacleWidth, obstacleHeight
        self.itersDiffuse, self.itersPressure, self.solverKind = itersDiffuse, itersPressure, solverKind
        self._device_id, self._use_graph, self._lib_path = device_id, use_cuda_graph, lib_path
        self.elapsedTime = 0.0
        self.runLog, self._currentRunID, self._smoothedFPS = None, -1, 0.0
        self.native: native.NativeSolver | None = None
        self.ResetSimulation()

    # ---- ResetSimulation, :213-235 (+ SetupObstacles :299) -------------------------------------
    def ResetSimulation(self):
        self.currentSize = _round_half_even(self.size * self.resolutionMultiplier)   # :216
        self.cellSize = f32(self.physicalSize) / f32(self.currentSize)               # :219
        self.dtScale = f32(128.0) / f32(self.currentSize) if self.autoAdjustParameters else f32(1.0)  # :222
        n = self.currentSize
        self.currentDepth = 1 if self.depth == 1 else _round_half_even(self.depth * self.resolutionMultiplier)
        if self.native is not None:
            self.native.close()
        self.native = native.NativeSolver(
            n, n, self.currentDepth, iters_diffuse=self.itersDiffuse, iters_pressure=self.itersPressure,
            solver_kind=self.solverKind, enable_obstacle=self.enableObstacle, cell_size=float(self.cellSize),
            raw_viscosity=self.viscosity, device_id=self._device_id, use_cuda_graph=self._use_graph,
            lib_path=self._lib_path)
        shape = (n, n) if self.currentDepth == 1 else (self.currentDepth, n, n)
        self.obstacles = np.zeros(shape, np.uint8)
        self.SetupObstacles()

    def SetPaused(self, Paused: bool):  # :149
        self.paused = Paused

    def GetSourcePosition(self):  # :979
        return (self.sourcePositionX * self.currentSize, self.sourcePositionY * self.currentSize)

    def SetSourcePosition(self, x, y):  # :984
        self.sourcePositionX = min(max(x / self.currentSize, 0.0), 1.0)
        self.sourcePositionY = min(max(y / self.currentSize, 0.0), 1.0)

    # ---- SetupObstacles / RecursiveFloodFill / IsInsideShape, :302-388 ---------------------------
    def _inside(self, x, y, size):
        n = self.currentSize
        cx, cy = f32(self.obstaclePositionX * n), f32(self.obstaclePositionY * n)
        if self.obstacleShape == "Circle":
            return f32(x - cx) * f32(x - cx) + f32(y - cy) * f32(y - cy) < f32(size) * f32(size)
        if self.obstacleShape == "Rectangle":
            hw, hh = f32(self.obstacleWidth * n * 0.5), f32(self.obstacleHeight * n * 0.5)
            return (cx - hw) < x < (cx + hw) and (cy - hh) < y < (cy + hh)
        if self.obstacleShape == "Airfoil":  # approximate NACA 0015, :369-383
            chord = f32(2 * self.obstacleWidth * n)
            t = f32(0.15)
            nx_ = f32(x - cx + chord / 2) / chord
            ny_ = f32(y - cy) / chord
            if nx_ < 0 or nx_ > 1 or abs(ny_) > t:
                return False
            half = 5 * t * (f32(0.2969) * f32(math.sqrt(nx_)) - f32(0.1260) * nx_ - f32(0.3516) * nx_ * nx_
                            + f32(0.2843) * nx_ ** 3 - f32(0.1015) * nx_ ** 4)
            return abs(ny_) <= half
        return False

    def obstacle_shape(self):
        """The reference's shape parameters in cells, computed as SetupObstacles / IsInsideShape compute them
        (:308-324, :355-370), for the native mask builder (fs_build_obstacles)."""
        n, nz = self.currentSize, self.currentDepth
        sh = native.FsObstacleShape()
        sh.kind = native.SHAPE_KINDS[self.obstacleShape]
        sh.center_x, sh.center_y = f32(self.obstaclePositionX) * f32(n), f32(self.obstaclePositionY) * f32(n)
        sh.center_z = f32(self.obstaclePositionZ) * f32(nz)
        sh.radius = f32(self.obstacleRadius) * f32(n)
        sh.width, sh.height = f32(self.obstacleWidth) * f32(n), f32(self.obstacleHeight) * f32(n)
        sh.depth = f32(self.obstacleWidth) * f32(nz)          # 3D extension: the extrusion spans obstacleWidth of the depth
        sh.seed_x, sh.seed_y = _round_half_even(self.obstaclePositionX * n), _round_half_even(self.obstaclePositionY * n)
        sh.seed_z = _round_half_even(self.obstaclePositionZ * nz) if nz > 1 else 0
        return sh

    def reference_mask(self):
        """SetupObstacles on the HOST, following the reference line by line (:302-388): the recursive 4-neighbour flood
        fill (iterative here) over IsInsideShape; 3D: sphere / extruded cross-section as include/fluidsolver.h states.
        Used by the tests as the restatement the device builder is checked against."""
        n, nz = self.currentSize, self.currentDepth
        sh = self.obstacle_shape()
        m2 = np.zeros((n, n), np.uint8)
        if not self.enableObstacle:
            return m2 if nz == 1 else np.zeros((nz, n, n), np.uint8)
        size = (self.obstacleRadius if self.obstacleShape == "Circle" else self.obstacleWidth) * n  # :316-324
        todo = deque([(sh.seed_x, sh.seed_y)])
        while todo:
            x, y = todo.pop()
            if x < 0 or x >= n or y < 0 or y >= n or m2[y, x] or not self._inside(x, y, size):
                continue
            m2[y, x] = 1
            todo.extend(((x + 1, y), (x - 1, y), (x, y + 1), (x, y - 1)))
        if nz == 1:
            return m2
        m3 = np.zeros((nz, n, n), np.uint8)
        if not (0 <= sh.seed_z < nz):
            return m3
        if self.obstacleShape == "Circle":
            cx, cy, cz, r = f32(sh.center_x), f32(sh.center_y), f32(sh.center_z), f32(sh.radius)
            z, y, x = np.meshgrid(np.arange(nz, dtype=f32), np.arange(n, dtype=f32), np.arange(n, dtype=f32), indexing="ij")
            d2 = ((x - cx) * (x - cx) + (y - cy) * (y - cy)) + (z - cz) * (z - cz)
            inside = d2 < r * r
            if inside[sh.seed_z, sh.seed_y, sh.seed_x] if (0 <= sh.seed_x < n and 0 <= sh.seed_y < n) else False:
                m3 = inside.astype(np.uint8)
            return m3
        half = f32(sh.depth) * f32(0.5)
        zs = np.arange(nz, dtype=f32)
        span = (zs > f32(sh.center_z) - half) & (zs < f32(sh.center_z) + half)
        if span[sh.seed_z]:
            m3[span] = m2
        return m3

    def SetupObstacles(self):
        """:302-327.  The mask is built ON THE DEVICE (fs_build_obstacles); `obstacles` is read back for the callers
        that look at it (the reference's UpdateVisualization reads the managed array, :765)."""
        if not self.enableObstacle:
            shape = (self.currentSize, self.currentSize) if self.currentDepth == 1 else (self.currentDepth, self.currentSize, self.currentSize)
            self.obstacles = np.zeros(shape, np.uint8)
            self.native.set_obstacles(self.obstacles)
            return
        self.obstacleCells = self.native.build_obstacles(self.obstacle_shape())
        self.obstacles = self.native.get_obstacles()

    # ---- sources -----------------------------------------------------------------------------------
    def AddDensity(self, x, y, amount, z=0.0):  # :723-729
        self.native.add_density(x, y, z, amount)

    def AddVelocity(self, x, y, amountX, amountY, z=0.0, amountZ=0.0):  # :731-738
        self.native.add_velocity(x, y, z, amountX, amountY, amountZ)

    def UpdateCustomSource(self):
        """:485-533 -- the disc of AddDensity/AddVelocity calls, sent as one batched native call."""
        n = self.currentSize
        sx, sy = f32(self.sourcePositionX * n), f32(self.sourcePositionY * n)
        pulse = abs(math.sin(self.elapsedTime * self.sourcePulseRate * math.pi)) if self.sourcePulsing else 1.0
        strength = f32(self.sourceStrength * pulse) * f32(self.resolutionMultiplier)
        rad = f32(self.sourceRadius * self.resolutionMultiplier)
        xs, ys, ds, ax, ay = [], [], [], [], []
        for i in range(max(0, math.floor(sx - rad)), min(n - 1, math.ceil(sx + rad)) + 1):
            for j in range(max(0, math.floor(sy - rad)), min(n - 1, math.ceil(sy + rad)) + 1):
                dist = f32(math.sqrt(f32((i - sx) * (i - sx) + (j - sy) * (j - sy))))
                if dist <= rad:
                    fall = f32(1.0) - dist / rad
                    xs.append(i); ys.append(j); ds.append(strength * fall)
                    if self.sourceEmitsVelocity:
                        ang = f32(self.sourceDirection) * f32(math.pi / 180.0)
                        ax.append(f32(math.cos(ang)) * f32(self.sourceVelocity) * f32(self.resolutionMultiplier) * fall)
                        ay.append(f32(math.sin(ang)) * f32(self.sourceVelocity) * f32(self.resolutionMultiplier) * fall)
        if xs:
            z = [self.sourcePositionZ * self.currentDepth] * len(xs) if self.currentDepth > 1 else None
            self.native.add_source_cells(xs, ys, z, ds, ax or None, ay or None, None)

    def AddForceToArea(self, center, force, radius):
        """:452-483 -- mouse drag: velocity with linear fall-off, density inside 0.3 r.  fp32 throughout, as
        Vector2.Distance and the C# float expressions evaluate; sent as ONE batched native call."""
        n = self.currentSize
        cx, cy, radius = f32(center[0]), f32(center[1]), f32(radius)
        fx, fy = f32(force[0]), f32(force[1])
        clamp = lambda v: min(max(v, 0), n - 1)
        xs, ys, ds, ax, ay = [], [], [], [], []
        for x in range(clamp(int(cx - radius)), clamp(int(cx + radius)) + 1):          # :454-457, (int) truncates
            for y in range(clamp(int(cy - radius)), clamp(int(cy + radius)) + 1):
                dx, dy = f32(x) - cx, f32(y) - cy
                dist = f32(math.sqrt(f32(dx * dx + dy * dy)))                           # Vector2.Distance
                if dist <= radius:
                    fall = f32(1) - dist / radius
                    xs.append(x); ys.append(y); ax.append(fx * fall); ay.append(fy * fall)
                    ds.append(f32(self.sourceStrength) * fall if dist < radius * f32(0.3) else f32(0))
        if xs:
            z = [self.sourcePositionZ * self.currentDepth] * len(xs) if self.currentDepth > 1 else None
            self.native.add_source_cells(xs, ys, z, ds, ax, ay, None)

    # ---- Simulate / Update, :390-450, :551-576 --------------------------------------------------------
    def effective_parameters(self):
        """:554-556"""
        if self.autoAdjustParameters:
            return (float(f32(self.timeStep) * self.dtScale), float(f32(self.viscosity) / f32(self.resolutionMultiplier)),
                    float(f32(self.diffusion) / f32(self.resolutionMultiplier)))
        return float(self.timeStep), float(self.viscosity), float(self.diffusion)

    def Simulate(self):
        dt, visc, diff = self.effective_parameters()
        self.native.step(dt, visc, diff)

    Step = Simulate

    def Update(self, deltaTime=1.0 / 60):
        """:390-450 without input/visualisation: sources first, then the step (+ LogCurrentMetrics :572-575 when a
        run log is attached)."""
        if self.paused:
            return
        self.elapsedTime += deltaTime
        if self.enableCustomSource:
            self.UpdateCustomSource()
        self.Simulate()
        if self.runLog is not None and self._currentRunID != -1:
            self.LogCurrentMetrics(deltaTime)

    # ---- persistence (SQL.cs through the portable sink, runlog.py) ---------------------------------------------
    def AttachRunLog(self, log):
        """Start(): `_currentRunID = SQL.SaveSimRunParams(...)` (:205); metrics follow every step (:572-575)."""
        self.runLog = log
        self._currentRunID = self.SaveCurrentConfiguration()
        return self._currentRunID

    def SaveCurrentConfiguration(self):  # :2004-2022
        if self.runLog is None:
            return -1
        return self.runLog.save_sim_run_params(
            self.size, self.diffusion, self.viscosity, self.timeStep, self.enableCustomSource, self.sourceStrength,
            self.sourcePositionX, self.sourcePositionY, self.enableObstacle, self.obstacleShape, self.obstaclePositionX,
            self.obstaclePositionY, self.obstacleRadius, self.obstacleWidth, self.obstacleHeight)

    def LogCurrentMetrics(self, deltaTime):  # :578-615: two device reductions instead of two host loops over the fields
        mean, mx = self.metrics()
        self._smoothedFPS = 0.9 * self._smoothedFPS + 0.1 * (1.0 / deltaTime)   # CalculateFrameRate :609-615
        if mx != 0 and mean != 0:                                                    # :597
            self.runLog.log_runtime_metrics(self._currentRunID, 0, mean, mx, self._smoothedFPS)

    # ---- readers (what UpdateVisualization/LogCurrentMetrics pull, :587-594, :761-768) -------------------
    def field(self, name):
        return self.native.get_field(name)

    def metrics(self):
        mean, mx, _ = self.native.metrics()
        return mean, mx

    def DrawStreamlines(self, streamlineDensity=4, streamlineScale=1.0, streamlineThickness=1.0, z_slice=None):
        """:886-959 -- glyph segments from the device (fs_streamlines), then the reference's host-side Bresenham
        drawing (:1765-1849).  Returns a (size, size) uint8 coverage mask (1 where streamlineColor is painted)."""
        n = self.currentSize
        skip = max(1, n // (streamlineDensity * 10))                                   # :892
        k = 0 if self.currentDepth == 1 else (self.currentDepth // 2 if z_slice is None else z_slice)
        seg = self.native.streamlines(skip, streamlineScale, k)
        tex = np.zeros((n, n), np.uint8)
        half = int(math.floor(streamlineThickness / 2))
        for sx, sy, ex, ey in seg:
            if sx < 0:
                continue
            x0, y0, x1, y1 = int(sx), int(sy), int(np.rint(ex)), int(np.rint(ey))      # math.round: half to even
            steep = abs(y1 - y0) > abs(x1 - x0)
            if steep:
                x0, y0, x1, y1 = y0, x0, y1, x1
            if x0 > x1:
                x0, x1, y0, y1 = x1, x0, y1, y0
            dx, dy = x1 - x0, abs(y1 - y0)
            err, y, ystep = dx // 2, y0, (1 if y0 < y1 else -1)
            for x in range(x0, x1 + 1):
                for tx in range(-half, half + 1):
                    for ty in range(-half, half + 1):
                        px, py = (y + tx, x + ty) if steep else (x + tx, y + ty)
                        if 0 <= px < n and 0 <= py < n:
                            tex[py, px] = 1
                err -= dy
                if err < 0:
                    y += ystep
                    err += dx
        return tex

    def UpdateVisualization(self, vis=None, z_slice=None):
        """:755-853 -- the colour mapping job on the device; returns (ny, nx, 4) RGBA floats (Color[] layout)."""
        vis = vis or native.FsVisParams.reference_defaults(self.currentSize)
        vis.source_x, vis.source_y = self.sourcePositionX * self.currentSize, self.sourcePositionY * self.currentSize
        vis.enable_custom_source = int(self.enableCustomSource)
        if self.currentDepth > 1:
            vis.z_slice = self.currentDepth // 2 if z_slice is None else z_slice
        return self.native.render_rgba(vis)

    def close(self):
        if self.native is not None:
            self.native.close()
            self.native = None
