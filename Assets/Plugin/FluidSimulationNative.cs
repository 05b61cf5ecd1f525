// FluidSimulationNative.cs -- drop-in replacement for the SOLVER part of the reference component
// (Assets/Scripts/FluidSim.cs, class FluidSimulation): same inspector field names, same public methods
// (SetPaused, GetSourcePosition, SetSourcePosition, SaveCurrentConfiguration), and the three operations
// the reference keeps private -- Simulate, AddDensity, AddVelocity -- exposed as Step/AddDensity/AddVelocity.
// All numerical work happens in libfluidsolver.so (CUDA, sm_100a) behind NativeFluidSolver.cs; this class
// only does what the reference does on the managed side: parameter scaling (FluidSim.cs:216-222, :554-556),
// source discs (:485-533, :452-483), the obstacle mask (:302-388) and the per-frame call order (:390-450).
// Rendering, input and SQLite logging stay in the reference's own scripts: they read the public arrays
// Density / Pressure / VelocityX / VelocityY that Update() refreshes (what UpdateVisualization and
// DrawStreamlines read from the private fields in the reference, :761-768, :919-920).
//
// NOT compiled in the build container (no C#/Unity toolchain).  The Python mirror
// 3dfluidsimulation_b200/solver.py has the same logic and is what the tests drive.
using System;
using System.Collections.Generic;
using UnityEngine;
using FluidSolverNative;

public class FluidSimulationNative : MonoBehaviour
{
    [Header("Simulation Parameters")]
    public bool paused = false;
    [Range(32, 512)] public int size = 128;
    [Tooltip("3D extension: number of z planes; 1 reproduces the reference's 2D solver")]
    public int depth = 1;
    public float physicalSize = 1.0f;
    [Range(0.1f, 10f)] public float resolutionMultiplier = 1.0f;
    public float diffusion = 0.0001f;
    public float viscosity = 0.0001f;
    public float timeStep = 0.1f;
    public bool autoAdjustParameters = true;
    public int itersDiffuse = 20, itersPressure = 20;
    public FsSolverKind solverKind = FsSolverKind.Jacobi;
    public int deviceId = 0;

    [Header("Customizable Source")]
    public bool enableCustomSource = false;
    [Range(1f, 500f)] public float sourceStrength = 100f;
    public bool sourceEmitsVelocity = false;
    [Range(0f, 360f)] public float sourceDirection = 0f;
    [Range(1f, 50f)] public float sourceVelocity = 10f;
    [Range(0.1f, 10f)] public float sourceRadius = 1f;
    [Range(0.1f, 5f)] public float sourcePulseRate = 1f;
    public bool sourcePulsing = false;
    [Range(0f, 1f)] public float sourcePositionX = 0.5f;
    [Range(0f, 1f)] public float sourcePositionY = 0.5f;
    [Range(0f, 1f)] public float sourcePositionZ = 0.5f;

    [Header("Obstacle Settings")]
    public bool enableObstacle = true;
    public enum ObstacleShape { Circle, Rectangle, Airfoil }
    public ObstacleShape obstacleShape = ObstacleShape.Circle;
    [Range(0f, 1f)] public float obstaclePositionX = 0.5f;
    [Range(0f, 1f)] public float obstaclePositionY = 0.5f;
    [Range(0.01f, 0.5f)] public float obstacleRadius = 0.1f;
    [Range(0.01f, 0.5f)] public float obstacleWidth = 0.2f;
    [Range(0.01f, 0.5f)] public float obstacleHeight = 0.2f;

    // what the visualisation / metrics code reads each frame
    public float[] Density, Pressure, VelocityX, VelocityY;
    public byte[] Obstacles;
    public int CurrentSize => currentSize;

    private SolverHandle solver;
    private int currentSize, currentDepth;
    private float cellSize, dtScale, elapsedTime;
    private readonly List<float> sx = new List<float>(), sy = new List<float>(), sz = new List<float>(),
                                 sd = new List<float>(), sax = new List<float>(), say = new List<float>();

    public void SetPaused(bool Paused) { paused = Paused; }
    public Vector2 GetSourcePosition() { return new Vector2(sourcePositionX * currentSize, sourcePositionY * currentSize); }
    public void SetSourcePosition(float x, float y)
    {
        sourcePositionX = Mathf.Clamp01(x / currentSize);
        sourcePositionY = Mathf.Clamp01(y / currentSize);
    }
    public void SaveCurrentConfiguration()
    {
        SQL.SaveSimRunParams(size, diffusion, viscosity, timeStep, enableCustomSource, sourceStrength, sourcePositionX,
            sourcePositionY, enableObstacle, obstacleShape.ToString(), obstaclePositionX, obstaclePositionY,
            obstacleRadius, obstacleWidth, obstacleHeight);
    }

    void Start() { ResetSimulation(); }
    void OnDestroy() { solver?.Dispose(); solver = null; }

    // ResetSimulation + SetupObstacles (FluidSim.cs:213-235, :299)
    public void ResetSimulation()
    {
        solver?.Dispose();
        currentSize = Mathf.RoundToInt(size * resolutionMultiplier);
        currentDepth = depth <= 1 ? 1 : Mathf.RoundToInt(depth * resolutionMultiplier);
        cellSize = physicalSize / currentSize;
        dtScale = autoAdjustParameters ? 128f / currentSize : 1f;
        var p = new FsParams
        {
            abiVersion = Native.AbiVersion, nx = currentSize, ny = currentSize, nz = currentDepth,
            itersDiffuse = itersDiffuse, itersPressure = itersPressure, solverKind = (int)solverKind,
            enableObstacle = enableObstacle ? 1 : 0, cellSize = cellSize, rawViscosity = viscosity,
            deviceId = deviceId, slabRank = 0, slabCount = 1, useCudaGraph = 1
        };
        int rc = Native.fs_create(ref p, out solver);
        if (rc != 0) throw new InvalidOperationException("fs_create failed: " + System.Runtime.InteropServices.Marshal.PtrToStringAnsi(Native.fs_last_error(null)));
        int total = currentSize * currentSize * currentDepth;
        Density = new float[total]; Pressure = new float[total]; VelocityX = new float[total]; VelocityY = new float[total];
        SetupObstacles();
    }

    // SetupObstacles / flood fill (FluidSim.cs:302-388), iterative instead of recursive; 3D: the 2D mask extruded
    public void SetupObstacles()
    {
        int n = currentSize;
        var plane = new byte[n * n];
        if (enableObstacle)
        {
            float extent = (obstacleShape == ObstacleShape.Circle ? obstacleRadius : obstacleWidth) * n;
            var todo = new Stack<Vector2Int>();
            todo.Push(new Vector2Int(Mathf.RoundToInt(obstaclePositionX * n), Mathf.RoundToInt(obstaclePositionY * n)));
            while (todo.Count > 0)
            {
                var c = todo.Pop();
                if (c.x < 0 || c.x >= n || c.y < 0 || c.y >= n || plane[c.x + c.y * n] != 0 || !InsideShape(c.x, c.y, extent)) continue;
                plane[c.x + c.y * n] = 1;
                todo.Push(new Vector2Int(c.x + 1, c.y)); todo.Push(new Vector2Int(c.x - 1, c.y));
                todo.Push(new Vector2Int(c.x, c.y + 1)); todo.Push(new Vector2Int(c.x, c.y - 1));
            }
        }
        Obstacles = new byte[n * n * currentDepth];
        for (int k = 0; k < currentDepth; k++) Array.Copy(plane, 0, Obstacles, k * n * n, n * n);
        Native.Check(Native.fs_set_obstacles(solver, Obstacles, Obstacles.Length), solver);
    }

    bool InsideShape(int x, int y, float extent)
    {
        float cx = obstaclePositionX * currentSize, cy = obstaclePositionY * currentSize;
        switch (obstacleShape)
        {
            case ObstacleShape.Circle:
                return (x - cx) * (x - cx) + (y - cy) * (y - cy) < extent * extent;
            case ObstacleShape.Rectangle:
                float hw = obstacleWidth * currentSize * 0.5f, hh = obstacleHeight * currentSize * 0.5f;
                return x > cx - hw && x < cx + hw && y > cy - hh && y < cy + hh;
            default: // NACA 0015 approximation, FluidSim.cs:369-383
                float chord = 2 * obstacleWidth * currentSize, t = 0.15f;
                float u = (x - cx + chord / 2) / chord, v = (y - cy) / chord;
                if (u < 0 || u > 1 || Math.Abs(v) > t) return false;
                float half = 5 * t * (0.2969f * Mathf.Sqrt(u) - 0.1260f * u - 0.3516f * u * u + 0.2843f * u * u * u - 0.1015f * u * u * u * u);
                return Math.Abs(v) <= half;
        }
    }

    // AddDensity / AddVelocity (FluidSim.cs:723-738): the native side truncates and clamps like the reference
    public void AddDensity(float x, float y, float amount, float z = 0f) { Native.Check(Native.fs_add_density(solver, x, y, z, amount), solver); }
    public void AddVelocity(float x, float y, float amountX, float amountY, float z = 0f, float amountZ = 0f)
    { Native.Check(Native.fs_add_velocity(solver, x, y, z, amountX, amountY, amountZ), solver); }

    // UpdateCustomSource (FluidSim.cs:485-533): the disc is sent as ONE batched native call
    void UpdateCustomSource()
    {
        float srcX = sourcePositionX * currentSize, srcY = sourcePositionY * currentSize;
        float pulse = sourcePulsing ? Mathf.Abs(Mathf.Sin(elapsedTime * sourcePulseRate * Mathf.PI)) : 1f;
        float strength = sourceStrength * pulse * resolutionMultiplier;
        float r = sourceRadius * resolutionMultiplier;
        sx.Clear(); sy.Clear(); sz.Clear(); sd.Clear(); sax.Clear(); say.Clear();
        for (int i = Mathf.Max(0, Mathf.FloorToInt(srcX - r)); i <= Mathf.Min(currentSize - 1, Mathf.CeilToInt(srcX + r)); i++)
            for (int j = Mathf.Max(0, Mathf.FloorToInt(srcY - r)); j <= Mathf.Min(currentSize - 1, Mathf.CeilToInt(srcY + r)); j++)
            {
                float dist = Mathf.Sqrt((i - srcX) * (i - srcX) + (j - srcY) * (j - srcY));
                if (dist > r) continue;
                float falloff = 1.0f - dist / r;
                sx.Add(i); sy.Add(j); sz.Add(sourcePositionZ * currentDepth); sd.Add(strength * falloff);
                float ang = sourceDirection * Mathf.Deg2Rad;
                sax.Add(sourceEmitsVelocity ? Mathf.Cos(ang) * sourceVelocity * resolutionMultiplier * falloff : 0f);
                say.Add(sourceEmitsVelocity ? Mathf.Sin(ang) * sourceVelocity * resolutionMultiplier * falloff : 0f);
            }
        if (sx.Count == 0) return;
        Native.Check(Native.fs_add_source_cells(solver, sx.Count, sx.ToArray(), sy.ToArray(), sz.ToArray(), sd.ToArray(),
            sourceEmitsVelocity ? sax.ToArray() : null, sourceEmitsVelocity ? say.ToArray() : null, null), solver);
    }

    // Simulate (FluidSim.cs:551-570): the scaling stays managed, the step is one native call
    public void Step()
    {
        float dt = autoAdjustParameters ? timeStep * dtScale : timeStep;
        float diff = autoAdjustParameters ? diffusion / resolutionMultiplier : diffusion;
        float visc = autoAdjustParameters ? viscosity / resolutionMultiplier : viscosity;
        Native.Check(Native.fs_step(solver, dt, visc, diff), solver);
    }

    // Update (FluidSim.cs:390-450): sources, step, then the fields the visualisation reads
    void Update()
    {
        if (paused) return;
        elapsedTime += Time.deltaTime;
        if (enableCustomSource) UpdateCustomSource();
        Step();
        Native.Check(Native.fs_get_field(solver, (int)FsField.Density, Density, Density.Length), solver);
        Native.Check(Native.fs_get_field(solver, (int)FsField.Pressure, Pressure, Pressure.Length), solver);
    }

    public void ReadVelocity()
    {
        Native.Check(Native.fs_get_field(solver, (int)FsField.VelocityX, VelocityX, VelocityX.Length), solver);
        Native.Check(Native.fs_get_field(solver, (int)FsField.VelocityY, VelocityY, VelocityY.Length), solver);
    }

    public void GetMetrics(out float meanDensity, out float maxSpeed)
    {
        Native.Check(Native.fs_get_metrics(solver, out meanDensity, out maxSpeed, out _), solver);
    }
}
