// FluidSimulationNative.cs -- drop-in replacement for the SOLVER part of the reference component
// (Assets/Scripts/FluidSim.cs, class FluidSimulation): same inspector field names, same public methods
// (SetPaused, GetSourcePosition, SetSourcePosition, SaveCurrentConfiguration), and the three operations
// the reference keeps private -- Simulate, AddDensity, AddVelocity -- exposed as Step/AddDensity/AddVelocity.
// All numerical work happens in libfluidsolver.so (CUDA, sm_100a) behind NativeFluidSolver.cs; this class
// only does what the reference does on the managed side: parameter scaling (FluidSim.cs:216-222, :554-556),
// source discs (:485-533, :452-483), the obstacle mask (:302-388) and the per-frame call order (:390-450).
// The per-frame consumers run on the device as well: UpdateVisualization (:755-866) is fs_render_rgba straight into the
// Color[] the reference fills, DrawStreamlines (:886-959) is fs_streamlines + the reference's host-side Bresenham
// drawing (:1765-1849), LogCurrentMetrics (:578-607) is fs_get_metrics; SetupObstacles (:302-388) is
// fs_build_obstacles.  No full field crosses the boundary in a frame; ReadFields() fetches them on demand.
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
    [Range(0f, 1f)] public float obstaclePositionZ = 0.5f;   // 3D extension
    public Color obstacleColor = Color.gray;

    [Header("Visualization")]            // FluidSim.cs:57-95
    public enum ColorMode { SingleColor, Gradient, DensityBased, PressureBased, Streamlines }
    public ColorMode colorMode = ColorMode.SingleColor;
    public Color fluidcolour = Color.white;
    [Range(0f, 1f)] public float colourIntensity = 1f;
    public Gradient colourGradient;
    public bool useLerp = false;
    public Color startColor = Color.white, endColor = Color.white;
    public Color lowPressureColor = Color.blue, neutralPressureColor = Color.white, highPressureColor = Color.red;
    public float lowPressureThreshold = -50f, highPressureThreshold = 50f;
    public Color lowDensityColor = Color.blue, mediumDensityColor = Color.green, highDensityColor = Color.red;
    [Range(0f, 500f)] public float mediumDensityThreshold = 50f;
    [Range(0f, 1000f)] public float highDensityThreshold = 200f;
    public bool visualizeSourcePosition = true;
    public Color sourcePositionColor = Color.yellow;
    public bool showStreamlines = false;
    [Range(1f, 5f)] public int streamlineDensity = 4;
    [Range(1f, 10f)] public float streamlineScale = 1.0f;
    public Color streamlineColor = Color.white;
    [Range(0.1f, 3f)] public float streamlineThickness = 1.0f;
    [Tooltip("3D extension: the z plane that is rendered (normalised)")] [Range(0f, 1f)] public float viewSliceZ = 0.5f;
    public bool moveSourceWithMouse = false;
    public KeyCode sourcePositionKey = KeyCode.LeftShift;
    public Texture2D fluidTexture, streamlineTexture;

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

    // SetupObstacles (FluidSim.cs:302-327): the flood fill over IsInsideShape runs on the device (fs_build_obstacles);
    // only the shape parameters, computed as the reference computes them (:308-324, :355-370), cross the boundary.
    public void SetupObstacles()
    {
        int n = currentSize;
        Obstacles = new byte[n * n * currentDepth];
        if (!enableObstacle) { Native.Check(Native.fs_set_obstacles(solver, Obstacles, Obstacles.Length), solver); return; }
        var shape = new FsObstacleShape
        {
            kind = (int)obstacleShape,
            centerX = obstaclePositionX * n, centerY = obstaclePositionY * n, centerZ = obstaclePositionZ * currentDepth,
            radius = obstacleRadius * n, width = obstacleWidth * n, height = obstacleHeight * n,
            depth = obstacleWidth * currentDepth,
            seedX = Mathf.RoundToInt(obstaclePositionX * n), seedY = Mathf.RoundToInt(obstaclePositionY * n),
            seedZ = currentDepth > 1 ? Mathf.RoundToInt(obstaclePositionZ * currentDepth) : 0
        };
        Native.Check(Native.fs_build_obstacles(solver, ref shape, out ObstacleCells), solver);
        Native.Check(Native.fs_get_obstacles(solver, Obstacles, Obstacles.Length), solver);
    }
    public long ObstacleCells;

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

    // AddForceToArea (FluidSim.cs:452-483): velocity with linear fall-off, density inside 0.3 r -- ONE batched native call
    public void AddForceToArea(Vector2 center, Vector2 force, float radius)
    {
        int minX = Mathf.Clamp((int)(center.x - radius), 0, currentSize - 1), maxX = Mathf.Clamp((int)(center.x + radius), 0, currentSize - 1);
        int minY = Mathf.Clamp((int)(center.y - radius), 0, currentSize - 1), maxY = Mathf.Clamp((int)(center.y + radius), 0, currentSize - 1);
        sx.Clear(); sy.Clear(); sz.Clear(); sd.Clear(); sax.Clear(); say.Clear();
        for (int x = minX; x <= maxX; x++)
            for (int y = minY; y <= maxY; y++)
            {
                float distance = Vector2.Distance(new Vector2(x, y), center);
                if (distance > radius) continue;
                float falloff = 1 - (distance / radius);
                sx.Add(x); sy.Add(y); sz.Add(sourcePositionZ * currentDepth);
                sax.Add(force.x * falloff); say.Add(force.y * falloff);
                sd.Add(distance < radius * 0.3f ? sourceStrength * falloff : 0f);
            }
        if (sx.Count == 0) return;
        Native.Check(Native.fs_add_source_cells(solver, sx.Count, sx.ToArray(), sy.ToArray(), sz.ToArray(), sd.ToArray(),
            sax.ToArray(), say.ToArray(), null), solver);
    }

    private Vector2 _prevMouseGridPos;
    private bool _isFirstDragFrame = true;
    public Func<Vector2> MousePositionInGrid;   // GetMousePositionInGrid (:535-549) stays with the scene's quad/camera code

    // Update (FluidSim.cs:390-450): source, mouse drag, Simulate, UpdateVisualization -- the reference's order
    void Update()
    {
        if (paused) return;
        elapsedTime += Time.deltaTime;
        Vector2 mouse = MousePositionInGrid != null ? MousePositionInGrid() : Vector2.zero;
        bool positioning = moveSourceWithMouse && Input.GetKey(sourcePositionKey);
        if (positioning) SetSourcePosition(mouse.x, mouse.y);
        if (enableCustomSource) UpdateCustomSource();
        if (MousePositionInGrid != null && Input.GetMouseButton(0) && !positioning)
        {
            if (!_isFirstDragFrame)
            {
                Vector2 mouseDelta = mouse - _prevMouseGridPos;
                float forceMagnitude = mouseDelta.magnitude * resolutionMultiplier;
                float scaledForce = Mathf.Pow(forceMagnitude, 1.5f) * 0.8f;
                AddForceToArea(mouse, mouseDelta.normalized * scaledForce, Mathf.Clamp(forceMagnitude * 0.5f, 2f, 10f));
            }
            _isFirstDragFrame = false;
            _prevMouseGridPos = mouse;
        }
        else _isFirstDragFrame = true;
        Step();
        UpdateVisualization();
        if (showStreamlines && colorMode != ColorMode.Streamlines) CombineTextures();
    }

    // UpdateVisualization (FluidSim.cs:755-866): UpdateVisualizationJob runs on the device, one RGBA plane comes back
    private Color[] colours;
    public void UpdateVisualization()
    {
        int n = currentSize;
        if (colours == null || colours.Length != n * n) colours = new Color[n * n];
        if (fluidTexture == null || fluidTexture.width != n) { fluidTexture = new Texture2D(n, n, TextureFormat.RGBAFloat, false); fluidTexture.filterMode = FilterMode.Point; }
        if (useLerp) fluidcolour = Color.Lerp(startColor, endColor, Mathf.PingPong(elapsedTime * 0.1f, 1f));   // :789-793
        var vp = new FsVisParams
        {
            colorMode = (int)colorMode, visualizeSourcePosition = visualizeSourcePosition ? 1 : 0, enableCustomSource = enableCustomSource ? 1 : 0,
            zSlice = currentDepth > 1 ? Mathf.Clamp(Mathf.RoundToInt(viewSliceZ * currentDepth), 0, currentDepth - 1) : 0,
            sourceX = sourcePositionX * n, sourceY = sourcePositionY * n, visualMarkerRadius = 3f,                 // :805-807
            colourIntensity = colourIntensity, mediumDensityThreshold = mediumDensityThreshold, highDensityThreshold = highDensityThreshold,
            lowPressureThreshold = lowPressureThreshold, highPressureThreshold = highPressureThreshold,
            fluidColor = fluidcolour, obstacleColor = obstacleColor, sourcePositionColor = sourcePositionColor,
            lowDensityColor = lowDensityColor, mediumDensityColor = mediumDensityColor, highDensityColor = highDensityColor,
            lowPressureColor = lowPressureColor, neutralPressureColor = neutralPressureColor, highPressureColor = highPressureColor,
            gradientColors = new Color[8], gradientTimes = new float[8]
        };
        if (colorMode == ColorMode.Gradient && colourGradient != null)                                            // :770-787
        {
            GradientColorKey[] keys = colourGradient.colorKeys;
            vp.gradientKeyCount = Mathf.Min(keys.Length, 8);
            for (int i = 0; i < vp.gradientKeyCount; i++) { vp.gradientColors[i] = keys[i].color; vp.gradientTimes[i] = keys[i].time; }
        }
        Native.Check(Native.fs_render_rgba(solver, ref vp, colours, (long)n * n * 4), solver);
        fluidTexture.SetPixels(colours);                                                                          // :849-850
        fluidTexture.Apply();
        if (showStreamlines || colorMode == ColorMode.Streamlines) DrawStreamlines();
        if (colorMode == ColorMode.Streamlines) CombineTextures();
    }

    void CombineTextures()                                                                                        // :868-884
    {
        Color[] fluidColors = fluidTexture.GetPixels(), streamColors = streamlineTexture.GetPixels();
        for (int i = 0; i < fluidColors.Length; i++) if (streamColors[i].a > 0) fluidColors[i] = streamColors[i];
        fluidTexture.SetPixels(fluidColors);
        fluidTexture.Apply();
    }

    // DrawStreamlines (FluidSim.cs:886-959): the two glyph jobs run on the device (fs_streamlines); Bresenham stays here
    private float[] segments;
    public void DrawStreamlines()
    {
        if (!showStreamlines && colorMode != ColorMode.Streamlines) return;
        int n = currentSize;
        int skip = Mathf.Max(1, n / (streamlineDensity * 10));
        int count = (n / skip) * (n / skip);
        if (streamlineTexture == null || streamlineTexture.width != n) { streamlineTexture = new Texture2D(n, n, TextureFormat.RGBA32, false); streamlineTexture.filterMode = FilterMode.Point; }
        if (segments == null || segments.Length != count * 4) segments = new float[count * 4];
        int z = currentDepth > 1 ? Mathf.Clamp(Mathf.RoundToInt(viewSliceZ * currentDepth), 0, currentDepth - 1) : 0;
        Native.Check(Native.fs_streamlines(solver, skip, streamlineScale, z, segments, count), solver);
        var colors = new Color[n * n];
        for (int i = 0; i < count; i++)
        {
            if (segments[4 * i] < 0) continue;                                                                     // :1772
            DrawBresenhamLine((int)segments[4 * i], (int)segments[4 * i + 1], (int)Math.Round(segments[4 * i + 2]), (int)Math.Round(segments[4 * i + 3]),
                              colors, streamlineColor, n, streamlineThickness);
        }
        streamlineTexture.SetPixels(colors);
        streamlineTexture.Apply();
    }

    static void DrawBresenhamLine(int x0, int y0, int x1, int y1, Color[] colors, Color lineColor, int size, float thickness) // :1783-1849
    {
        bool steep = Mathf.Abs(y1 - y0) > Mathf.Abs(x1 - x0);
        if (steep) { int t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
        if (x0 > x1) { int t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
        int dx = x1 - x0, dy = Mathf.Abs(y1 - y0), error = dx / 2, y = y0, ystep = (y0 < y1) ? 1 : -1;
        int halfThick = (int)Mathf.Floor(thickness / 2);
        for (int x = x0; x <= x1; x++)
        {
            for (int tx = -halfThick; tx <= halfThick; tx++)
                for (int ty = -halfThick; ty <= halfThick; ty++)
                {
                    int drawX = steep ? y + tx : x + tx, drawY = steep ? x + ty : y + ty;
                    if (drawX >= 0 && drawX < size && drawY >= 0 && drawY < size) colors[drawX + drawY * size] = lineColor;
                }
            error -= dy;
            if (error < 0) { y += ystep; error += dx; }
        }
    }

    /// <summary>Full fields on demand (the reference keeps them in managed arrays; here they live on the device).</summary>
    public void ReadFields()
    {
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
