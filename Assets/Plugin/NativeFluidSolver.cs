// NativeFluidSolver.cs -- P/Invoke binding of include/fluidsolver.h (libfluidsolver.so / fluidsolver.dll).
//
// Lives where the reference keeps its only other native plugin (Assets/Plugin/sqlite3.dll, consumed by
// Mono.Data.Sqlite): the native library goes to Assets/Plugin/x86_64/libfluidsolver.so, cdecl, one entry
// point per line of the header.  NOT compiled in the build container (no C# toolchain there); the same
// entry points are exercised by the Python ctypes binding (3dfluidsimulation_b200/native.py) in the tests.
//
// Marshalling notes
//   * float[] / byte[] are blittable: the marshaller pins them for the duration of the call, which is all the
//     plugin requires (host pointers are only touched during the call).
//   * bool[] is NOT blittable (4-byte BOOL by default): the obstacle mask crosses the boundary as byte[].
//   * the solver handle is an opaque IntPtr wrapped in a SafeHandle so that a domain reload frees the GPU memory.
using System;
using System.Runtime.InteropServices;
using UnityEngine;

namespace FluidSolverNative
{
    public enum FsStatus { Ok = 0, BadArgument = -1, Cuda = -2, OutOfMemory = -3, Unsupported = -4, Comm = -5 }

    public enum FsField
    {
        Density = 0, VelocityX = 1, VelocityY = 2, VelocityZ = 3,
        VelocityX0 = 4, VelocityY0 = 5, VelocityZ0 = 6, Pressure = 7, Divergence = 8
    }

    public enum FsSolverKind { Jacobi = 0, RedBlack = 1 }

    [StructLayout(LayoutKind.Sequential)]
    public struct FsParams
    {
        public int abiVersion;      // FS_ABI_VERSION = 1
        public int nx, ny, nz;      // nz == 1: the reference's 2D solver
        public int itersDiffuse;    // 20 in the reference
        public int itersPressure;   // 20 in the reference
        public int solverKind;      // FsSolverKind
        public int enableObstacle;
        public float cellSize;      // physicalSize / currentSize
        public float rawViscosity;  // the drag uses the unscaled viscosity
        public int deviceId;
        public int slabRank, slabCount;
        public int useCudaGraph;
        public int reserved0, reserved1, reserved2, reserved3;
        public const int NativeSize = 72;
    }

    /// <summary>fs_vis_params (include/fluidsolver.h): the parameters UpdateVisualization passes to UpdateVisualizationJob
    /// (FluidSim.cs:799-829).  Field for field and in the same order as the C struct; NativeSize / offsets are checked
    /// against a compiled C probe by tests/test_abi_and_host.py::test_csharp_struct_layouts.</summary>
    [StructLayout(LayoutKind.Sequential)]
    public struct FsVisParams
    {
        public const int NativeSize = 356;
        public int colorMode;                 // offset 0   ColorMode: 0 SingleColor, 1 Gradient, 2 DensityBased, 3 PressureBased, 4 Streamlines
        public int visualizeSourcePosition;   // offset 4
        public int enableCustomSource;        // offset 8
        public int gradientKeyCount;          // offset 12  0..8
        public int zSlice;                    // offset 16
        public float sourceX, sourceY;        // offset 20, 24   sourcePosition * currentSize
        public float visualMarkerRadius;      // offset 28
        public float colourIntensity;         // offset 32
        public float mediumDensityThreshold, highDensityThreshold; // offset 36, 40
        public float lowPressureThreshold, highPressureThreshold; // offset 44, 48
        public Color fluidColor, obstacleColor, sourcePositionColor;        // offset 52, 68, 84   (Color = 4 floats r,g,b,a)
        public Color lowDensityColor, mediumDensityColor, highDensityColor; // offset 100, 116, 132
        public Color lowPressureColor, neutralPressureColor, highPressureColor; // offset 148, 164, 180
        [MarshalAs(UnmanagedType.ByValArray, SizeConst = 8)] public Color[] gradientColors; // offset 196
        [MarshalAs(UnmanagedType.ByValArray, SizeConst = 8)] public float[] gradientTimes;  // offset 324
    }

    /// <summary>fs_obstacle_shape: SetupObstacles / IsInsideShape parameters in cells (FluidSim.cs:302-388).</summary>
    [StructLayout(LayoutKind.Sequential)]
    public struct FsObstacleShape
    {
        public const int NativeSize = 56;
        public int kind;                       // offset 0   0 Circle, 1 Rectangle, 2 Airfoil
        public float centerX, centerY, centerZ; // offset 4, 8, 12
        public float radius;                   // offset 16
        public float width, height;            // offset 20, 24
        public float depth;                    // offset 28
        public int seedX, seedY, seedZ;        // offset 32, 36, 40
        public int reserved0, reserved1, reserved2; // offset 44
    }

    public sealed class SolverHandle : SafeHandle
    {
        public SolverHandle() : base(IntPtr.Zero, true) { }
        public override bool IsInvalid => handle == IntPtr.Zero;
        protected override bool ReleaseHandle() { Native.fs_destroy(handle); return true; }
    }

    public static class Native
    {
        private const string Lib = "fluidsolver";
        private const CallingConvention CC = CallingConvention.Cdecl;
        public const int AbiVersion = 1;

        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_abi_version();
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_create(ref FsParams p, out SolverHandle solver);
        [DllImport(Lib, CallingConvention = CC)] public static extern void fs_destroy(IntPtr solver);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_reset(SolverHandle s);
        [DllImport(Lib, CallingConvention = CC)] public static extern IntPtr fs_last_error(SolverHandle s);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_slab_range(SolverHandle s, out int zBegin, out int zEnd, out long ownedVoxels);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_set_obstacles(SolverHandle s, byte[] mask, long n);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_slab_halo_range(SolverHandle s, out int zBegin, out int zEnd);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_set_obstacles_slab(SolverHandle s, byte[] mask, long n, int globalAny, int globalInterior);
        // device-side SetupObstacles + RecursiveFloodFill (FluidSim.cs:302-388): no mask crosses the boundary
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_build_obstacles(SolverHandle s, ref FsObstacleShape shape, out long obstacleCells);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_get_obstacles(SolverHandle s, [Out] byte[] dst, long n);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_add_density(SolverHandle s, float x, float y, float z, float amount);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_add_velocity(SolverHandle s, float x, float y, float z, float ax, float ay, float az);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_add_source_cells(SolverHandle s, long count, float[] x, float[] y, float[] z, float[] density, float[] ax, float[] ay, float[] az);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_add_sources(SolverHandle s, float[] density, float[] vx, float[] vy, float[] vz);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_step(SolverHandle s, float dt, float visc, float diff);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_sync(SolverHandle s);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_get_field(SolverHandle s, int field, float[] dst, long n);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_set_field(SolverHandle s, int field, float[] src, long n);
        // pipelined readback: dst must be pinned (GCHandle.Alloc(..., GCHandleType.Pinned)) until fs_wait_transfers returns
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_get_field_async(SolverHandle s, int field, IntPtr dst, long n);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_wait_transfers(SolverHandle s);
        // device-side UpdateVisualizationJob: one xy plane as RGBA floats, straight into the Color[] the reference fills (:849)
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_render_rgba(SolverHandle s, ref FsVisParams visParams, [Out] Color[] outRgba, long n);
        // device-side StreamlineCalculationJob + StreamlineDrawJob: count*4 floats (x0, y0, x1, y1), -1 = invalid glyph
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_streamlines(SolverHandle s, int skip, float scale, int zSlice, [Out] float[] segments, long count);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_get_metrics(SolverHandle s, out float meanDensity, out float maxSpeed, out double sumDensity);

        // operator entry points (one reference job chain each); used by tests and by hosts that compose their own step
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_op_set_bnd(SolverHandle s, int field, int b);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_op_diffuse(SolverHandle s, int dst, int src, int b, float diff, float dt);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_op_smooth(SolverHandle s, int dst, int src, int b, float a, float c, int iters);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_op_lin_solve(SolverHandle s, int dst, int rhs, int b, float a, float c, int iters, int solverKind);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_op_project(SolverHandle s, int useV0Fields);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_op_advect(SolverHandle s, int dst, int src, int b, int useV0Fields, float dt);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_op_advect_velocity(SolverHandle s, float dt);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_op_enforce_obstacles(SolverHandle s);

        // measurement + multi-GPU wiring
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_timer_start(SolverHandle s);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_timer_stop(SolverHandle s, out float elapsedMs);
        [DllImport(Lib, CallingConvention = CC)] public static extern long fs_launch_count(SolverHandle s);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_bench_sweep(SolverHandle s, int kindAndFill, int b, int reps, out float avgMs, out double algoBytes);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_selftest_division(SolverHandle s, float divisor, ulong firstBits, ulong count, out ulong mismatches);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_halo_export(SolverHandle s, byte[] blob, long blobBytes);
        [DllImport(Lib, CallingConvention = CC)] public static extern int fs_halo_connect(SolverHandle s, byte[] lowerBlob, byte[] upperBlob, int sameProcess);

        /// <summary>Turns a negative status into an exception carrying fs_last_error (never thrown by the native side).</summary>
        public static void Check(int status, SolverHandle s)
        {
            if (status == 0) return;
            IntPtr msg = fs_last_error(s);
            throw new InvalidOperationException($"fluidsolver: {(FsStatus)status}: {Marshal.PtrToStringAnsi(msg)}");
        }
    }
}
