/*
 * fluidsolver.h -- C ABI of libfluidsolver.so, the B200 (sm_100a) native plugin that replaces the
 * solver hot path of ChrisWangstpauls/3DFluidSimulation (Assets/Scripts/FluidSim.cs).
 *
 * The reference has no plugin boundary for the solver: it is private methods on a MonoBehaviour.
 * The boundary is therefore cut at the reference's internal seam: every
 * entry point below names the FluidSim.cs member(s) it replaces.  Plain C types only; cdecl; the
 * library never throws across the ABI.  The reference-side binding (C# P/Invoke) is
 * Assets/Plugin/NativeFluidSolver.cs, described in INTEGRATION.md.
 *
 * Conventions
 *   - return value: FS_OK (0) or a negative fs_status; fs_last_error() gives the message.
 *   - layout: fp32, idx = x + y*nx + z*nx*ny (FluidSim.cs:749-752 plus a z stride).
 *   - nz == 1 selects the reference's 2D solver exactly (z terms skipped, coefficients unchanged).
 *   - "N" (the reference's `size`/`currentSize`) is nx.
 *   - host pointers are only read/written during the call; the plugin owns all device memory.
 *   - one host thread per handle; distinct handles are independent.
 *   - fs_step is asynchronous (returns after enqueueing); fs_get_* / fs_sync synchronise.
 */
#ifndef FLUIDSOLVER_H
#define FLUIDSOLVER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS_ABI_VERSION 1

typedef enum fs_status {
    FS_OK = 0,
    FS_ERR_BAD_ARGUMENT = -1,
    FS_ERR_CUDA = -2,
    FS_ERR_OUT_OF_MEMORY = -3,
    FS_ERR_UNSUPPORTED = -4,
    FS_ERR_COMM = -5
} fs_status;

/* Field ids for fs_get_field / fs_set_field.  FluidSim.cs:112-117 (+ z components). */
typedef enum fs_field {
    FS_DENSITY = 0,  /* density      :112 */
    FS_VX = 1,       /* velocityX    :113 */
    FS_VY = 2,       /* velocityY    :114 */
    FS_VZ = 3,       /* (3D only) */
    FS_VX0 = 4,      /* velocityX0   :115  scratch: contents after fs_step are unspecified */
    FS_VY0 = 5,      /* velocityY0   :116  scratch */
    FS_VZ0 = 6,      /* scratch */
    FS_PRESSURE = 7, /* pressure     :117, written by ProjectWithJobs :1509 */
    FS_DIVERGENCE = 8, /* nativeDiv of the last projection (:1428), for diagnostics/tests */
    FS_FIELD_COUNT = 9
} fs_field;

typedef enum fs_solver_kind {
    FS_JACOBI = 0,    /* the reference's double-buffered relaxation (FluidSim.cs:1188-1233) */
    FS_RED_BLACK = 1  /* red-black Gauss-Seidel for the pressure solve (BASELINE config 5) */
} fs_solver_kind;

/* Creation parameters.  Replaces ResetSimulation's sizing/allocation, FluidSim.cs:213-235. */
typedef struct fs_params {
    int32_t abi_version;     /* must be FS_ABI_VERSION */
    int32_t nx, ny, nz;      /* GLOBAL grid; the reference uses nx == ny == currentSize, nz == 1 */
    int32_t iters_diffuse;   /* iterations of each Diffuse pass; 20 in the reference (:1310, :1378) */
    int32_t iters_pressure;  /* iterations of the pressure solve; 20 in the reference (:1594) */
    int32_t solver_kind;     /* fs_solver_kind */
    int32_t enable_obstacle; /* FluidSim.cs:97, :567: run the obstacle post-pass inside fs_step */
    float cell_size;         /* FluidSim.cs:219  physicalSize / currentSize */
    float raw_viscosity;     /* FluidSim.cs:664: the drag uses the UNSCALED viscosity */
    int32_t device_id;       /* CUDA device ordinal of this handle */
    int32_t slab_rank;       /* z-slab decomposition: this handle owns slab slab_rank ... */
    int32_t slab_count;      /* ... of slab_count (1 = whole grid on one GPU) */
    int32_t use_cuda_graph;  /* 1: capture fs_step's launch sequence in a CUDA graph and replay it */
    int32_t reserved[4];     /* must be 0 */
} fs_params;

typedef struct fs_solver fs_solver; /* opaque */

/* ---- lifetime: ResetSimulation/ResetJobBuffers/OnDestroy, FluidSim.cs:213-235, :990-1032 ---- */
int fs_create(const fs_params *params, fs_solver **out);
void fs_destroy(fs_solver *s);
int fs_reset(fs_solver *s); /* all fields and the mask to zero (:225-232) */
const char *fs_last_error(const fs_solver *s); /* s may be NULL: error of the last failed fs_create */
int fs_abi_version(void);

/* ---- geometry of this handle's slab ----------------------------------------------------------- */
/* Owned global z range [z_begin, z_end) and the number of owned voxels nx*ny*(z_end-z_begin). */
int fs_slab_range(const fs_solver *s, int32_t *z_begin, int32_t *z_end, int64_t *owned_voxels);

/* ---- obstacles: SetupObstacles writes bool[] obstacles, FluidSim.cs:302-327 -------------------
 * mask: one byte per cell of the GLOBAL grid (0 = fluid), n = nx*ny*nz.  C# bool[] is not
 * blittable: pass byte[]. */
int fs_set_obstacles(fs_solver *s, const uint8_t *mask, int64_t n);

/* Same, for a handle that only knows its own slab (no rank needs the GLOBAL mask: 1 GiB at 1024^3): mask holds the
 * planes [z0, z1) reported by fs_slab_halo_range (the owned planes plus the ghost planes), n = nx*ny*(z1-z0).
 * global_any / global_interior: whether ANY slab has an obstacle cell / an obstacle cell off the domain ring -- every
 * slab must pass the same values (they select which halo operations exist). */
int fs_slab_halo_range(const fs_solver *s, int32_t *z_begin, int32_t *z_end);
int fs_set_obstacles_slab(fs_solver *s, const uint8_t *mask, int64_t n, int32_t global_any, int32_t global_interior);

/* ---- obstacle mask builder on the device: SetupObstacles + RecursiveFloodFill +
 * IsInsideShape, FluidSim.cs:302-388, generalised to 3D.  Each handle builds ONLY its own slab; no mask crosses the ABI.
 *   kind 0 Circle    (x-cx)^2 + (y-cy)^2 [+ (z-cz)^2 in 3D: a sphere] < radius^2                        :360-361
 *   kind 1 Rectangle x, y strictly inside center +- width/2, height/2 [3D: and z inside center_z +- depth/2: a box] :363-367
 *   kind 2 Airfoil   the reference's NACA-0015 approximation, chord = 2*width [3D: extruded over depth]  :369-383
 * and the reference's 4-neighbour flood fill from the seed cell (RoundToInt(position*size), :308-309): only cells
 * of the shape that are connected to the seed become obstacles (nothing if the seed itself is outside).  In 3D the
 * fill runs on the xy cross-section and is extruded (Rectangle / Airfoil); a sphere is its own fill.
 * All lengths are in cells, computed by the host exactly as the reference does (obstacleRadius * currentSize, ...). */
typedef struct fs_obstacle_shape {
    int32_t kind;
    float center_x, center_y, center_z; /* obstaclePositionX/Y * currentSize (:355-356); Z: * depth */
    float radius;                        /* Circle: obstacleRadius * currentSize (:316) */
    float width, height;                 /* obstacleWidth * currentSize, obstacleHeight * currentSize */
    float depth;                         /* 3D Rectangle / Airfoil: extent in z, cells; ignored otherwise */
    int32_t seed_x, seed_y, seed_z;      /* flood fill start */
    int32_t reserved[3];                 /* must be 0 */
} fs_obstacle_shape;
/* obstacle_cells (may be NULL): number of obstacle cells of the GLOBAL grid. */
int fs_build_obstacles(fs_solver *s, const fs_obstacle_shape *shape, int64_t *obstacle_cells);
/* The mask of this handle's OWNED planes as the device holds it (n = owned voxels): what UpdateVisualization reads (:765). */
int fs_get_obstacles(fs_solver *s, uint8_t *out, int64_t n);

/* ---- sources: AddDensity / AddVelocity, FluidSim.cs:723-738 (cell = clamp((int)coord)) --------
 * Global coordinates; a handle ignores cells outside its slab.  z, az ignored when nz == 1. */
int fs_add_density(fs_solver *s, float x, float y, float z, float amount);
int fs_add_velocity(fs_solver *s, float x, float y, float z, float ax, float ay, float az);
/* Batched form of the loops in UpdateCustomSource/AddForceToArea (FluidSim.cs:452-533): count
 * cells, arrays of length count (vx/vy/vz/density amounts may be NULL). */
int fs_add_source_cells(fs_solver *s, int64_t count, const float *x, const float *y, const float *z,
                        const float *density, const float *ax, const float *ay, const float *az);
/* Dense add: field[i] += src[i] over this handle's OWNED voxels (any pointer may be NULL). */
int fs_add_sources(fs_solver *s, const float *density, const float *vx, const float *vy, const float *vz);

/* ---- the step: Simulate(), FluidSim.cs:551-570 = VelocityStep :703-714 + DensityStep :716-721
 * + EnforceObstacleBoundaries :617-673 when enable_obstacle.  dt/visc/diff are the already scaled
 * effective values of :554-556 (the scaling stays on the C# side). */
int fs_step(fs_solver *s, float dt, float visc, float diff);
int fs_sync(fs_solver *s);

/* ---- field access: what UpdateVisualization / DrawStreamlines / tests read (:761-768, :919) ----
 * n must equal the handle's owned voxel count; planes are in global z order. */
int fs_get_field(fs_solver *s, int32_t field, float *out, int64_t n);
int fs_set_field(fs_solver *s, int32_t field, const float *in, int64_t n);
/* Pipelined form of fs_get_field for per-frame readback (what Update() does after Simulate(), :443): the field is
 * snapshotted in stream order and copied to `out` on a copy stream while later steps compute.  `out` must stay
 * valid (pinned for full speed) and must not be read until fs_wait_transfers() returns.  One transfer per field id
 * may be in flight; a second call for the same field waits for the first on the device, never on the host. */
int fs_get_field_async(fs_solver *s, int32_t field, float *out, int64_t n);
int fs_wait_transfers(fs_solver *s);

/* ---- visualisation colour mapping: UpdateVisualizationJob, FluidSim.cs:1851-2002,
 * with the parameters UpdateVisualization passes (:799-829).  Renders one xy plane (the whole grid when nz == 1,
 * plane z_slice otherwise) to RGBA floats in the reference's Color[] layout, on the device, so that a frame needs
 * nx*ny*16 bytes of readback instead of the density and pressure fields.  colour arrays are r,g,b,a. */
typedef struct fs_vis_params {
    int32_t color_mode;                /* ColorMode :32: 0 SingleColor, 1 Gradient, 2 DensityBased, 3 PressureBased, 4 Streamlines */
    int32_t visualize_source_position; /* :54 */
    int32_t enable_custom_source;      /* :35 */
    int32_t gradient_key_count;        /* 0..8 */
    int32_t z_slice;                   /* global z plane to render (ignored when nz == 1); must be owned by this handle */
    float source_x, source_y;          /* sourcePosition * currentSize :805-806 */
    float visual_marker_radius;        /* 3 in the reference :807 */
    float colour_intensity;
    float medium_density_threshold, high_density_threshold;
    float low_pressure_threshold, high_pressure_threshold;
    float fluid_color[4], obstacle_color[4], source_position_color[4];
    float low_density_color[4], medium_density_color[4], high_density_color[4];
    float low_pressure_color[4], neutral_pressure_color[4], high_pressure_color[4];
    float gradient_colors[8][4];
    float gradient_times[8];
} fs_vis_params;
/* out_rgba: nx*ny*4 floats, n = nx*ny*4. */
int fs_render_rgba(fs_solver *s, const fs_vis_params *vp, float *out_rgba, int64_t n);

/* ---- streamline glyphs: StreamlineCalculationJob + StreamlineDrawJob,
 * FluidSim.cs:1668-1763, on the device for plane z_slice (ignored when nz == 1).  skip = max(1, size/(streamlineDensity*10))
 * and scale = streamlineScale as in DrawStreamlines (:892, :925).  out_segments: count*4 floats
 * (startX, startY, endX, endY), (-1,-1,-1,-1) for glyphs the reference marks invalid; count must equal
 * (nx/skip)*(ny/skip).  The Bresenham drawing of the segments (:1765-1849) stays on the host, as in the reference. */
int fs_streamlines(fs_solver *s, int32_t skip, float scale, int32_t z_slice, float *out_segments, int64_t count);

/* ---- metrics: LogCurrentMetrics, FluidSim.cs:582-594 (mean density, max |V|) over owned voxels;
 * sum_density is returned so that slabs can be combined. */
int fs_get_metrics(fs_solver *s, float *mean_density, float *max_speed, double *sum_density);

/* ---- operator entry points (one reference job chain each) --------------------------------------
 * They act on the handle's fields and are what the parity tests drive kernel by kernel.
 *   fs_op_set_bnd       BoundaryJob                          :1235-1289
 *   fs_op_diffuse       Diffuse = DiffuseWithJobs + LinearSolveWithJobs   :740-745, :1292-1415
 *   fs_op_smooth        DiffuseWithJobs only (pass 1)        :1292-1357
 *   fs_op_lin_solve     LinearSolveWithJobs only (pass 2)    :1359-1415 (dst holds the initial guess)
 *   fs_op_project       ProjectWithJobs on (vx,vy,vz) fields :1417-1521 (+ :1578-1637)
 *   fs_op_advect        AdvectWithJobs                       :1523-1576
 *   fs_op_enforce_obstacles  EnforceObstacleBoundaries       :617-673
 * b: 0 scalar, 1/2/3 velocity component.  dst/src are fs_field ids; dst != src. */
int fs_op_set_bnd(fs_solver *s, int32_t field, int32_t b);
int fs_op_diffuse(fs_solver *s, int32_t dst, int32_t src, int32_t b, float diff, float dt);
int fs_op_smooth(fs_solver *s, int32_t dst, int32_t src, int32_t b, float a, float c, int32_t iters);
int fs_op_lin_solve(fs_solver *s, int32_t dst, int32_t rhs, int32_t b, float a, float c, int32_t iters,
                    int32_t solver_kind);
int fs_op_project(fs_solver *s, int32_t use_v0_fields); /* 0: (VX,VY,VZ); 1: (VX0,VY0,VZ0) */
int fs_op_advect(fs_solver *s, int32_t dst, int32_t src, int32_t b, int32_t use_v0_fields, float dt);
int fs_op_advect_velocity(fs_solver *s, float dt); /* (VX,VY,VZ) <- advect (VX0,VY0,VZ0) along itself, :710-711 */
int fs_op_enforce_obstacles(fs_solver *s);

/* ---- measurement helpers (CUDA events on the solver's own stream) ------------------------------ */
int fs_timer_start(fs_solver *s);
int fs_timer_stop(fs_solver *s, float *elapsed_ms); /* records, synchronises, returns the interval */
int64_t fs_launch_count(const fs_solver *s);        /* kernels launched by this handle so far */
/* Average duration of `reps` back-to-back launches of one kernel of the step, for the roofline figures.
 * kind_and_fill = kind + 16*fill.  Relaxation sweeps on the scratch fields (in = VX0, rhs = VY0, out = ping-pong):
 *   0 pass-1 smoother, 1 Jacobi, 2 red-black full sweep as two colour launches,
 *   7 fused Jacobi pair (two iterations per launch), 8 fused smoother pair, 9 fused red-black full sweep;
 * once-per-step kernels on the live fields (call them after the timed steps; scratch fields are overwritten):
 *   3 advect of one scalar (density), 4 fused advect of the velocity components, 5 divergence, 6 gradient subtract.
 * fill 0 = operands as the last step left them, 1 = overwrite the sweep operands with uniform random numbers first,
 * 2 = zeros.  Also returns the algorithmic bytes per launch (SURVEY.md section 8d per-voxel figures x owned voxels). */
#define FS_BENCH_KIND_MAX 9
int fs_bench_sweep(fs_solver *s, int32_t kind_and_fill, int32_t b, int32_t reps, float *avg_ms, double *algo_bytes);

/* Self-test of the sweep kernels' constant-divisor division (csrc/fs_kernels.cuh fs_div) against IEEE
 * division, bit for bit, over numerator bit patterns [first_bits, first_bits + count).  Returns the
 * number of mismatching quotients (0 expected). */
int fs_selftest_division(fs_solver *s, float divisor, uint64_t first_bits, uint64_t count, uint64_t *mismatches);

/* ---- multi-GPU wiring (z-slabs, one handle per GPU) ------------------- */
#define FS_IPC_BLOB_BYTES 1024
/* Export this handle's halo mailbox as an opaque blob (CUDA IPC handles inside). */
int fs_halo_export(fs_solver *s, void *blob, int64_t blob_bytes);
/* Connect the lower (slab_rank-1) and upper (slab_rank+1) neighbours; NULL for none. same_process=1
 * when the neighbour handle lives in this process (pointers are passed instead of IPC handles). */
int fs_halo_connect(fs_solver *s, const void *lower_blob, const void *upper_blob, int32_t same_process);

#ifdef __cplusplus
}
#endif
#endif /* FLUIDSOLVER_H */
