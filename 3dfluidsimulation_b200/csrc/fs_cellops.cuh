// fs_cellops.cuh -- per-cell arithmetic of the stable-fluids step, shared by every kernel.
//
// Each function is the body of one reference Burst job (Assets/Scripts/FluidSim.cs) generalised to
// 3D as specified in DESIGN.md section 2, evaluated for ONE interior cell, plus the set_bnd
// (BoundaryJob, :1235-1289) ring cells that derive from that cell ("ring scatter"): the thread
// that owns interior cell (1,j,k) also writes face cell (0,j,k), and so on for edges and corners.
// fp32, reference association order, z terms appended last, true division.  Compiled with
// -fmad=false so that nvcc does not contract a*s + r into an FMA (Burst's strict float mode).
//
// The functions are __host__ __device__ and free of CUDA-only constructs so that the CPU test
// suite can compile them with g++ (tests/host_emul) and check the ring/obstacle logic against the
// oracle without a GPU.  The product only ever runs them on the device.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/fluidsolver.h"

#if defined(__CUDACC__)
#define FS_HD __host__ __device__ __forceinline__
#else
#define FS_HD inline
#endif

// Obstacle flag byte, built once per fs_set_obstacles (kernel build_flags):
//   bit0 self is obstacle; bit1/2 x-1/x+1 neighbour is obstacle; bit3/4 y-1/y+1; bit5/6 z-1/z+1.
// Out-of-grid neighbours count as fluid (they are never consulted for interior cells).
enum : uint8_t {
    FS_OB_SELF = 1, FS_OB_XM = 2, FS_OB_XP = 4, FS_OB_YM = 8, FS_OB_YP = 16, FS_OB_ZM = 32, FS_OB_ZP = 64
};

enum { FS_MODE_SMOOTH = 0, FS_MODE_JACOBI = 1 };
// fused two-stage sweeps (fs_kernels.cuh relax_pair): two smoother / Jacobi iterations, or the two colour passes of one
// red-black sweep, per pass over HBM
enum { FS_PAIR_SMOOTH = 0, FS_PAIR_JACOBI = 1, FS_PAIR_RED_BLACK = 2 };
// how a single sweep on z-slabs relates to the halo exchange (executor relax / relax_n; 0 / 1 also read as false / true)
enum { FS_X_NONE = 0, FS_X_EXCHANGE = 1, FS_X_EXTEND = 2, FS_X_EXCHANGE_OPEN = 3 };

// Ghost planes per internal slab side.  Two: an extended sweep (FS_X_EXTEND) or the first stage of a fused two-stage
// sweep is evaluated one plane beyond the owned range, which reads one plane further; every halo operation moves
// FS_GHOST planes each way, so two sweeps can share one operation.
#define FS_GHOST 2

// Geometry of one z-slab.  Local arrays hold planes [zoff, zoff+nzl) of the global grid; the slab
// owns local planes [kb, ke) and the rest (FS_GHOST per internal side) are ghosts filled by the halo exchange.
struct FsGrid {
    int nx, ny, nz;   // GLOBAL dimensions; N ("size") = nx
    int hz;           // nz > 1
    int zoff;         // global z of local plane 0
    int nzl;          // local planes allocated
    int kb, ke;       // owned local planes [kb, ke)
    long long sy, sz; // strides nx, nx*ny
};

FS_HD long long fs_idx(const FsGrid &g, int i, int j, int kl) { return i + j * g.sy + kl * g.sz; }

// Value of a set_bnd ring cell whose nearest interior cell holds v (pre-obstacle-mirroring value).
// fx/fy/fz: the ring cell lies on an x/y/z boundary plane.  b: field kind (1/2/3 negate across
// that axis).  Faces :1246-1252, 2D corners :1255-1258; 3D edges = 0.5*(two adjacent face cells),
// 3D corners = (three adjacent edge cells)/3, in x,y,z order.
FS_HD float fs_ring_value(float v, int fx, int fy, int fz, int b) {
    const float sx = b == 1 ? -v : v, sy = b == 2 ? -v : v, sz = b == 3 ? -v : v;
    const int n = fx + fy + fz;
    if (n == 0) return v;
    if (n == 1) return fx ? sx : (fy ? sy : sz);
    if (n == 2) {
        if (fx && fy) return 0.5f * (sy + sx); // (1,0,k) is a y-face cell, (0,1,k) an x-face cell
        if (fx && fz) return 0.5f * (sz + sx);
        return 0.5f * (sz + sy);
    }
    const float ex = 0.5f * (sz + sy), ey = 0.5f * (sz + sx), ez = 0.5f * (sy + sx);
    return ((ex + ey) + ez) / 3.0f;
}

// Calls emit(ii, jj, kl, fx, fy, fz) for interior cell (i,j,kl) itself (flags 0) and for every ring
// cell whose nearest interior cell it is.  k is the GLOBAL z of local plane kl.
template <class Emit>
FS_HD void fs_ring_scatter(const FsGrid &g, int i, int j, int kl, Emit emit) {
    const int k = kl + g.zoff;
    int xs[3], ys[3], zs[3], fxs[3], fys[3], fzs[3];
    int nxs = 0, nys = 0, nzs = 0;
    xs[nxs] = i; fxs[nxs++] = 0;
    if (i == 1) { xs[nxs] = 0; fxs[nxs++] = 1; }
    if (i == g.nx - 2) { xs[nxs] = g.nx - 1; fxs[nxs++] = 1; }
    ys[nys] = j; fys[nys++] = 0;
    if (j == 1) { ys[nys] = 0; fys[nys++] = 1; }
    if (j == g.ny - 2) { ys[nys] = g.ny - 1; fys[nys++] = 1; }
    zs[nzs] = kl; fzs[nzs++] = 0;
    if (g.hz) {
        if (k == 1) { zs[nzs] = kl - 1; fzs[nzs++] = 1; }
        if (k == g.nz - 2) { zs[nzs] = kl + 1; fzs[nzs++] = 1; }
    }
    for (int c = 0; c < nzs; c++)
        for (int bq = 0; bq < nys; bq++)
            for (int a = 0; a < nxs; a++) emit(xs[a], ys[bq], zs[c], fxs[a], fys[bq], fzs[c]);
}

// ---- obstacle mirroring fused into the sweeps -----------------------------------------------------------------
// BoundaryJob's obstacle pass (:1261-1287) after a relaxation sweep sets an interior obstacle cell of a velocity field
// (b = 1/2/3) to the mean of -x over its NON-obstacle neighbours along axis b, x being the values the sweep just wrote.
// Those two values are recomputed here from the sweep's inputs (a non-obstacle interior neighbour n gets
// (r[n] + a*nbsum(in, n))/c; a face cell of the ring gets set_bnd's value -v_self, v_self being the obstacle cell's own
// pre-mirror value), so the thread that owns the obstacle cell can write the mirrored value in the same launch: no
// separate mirror kernel and, on z-slabs, no extra halo round trip (the z neighbour of a boundary-plane cell lies in the
// ghost zone, whose FS_GHOST = 2 planes hold everything the recomputation reads).
template <int MODE>
FS_HD float fs_relax_new(const FsGrid &g, const float *in, const float *rhs, float a, float c, bool in_zero, long long n) {
    float s, r;
    if (in_zero) {
        s = ((0.0f + 0.0f) + 0.0f) + 0.0f;
        if (g.hz) s = (s + 0.0f) + 0.0f;
        r = rhs[n];
    } else {
        s = ((in[n + 1] + in[n - 1]) + in[n + g.sy]) + in[n - g.sy];
        if (g.hz) s = (s + in[n + g.sz]) + in[n - g.sz];
        r = MODE == FS_MODE_JACOBI ? rhs[n] : in[n];
    }
    return (r + a * s) / c;
}
// f: the cell's obstacle flag byte; v_self: its pre-mirror value (what the ring cells next to it are derived from).
template <int MODE>
FS_HD float fs_mirror_fused(const FsGrid &g, const float *in, const float *rhs, uint8_t f, float a, float c, int b,
                            bool in_zero, float v_self, int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    const long long step = b == 1 ? 1 : (b == 2 ? g.sy : g.sz);
    const uint8_t lo = b == 1 ? FS_OB_XM : (b == 2 ? FS_OB_YM : FS_OB_ZM);
    const uint8_t hi = b == 1 ? FS_OB_XP : (b == 2 ? FS_OB_YP : FS_OB_ZP);
    const int pos = b == 1 ? i : (b == 2 ? j : kl + g.zoff), last = b == 1 ? g.nx - 1 : (b == 2 ? g.ny - 1 : g.nz - 1);
    float m = 0.0f;
    int count = 0;
    if (!(f & lo)) { m += -(pos - 1 == 0 ? -v_self : fs_relax_new<MODE>(g, in, rhs, a, c, in_zero, idx - step)); count++; }
    if (!(f & hi)) { m += -(pos + 1 == last ? -v_self : fs_relax_new<MODE>(g, in, rhs, a, c, in_zero, idx + step)); count++; }
    return count > 0 ? m / (float)count : 0.0f;
}
FS_HD bool fs_mirrors(const FsGrid &g, int b) { return b != 0 && (b != 3 || g.hz); }

// ---- relaxation sweeps ---------------------------------------------------------------------------
// MODE SMOOTH: DiffuseJob :1045-1068  out = (in[c] + a*nbsum(in))/c, obstacle cells keep the stale
//              content of the output buffer (`stale`: x0 during the first two iterations because both
//              reference buffers start as copies of x0 (:1299-1300); afterwards the buffer itself,
//              signalled by stale == nullptr).
// MODE JACOBI: LinearSolveIterationJob :1200-1231  out = (rhs[c] + a*nbsum(in))/c, obstacle: copy.
// in_zero: the read buffer is identically zero (first pressure iteration, :1094/:1427) and is not loaded.
template <int MODE>
FS_HD void fs_relax_cell(const FsGrid &g, const float *in, const float *rhs, const float *stale, float *out,
                         const uint8_t *flags, float a, float c, int b, bool in_zero, int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    const bool obst = flags && (flags[idx] & FS_OB_SELF);
    float v;
    bool write_self = true;
    if (obst) {
        if (MODE == FS_MODE_JACOBI) {
            v = in_zero ? 0.0f : in[idx];
        } else if (stale) {
            v = stale[idx];
        } else {
            v = out[idx];
            write_self = false;
        }
    } else {
        float s;
        if (in_zero) {
            s = ((0.0f + 0.0f) + 0.0f) + 0.0f;
            if (g.hz) s = (s + 0.0f) + 0.0f;
        } else {
            s = ((in[idx + 1] + in[idx - 1]) + in[idx + g.sy]) + in[idx - g.sy];
            if (g.hz) s = (s + in[idx + g.sz]) + in[idx - g.sz];
        }
        const float r = MODE == FS_MODE_JACOBI ? rhs[idx] : in[idx];
        v = (r + a * s) / c;
    }
    fs_ring_scatter(g, i, j, kl, [&](int ii, int jj, int kk, int fx, int fy, int fz) {
        if (fx | fy | fz) out[fs_idx(g, ii, jj, kk)] = fs_ring_value(v, fx, fy, fz, b);
        else if (write_self) out[idx] = v;
    });
    // the obstacle pass of BoundaryJob, fused (the ring cells above were derived from the pre-mirror value, as in the
    // reference's order faces -> corners -> obstacles)
    if (obst && fs_mirrors(g, b)) out[idx] = fs_mirror_fused<MODE>(g, in, rhs, flags[idx], a, c, b, in_zero, v, i, j, kl);
}

// Red-black half sweep (colour = (i+j+k)&1 with GLOBAL k), in place, no ring writes; set_bnd is a
// separate pass after both colours (oracle fo_lin_solve_rb).
FS_HD void fs_rb_cell(const FsGrid &g, float *x, const float *rhs, const uint8_t *flags, float a, float c,
                      int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    if (flags && (flags[idx] & FS_OB_SELF)) return;
    float s = ((x[idx + 1] + x[idx - 1]) + x[idx + g.sy]) + x[idx - g.sy];
    if (g.hz) s = (s + x[idx + g.sz]) + x[idx - g.sz];
    x[idx] = (rhs[idx] + a * s) / c;
}

// set_bnd faces/edges/corners only, from the interior values already in x (used after red-black and
// by fs_op_set_bnd).  Interior cell is left untouched.
FS_HD void fs_bnd_cell(const FsGrid &g, float *x, int b, int i, int j, int kl) {
    const float v = x[fs_idx(g, i, j, kl)];
    fs_ring_scatter(g, i, j, kl, [&](int ii, int jj, int kk, int fx, int fy, int fz) {
        if (fx | fy | fz) x[fs_idx(g, ii, jj, kk)] = fs_ring_value(v, fx, fy, fz, b);
    });
}

// Obstacle mirroring of BoundaryJob :1261-1287 for one interior obstacle cell (b = 1, 2 or 3).
// Runs after the sweep that produced x (it reads the NEW values of fluid neighbours and ring cells).
FS_HD void fs_mirror_cell(const FsGrid &g, float *x, const uint8_t *flags, int b, long long idx) {
    const long long step = b == 1 ? 1 : (b == 2 ? g.sy : g.sz);
    const uint8_t f = flags[idx];
    const uint8_t lo = b == 1 ? FS_OB_XM : (b == 2 ? FS_OB_YM : FS_OB_ZM);
    const uint8_t hi = b == 1 ? FS_OB_XP : (b == 2 ? FS_OB_YP : FS_OB_ZP);
    float m = 0.0f;
    int count = 0;
    if (!(f & lo)) { m += -x[idx - step]; count++; }
    if (!(f & hi)) { m += -x[idx + step]; count++; }
    x[idx] = count > 0 ? m / (float)count : 0.0f;
}

// ---- projection -----------------------------------------------------------------------------------
// ProjectDivergenceJob :1080-1095 (obstacle cells included) + BoundaryJob(b=0) on div.  p is not
// written: the first pressure iteration runs with in_zero instead (:1094, :1427).
FS_HD void fs_divergence_cell(const FsGrid &g, float *div, const float *vx, const float *vy, const float *vz,
                              int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    float s = ((vx[idx + 1] - vx[idx - 1]) + vy[idx + g.sy]) - vy[idx - g.sy];
    if (g.hz) s = (s + vz[idx + g.sz]) - vz[idx - g.sz];
    const float v = -0.5f * s / (float)g.nx;
    fs_ring_scatter(g, i, j, kl, [&](int ii, int jj, int kk, int fx, int fy, int fz) {
        div[fs_idx(g, ii, jj, kk)] = fs_ring_value(v, fx, fy, fz, 0);
    });
}

// ProjectVelocityAdjustJob :1107-1122 + BoundaryJob(b=1/2/3) faces; in place.  Only the owning
// thread touches a velocity cell and its ring cells, so in-place is race free.
FS_HD void fs_gradient_cell(const FsGrid &g, float *vx, float *vy, float *vz, const float *p,
                            const uint8_t *flags, int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    const bool obst = flags && (flags[idx] & FS_OB_SELF);
    float nvx = vx[idx], nvy = vy[idx], nvz = g.hz ? vz[idx] : 0.0f;
    if (!obst) {
        nvx = nvx - 0.5f * (p[idx + 1] - p[idx - 1]) * (float)g.nx;
        nvy = nvy - 0.5f * (p[idx + g.sy] - p[idx - g.sy]) * (float)g.nx;
        if (g.hz) nvz = nvz - 0.5f * (p[idx + g.sz] - p[idx - g.sz]) * (float)g.nx;
    }
    fs_ring_scatter(g, i, j, kl, [&](int ii, int jj, int kk, int fx, int fy, int fz) {
        const long long o = fs_idx(g, ii, jj, kk);
        if (fx | fy | fz) {
            vx[o] = fs_ring_value(nvx, fx, fy, fz, 1);
            vy[o] = fs_ring_value(nvy, fx, fy, fz, 2);
            if (g.hz) vz[o] = fs_ring_value(nvz, fx, fy, fz, 3);
        } else if (!obst) {
            vx[o] = nvx;
            vy[o] = nvy;
            if (g.hz) vz[o] = nvz;
        }
    });
}

// ---- advection -------------------------------------------------------------------------------------
// AdvectJob :1138-1185: back-trace, clamp, bi/trilinear weights.
struct FsAdvectWeights {
    int i0, j0, k0;
    float s0, s1, t0, t1, u0, u1;
};

FS_HD FsAdvectWeights fs_advect_weights(const FsGrid &g, float dt0, float velx, float vely, float velz, int i, int j,
                                        int k) {
    FsAdvectWeights w;
    float x = (float)i - dt0 * velx;
    float y = (float)j - dt0 * vely;
    if (x < 0.5f) x = 0.5f;
    if (x > (float)g.nx - 1.5f) x = (float)g.nx - 1.5f;
    w.i0 = (int)x;
    if (y < 0.5f) y = 0.5f;
    if (y > (float)g.ny - 1.5f) y = (float)g.ny - 1.5f;
    w.j0 = (int)y;
    w.s1 = x - (float)w.i0;
    w.s0 = 1.0f - w.s1;
    w.t1 = y - (float)w.j0;
    w.t0 = 1.0f - w.t1;
    w.k0 = 0;
    w.u0 = 1.0f;
    w.u1 = 0.0f;
    if (g.hz) {
        float z = (float)k - dt0 * velz;
        if (z < 0.5f) z = 0.5f;
        if (z > (float)g.nz - 1.5f) z = (float)g.nz - 1.5f;
        w.k0 = (int)z;
        w.u1 = z - (float)w.k0;
        w.u0 = 1.0f - w.u1;
    }
    return w;
}

// `plane(kk)` returns the base pointer of GLOBAL plane kk of the advected field (it resolves slab ownership:
// local planes, or the neighbour slab's copy through peer memory), so that the 4 gathers of one plane
// share one address computation.
template <class Plane>
FS_HD float fs_advect_interp(const FsGrid &g, const FsAdvectWeights &w, Plane plane) {
    const long long o = w.i0 + w.j0 * g.sy;
    const float *p0 = plane(w.k0);
    const float lo = w.s0 * (w.t0 * p0[o] + w.t1 * p0[o + g.sy]) +
                     w.s1 * (w.t0 * p0[o + 1] + w.t1 * p0[o + g.sy + 1]); // :1183-1184
    if (!g.hz) return lo;
    const float *p1 = plane(w.k0 + 1);
    const float hi = w.s0 * (w.t0 * p1[o] + w.t1 * p1[o + g.sy]) +
                     w.s1 * (w.t0 * p1[o + 1] + w.t1 * p1[o + g.sy + 1]);
    return w.u0 * lo + w.u1 * hi;
}

// One scalar field (b = 0 density, or a single velocity component): obstacle cells get 0 because the
// reference's output array is freshly cleared (:1529, :1148-1156); ring via set_bnd(b).
template <class Samp>
FS_HD void fs_advect_cell(const FsGrid &g, float *d, Samp samp, const float *velx, const float *vely,
                          const float *velz, const uint8_t *flags, float dt0, int b, int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    float v = 0.0f;
    if (!(flags && (flags[idx] & FS_OB_SELF))) {
        const FsAdvectWeights w =
            fs_advect_weights(g, dt0, velx[idx], vely[idx], g.hz ? velz[idx] : 0.0f, i, j, kl + g.zoff);
        v = fs_advect_interp(g, w, samp);
    }
    fs_ring_scatter(g, i, j, kl, [&](int ii, int jj, int kk, int fx, int fy, int fz) {
        d[fs_idx(g, ii, jj, kk)] = fs_ring_value(v, fx, fy, fz, b);
    });
}

// Self-advection of the velocity (VelocityStep :710-711): all components share one back-trace
// because every AdvectWithJobs call there uses the same (velocityX0, velocityY0) as carrier.
template <class SampX, class SampY, class SampZ>
FS_HD void fs_advect_velocity_cell(const FsGrid &g, float *dx, float *dy, float *dz, SampX sampx, SampY sampy,
                                   SampZ sampz, const float *velx, const float *vely, const float *velz,
                                   const uint8_t *flags, float dt0, int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    if (!(flags && (flags[idx] & FS_OB_SELF))) {
        const FsAdvectWeights w =
            fs_advect_weights(g, dt0, velx[idx], vely[idx], g.hz ? velz[idx] : 0.0f, i, j, kl + g.zoff);
        ax = fs_advect_interp(g, w, sampx);
        ay = fs_advect_interp(g, w, sampy);
        if (g.hz) az = fs_advect_interp(g, w, sampz);
    }
    fs_ring_scatter(g, i, j, kl, [&](int ii, int jj, int kk, int fx, int fy, int fz) {
        const long long o = fs_idx(g, ii, jj, kk);
        dx[o] = fs_ring_value(ax, fx, fy, fz, 1);
        dy[o] = fs_ring_value(ay, fx, fy, fz, 2);
        if (g.hz) dz[o] = fs_ring_value(az, fx, fy, fz, 3);
    });
}

// ---- obstacle post-pass ---------------------------------------------------------------------------
// EnforceObstacleBoundaries + ApplyDragNearObstacle :617-673, per cell: obstacle -> V = 0; fluid cell
// -> V *= f(|V|) once per INTERIOR obstacle neighbour, |V| recomputed each time.
FS_HD float fs_drag_factor(float U, float cell, float rawvisc) {
    const float visc = rawvisc > 1e-5f ? rawvisc : 1e-5f;
    const float Re = (U * cell) / visc;
    float t = 1.0f - (float)exp((double)(-Re * 0.01f)); // Mathf.Exp is (float)Math.Exp((double)x)
    if (t < 0.0f) t = 0.0f;
    if (t > 1.0f) t = 1.0f;
    return 0.8f + (0.98f - 0.8f) * t;
}

FS_HD void fs_enforce_cell(const FsGrid &g, float *vx, float *vy, float *vz, const uint8_t *flags, float cell,
                           float rawvisc, int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    const uint8_t f = flags[idx];
    if (f == 0) return;
    if (f & FS_OB_SELF) {
        vx[idx] = 0.0f;
        vy[idx] = 0.0f;
        if (g.hz) vz[idx] = 0.0f;
        return;
    }
    const int k = kl + g.zoff;
    int n = 0;
    if ((f & FS_OB_XM) && i - 1 >= 1) n++;
    if ((f & FS_OB_XP) && i + 1 <= g.nx - 2) n++;
    if ((f & FS_OB_YM) && j - 1 >= 1) n++;
    if ((f & FS_OB_YP) && j + 1 <= g.ny - 2) n++;
    if (g.hz && (f & FS_OB_ZM) && k - 1 >= 1) n++;
    if (g.hz && (f & FS_OB_ZP) && k + 1 <= g.nz - 2) n++;
    if (n == 0) return;
    float ax = vx[idx], ay = vy[idx], az = g.hz ? vz[idx] : 0.0f;
    for (int r = 0; r < n; r++) {
        float q = ax * ax + ay * ay;
        if (g.hz) q = q + az * az;
        const float U = sqrtf(q); // == (float)sqrt((double)q): sqrt is correctly rounded in both
        const float fac = fs_drag_factor(U, cell, rawvisc);
        ax *= fac;
        ay *= fac;
        az *= fac;
    }
    vx[idx] = ax;
    vy[idx] = ay;
    if (g.hz) vz[idx] = az;
}

// Flag byte of one cell from the raw 0/1 mask (local array incl. ghost planes).
FS_HD uint8_t fs_flags_cell(const FsGrid &g, const uint8_t *mask, int i, int j, int kl) {
    const long long idx = fs_idx(g, i, j, kl);
    uint8_t f = mask[idx] ? FS_OB_SELF : 0;
    if (i > 0 && mask[idx - 1]) f |= FS_OB_XM;
    if (i < g.nx - 1 && mask[idx + 1]) f |= FS_OB_XP;
    if (j > 0 && mask[idx - g.sy]) f |= FS_OB_YM;
    if (j < g.ny - 1 && mask[idx + g.sy]) f |= FS_OB_YP;
    if (kl > 0 && mask[idx - g.sz]) f |= FS_OB_ZM;
    if (kl < g.nzl - 1 && mask[idx + g.sz]) f |= FS_OB_ZP;
    return f;
}

// ---- obstacle shapes (next row N4) --------------------------------------------------------------------------------
// IsInsideShape, FluidSim.cs:353-388, for cell (x, y[, z]).  The 2D expressions are the reference's, term by term;
// in 3D a Circle becomes a sphere (z term appended last) and Rectangle / Airfoil are extruded over
// center_z +- depth/2 (strict, like the reference's x / y tests).
FS_HD bool fs_shape_inside_xy(const fs_obstacle_shape &sh, int x, int y) { // Rectangle / Airfoil cross-section
    const float centerX = sh.center_x, centerY = sh.center_y;
    if (sh.kind == 1) {
        const float halfWidth = sh.width * 0.5f, halfHeight = sh.height * 0.5f;
        return x > (centerX - halfWidth) && x < (centerX + halfWidth) && y > (centerY - halfHeight) && y < (centerY + halfHeight);
    }
    const float chord = 2 * sh.width;
    const float thickness = 0.15f;
    const float normX = (x - centerX + chord / 2) / chord;
    const float normY = (y - centerY) / chord;
    if (normX < 0 || normX > 1 || fabsf(normY) > thickness) return false;
    const float halfThickness = 5 * thickness * (0.2969f * sqrtf(normX) - 0.1260f * normX - 0.3516f * normX * normX +
                                                 0.2843f * normX * normX * normX - 0.1015f * normX * normX * normX * normX);
    return fabsf(normY) <= halfThickness;
}
FS_HD bool fs_shape_inside_circle(const fs_obstacle_shape &sh, bool hz, int x, int y, int z) {
    float d2 = (x - sh.center_x) * (x - sh.center_x) + (y - sh.center_y) * (y - sh.center_y);
    if (hz) d2 = d2 + (z - sh.center_z) * (z - sh.center_z);
    return d2 < sh.radius * sh.radius;
}
FS_HD bool fs_shape_in_span(const fs_obstacle_shape &sh, bool hz, int z) {
    if (!hz) return true;
    const float halfDepth = sh.depth * 0.5f;
    return z > (sh.center_z - halfDepth) && z < (sh.center_z + halfDepth);
}
// Final mask value of cell (x, y, z) given the flood-filled cross-section `reach` (nx*ny bytes; Rectangle / Airfoil, and
// the 2D Circle) -- or, for the 3D sphere, the inside test itself gated by "the seed cell is inside".
FS_HD uint8_t fs_shape_mask(const fs_obstacle_shape &sh, bool hz, int nx, const uint8_t *reach, bool seed_ok, int x, int y, int z) {
    if (sh.kind == 0 && hz) return (seed_ok && fs_shape_inside_circle(sh, true, x, y, z)) ? 1 : 0;
    return (reach[x + (long long)y * nx] && fs_shape_in_span(sh, hz && sh.kind != 0, z)) ? 1 : 0;
}

// ---- visualisation colour mapping (next row N2) ------------------------------------------------------------------
// UpdateVisualizationJob.Execute, FluidSim.cs:1888-1979, for one cell of one xy plane.  Color.Lerp clamps t to
// [0,1]; Color.black = (0,0,0,1); the "very high pressure" colour is (1, 0.5, 0, 1) (:1962).
struct FsColor { float r, g, b, a; };
FS_HD FsColor fs_color4(const float *c) { FsColor o = {c[0], c[1], c[2], c[3]}; return o; }
FS_HD FsColor fs_color_lerp(FsColor a, FsColor b, float t) {
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    FsColor o = {a.r + (b.r - a.r) * t, a.g + (b.g - a.g) * t, a.b + (b.b - a.b) * t, a.a + (b.a - a.a) * t};
    return o;
}
FS_HD FsColor fs_eval_gradient(const fs_vis_params &vp, float time) { // :1981-2001
    const int n = vp.gradient_key_count;
    if (n <= 0) { FsColor w = {1.0f, 1.0f, 1.0f, 1.0f}; return w; }
    if (time <= vp.gradient_times[0]) return fs_color4(vp.gradient_colors[0]);
    if (time >= vp.gradient_times[n - 1]) return fs_color4(vp.gradient_colors[n - 1]);
    int index = 0;
    while (index < n - 1 && time > vp.gradient_times[index + 1]) index++;
    const float t = (time - vp.gradient_times[index]) / (vp.gradient_times[index + 1] - vp.gradient_times[index]);
    return fs_color_lerp(fs_color4(vp.gradient_colors[index]), fs_color4(vp.gradient_colors[index + 1]), t);
}
FS_HD FsColor fs_visualize_cell(const fs_vis_params &vp, float d, float p, bool obstacle, int i, int j) {
    if (obstacle) return fs_color4(vp.obstacle_color); // :1894-1899
    const float normalizedD = d * vp.colour_intensity;
    FsColor px;
    if (vp.color_mode == 2) { // DensityBased :1908-1928
        if (d < vp.medium_density_threshold) {
            const FsColor black = {0.0f, 0.0f, 0.0f, 1.0f};
            px = fs_color_lerp(black, fs_color4(vp.low_density_color), d / vp.medium_density_threshold);
        } else if (d < vp.high_density_threshold) {
            const float t = (d - vp.medium_density_threshold) / (vp.high_density_threshold - vp.medium_density_threshold);
            px = fs_color_lerp(fs_color4(vp.low_density_color), fs_color4(vp.medium_density_color), t);
        } else {
            float t = (d - vp.high_density_threshold) / vp.high_density_threshold;
            t = t < 1.0f ? t : 1.0f;
            px = fs_color_lerp(fs_color4(vp.medium_density_color), fs_color4(vp.high_density_color), t);
        }
    } else if (vp.color_mode == 1) { // Gradient :1930-1934
        const float c = normalizedD < 0.0f ? 0.0f : (normalizedD > 1.0f ? 1.0f : normalizedD);
        px = fs_eval_gradient(vp, c);
    } else if (vp.color_mode == 3) { // PressureBased :1947-1964
        if (p < vp.low_pressure_threshold) {
            const float t = p / vp.low_pressure_threshold;
            px = fs_color_lerp(fs_color4(vp.low_pressure_color), fs_color4(vp.neutral_pressure_color), 1.0f + t);
        } else if (p <= vp.high_pressure_threshold) {
            const float t = (p - vp.low_pressure_threshold) / (vp.high_pressure_threshold - vp.low_pressure_threshold);
            px = fs_color_lerp(fs_color4(vp.neutral_pressure_color), fs_color4(vp.high_pressure_color), t);
        } else {
            float t = (p - vp.high_pressure_threshold) / vp.high_pressure_threshold;
            t = t < 1.0f ? t : 1.0f;
            const FsColor orange = {1.0f, 0.5f, 0.0f, 1.0f};
            px = fs_color_lerp(fs_color4(vp.high_pressure_color), orange, t);
        }
    } else { // SingleColor and default (Streamlines) :1936-1945
        const FsColor c = {vp.fluid_color[0] * normalizedD, vp.fluid_color[1] * normalizedD, vp.fluid_color[2] * normalizedD,
                           vp.fluid_color[3]};
        px = c;
    }
    if (vp.visualize_source_position && vp.enable_custom_source) { // :1970-1978
        const float dx = (float)i - vp.source_x, dy = (float)j - vp.source_y;
        const float distSq = dx * dx + dy * dy;
        if (distSq < vp.visual_marker_radius * vp.visual_marker_radius) px = fs_color4(vp.source_position_color);
    }
    return px;
}

// ---- streamline glyphs (next row N3) ---------------------------------------------------------------------------
// StreamlineCalculationJob.Execute (:1680-1727) followed by StreamlineDrawJob.Execute (:1739-1762) for glyph `index`
// of one xy plane: out = (startX, startY, endX, endY), or (-1,-1,-1,-1) for a glyph the reference marks invalid
// (outside the interior, obstacle cell, |V| < 0.01).  cols = nx / skip generalises the reference's size / skip.
FS_HD void fs_streamline_glyph(int nx, int ny, int skip, float scale, const float *vx, const float *vy,
                               const uint8_t *mask, int index, float out[4]) {
    const int cols = nx / skip;
    const int x = index % cols, y = index / cols;
    const int i = x * skip + skip, j = y * skip + skip;
    out[0] = out[1] = out[2] = out[3] = -1.0f;
    if (i <= 0 || i >= nx - 1 || j <= 0 || j >= ny - 1) return;
    const long long idx = i + (long long)j * nx;
    if (mask[idx]) return;
    const float ux = vx[idx], uy = vy[idx];
    const float magnitude = sqrtf(ux * ux + uy * uy);
    if (magnitude < 0.01f) return;
    const float cap = (float)(skip - 1), want = magnitude * scale;
    const float lineLength = cap < want ? cap : want; // math.min(skip - 1, magnitude * streamlineScale)
    if (lineLength <= 0.0f) return;                   // StreamlineDrawJob: w <= 0 is invalid
    const float angle = atan2f(uy, ux);
    out[0] = (float)i;
    out[1] = (float)j;
    out[2] = (float)i + cosf(angle) * lineLength;
    out[3] = (float)j + sinf(angle) * lineLength;
}
