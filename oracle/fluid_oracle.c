/*
 * fluid_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the stable-fluids hot path: the 3D generalisation of the
 * reference's 2D solver (Assets/Scripts/FluidSim.cs), as specified in
 * SURVEY.md section 8a ("3D generalisation of each row") and DESIGN.md
 * section 2.  With nz == 1 it IS the reference's 2D algorithm: every z term
 * is skipped (not added as zero) and the result is bit-identical to the
 * literal restatement in oracle/ref2d.c (test K9).
 *
 * PARITY UNPINNED by the reference (no C# toolchain here, no reference
 * tests/golden vectors exist): see the header of ref2d.c.  This file is
 * pinned by ref2d.c (nz == 1), by the numpy restatement's golden vectors
 * under tests/golden/ and by the hand-derived known-answer tests.
 *
 * Conventions
 *   layout      idx = x + y*nx + z*nx*ny, x fastest (FluidSim.cs:749-752 + z)
 *   N ("size")  = nx; it is the scalar in a = dt*diff*(N-2)^2 (:743),
 *                 dt0 = dt*(N-2) (:1526) and the divergence / gradient scale (:1092, :1120)
 *   b           0 scalar, 1 x-velocity, 2 y-velocity, 3 z-velocity (set_bnd sign)
 *   sums        reference association first, z terms appended last:
 *                 (((R+L)+T)+B)+U)+D ; divergence ((((xR-xL)+yT)-yB)+zU)-zD
 *   edges       3D only: 0.5*(two adjacent face cells); corners (3D): (fx+fy+fz)/3
 *                 of the three adjacent edge cells; 2D corners as :1255-1258
 *
 * fp32, no FMA contraction (-ffp-contract=off), true division.  Sweeps are
 * OpenMP-parallel (Jacobi is order independent, so results do not depend on
 * the thread count); set_bnd's obstacle pass is order independent as well
 * (reads fluid cells, writes obstacle cells).
 */
#include "fluid_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long long i64;
#define ID(x, y, z) ((i64)(x) + (i64)(y) * nx + (i64)(z) * nx * ny)

static inline int has_z(int nz) { return nz > 1; }

/* ---- set_bnd: FluidSim.cs:1235-1289 generalised ------------------------------------------- */
void fo_set_bnd(int nx, int ny, int nz, int b, float *x, const uint8_t *obs) {
    const int hz = has_z(nz);
    const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0; /* interior z range (inclusive) */
    /* faces (:1246-1252) */
#pragma omp parallel for schedule(static)
    for (int k = k0; k <= k1; k++) {
        for (int j = 1; j <= ny - 2; j++) {
            x[ID(0, j, k)] = b == 1 ? -x[ID(1, j, k)] : x[ID(1, j, k)];
            x[ID(nx - 1, j, k)] = b == 1 ? -x[ID(nx - 2, j, k)] : x[ID(nx - 2, j, k)];
        }
        for (int i = 1; i <= nx - 2; i++) {
            x[ID(i, 0, k)] = b == 2 ? -x[ID(i, 1, k)] : x[ID(i, 1, k)];
            x[ID(i, ny - 1, k)] = b == 2 ? -x[ID(i, ny - 2, k)] : x[ID(i, ny - 2, k)];
        }
    }
    if (hz) {
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= ny - 2; j++)
            for (int i = 1; i <= nx - 2; i++) {
                x[ID(i, j, 0)] = b == 3 ? -x[ID(i, j, 1)] : x[ID(i, j, 1)];
                x[ID(i, j, nz - 1)] = b == 3 ? -x[ID(i, j, nz - 2)] : x[ID(i, j, nz - 2)];
            }
    }
    /* edges along z == the 2D corners (:1255-1258), for every interior plane */
    for (int k = k0; k <= k1; k++) {
        x[ID(0, 0, k)] = 0.5f * (x[ID(1, 0, k)] + x[ID(0, 1, k)]);
        x[ID(0, ny - 1, k)] = 0.5f * (x[ID(1, ny - 1, k)] + x[ID(0, ny - 2, k)]);
        x[ID(nx - 1, 0, k)] = 0.5f * (x[ID(nx - 2, 0, k)] + x[ID(nx - 1, 1, k)]);
        x[ID(nx - 1, ny - 1, k)] = 0.5f * (x[ID(nx - 2, ny - 1, k)] + x[ID(nx - 1, ny - 2, k)]);
    }
    if (hz) {
        const int zs[2] = {0, nz - 1}, zi[2] = {1, nz - 2};
        for (int s = 0; s < 2; s++) {
            const int z = zs[s], zn = zi[s];
            for (int j = 1; j <= ny - 2; j++) { /* edges along y: x-adjacent + z-adjacent */
                x[ID(0, j, z)] = 0.5f * (x[ID(1, j, z)] + x[ID(0, j, zn)]);
                x[ID(nx - 1, j, z)] = 0.5f * (x[ID(nx - 2, j, z)] + x[ID(nx - 1, j, zn)]);
            }
            for (int i = 1; i <= nx - 2; i++) { /* edges along x: y-adjacent + z-adjacent */
                x[ID(i, 0, z)] = 0.5f * (x[ID(i, 1, z)] + x[ID(i, 0, zn)]);
                x[ID(i, ny - 1, z)] = 0.5f * (x[ID(i, ny - 2, z)] + x[ID(i, ny - 1, zn)]);
            }
        }
        for (int s = 0; s < 2; s++) { /* 8 corners: (x-adj + y-adj + z-adj) / 3 */
            const int z = zs[s], zn = zi[s];
            x[ID(0, 0, z)] = (x[ID(1, 0, z)] + x[ID(0, 1, z)] + x[ID(0, 0, zn)]) / 3.0f;
            x[ID(nx - 1, 0, z)] = (x[ID(nx - 2, 0, z)] + x[ID(nx - 1, 1, z)] + x[ID(nx - 1, 0, zn)]) / 3.0f;
            x[ID(0, ny - 1, z)] = (x[ID(1, ny - 1, z)] + x[ID(0, ny - 2, z)] + x[ID(0, ny - 1, zn)]) / 3.0f;
            x[ID(nx - 1, ny - 1, z)] =
                (x[ID(nx - 2, ny - 1, z)] + x[ID(nx - 1, ny - 2, z)] + x[ID(nx - 1, ny - 1, zn)]) / 3.0f;
        }
    }
    /* obstacle mirroring (:1261-1287); b == 0 leaves obstacle cells untouched */
    if (b == 0) return;
    if (b == 3 && !hz) return;
    const i64 step = b == 1 ? 1 : (b == 2 ? nx : (i64)nx * ny);
#pragma omp parallel for schedule(static)
    for (int k = k0; k <= k1; k++)
        for (int j = 1; j <= ny - 2; j++)
            for (int i = 1; i <= nx - 2; i++) {
                const i64 idx = ID(i, j, k);
                if (!obs[idx]) continue;
                float m = 0;
                int count = 0;
                if (!obs[idx - step]) { m += -x[idx - step]; count++; }
                if (!obs[idx + step]) { m += -x[idx + step]; count++; }
                x[idx] = count > 0 ? m / count : 0;
            }
}

/* ---- pass-1 smoother: FluidSim.cs:1292-1357 + :1034-1069 ---------------------------------- */
void fo_diffuse_smooth(int nx, int ny, int nz, int b, float *x, const float *x0, float a, float c,
                       const uint8_t *obs, int iters) {
    const i64 total = (i64)nx * ny * nz;
    const int hz = has_z(nz);
    const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0;
    const i64 sy = nx, sz = (i64)nx * ny;
    float *buf1 = malloc(sizeof(float) * total), *buf2 = malloc(sizeof(float) * total);
    memcpy(buf1, x0, sizeof(float) * total); /* :1299-1300 */
    memcpy(buf2, x0, sizeof(float) * total);
    float *in = buf1, *out = buf2;
    for (int it = 0; it < iters; it++) {
#pragma omp parallel for schedule(static)
        for (int k = k0; k <= k1; k++)
            for (int j = 1; j <= ny - 2; j++)
                for (int i = 1; i <= nx - 2; i++) {
                    const i64 idx = ID(i, j, k);
                    if (obs[idx]) continue; /* :1055: not written */
                    float s = in[idx + 1] + in[idx - 1] + in[idx + sy] + in[idx - sy]; /* :1063-1066 */
                    if (hz) s = s + in[idx + sz] + in[idx - sz];
                    out[idx] = (in[idx] + a * s) / c; /* :1062: in[idx], not x0[idx] */
                }
        fo_set_bnd(nx, ny, nz, b, out, obs);
        float *t = in; in = out; out = t;
    }
    memcpy(x, in, sizeof(float) * total); /* :1348 */
    free(buf1);
    free(buf2);
}

/* ---- Jacobi: FluidSim.cs:1188-1233 iteration, loops :1359-1415 / :1578-1637 ---------------- */
static void fo_jacobi_iteration(int nx, int ny, int nz, const float *x0, const float *xr, float *xw,
                                float a, float c, const uint8_t *obs) {
    const int hz = has_z(nz);
    const i64 sy = nx, sz = (i64)nx * ny;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const i64 idx = ID(i, j, k);
                const int ring = i <= 0 || i >= nx - 1 || j <= 0 || j >= ny - 1 || (hz && (k <= 0 || k >= nz - 1));
                if (ring || obs[idx]) { xw[idx] = xr[idx]; continue; } /* :1206-1218 */
                float s = xr[idx + 1] + xr[idx - 1] + xr[idx + sy] + xr[idx - sy];  /* :1228-1229 */
                if (hz) s = s + xr[idx + sz] + xr[idx - sz];
                xw[idx] = (x0[idx] + a * s) / c; /* :1227-1230 */
            }
}

void fo_lin_solve(int nx, int ny, int nz, int b, float *x, const float *x0, float a, float c,
                  const uint8_t *obs, int iters) {
    const i64 total = (i64)nx * ny * nz;
    float *b1 = malloc(sizeof(float) * total), *b2 = calloc(total, sizeof(float));
    memcpy(b1, x, sizeof(float) * total); /* :1367 initial guess */
    float *rd = b1, *wr = b2;
    for (int it = 0; it < iters; it++) {
        fo_jacobi_iteration(nx, ny, nz, x0, rd, wr, a, c, obs);
        fo_set_bnd(nx, ny, nz, b, wr, obs); /* :1399 */
        float *t = rd; rd = wr; wr = t;
    }
    memcpy(x, rd, sizeof(float) * total); /* :1408 */
    free(b1);
    free(b2);
}

/* Red-black Gauss-Seidel variant (BASELINE.json config 5; not in the reference, which is Jacobi).
 * In place; colour = (i+j+k)&1, colour 0 first; set_bnd after each full sweep.  Used only for the
 * residual-equivalence check documented in DESIGN.md section 2.9. */
void fo_lin_solve_rb(int nx, int ny, int nz, int b, float *x, const float *x0, float a, float c,
                     const uint8_t *obs, int iters) {
    const int hz = has_z(nz);
    const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0;
    const i64 sy = nx, sz = (i64)nx * ny;
    for (int it = 0; it < iters; it++) {
        for (int colour = 0; colour < 2; colour++) {
#pragma omp parallel for schedule(static)
            for (int k = k0; k <= k1; k++)
                for (int j = 1; j <= ny - 2; j++)
                    for (int i = 1; i <= nx - 2; i++) {
                        if (((i + j + k) & 1) != colour) continue;
                        const i64 idx = ID(i, j, k);
                        if (obs[idx]) continue;
                        float s = x[idx + 1] + x[idx - 1] + x[idx + sy] + x[idx - sy];
                        if (hz) s = s + x[idx + sz] + x[idx - sz];
                        x[idx] = (x0[idx] + a * s) / c;
                    }
        }
        fo_set_bnd(nx, ny, nz, b, x, obs);
    }
}

/* ---- Diffuse: FluidSim.cs:740-745 --------------------------------------------------------- */
void fo_diffuse_coeffs(int n, float diff, float dt, float *a, float *c) {
    *a = dt * diff * (n - 2) * (n - 2); /* :743, left to right in fp32 */
    *c = 1 + 6 * *a;                    /* :744 */
}

void fo_diffuse(int nx, int ny, int nz, int b, float *x, const float *x0, float diff, float dt,
                const uint8_t *obs, int iters) {
    float a, c;
    fo_diffuse_coeffs(nx, diff, dt, &a, &c);
    fo_diffuse_smooth(nx, ny, nz, b, x, x0, a, c, obs, iters);
    fo_lin_solve(nx, ny, nz, b, x, x0, a, c, obs, iters);
}

/* ---- Project: FluidSim.cs:1417-1521 ------------------------------------------------------- */
void fo_divergence(int nx, int ny, int nz, float *div, const float *vx, const float *vy, const float *vz,
                   const uint8_t *obs) {
    const i64 total = (i64)nx * ny * nz;
    const int hz = has_z(nz);
    const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0;
    const i64 sy = nx, sz = (i64)nx * ny;
    memset(div, 0, sizeof(float) * total); /* :1428 */
#pragma omp parallel for schedule(static)
    for (int k = k0; k <= k1; k++)
        for (int j = 1; j <= ny - 2; j++)
            for (int i = 1; i <= nx - 2; i++) {
                const i64 idx = ID(i, j, k);
                float s = vx[idx + 1] - vx[idx - 1] + vy[idx + sy] - vy[idx - sy]; /* :1089-1092 */
                if (hz) s = s + vz[idx + sz] - vz[idx - sz];
                div[idx] = -0.5f * s / nx;
            }
    fo_set_bnd(nx, ny, nz, 0, div, obs);
}

void fo_subtract_gradient(int nx, int ny, int nz, float *vx, float *vy, float *vz, const float *p,
                          const uint8_t *obs) {
    const int hz = has_z(nz);
    const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0;
    const i64 sy = nx, sz = (i64)nx * ny;
#pragma omp parallel for schedule(static)
    for (int k = k0; k <= k1; k++)
        for (int j = 1; j <= ny - 2; j++)
            for (int i = 1; i <= nx - 2; i++) {
                const i64 idx = ID(i, j, k);
                if (obs[idx]) continue; /* :1117 */
                vx[idx] -= 0.5f * (p[idx + 1] - p[idx - 1]) * nx;   /* :1120 */
                vy[idx] -= 0.5f * (p[idx + sy] - p[idx - sy]) * nx; /* :1121 */
                if (hz) vz[idx] -= 0.5f * (p[idx + sz] - p[idx - sz]) * nx;
            }
    fo_set_bnd(nx, ny, nz, 1, vx, obs);
    fo_set_bnd(nx, ny, nz, 2, vy, obs);
    if (hz) fo_set_bnd(nx, ny, nz, 3, vz, obs);
}

void fo_project(int nx, int ny, int nz, float *vx, float *vy, float *vz, float *p, const uint8_t *obs,
                int iters, int red_black) {
    const i64 total = (i64)nx * ny * nz;
    float *div = malloc(sizeof(float) * total);
    fo_divergence(nx, ny, nz, div, vx, vy, vz, obs);
    memset(p, 0, sizeof(float) * total); /* :1427, :1094; set_bnd(0) of zeros is zeros */
    if (red_black)
        fo_lin_solve_rb(nx, ny, nz, 0, p, div, 1.0f, 6.0f, obs, iters);
    else
        fo_lin_solve(nx, ny, nz, 0, p, div, 1.0f, 6.0f, obs, iters); /* :1581-1582 */
    fo_subtract_gradient(nx, ny, nz, vx, vy, vz, p, obs);
    free(div);
}

/* ---- Advect: FluidSim.cs:1523-1576 + :1125-1186 ------------------------------------------- */
void fo_advect(int nx, int ny, int nz, int b, float *d, const float *d0, const float *vx, const float *vy,
               const float *vz, float dt, const uint8_t *obs) {
    const i64 total = (i64)nx * ny * nz;
    const int hz = has_z(nz);
    const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0;
    const float dt0 = dt * (nx - 2); /* :1526 */
    float *out = calloc(total, sizeof(float)); /* :1529 */
#pragma omp parallel for schedule(static)
    for (int k = k0; k <= k1; k++)
        for (int j = 1; j <= ny - 2; j++)
            for (int i = 1; i <= nx - 2; i++) {
                const i64 idx = ID(i, j, k);
                if (obs[idx]) continue; /* :1148-1156: 0 for every b, the output being fresh */
                float x = i - dt0 * vx[idx];
                float y = j - dt0 * vy[idx];
                if (x < 0.5f) x = 0.5f;
                if (x > nx - 1.5f) x = nx - 1.5f;
                const int i0 = (int)x, i1 = i0 + 1;
                if (y < 0.5f) y = 0.5f;
                if (y > ny - 1.5f) y = ny - 1.5f;
                const int j0 = (int)y, j1 = j0 + 1;
                const float s1 = x - i0, s0 = 1 - s1, t1 = y - j0, t0 = 1 - t1;
                if (!hz) {
                    out[idx] = s0 * (t0 * d0[ID(i0, j0, 0)] + t1 * d0[ID(i0, j1, 0)]) +
                               s1 * (t0 * d0[ID(i1, j0, 0)] + t1 * d0[ID(i1, j1, 0)]); /* :1183-1184 */
                } else {
                    float z = k - dt0 * vz[idx];
                    if (z < 0.5f) z = 0.5f;
                    if (z > nz - 1.5f) z = nz - 1.5f;
                    const int kk0 = (int)z, kk1 = kk0 + 1;
                    const float u1 = z - kk0, u0 = 1 - u1;
                    const float lo = s0 * (t0 * d0[ID(i0, j0, kk0)] + t1 * d0[ID(i0, j1, kk0)]) +
                                     s1 * (t0 * d0[ID(i1, j0, kk0)] + t1 * d0[ID(i1, j1, kk0)]);
                    const float hi = s0 * (t0 * d0[ID(i0, j0, kk1)] + t1 * d0[ID(i0, j1, kk1)]) +
                                     s1 * (t0 * d0[ID(i1, j0, kk1)] + t1 * d0[ID(i1, j1, kk1)]);
                    out[idx] = u0 * lo + u1 * hi;
                }
            }
    fo_set_bnd(nx, ny, nz, b, out, obs);
    memcpy(d, out, sizeof(float) * total);
    free(out);
}

/* ---- obstacle post-pass: FluidSim.cs:617-673 ----------------------------------------------
 * Per fluid interior cell: apply V *= f(|V|) once per INTERIOR obstacle neighbour (4 in 2D, 6 in
 * 3D), recomputing |V| each time; obstacle interior cells get V = 0.  Equivalent to the
 * reference's sequential loop because each application touches one fluid cell only. */
static inline float fo_drag_factor(float U, float cell, float rawvisc) {
    const float visc = rawvisc > 1e-5f ? rawvisc : 1e-5f; /* Mathf.Max(viscosity, 1e-5f) :664 */
    const float Re = (U * cell) / visc;
    float t = 1.0f - (float)exp((double)(-Re * 0.01f)); /* Mathf.Exp :667 */
    if (t < 0.0f) t = 0.0f;
    if (t > 1.0f) t = 1.0f; /* Mathf.Lerp clamps */
    return 0.8f + (0.98f - 0.8f) * t;
}

void fo_enforce_obstacles(int nx, int ny, int nz, float *vx, float *vy, float *vz, const uint8_t *obs,
                          float cell, float rawvisc) {
    const int hz = has_z(nz);
    const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0;
#pragma omp parallel for schedule(static)
    for (int k = k0; k <= k1; k++)
        for (int j = 1; j <= ny - 2; j++)
            for (int i = 1; i <= nx - 2; i++) {
                const i64 idx = ID(i, j, k);
                if (obs[idx]) {
                    vx[idx] = 0;
                    vy[idx] = 0;
                    if (hz) vz[idx] = 0;
                    continue;
                }
                int n = 0; /* interior obstacle neighbours */
                if (i - 1 >= 1 && obs[ID(i - 1, j, k)]) n++;
                if (i + 1 <= nx - 2 && obs[ID(i + 1, j, k)]) n++;
                if (j - 1 >= 1 && obs[ID(i, j - 1, k)]) n++;
                if (j + 1 <= ny - 2 && obs[ID(i, j + 1, k)]) n++;
                if (hz && k - 1 >= 1 && obs[ID(i, j, k - 1)]) n++;
                if (hz && k + 1 <= nz - 2 && obs[ID(i, j, k + 1)]) n++;
                for (int r = 0; r < n; r++) {
                    float q = vx[idx] * vx[idx] + vy[idx] * vy[idx];
                    if (hz) q = q + vz[idx] * vz[idx];
                    const float U = (float)sqrt((double)q); /* Mathf.Sqrt :661 */
                    const float f = fo_drag_factor(U, cell, rawvisc);
                    vx[idx] *= f;
                    vy[idx] *= f;
                    if (hz) vz[idx] *= f;
                }
            }
}

/* ---- sources: FluidSim.cs:723-738 --------------------------------------------------------- */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
i64 fo_cell_index(int nx, int ny, int nz, float x, float y, float z) {
    const int i = clampi((int)x, 0, nx - 1), j = clampi((int)y, 0, ny - 1);
    const int k = has_z(nz) ? clampi((int)z, 0, nz - 1) : 0;
    return ID(i, j, k);
}

/* ---- full step: FluidSim.cs:551-570, :703-721 --------------------------------------------- */
void fo_step(fo_state *s, float dt, float visc, float diff) {
    const int nx = s->nx, ny = s->ny, nz = s->nz, hz = has_z(nz);
    const i64 total = (i64)nx * ny * nz;
    const uint8_t *obs = s->obstacles;
    /* VelocityStep :703-714 */
    fo_diffuse(nx, ny, nz, 1, s->vx0, s->vx, visc, dt, obs, s->iters_diffuse);
    fo_diffuse(nx, ny, nz, 2, s->vy0, s->vy, visc, dt, obs, s->iters_diffuse);
    if (hz) fo_diffuse(nx, ny, nz, 3, s->vz0, s->vz, visc, dt, obs, s->iters_diffuse);
    fo_project(nx, ny, nz, s->vx0, s->vy0, s->vz0, s->pressure, obs, s->iters_pressure, s->red_black);
    {   /* :710-711: every component is advected along the same (projected, diffused) field */
        fo_advect(nx, ny, nz, 1, s->vx, s->vx0, s->vx0, s->vy0, s->vz0, dt, obs);
        fo_advect(nx, ny, nz, 2, s->vy, s->vy0, s->vx0, s->vy0, s->vz0, dt, obs);
        if (hz) fo_advect(nx, ny, nz, 3, s->vz, s->vz0, s->vx0, s->vy0, s->vz0, dt, obs);
    }
    fo_project(nx, ny, nz, s->vx, s->vy, s->vz, s->pressure, obs, s->iters_pressure, s->red_black);
    /* DensityStep :716-721 */
    float *tmp = calloc(total, sizeof(float));
    fo_diffuse(nx, ny, nz, 0, tmp, s->density, diff, dt, obs, s->iters_diffuse);
    fo_advect(nx, ny, nz, 0, s->density, tmp, s->vx, s->vy, s->vz, dt, obs);
    free(tmp);
    if (s->enable_obstacle) /* :567-570 */
        fo_enforce_obstacles(nx, ny, nz, s->vx, s->vy, s->vz, obs, s->cell_size, s->raw_viscosity);
}

/* ---- metrics: FluidSim.cs:582-594 (next row N1) ------------------------------------------- */
void fo_metrics(const fo_state *s, float *mean_density, float *max_speed) {
    const i64 total = (i64)s->nx * s->ny * s->nz;
    const int hz = has_z(s->nz);
    double acc = 0; /* NOTE: the reference accumulates in float (:587); double here is the truth value */
    float mx = 0;
    for (i64 i = 0; i < total; i++) {
        acc += s->density[i];
        float q = s->vx[i] * s->vx[i] + s->vy[i] * s->vy[i];
        if (hz) q = q + s->vz[i] * s->vz[i];
        const float m = sqrtf(q);
        if (m > mx) mx = m;
    }
    *mean_density = (float)(acc / (double)total);
    *max_speed = mx;
}

/* ---- visualisation colour mapping: UpdateVisualizationJob.Execute, FluidSim.cs:1888-2001 (next row N2) ----------
 * One xy plane of size nx*ny; out is nx*ny Color (r,g,b,a floats), as the reference's NativeArray<Color>. */
typedef struct { float r, g, b, a; } fo_color;
static fo_color fo_c4(const float *c) { fo_color o = {c[0], c[1], c[2], c[3]}; return o; }
static fo_color fo_lerp(fo_color a, fo_color b, float t) { /* Color.Lerp: t clamped to [0,1] */
    if (t < 0.0f) t = 0.0f;
    if (t > 1.0f) t = 1.0f;
    fo_color o = {a.r + (b.r - a.r) * t, a.g + (b.g - a.g) * t, a.b + (b.b - a.b) * t, a.a + (b.a - a.a) * t};
    return o;
}
static fo_color fo_gradient(const fo_vis_params *vp, float time) { /* EvaluateGradient :1981-2001 */
    const int n = vp->gradient_key_count;
    if (n <= 0) { fo_color w = {1, 1, 1, 1}; return w; }
    if (time <= vp->gradient_times[0]) return fo_c4(vp->gradient_colors[0]);
    if (time >= vp->gradient_times[n - 1]) return fo_c4(vp->gradient_colors[n - 1]);
    int index = 0;
    while (index < n - 1 && time > vp->gradient_times[index + 1]) index++;
    float t = (time - vp->gradient_times[index]) / (vp->gradient_times[index + 1] - vp->gradient_times[index]);
    return fo_lerp(fo_c4(vp->gradient_colors[index]), fo_c4(vp->gradient_colors[index + 1]), t);
}
void fo_visualize(int nx, int ny, const float *density, const float *pressure, const uint8_t *obstacles,
                  const fo_vis_params *vp, float *out) {
    for (int index = 0; index < nx * ny; index++) {
        const int i = index % nx, j = index / nx;
        fo_color px;
        if (obstacles[index]) { px = fo_c4(vp->obstacle_color); goto store; } /* :1894-1899 */
        {
            const float d = density[index];
            const float normalizedD = d * vp->colour_intensity;
            switch (vp->color_mode) {
            case 2: /* DensityBased */
                if (d < vp->medium_density_threshold) {
                    fo_color black = {0, 0, 0, 1};
                    px = fo_lerp(black, fo_c4(vp->low_density_color), d / vp->medium_density_threshold);
                } else if (d < vp->high_density_threshold) {
                    float t = (d - vp->medium_density_threshold) / (vp->high_density_threshold - vp->medium_density_threshold);
                    px = fo_lerp(fo_c4(vp->low_density_color), fo_c4(vp->medium_density_color), t);
                } else {
                    float t = fminf(1.0f, (d - vp->high_density_threshold) / vp->high_density_threshold);
                    px = fo_lerp(fo_c4(vp->medium_density_color), fo_c4(vp->high_density_color), t);
                }
                break;
            case 1: { /* Gradient */
                float c = normalizedD < 0.0f ? 0.0f : (normalizedD > 1.0f ? 1.0f : normalizedD);
                px = fo_gradient(vp, c);
                break;
            }
            case 3: { /* PressureBased */
                const float p = pressure[index];
                if (p < vp->low_pressure_threshold) {
                    float t = p / vp->low_pressure_threshold;
                    px = fo_lerp(fo_c4(vp->low_pressure_color), fo_c4(vp->neutral_pressure_color), 1.0f + t);
                } else if (p <= vp->high_pressure_threshold) {
                    float t = (p - vp->low_pressure_threshold) / (vp->high_pressure_threshold - vp->low_pressure_threshold);
                    px = fo_lerp(fo_c4(vp->neutral_pressure_color), fo_c4(vp->high_pressure_color), t);
                } else {
                    float t = fminf(1.0f, (p - vp->high_pressure_threshold) / vp->high_pressure_threshold);
                    fo_color orange = {1.0f, 0.5f, 0.0f, 1.0f};
                    px = fo_lerp(fo_c4(vp->high_pressure_color), orange, t);
                }
                break;
            }
            default: /* SingleColor, Streamlines */
                px.r = vp->fluid_color[0] * normalizedD;
                px.g = vp->fluid_color[1] * normalizedD;
                px.b = vp->fluid_color[2] * normalizedD;
                px.a = vp->fluid_color[3];
                break;
            }
            if (vp->visualize_source_position && vp->enable_custom_source) { /* :1970-1978 */
                float distSq = (i - vp->source_x) * (i - vp->source_x) + (j - vp->source_y) * (j - vp->source_y);
                if (distSq < vp->visual_marker_radius * vp->visual_marker_radius) px = fo_c4(vp->source_position_color);
            }
        }
    store:
        out[4 * index] = px.r; out[4 * index + 1] = px.g; out[4 * index + 2] = px.b; out[4 * index + 3] = px.a;
    }
}

/* ---- streamline glyphs: StreamlineCalculationJob + StreamlineDrawJob, FluidSim.cs:1680-1762 (next row N3) -------
 * out: count*4 floats (startX, startY, endX, endY), -1 x4 for invalid glyphs; count = (nx/skip)*(ny/skip). */
void fo_streamlines(int nx, int ny, int skip, float streamlineScale, const float *velocX, const float *velocY,
                    const uint8_t *obstacles, float *out) {
    const int cols = nx / skip, rows = ny / skip;
    for (int index = 0; index < cols * rows; index++) {
        float sx = 0, sy = 0, sz = 0, sw = 0; /* the float4 `streamline` of :1726 */
        const int x = index % cols, y = index / cols;
        const int i = x * skip + skip, j = y * skip + skip;
        sx = (float)i; sy = (float)j;
        if (!(i <= 0 || i >= nx - 1 || j <= 0 || j >= ny - 1) && !obstacles[i + j * nx]) {
            const float vx = velocX[i + j * nx], vy = velocY[i + j * nx];
            const float magnitude = sqrtf(vx * vx + vy * vy);
            if (!(magnitude < 0.01f)) {
                sw = fminf((float)(skip - 1), magnitude * streamlineScale);
                sz = atan2f(vy, vx);
            }
        }
        float *o = out + 4 * index;
        if (sw <= 0) { o[0] = o[1] = o[2] = o[3] = -1.0f; continue; } /* :1744-1749 */
        const int startX = (int)sx, startY = (int)sy;
        o[0] = (float)startX; o[1] = (float)startY;
        o[2] = startX + cosf(sz) * sw;
        o[3] = startY + sinf(sz) * sw;
    }
}
