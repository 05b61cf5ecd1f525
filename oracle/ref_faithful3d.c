/*
 * ref_faithful3d.c -- TEST / BENCHMARK INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The CPU baseline "as the reference executes it" (BASELINE.md section 4): the same
 * arithmetic as fluid_oracle.c (bit identical, tested), but with the EXECUTION STRUCTURE of
 * Assets/Scripts/FluidSim.cs kept:
 *   - every Burst job is a flat index loop `Schedule(total, 64)` (:1324, :1393, :1443, :1480, :1550, :1607):
 *     OpenMP `schedule(static, 64)` over index = x + y*nx + z*nx*ny with the div/mod of :1047-1048;
 *   - BoundaryJob is an IJob: ONE thread, after every sweep, and its obstacle loop visits every interior
 *     cell whatever b is (:1261-1287);
 *   - every *WithJobs call allocates fresh native arrays and copies the managed fields in and out
 *     (:1299-1301, :1367-1369, :1425-1429, :1506-1509, :1529-1533, :1565) and ApplyBoundaryConditions
 *     re-copies the obstacle mask on EVERY iteration (:1645);
 *   - a blocking Complete() after each job (:1339, :1396, :1608) = the implicit barrier of each omp loop.
 * fluid_oracle.c is the "tidy" port (planes split over threads, parallel boundary pass, no gratuitous copies).
 * bench.py times both and labels them "port-faithful" / "port-tidy".
 *
 * PARITY UNPINNED by the reference, like the rest of oracle/ (no C# toolchain here, no reference tests).
 * 3D generalisation rules: DESIGN.md section 2 (z terms appended last, edges/corners as in fluid_oracle.c).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fluid_oracle.h"

typedef long long i64;
#define ID(x, y, z) ((i64)(x) + (i64)(y) * nx + (i64)(z) * nx * ny)

static float *rf_new_copy(const float *src, i64 n) { /* new NativeArray<float>(managed, TempJob) */
    float *p = malloc(sizeof(float) * n);
    memcpy(p, src, sizeof(float) * n);
    return p;
}
static float *rf_new_clear(i64 n) { return calloc(n, sizeof(float)); } /* new NativeArray<float>(n, TempJob) */
static uint8_t *rf_new_mask(const uint8_t *src, i64 n) {
    uint8_t *p = malloc(n);
    memcpy(p, src, n);
    return p;
}

/* BoundaryJob.Execute :1235-1289, single threaded (IJob). */
static void rf_boundary(int nx, int ny, int nz, int b, float *x, const uint8_t *obs) {
    const int hz = nz > 1;
    const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0;
    for (int k = k0; k <= k1; k++) {
        for (int j = 1; j <= ny - 2; j++) { /* :1246-1252 */
            x[ID(0, j, k)] = b == 1 ? -x[ID(1, j, k)] : x[ID(1, j, k)];
            x[ID(nx - 1, j, k)] = b == 1 ? -x[ID(nx - 2, j, k)] : x[ID(nx - 2, j, k)];
        }
        for (int i = 1; i <= nx - 2; i++) {
            x[ID(i, 0, k)] = b == 2 ? -x[ID(i, 1, k)] : x[ID(i, 1, k)];
            x[ID(i, ny - 1, k)] = b == 2 ? -x[ID(i, ny - 2, k)] : x[ID(i, ny - 2, k)];
        }
    }
    if (hz)
        for (int j = 1; j <= ny - 2; j++)
            for (int i = 1; i <= nx - 2; i++) {
                x[ID(i, j, 0)] = b == 3 ? -x[ID(i, j, 1)] : x[ID(i, j, 1)];
                x[ID(i, j, nz - 1)] = b == 3 ? -x[ID(i, j, nz - 2)] : x[ID(i, j, nz - 2)];
            }
    for (int k = k0; k <= k1; k++) { /* :1255-1258 */
        x[ID(0, 0, k)] = 0.5f * (x[ID(1, 0, k)] + x[ID(0, 1, k)]);
        x[ID(0, ny - 1, k)] = 0.5f * (x[ID(1, ny - 1, k)] + x[ID(0, ny - 2, k)]);
        x[ID(nx - 1, 0, k)] = 0.5f * (x[ID(nx - 2, 0, k)] + x[ID(nx - 1, 1, k)]);
        x[ID(nx - 1, ny - 1, k)] = 0.5f * (x[ID(nx - 2, ny - 1, k)] + x[ID(nx - 1, ny - 2, k)]);
    }
    if (hz) {
        const int zs[2] = {0, nz - 1}, zi[2] = {1, nz - 2};
        for (int s = 0; s < 2; s++) {
            const int z = zs[s], zn = zi[s];
            for (int j = 1; j <= ny - 2; j++) {
                x[ID(0, j, z)] = 0.5f * (x[ID(1, j, z)] + x[ID(0, j, zn)]);
                x[ID(nx - 1, j, z)] = 0.5f * (x[ID(nx - 2, j, z)] + x[ID(nx - 1, j, zn)]);
            }
            for (int i = 1; i <= nx - 2; i++) {
                x[ID(i, 0, z)] = 0.5f * (x[ID(i, 1, z)] + x[ID(i, 0, zn)]);
                x[ID(i, ny - 1, z)] = 0.5f * (x[ID(i, ny - 2, z)] + x[ID(i, ny - 1, zn)]);
            }
        }
        for (int s = 0; s < 2; s++) {
            const int z = zs[s], zn = zi[s];
            x[ID(0, 0, z)] = (x[ID(1, 0, z)] + x[ID(0, 1, z)] + x[ID(0, 0, zn)]) / 3.0f;
            x[ID(nx - 1, 0, z)] = (x[ID(nx - 2, 0, z)] + x[ID(nx - 1, 1, z)] + x[ID(nx - 1, 0, zn)]) / 3.0f;
            x[ID(0, ny - 1, z)] = (x[ID(1, ny - 1, z)] + x[ID(0, ny - 2, z)] + x[ID(0, ny - 1, zn)]) / 3.0f;
            x[ID(nx - 1, ny - 1, z)] =
                (x[ID(nx - 2, ny - 1, z)] + x[ID(nx - 1, ny - 2, z)] + x[ID(nx - 1, ny - 1, zn)]) / 3.0f;
        }
    }
    /* :1261-1287: the scan over every interior cell happens for every b; only b = 1/2/3 write */
    for (int k = k0; k <= k1; k++)
        for (int j = 1; j <= ny - 2; j++)
            for (int i = 1; i <= nx - 2; i++) {
                const i64 idx = ID(i, j, k);
                if (!obs[idx]) continue;
                const i64 step = b == 1 ? 1 : (b == 2 ? nx : (i64)nx * ny);
                if (b == 1 || b == 2 || (b == 3 && hz)) {
                    float m = 0;
                    int count = 0;
                    if (!obs[idx - step]) { m += -x[idx - step]; count++; }
                    if (!obs[idx + step]) { m += -x[idx + step]; count++; }
                    x[idx] = count > 0 ? m / count : 0;
                }
            }
}

/* ApplyBoundaryConditions :1639-1655: a fresh copy of the mask for one BoundaryJob. */
static void rf_apply_boundary(int nx, int ny, int nz, int b, float *x, const uint8_t *managed_obs) {
    const i64 total = (i64)nx * ny * nz;
    uint8_t *tmp = rf_new_mask(managed_obs, total); /* :1645 */
    rf_boundary(nx, ny, nz, b, x, tmp);
    free(tmp);
}

#define RF_FOR_EACH_INDEX                                                                                              \
    _Pragma("omp parallel for schedule(static, 64)") for (i64 index = 0; index < total; index++)
#define RF_IJK                                                                                                         \
    const int i = (int)(index % nx), j = (int)((index / nx) % ny), k = (int)(index / ((i64)nx * ny));               \
    const int ring = i <= 0 || i >= nx - 1 || j <= 0 || j >= ny - 1 || (hz && (k <= 0 || k >= nz - 1));              \
    (void)k

/* DiffuseWithJobs :1292-1357 + DiffuseJob :1034-1069 */
static void rf_diffuse_with_jobs(int nx, int ny, int nz, int b, float *x, const float *x0, float diff, float dt,
                                 const uint8_t *obs, int iters) {
    const i64 total = (i64)nx * ny * nz, sy = nx, sz = (i64)nx * ny;
    const int hz = nz > 1;
    const float a = dt * diff * (nx - 2) * (nx - 2), c = 1 + 6 * a;
    float *buffer1 = rf_new_copy(x0, total), *buffer2 = rf_new_copy(x0, total); /* :1299-1300 */
    uint8_t *nobs = rf_new_mask(obs, total);                                     /* :1301 */
    float *input = buffer1, *output = buffer2;
    for (int it = 0; it < iters; it++) {
        RF_FOR_EACH_INDEX {
            RF_IJK;
            if (ring || nobs[index]) continue;
            float s = input[index + 1] + input[index - 1] + input[index + sy] + input[index - sy];
            if (hz) s = s + input[index + sz] + input[index - sz];
            output[index] = (input[index] + a * s) / c;
        }
        rf_boundary(nx, ny, nz, b, output, nobs); /* :1327-1339 */
        float *t = input; input = output; output = t;
    }
    memcpy(x, input, sizeof(float) * total); /* :1348 */
    free(buffer1); free(buffer2); free(nobs);
}

/* LinearSolveIterationJob :1188-1233 */
static void rf_linsolve_job(int nx, int ny, int nz, const float *x0, const float *xr, float *xw, const uint8_t *obs,
                            float a, float c) {
    const i64 total = (i64)nx * ny * nz, sy = nx, sz = (i64)nx * ny;
    const int hz = nz > 1;
    RF_FOR_EACH_INDEX {
        RF_IJK;
        if (ring || obs[index]) { xw[index] = xr[index]; continue; }
        float s = xr[index + 1] + xr[index - 1] + xr[index + sy] + xr[index - sy];
        if (hz) s = s + xr[index + sz] + xr[index - sz];
        xw[index] = (x0[index] + a * s) / c;
    }
}

/* persistent job buffers of :990-1000 (InitializeJobBuffers) */
static float *g_job1, *g_job2;
static uint8_t *g_jobobs;
static i64 g_job_n;
static void rf_init_job_buffers(i64 total) {
    if (g_job_n == total) return;
    free(g_job1); free(g_job2); free(g_jobobs);
    g_job1 = rf_new_clear(total); g_job2 = rf_new_clear(total); g_jobobs = calloc(total, 1);
    g_job_n = total;
}

/* LinearSolveWithJobs :1359-1415 */
static void rf_linear_solve_with_jobs(int nx, int ny, int nz, int b, float *x, const float *x0, float a, float c,
                                      const uint8_t *obs, int iters) {
    const i64 total = (i64)nx * ny * nz;
    rf_init_job_buffers(total);
    memcpy(g_job1, x, sizeof(float) * total);   /* :1367 */
    float *tempX0 = rf_new_copy(x0, total);      /* :1368 */
    memcpy(g_jobobs, obs, total);                /* :1369 */
    float *rd = g_job1, *wr = g_job2;
    for (int it = 0; it < iters; it++) {
        rf_linsolve_job(nx, ny, nz, tempX0, rd, wr, g_jobobs, a, c);
        rf_apply_boundary(nx, ny, nz, b, wr, obs); /* :1399 -> :1645 copies the mask again */
        float *t = rd; rd = wr; wr = t;
    }
    memcpy(x, rd, sizeof(float) * total); /* :1408 */
    free(tempX0);
}

static void rf_diffuse(int nx, int ny, int nz, int b, float *x, const float *x0, float diff, float dt,
                       const uint8_t *obs, int iters) { /* :740-745 */
    rf_diffuse_with_jobs(nx, ny, nz, b, x, x0, diff, dt, obs, iters);
    const float a = dt * diff * (nx - 2) * (nx - 2);
    rf_linear_solve_with_jobs(nx, ny, nz, b, x, x0, a, 1 + 6 * a, obs, iters);
}

/* ProjectWithJobs :1417-1521 (+ PressureSolveWithJobs :1578-1637) */
static void rf_project_with_jobs(int nx, int ny, int nz, float *vx, float *vy, float *vz, float *p, float *pressure,
                                 const uint8_t *obs, int iters) {
    const i64 total = (i64)nx * ny * nz, sy = nx, sz = (i64)nx * ny;
    const int hz = nz > 1;
    rf_init_job_buffers(total);
    float *nvx = rf_new_copy(vx, total), *nvy = rf_new_copy(vy, total), *nvz = hz ? rf_new_copy(vz, total) : NULL;
    float *np_ = rf_new_clear(total), *ndiv = rf_new_clear(total); /* :1427-1428 */
    uint8_t *nobs = rf_new_mask(obs, total);
    RF_FOR_EACH_INDEX { /* ProjectDivergenceJob :1080-1095 */
        RF_IJK;
        if (ring) continue;
        float s = nvx[index + 1] - nvx[index - 1] + nvy[index + sy] - nvy[index - sy];
        if (hz) s = s + nvz[index + sz] - nvz[index - sz];
        ndiv[index] = -0.5f * s / nx;
        np_[index] = 0;
    }
    rf_boundary(nx, ny, nz, 0, ndiv, nobs); /* :1446-1466 */
    rf_boundary(nx, ny, nz, 0, np_, nobs);
    {
        float *tempBuffer = rf_new_clear(total); /* :1586 */
        float *rd = np_, *wr = tempBuffer;
        for (int it = 0; it < iters; it++) {
            rf_linsolve_job(nx, ny, nz, ndiv, rd, wr, nobs, 1.0f, 6.0f);
            rf_boundary(nx, ny, nz, 0, wr, nobs);
            float *t = rd; rd = wr; wr = t;
        }
        if (rd != np_) memcpy(np_, rd, sizeof(float) * total); /* :1627-1631 */
        free(tempBuffer);
    }
    RF_FOR_EACH_INDEX { /* ProjectVelocityAdjustJob :1107-1122 */
        RF_IJK;
        if (ring || nobs[index]) continue;
        nvx[index] -= 0.5f * (np_[index + 1] - np_[index - 1]) * nx;
        nvy[index] -= 0.5f * (np_[index + sy] - np_[index - sy]) * nx;
        if (hz) nvz[index] -= 0.5f * (np_[index + sz] - np_[index - sz]) * nx;
    }
    rf_boundary(nx, ny, nz, 1, nvx, nobs); /* :1483-1503 */
    rf_boundary(nx, ny, nz, 2, nvy, nobs);
    if (hz) rf_boundary(nx, ny, nz, 3, nvz, nobs);
    memcpy(vx, nvx, sizeof(float) * total); /* :1506-1509 */
    memcpy(vy, nvy, sizeof(float) * total);
    if (hz) memcpy(vz, nvz, sizeof(float) * total);
    memcpy(p, np_, sizeof(float) * total);
    memcpy(pressure, np_, sizeof(float) * total);
    free(nvx); free(nvy); free(nvz); free(np_); free(ndiv); free(nobs);
}

/* AdvectWithJobs :1523-1576 + AdvectJob :1125-1186 */
static void rf_advect_with_jobs(int nx, int ny, int nz, int b, float *d, const float *d0, const float *vx,
                                const float *vy, const float *vz, float dt, const uint8_t *obs) {
    const i64 total = (i64)nx * ny * nz;
    const int hz = nz > 1;
    const float dt0 = dt * (nx - 2);
    float *nd = rf_new_clear(total), *nd0 = rf_new_copy(d0, total); /* :1529-1530 */
    float *nvx = rf_new_copy(vx, total), *nvy = rf_new_copy(vy, total), *nvz = hz ? rf_new_copy(vz, total) : NULL;
    uint8_t *nobs = rf_new_mask(obs, total);
    RF_FOR_EACH_INDEX {
        RF_IJK;
        if (ring || nobs[index]) continue; /* :1148-1156: the output is fresh, so 0 */
        float x = i - dt0 * nvx[index];
        float y = j - dt0 * nvy[index];
        if (x < 0.5f) x = 0.5f;
        if (x > nx - 1.5f) x = nx - 1.5f;
        const int i0 = (int)x, i1 = i0 + 1;
        if (y < 0.5f) y = 0.5f;
        if (y > ny - 1.5f) y = ny - 1.5f;
        const int j0 = (int)y, j1 = j0 + 1;
        const float s1 = x - i0, s0 = 1 - s1, t1 = y - j0, t0 = 1 - t1;
        if (!hz) {
            nd[index] = s0 * (t0 * nd0[ID(i0, j0, 0)] + t1 * nd0[ID(i0, j1, 0)]) +
                        s1 * (t0 * nd0[ID(i1, j0, 0)] + t1 * nd0[ID(i1, j1, 0)]);
        } else {
            float z = k - dt0 * nvz[index];
            if (z < 0.5f) z = 0.5f;
            if (z > nz - 1.5f) z = nz - 1.5f;
            const int kk0 = (int)z, kk1 = kk0 + 1;
            const float u1 = z - kk0, u0 = 1 - u1;
            const float lo = s0 * (t0 * nd0[ID(i0, j0, kk0)] + t1 * nd0[ID(i0, j1, kk0)]) +
                             s1 * (t0 * nd0[ID(i1, j0, kk0)] + t1 * nd0[ID(i1, j1, kk0)]);
            const float hi = s0 * (t0 * nd0[ID(i0, j0, kk1)] + t1 * nd0[ID(i0, j1, kk1)]) +
                             s1 * (t0 * nd0[ID(i1, j0, kk1)] + t1 * nd0[ID(i1, j1, kk1)]);
            nd[index] = u0 * lo + u1 * hi;
        }
    }
    rf_boundary(nx, ny, nz, b, nd, nobs); /* :1553-1562 */
    memcpy(d, nd, sizeof(float) * total);  /* :1565 */
    free(nd); free(nd0); free(nvx); free(nvy); free(nvz); free(nobs);
}

/* Simulate :551-570 = VelocityStep :703-714 + DensityStep :716-721 + EnforceObstacleBoundaries :617-673.
 * The obstacle post-pass is the reference's serial main-thread loop (fo_enforce_obstacles is its per-cell form;
 * here it runs on one thread). */
void rf_step(fo_state *s, float dt, float visc, float diff) {
    const int nx = s->nx, ny = s->ny, nz = s->nz, hz = nz > 1;
    const i64 total = (i64)nx * ny * nz;
    const uint8_t *obs = s->obstacles;
    const int kd = s->iters_diffuse, kp = s->iters_pressure;
    rf_diffuse(nx, ny, nz, 1, s->vx0, s->vx, visc, dt, obs, kd);
    rf_diffuse(nx, ny, nz, 2, s->vy0, s->vy, visc, dt, obs, kd);
    if (hz) rf_diffuse(nx, ny, nz, 3, s->vz0, s->vz, visc, dt, obs, kd);
    /* :708: p lands in velocityX (clobbered, the advect below rewrites it) and in `pressure` */
    {
        float *p = rf_new_clear(total);
        rf_project_with_jobs(nx, ny, nz, s->vx0, s->vy0, s->vz0, p, s->pressure, obs, kp);
        free(p);
    }
    rf_advect_with_jobs(nx, ny, nz, 1, s->vx, s->vx0, s->vx0, s->vy0, s->vz0, dt, obs);
    rf_advect_with_jobs(nx, ny, nz, 2, s->vy, s->vy0, s->vx0, s->vy0, s->vz0, dt, obs);
    if (hz) rf_advect_with_jobs(nx, ny, nz, 3, s->vz, s->vz0, s->vx0, s->vy0, s->vz0, dt, obs);
    {
        float *p = rf_new_clear(total);
        rf_project_with_jobs(nx, ny, nz, s->vx, s->vy, s->vz, p, s->pressure, obs, kp);
        free(p);
    }
    float *densityTemp = rf_new_clear(total); /* :718 */
    rf_diffuse(nx, ny, nz, 0, densityTemp, s->density, diff, dt, obs, kd);
    rf_advect_with_jobs(nx, ny, nz, 0, s->density, densityTemp, s->vx, s->vy, s->vz, dt, obs);
    free(densityTemp);
    if (s->enable_obstacle) {
        /* one thread, like the reference's main-thread loop */
        const int k0 = hz ? 1 : 0, k1 = hz ? nz - 2 : 0;
        for (int k = k0; k <= k1; k++)
            for (int j = 1; j <= ny - 2; j++)
                for (int i = 1; i <= nx - 2; i++) {
                    const i64 idx = ID(i, j, k);
                    if (obs[idx]) {
                        s->vx[idx] = 0; s->vy[idx] = 0;
                        if (hz) s->vz[idx] = 0;
                        continue;
                    }
                    int n = 0;
                    if (i - 1 >= 1 && obs[ID(i - 1, j, k)]) n++;
                    if (i + 1 <= nx - 2 && obs[ID(i + 1, j, k)]) n++;
                    if (j - 1 >= 1 && obs[ID(i, j - 1, k)]) n++;
                    if (j + 1 <= ny - 2 && obs[ID(i, j + 1, k)]) n++;
                    if (hz && k - 1 >= 1 && obs[ID(i, j, k - 1)]) n++;
                    if (hz && k + 1 <= nz - 2 && obs[ID(i, j, k + 1)]) n++;
                    for (int r = 0; r < n; r++) {
                        float q = s->vx[idx] * s->vx[idx] + s->vy[idx] * s->vy[idx];
                        if (hz) q = q + s->vz[idx] * s->vz[idx];
                        const float U = (float)sqrt((double)q);
                        const float vsc = s->raw_viscosity > 1e-5f ? s->raw_viscosity : 1e-5f;
                        const float Re = (U * s->cell_size) / vsc;
                        float t = 1.0f - (float)exp((double)(-Re * 0.01f));
                        if (t < 0.0f) t = 0.0f;
                        if (t > 1.0f) t = 1.0f;
                        const float f = 0.8f + (0.98f - 0.8f) * t;
                        s->vx[idx] *= f; s->vy[idx] *= f;
                        if (hz) s->vz[idx] *= f;
                    }
                }
    }
}
