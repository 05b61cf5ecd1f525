/*
 * ref2d.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Literal, single-threaded C restatement of the 2D solver hot path of
 * ChrisWangstpauls/3DFluidSimulation (Assets/Scripts/FluidSim.cs).  Every
 * function cites the reference lines it follows and keeps the reference's
 * evaluation order, buffer initialisation and quirks.
 *
 * PARITY UNPINNED: the reference is Unity C# (Burst jobs).  No C# toolchain
 * (dotnet/mono/csc/Unity) exists in the build container, the shipped WebGL
 * player lacks its .wasm, and the reference repository contains no tests,
 * golden vectors or fixtures for this path.  This file is therefore pinned
 * only by (a) hand-derived known-answer tests (tests/test_oracle_known_answers.py,
 * SURVEY.md section 8c K1-K8) and (b) an independently written numpy
 * restatement (oracle/np_restatement.py) whose outputs are committed under
 * tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this code.  The product (libfluidsolver.so)
 * never links or calls it.
 *
 * Arithmetic: fp32 throughout, no FMA contraction (compiled with
 * -ffp-contract=off), true division -- the semantics of Burst's default
 * (strict) float mode.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IX(x, y, size) ((x) + (y) * (size)) /* FluidSim.cs:749-752 */

/* FluidSim.cs:1235-1289  BoundaryJob.Execute */
void r2_boundary(int size, int b, float *x, const uint8_t *obstacles) {
    for (int i = 1; i < size - 1; i++) { /* :1246-1252 */
        x[IX(0, i, size)] = b == 1 ? -x[IX(1, i, size)] : x[IX(1, i, size)];
        x[IX(size - 1, i, size)] = b == 1 ? -x[IX(size - 2, i, size)] : x[IX(size - 2, i, size)];
        x[IX(i, 0, size)] = b == 2 ? -x[IX(i, 1, size)] : x[IX(i, 1, size)];
        x[IX(i, size - 1, size)] = b == 2 ? -x[IX(i, size - 2, size)] : x[IX(i, size - 2, size)];
    }
    /* :1255-1258 corners */
    x[IX(0, 0, size)] = 0.5f * (x[IX(1, 0, size)] + x[IX(0, 1, size)]);
    x[IX(0, size - 1, size)] = 0.5f * (x[IX(1, size - 1, size)] + x[IX(0, size - 2, size)]);
    x[IX(size - 1, 0, size)] = 0.5f * (x[IX(size - 2, 0, size)] + x[IX(size - 1, 1, size)]);
    x[IX(size - 1, size - 1, size)] =
        0.5f * (x[IX(size - 2, size - 1, size)] + x[IX(size - 1, size - 2, size)]);
    /* :1261-1287 obstacle mirroring */
    for (int i = 1; i < size - 1; i++) {
        for (int j = 1; j < size - 1; j++) {
            int idx = IX(i, j, size);
            if (!obstacles[idx]) continue;
            if (b == 1) {
                float m = 0;
                int count = 0;
                if (!obstacles[IX(i - 1, j, size)]) { m += -x[IX(i - 1, j, size)]; count++; }
                if (!obstacles[IX(i + 1, j, size)]) { m += -x[IX(i + 1, j, size)]; count++; }
                x[idx] = count > 0 ? m / count : 0;
            } else if (b == 2) {
                float m = 0;
                int count = 0;
                if (!obstacles[IX(i, j - 1, size)]) { m += -x[IX(i, j - 1, size)]; count++; }
                if (!obstacles[IX(i, j + 1, size)]) { m += -x[IX(i, j + 1, size)]; count++; }
                x[idx] = count > 0 ? m / count : 0;
            }
        }
    }
}

/* FluidSim.cs:1292-1357 DiffuseWithJobs + :1034-1069 DiffuseJob */
void r2_diffuse_with_jobs(int size, int b, float *x, const float *x0, float diff, float dt,
                          const uint8_t *obstacles, int iters) {
    int total = size * size;
    float a = dt * diff * (size - 2) * (size - 2); /* :1295 */
    float c = 1 + 6 * a;                           /* :1296 */
    float *buffer1 = malloc(sizeof(float) * total); /* :1299-1300: both start as x0 */
    float *buffer2 = malloc(sizeof(float) * total);
    memcpy(buffer1, x0, sizeof(float) * total);
    memcpy(buffer2, x0, sizeof(float) * total);
    float *input = buffer1, *output = buffer2;
    for (int k = 0; k < iters; k++) { /* :1310 (20 in the reference) */
        for (int index = 0; index < total; index++) {
            int i = index % size, j = index / size;
            if (i <= 0 || i >= size - 1 || j <= 0 || j >= size - 1) continue; /* :1051 */
            if (obstacles[index]) continue;                                   /* :1055 */
            output[index] = (input[index] + a * (input[IX(i + 1, j, size)] + input[IX(i - 1, j, size)] +
                                                 input[IX(i, j + 1, size)] + input[IX(i, j - 1, size)])) /
                            c; /* :1062-1067 */
        }
        r2_boundary(size, b, output, obstacles); /* :1327-1339 */
        float *t = input; input = output; output = t; /* :1342-1344 */
    }
    memcpy(x, input, sizeof(float) * total); /* :1348 */
    free(buffer1);
    free(buffer2);
}

/* FluidSim.cs:1188-1233 LinearSolveIterationJob, shared by :1359-1415 and :1578-1637 */
static void r2_linsolve_iteration(int size, const float *x0, const float *xRead, float *xWrite,
                                  const uint8_t *obstacles, float a, float c) {
    int total = size * size;
    for (int index = 0; index < total; index++) {
        int i = index % size, j = index / size;
        if (i <= 0 || i >= size - 1 || j <= 0 || j >= size - 1) { xWrite[index] = xRead[index]; continue; }
        if (obstacles[index]) { xWrite[index] = xRead[index]; continue; }
        int left = IX(i - 1, j, size), right = IX(i + 1, j, size);
        int top = IX(i, j + 1, size), bottom = IX(i, j - 1, size);
        xWrite[index] = (x0[index] + a * (xRead[right] + xRead[left] + xRead[top] + xRead[bottom])) / c;
    }
}

/* FluidSim.cs:1359-1415 LinearSolveWithJobs.  jobBuffer2's stale content is
 * irrelevant because every cell of the write buffer is assigned (:1209/:1216/:1227). */
void r2_linear_solve_with_jobs(int size, int b, float *x, const float *x0, float a, float c,
                               const uint8_t *obstacles, int iters) {
    int total = size * size;
    float *b1 = malloc(sizeof(float) * total), *b2 = calloc(total, sizeof(float));
    memcpy(b1, x, sizeof(float) * total); /* :1367 */
    float *rd = b1, *wr = b2;
    for (int k = 0; k < iters; k++) { /* :1378 */
        r2_linsolve_iteration(size, x0, rd, wr, obstacles, a, c);
        r2_boundary(size, b, wr, obstacles); /* :1399 */
        float *t = rd; rd = wr; wr = t;
    }
    memcpy(x, rd, sizeof(float) * total); /* :1408 */
    free(b1);
    free(b2);
}

/* FluidSim.cs:740-745 Diffuse */
void r2_diffuse(int size, int b, float *x, const float *x0, float diff, float dt,
                const uint8_t *obstacles, int iters) {
    r2_diffuse_with_jobs(size, b, x, x0, diff, dt, obstacles, iters);
    float a = dt * diff * (size - 2) * (size - 2);
    r2_linear_solve_with_jobs(size, b, x, x0, a, 1 + 6 * a, obstacles, iters);
}

/* FluidSim.cs:1417-1521 ProjectWithJobs (+ :1071-1123 jobs, :1578-1637 pressure solve).
 * velocX/velocY updated in place; p receives the pressure (also what the
 * reference copies into `pressure`, :1509).  The managed `div` argument of the
 * reference is never touched, so it is not a parameter here. */
void r2_project_with_jobs(int size, float *velocX, float *velocY, float *p, const uint8_t *obstacles,
                          int iters) {
    int total = size * size;
    float *nP = calloc(total, sizeof(float));   /* :1427 cleared */
    float *nDiv = calloc(total, sizeof(float)); /* :1428 cleared */
    for (int index = 0; index < total; index++) { /* :1080-1095 */
        int i = index % size, j = index / size;
        if (i <= 0 || i >= size - 1 || j <= 0 || j >= size - 1) continue;
        nDiv[index] = -0.5f * (velocX[IX(i + 1, j, size)] - velocX[IX(i - 1, j, size)] +
                               velocY[IX(i, j + 1, size)] - velocY[IX(i, j - 1, size)]) / size;
        nP[index] = 0;
    }
    r2_boundary(size, 0, nDiv, obstacles); /* :1446-1466 */
    r2_boundary(size, 0, nP, obstacles);
    { /* :1578-1637 PressureSolveWithJobs, a = 1, c = 6 */
        float a = 1.0f, c = 6.0f;
        float *tmp = calloc(total, sizeof(float));
        float *rd = nP, *wr = tmp;
        for (int k = 0; k < iters; k++) {
            r2_linsolve_iteration(size, nDiv, rd, wr, obstacles, a, c);
            r2_boundary(size, 0, wr, obstacles);
            float *t = rd; rd = wr; wr = t;
        }
        if (rd != nP) memcpy(nP, rd, sizeof(float) * total); /* :1628-1631 */
        free(tmp);
    }
    for (int index = 0; index < total; index++) { /* :1107-1122 */
        int i = index % size, j = index / size;
        if (i <= 0 || i >= size - 1 || j <= 0 || j >= size - 1) continue;
        if (obstacles[index]) continue;
        velocX[index] -= 0.5f * (nP[IX(i + 1, j, size)] - nP[IX(i - 1, j, size)]) * size;
        velocY[index] -= 0.5f * (nP[IX(i, j + 1, size)] - nP[IX(i, j - 1, size)]) * size;
    }
    r2_boundary(size, 1, velocX, obstacles); /* :1483-1503 */
    r2_boundary(size, 2, velocY, obstacles);
    memcpy(p, nP, sizeof(float) * total); /* :1508 */
    free(nP);
    free(nDiv);
}

/* FluidSim.cs:1523-1576 AdvectWithJobs + :1125-1186 AdvectJob.  d must not alias d0. */
void r2_advect_with_jobs(int size, int b, float *d, const float *d0, const float *velocX,
                         const float *velocY, float dt, const uint8_t *obstacles) {
    int total = size * size;
    float dt0 = dt * (size - 2);              /* :1526 */
    float *nD = calloc(total, sizeof(float)); /* :1529 fresh, zeroed output */
    for (int index = 0; index < total; index++) {
        int i = index % size, j = index / size;
        if (i <= 0 || i >= size - 1 || j <= 0 || j >= size - 1) continue;
        if (obstacles[index] && (b == 1 || b == 2)) { nD[index] = 0; continue; } /* :1148-1152 */
        if (obstacles[index]) continue; /* :1155 -> stays 0 because nD is fresh */
        float x = i - dt0 * velocX[index];
        float y = j - dt0 * velocY[index];
        if (x < 0.5f) x = 0.5f;
        if (x > size - 1.5f) x = size - 1.5f;
        int i0 = (int)x, i1 = i0 + 1;
        if (y < 0.5f) y = 0.5f;
        if (y > size - 1.5f) y = size - 1.5f;
        int j0 = (int)y, j1 = j0 + 1;
        float s1 = x - i0, s0 = 1 - s1, t1 = y - j0, t0 = 1 - t1;
        nD[index] = s0 * (t0 * d0[IX(i0, j0, size)] + t1 * d0[IX(i0, j1, size)]) +
                    s1 * (t0 * d0[IX(i1, j0, size)] + t1 * d0[IX(i1, j1, size)]); /* :1183-1184 */
    }
    r2_boundary(size, b, nD, obstacles); /* :1553-1562 */
    memcpy(d, nD, sizeof(float) * total); /* :1565 */
    free(nD);
}

/* FluidSim.cs:617-673 EnforceObstacleBoundaries + ApplyDragNearObstacle.
 * Mathf.Sqrt/Exp are (float) of the double functions; Mathf.Lerp clamps t to [0,1]. */
void r2_enforce_obstacles(int size, float *velocityX, float *velocityY, const uint8_t *obstacles,
                          float cellSize, float viscosity) {
    static const int di[4] = {-1, 1, 0, 0}, dj[4] = {0, 0, -1, 1};
    for (int i = 1; i < size - 1; i++) {
        for (int j = 1; j < size - 1; j++) {
            int idx = IX(i, j, size);
            if (!obstacles[idx]) continue;
            velocityX[idx] = 0;
            velocityY[idx] = 0;
            for (int n = 0; n < 4; n++) {
                int ni = i + di[n], nj = j + dj[n];
                if (ni < 1 || ni >= size - 1 || nj < 1 || nj >= size - 1) continue;
                int nidx = IX(ni, nj, size);
                if (obstacles[nidx]) continue;
                float U = (float)sqrt((double)(velocityX[nidx] * velocityX[nidx] + velocityY[nidx] * velocityY[nidx]));
                float visc = viscosity > 1e-5f ? viscosity : 1e-5f;
                float Re = (U * cellSize) / visc;
                float t = 1.0f - (float)exp((double)(-Re * 0.01f));
                if (t < 0.0f) t = 0.0f;
                if (t > 1.0f) t = 1.0f;
                float drag = 0.8f + (0.98f - 0.8f) * t;
                velocityX[nidx] *= drag;
                velocityY[nidx] *= drag;
            }
        }
    }
}

/* FluidSim.cs:703-721 VelocityStep + DensityStep, :551-570 Simulate (dt/visc/diff are the
 * already-scaled effective values of :554-556).  State arrays as in :112-117. */
void r2_simulate(int size, float *density, float *velocityX, float *velocityY, float *velocityX0,
                 float *velocityY0, float *pressure, const uint8_t *obstacles, float dt, float visc,
                 float diff, int enableObstacle, float cellSize, float rawViscosity, int iters) {
    int total = size * size;
    r2_diffuse(size, 1, velocityX0, velocityX, visc, dt, obstacles, iters); /* :705 */
    r2_diffuse(size, 2, velocityY0, velocityY, visc, dt, obstacles, iters); /* :706 */
    r2_project_with_jobs(size, velocityX0, velocityY0, velocityX, obstacles, iters); /* :708 p -> velocityX */
    memcpy(pressure, velocityX, sizeof(float) * total);                       /* :1509 */
    {   /* :710-711.  AdvectWithJobs copies every input to native arrays first (:1530-1532),
           so d aliasing is impossible; here d != d0 already. */
        r2_advect_with_jobs(size, 1, velocityX, velocityX0, velocityX0, velocityY0, dt, obstacles);
        r2_advect_with_jobs(size, 2, velocityY, velocityY0, velocityX0, velocityY0, dt, obstacles);
    }
    r2_project_with_jobs(size, velocityX, velocityY, velocityX0, obstacles, iters); /* :713 p -> velocityX0 */
    memcpy(pressure, velocityX0, sizeof(float) * total);
    float *densityTemp = calloc(total, sizeof(float));                               /* :718 */
    r2_diffuse(size, 0, densityTemp, density, diff, dt, obstacles, iters);           /* :719 */
    r2_advect_with_jobs(size, 0, density, densityTemp, velocityX, velocityY, dt, obstacles); /* :720 */
    free(densityTemp);
    if (enableObstacle) r2_enforce_obstacles(size, velocityX, velocityY, obstacles, cellSize, rawViscosity); /* :567-570 */
}

/* FluidSim.cs:723-738 AddDensity / AddVelocity: (int) truncation then clamp. */
static int r2_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
void r2_add_density(int size, float *density, float x, float y, float amount) {
    int i = r2_clampi((int)x, 0, size - 1), j = r2_clampi((int)y, 0, size - 1);
    density[IX(i, j, size)] += amount;
}
void r2_add_velocity(int size, float *velocityX, float *velocityY, float x, float y, float ax, float ay) {
    int i = r2_clampi((int)x, 0, size - 1), j = r2_clampi((int)y, 0, size - 1);
    velocityX[IX(i, j, size)] += ax;
    velocityY[IX(i, j, size)] += ay;
}
