// fs_core.h -- orchestration of the stable-fluids step (the reference's L2 layer:
// Simulate/VelocityStep/DensityStep/Diffuse/ProjectWithJobs/AdvectWithJobs, FluidSim.cs:551-576,
// :703-745, :1292-1655), written once and parameterised on an executor that runs the sweeps.
//
//   libfluidsolver.so          = SolverCore<CudaExec>   (fluidsolver.cu; the product, sm_100a kernels)
//   tests/host_emul/*.so       = SolverCore<HostExec>   (test scaffolding only: lets the CPU-only test
//                                 suite exercise this orchestration and fs_cellops.cuh without a GPU)
//
// Buffer plan (no per-call allocate-and-copy as in the reference, :1299-1301, :1425-1429, :1529-1533):
// 11 persistent fields (9 in 2D) whose roles rotate by pointer swap; 1 flag byte per voxel.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/fluidsolver.h"
#include "fs_cellops.cuh"

template <class Exec>
struct SolverCore {
    fs_params prm;
    FsGrid g;
    Exec ex;
    long long nloc = 0;   // voxels in a local array (incl. ghost planes)
    long long nowned = 0; // owned voxels
    int zb = 0, ze = 0;   // owned GLOBAL planes [zb, ze)
    // logical fields (device pointers); roles rotate by swapping
    float *density = nullptr, *dens0 = nullptr;
    float *vx = nullptr, *vy = nullptr, *vz = nullptr, *vx0 = nullptr, *vy0 = nullptr, *vz0 = nullptr;
    float *pressure = nullptr, *div = nullptr, *tmp = nullptr;
    std::vector<float *> allocated;
    uint8_t *mask = nullptr, *flags = nullptr;
    long long *obst_list = nullptr; // local indices of OWNED interior obstacle cells
    long long n_obst = 0;
    bool any_obstacle = false;
    // GLOBAL obstacle presence: every slab must take the same decisions about which halo operations exist,
    // otherwise the ranks' op sequences (and sequence numbers) diverge and the exchange deadlocks.
    bool g_any_obstacle = false, g_interior_obstacle = false;
    std::string err;

    // ---- lifetime ----------------------------------------------------------------------------
    int init(const fs_params &p) {
        prm = p;
        if (p.abi_version != FS_ABI_VERSION) return fail(FS_ERR_BAD_ARGUMENT, "abi_version mismatch");
        if (p.nx < 3 || p.ny < 3 || (p.nz != 1 && p.nz < 3)) return fail(FS_ERR_BAD_ARGUMENT, "nx, ny >= 3 and nz == 1 or nz >= 3 required");
        if (p.iters_diffuse < 0 || p.iters_pressure < 0) return fail(FS_ERR_BAD_ARGUMENT, "negative iteration count");
        if (p.slab_count < 1 || p.slab_rank < 0 || p.slab_rank >= p.slab_count) return fail(FS_ERR_BAD_ARGUMENT, "bad slab_rank/slab_count");
        if (p.slab_count > 1 && (p.nz == 1 || p.nz / p.slab_count < FS_GHOST)) return fail(FS_ERR_BAD_ARGUMENT, "z-slabs need nz/slab_count >= 2");
        if (p.solver_kind != FS_JACOBI && p.solver_kind != FS_RED_BLACK) return fail(FS_ERR_BAD_ARGUMENT, "unknown solver_kind");
        g.nx = p.nx; g.ny = p.ny; g.nz = p.nz; g.hz = p.nz > 1;
        g.sy = p.nx; g.sz = (long long)p.nx * p.ny;
        // balanced contiguous z partition
        const int P = p.slab_count, r = p.slab_rank;
        zb = (int)((long long)p.nz * r / P);
        ze = (int)((long long)p.nz * (r + 1) / P);
        // FS_GHOST ghost planes per internal side: the fused two-stage sweeps read two planes beyond the owned range
        const int lo = r > 0 ? zb - FS_GHOST : zb, hi = r < P - 1 ? ze + FS_GHOST : ze;
        g.zoff = lo; g.nzl = hi - lo; g.kb = zb - lo; g.ke = ze - lo;
        nloc = g.sz * g.nzl;
        nowned = g.sz * (ze - zb);
        ex.use_graph = p.use_cuda_graph != 0;
        int rc = ex.open(p.device_id);
        if (rc) return fail(FS_ERR_CUDA, ex.error());
        float **fields[] = {&density, &dens0, &vx, &vy, &vx0, &vy0, &pressure, &div, &tmp, &vz, &vz0};
        const int nf = g.hz ? 11 : 9;
        for (int i = 0; i < nf; i++) {
            *fields[i] = (float *)ex.alloc(sizeof(float) * nloc);
            if (!*fields[i]) return fail(FS_ERR_OUT_OF_MEMORY, ex.error());
            allocated.push_back(*fields[i]);
        }
        mask = (uint8_t *)ex.alloc(nloc);
        flags = (uint8_t *)ex.alloc(nloc);
        if (!mask || !flags) return fail(FS_ERR_OUT_OF_MEMORY, ex.error());
        return reset();
    }

    void destroy() {
        ex.sync();
        for (float *p : allocated) ex.free(p);
        allocated.clear();
        ex.free(mask); ex.free(flags); ex.free(obst_list);
        mask = flags = nullptr; obst_list = nullptr;
        ex.close();
    }

    int reset() { // FluidSim.cs:225-232: every field and the mask start at zero
        for (float *p : allocated) ex.zero(p, sizeof(float) * nloc);
        ex.zero(mask, nloc);
        ex.zero(flags, nloc);
        ex.free(obst_list);
        obst_list = nullptr; n_obst = 0; any_obstacle = false;
        g_any_obstacle = g_interior_obstacle = false;
        // captured graphs hold the freed obstacle list and the mirror / enforce / halo decisions taken with it
        ex.invalidate_graph();
        return check();
    }

    int fail(int code, const std::string &m) { err = m; return code; }
    int check() {
        ex.halo_commit();
        if (ex.failed()) return fail(FS_ERR_CUDA, ex.error());
        return FS_OK;
    }
    const uint8_t *fl() const { return any_obstacle ? flags : nullptr; }

    // ---- obstacles (SetupObstacles output, FluidSim.cs:302-327) --------------------------------
    // The mask of the local planes is on the device: derive the flag bytes, the local "any obstacle" bit and the list of
    // owned interior obstacle cells there too (no host pass over the grid), and record the GLOBAL presence bits.
    int finish_obstacles(bool global_any, bool global_interior) {
        ex.build_flags(g, mask, flags);
        ex.free(obst_list);
        obst_list = nullptr;
        n_obst = 0;
        const int k0 = (g.hz ? std::max(zb, 1) : 0) - g.zoff, k1 = (g.hz ? std::min(ze, g.nz - 1) : 1) - g.zoff;
        bool any = false;
        if (!ex.scan_obstacles(g, mask, k0, k1, &any, &obst_list, &n_obst)) return fail(FS_ERR_OUT_OF_MEMORY, ex.error());
        any_obstacle = any;
        g_any_obstacle = global_any;
        g_interior_obstacle = global_interior;
        ex.invalidate_graph();
        return check();
    }
    int set_obstacles(const uint8_t *gmask, long long n) {
        if (!gmask || n != g.sz * g.nz) return fail(FS_ERR_BAD_ARGUMENT, "mask must have nx*ny*nz bytes");
        ex.upload(mask, gmask + g.sz * g.zoff, nloc);
        // global presence: one early-exit pass on the host (the typical mask has an obstacle and stops at once)
        bool ga = false, gi = false;
        const long long total = g.sz * g.nz;
        for (long long i = 0; i < total && !ga; i++) ga = gmask[i] != 0;
        const int gk0 = g.hz ? 1 : 0, gk1 = g.hz ? g.nz - 1 : 1;
        for (int k = gk0; k < gk1 && ga && !gi; k++)
            for (int j = 1; j <= g.ny - 2 && !gi; j++) {
                const uint8_t *row = gmask + j * g.sy + k * g.sz;
                for (int i = 1; i <= g.nx - 2; i++)
                    if (row[i]) { gi = true; break; }
            }
        return finish_obstacles(ga, gi);
    }
    int set_obstacles_slab(const uint8_t *lmask, long long n, bool global_any, bool global_interior) {
        if (!lmask || n != nloc) return fail(FS_ERR_BAD_ARGUMENT, "mask must hold the planes of fs_slab_halo_range");
        ex.upload(mask, lmask, nloc);
        return finish_obstacles(global_any, global_interior);
    }
    // Device-side SetupObstacles (next row N4): every handle builds its own planes from the shape parameters.
    int build_obstacles(const fs_obstacle_shape *sh, int64_t *cells) {
        if (!sh || sh->kind < 0 || sh->kind > 2) return fail(FS_ERR_BAD_ARGUMENT, "unknown obstacle shape");
        long long total = 0, interior = 0;
        if (!ex.build_shape(g, *sh, mask, &total, &interior)) return fail(FS_ERR_OUT_OF_MEMORY, ex.error());
        if (cells) *cells = total;
        return finish_obstacles(total > 0, interior > 0);
    }
    int get_obstacles(uint8_t *out, long long n) {
        if (!out || n != nowned) return fail(FS_ERR_BAD_ARGUMENT, "n must equal the owned voxel count");
        ex.download(out, mask + g.sz * g.kb, (size_t)nowned);
        return check();
    }

    // ---- sources (AddDensity/AddVelocity, FluidSim.cs:723-738) ----------------------------------
    static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
    // local index of the cell the reference would pick, or -1 when it lies outside this slab's planes
    long long source_cell(float x, float y, float z) const {
        const int i = clampi((int)x, 0, g.nx - 1), j = clampi((int)y, 0, g.ny - 1);
        const int k = g.hz ? clampi((int)z, 0, g.nz - 1) : 0;
        if (k < g.zoff || k >= g.zoff + g.nzl) return -1;
        return fs_idx(g, i, j, k - g.zoff);
    }
    // The cell list goes through a small ring of host staging buffers (pinned on the GPU build): the call returns as
    // soon as the copy and the scatter kernel are enqueued -- no stream synchronisation per frame.
    int add_cells(long long count, const float *x, const float *y, const float *z, const float *d, const float *ax,
                  const float *ay, const float *az) {
        if (count < 0 || !x || !y) return fail(FS_ERR_BAD_ARGUMENT, "bad source list");
        if (count == 0) return FS_OK;
        long long *idx = nullptr;
        float *amt[4] = {};
        if (!ex.source_stage(count, &idx, amt)) return fail(FS_ERR_OUT_OF_MEMORY, ex.error());
        long long m = 0;
        for (long long n = 0; n < count; n++) {
            const long long c = source_cell(x[n], y[n], z ? z[n] : 0.0f);
            if (c < 0) continue;
            idx[m] = c;
            amt[0][m] = d ? d[n] : 0.0f;
            amt[1][m] = ax ? ax[n] : 0.0f;
            amt[2][m] = ay ? ay[n] : 0.0f;
            amt[3][m] = az ? az[n] : 0.0f;
            m++;
        }
        float *dst[4] = {d ? density : nullptr, ax ? vx : nullptr, ay ? vy : nullptr, (az && g.hz) ? vz : nullptr};
        ex.scatter_add_staged(dst, count, m);
        return check();
    }
    int add_dense(const float *d, const float *ax, const float *ay, const float *az) {
        const float *src[4] = {d, ax, ay, g.hz ? az : nullptr};
        float *dst[4] = {density, vx, vy, vz};
        for (int f = 0; f < 4; f++) {
            if (!src[f]) continue;
            ex.upload(tmp + g.sz * g.kb, src[f], sizeof(float) * nowned); // tmp is scratch between steps
            ex.axpy(dst[f] + g.sz * g.kb, tmp + g.sz * g.kb, nowned);
            ex.halo(g, dst[f]);
        }
        return check();
    }

    // ---- Diffuse (FluidSim.cs:740-745) ----------------------------------------------------------
    static void coeffs(int n, float diff, float dt, float *a, float *c) {
        const float av = dt * diff * (float)(n - 2) * (float)(n - 2); // :743, left to right
        *a = av;
        *c = 1.0f + 6.0f * av; // :744
    }
    void mirror(float *x, int b) {
        if (b != 0 && n_obst && (b != 3 || g.hz)) ex.mirror(g, x, flags, obst_list, n_obst, b);
    }
    // One relaxation sweep with all of BoundaryJob fused into it: the set_bnd ring (ring scatter) and, for velocity
    // components, the obstacle mirroring (fs_mirror_fused recomputes the two neighbour values it needs).  On z-slabs the
    // executor overlaps the push of the boundary planes with the interior of the same sweep.
    void relax_op(int mode, const float *in, const float *rhs, const float *stale, float *out, float a, float c, int b,
                  bool in_zero, int xmode) {
        ex.relax(mode, g, in, rhs, stale, out, fl(), a, c, b, in_zero, xmode);
    }
    // z-slabs: sweeps without obstacle mirroring go in twos per halo operation (the ghost zone is two planes deep): sweep
    // `it` of `iters` is an extended one when another sweep follows it, the exchanging one after it leaves its fork open
    // when a further couple follows.  Every rank takes the same decisions (they only depend on global facts).
    bool extendable(int b) const { return ex.can_extend(g) && !needs_mirror(b); }
    static int sweep_xmode(bool ext, int it, int iters) {
        if (!ext) return FS_X_EXCHANGE;
        if ((it & 1) == 0) return it + 1 < iters ? FS_X_EXTEND : FS_X_EXCHANGE;
        return it + 2 < iters ? FS_X_EXCHANGE_OPEN : FS_X_EXCHANGE;
    }
    // Whether sweeps of field kind b may be fused in pairs: no obstacle mirroring between the two stages.
    bool needs_mirror(int b) const { return b != 0 && g_interior_obstacle && (b != 3 || g.hz); }
    bool pair_ok(int b, float c, int kind) const { return !needs_mirror(b) && ex.pair_supported(g, c, kind); }
    // pass 1, DiffuseWithJobs :1292-1357.  Result ends in `x` (roles of x and tmp may swap).
    void smooth(int b, float *&x, const float *x0, float a, float c, int iters) {
        if (iters == 0) { ex.copy(x, x0, sizeof(float) * nloc); return; }
        float *A = tmp, *B = x;
        const float *in = x0;
        const bool pairs = pair_ok(b, c, FS_PAIR_SMOOTH);
        // Obstacle cells keep the stale content of the write buffer (DiffuseJob does not write them, :1055).  Both
        // reference buffers start as copies of x0 (:1299-1300) and, without mirroring, nothing ever changes those cells:
        // they hold x0 throughout, so `stale = x0` is exact for every iteration.  With mirroring (b != 0 and interior
        // obstacles) the stale content is the mirrored value of two iterations ago: x0 for the first two, then the buffer.
        const bool mir = needs_mirror(b);
        const bool ext = !pairs && extendable(b);
        int sweeps = 0;
        for (int it = 0; it < iters;) {
            float *out = (sweeps & 1) ? B : A;
            if (pairs && it + 2 <= iters && ex.relax_pair(FS_PAIR_SMOOTH, g, in, nullptr, out, fl(), a, c, b, false, true)) {
                it += 2;
            } else {
                relax_op(FS_MODE_SMOOTH, in, nullptr, (!mir || it < 2) ? x0 : nullptr, out, a, c, b, false, sweep_xmode(ext, it, iters));
                it += 1;
            }
            in = out;
            sweeps++;
        }
        ex.relax_end();
        float *res = const_cast<float *>(in);
        tmp = res == A ? B : A;
        x = res;
    }
    // pass 2 / pressure, LinearSolveWithJobs :1359-1415, PressureSolveWithJobs :1578-1637
    void lin_solve(int b, float *&x, const float *rhs, float a, float c, int iters, bool zero_guess) {
        if (iters == 0) { if (zero_guess) ex.zero(x, sizeof(float) * nloc); return; }
        float *rd = x, *wr = tmp;
        const bool pairs = pair_ok(b, c, FS_PAIR_JACOBI);
        const bool ext = !pairs && extendable(b); // (the right-hand side must then be valid one plane into the ghost zone)
        for (int it = 0; it < iters;) {
            const bool iz = zero_guess && it == 0;
            if (pairs && it + 2 <= iters && ex.relax_pair(FS_PAIR_JACOBI, g, rd, rhs, wr, fl(), a, c, b, iz, true)) {
                it += 2;
            } else {
                relax_op(FS_MODE_JACOBI, rd, rhs, nullptr, wr, a, c, b, iz, sweep_xmode(ext, it, iters));
                it += 1;
            }
            std::swap(rd, wr);
        }
        ex.relax_end();
        x = rd;
        tmp = wr;
    }
    // Red-black Gauss-Seidel (BASELINE config 5).  Fused form: both colour passes and set_bnd in one out-of-place pass
    // (ping-pong like Jacobi).  Fallback: in place, one launch per colour.
    void lin_solve_rb(int b, float *&x, const float *rhs, float a, float c, int iters, bool zero_guess) {
        if (pair_ok(b, c, FS_PAIR_RED_BLACK) && iters > 0) {
            float *rd = x, *wr = tmp;
            for (int it = 0; it < iters; it++) {
                ex.relax_pair(FS_PAIR_RED_BLACK, g, rd, rhs, wr, fl(), a, c, b, zero_guess && it == 0, true);
                std::swap(rd, wr);
            }
            x = rd;
            tmp = wr;
            return;
        }
        if (zero_guess) {
            ex.zero(x, sizeof(float) * nloc);
            // slabs: a faster neighbour must not store its first boundary plane into this slab's ghost plane before
            // the memset above has run (it would be wiped): nobody passes the fence before everybody has zeroed
            ex.halo_fence();
        }
        for (int it = 0; it < iters; it++) {
            bool ringed = false;
            for (int colour = 0; colour < 2; colour++) {
                ringed = ex.rb_half(g, x, rhs, fl(), a, c, colour, b);
                ex.halo(g, x);
            }
            if (!ringed) ex.bnd(g, x, b); // the float4 kernel writes the set_bnd ring in its colour-1 launch
            mirror(x, b);
            if (b != 0) ex.halo(g, x);
        }
    }
    // ---- batched Diffuse of the velocity components (VelocityStep :705-706 + the z component) ---------------------------
    // The components share a, c and the obstacle flags and are independent until the projection, so each of their 2*K_d
    // sweeps is ONE launch over all of them (and one halo operation on z-slabs) instead of one per component.  Ping-pong
    // partners: tmp, pressure and div, all dead at this point of the step (ProjectWithJobs rewrites the latter two).
    void diffuse_velocity(float visc, float dt) {
        float a, c;
        coeffs(g.nx, visc, dt, &a, &c);
        const int nf = g.hz ? 3 : 2, iters = prm.iters_diffuse;
        if (pair_ok(1, c, FS_PAIR_SMOOTH) || iters == 0) { // fused pairs are issued per field
            diffuse(1, vx0, vx, visc, dt);
            diffuse(2, vy0, vy, visc, dt);
            if (g.hz) diffuse(3, vz0, vz, visc, dt);
            return;
        }
        const int b[3] = {1, 2, 3};
        float **x[3] = {&vx0, &vy0, &vz0}, **scratch[3] = {&tmp, &pressure, &div};
        const float *x0[3] = {vx, vy, vz};
        // pass 1 (DiffuseWithJobs): x0 -> scratch -> x -> ...
        const float *in[3] = {x0[0], x0[1], x0[2]}, *stale[3];
        float *out[3], *A[3], *B[3];
        for (int f = 0; f < nf; f++) { A[f] = *scratch[f]; B[f] = *x[f]; }
        const bool ext = extendable(1); // (needs_mirror is the same for every component that exists)
        for (int it = 0; it < iters; it++) {
            for (int f = 0; f < nf; f++) {
                out[f] = (it & 1) ? B[f] : A[f];
                stale[f] = (!needs_mirror(b[f]) || it < 2) ? x0[f] : nullptr; // see smooth()
            }
            ex.relax_n(FS_MODE_SMOOTH, g, nf, in, nullptr, stale, out, fl(), a, c, b, false, sweep_xmode(ext, it, iters));
            for (int f = 0; f < nf; f++) in[f] = out[f];
        }
        // pass 2 (LinearSolveWithJobs) seeded with pass 1's result, rhs = x0
        float *rd[3], *wr[3];
        for (int f = 0; f < nf; f++) { rd[f] = const_cast<float *>(in[f]); wr[f] = rd[f] == A[f] ? B[f] : A[f]; }
        for (int it = 0; it < iters; it++) {
            const float *rdc[3] = {rd[0], rd[1], rd[2]};
            ex.relax_n(FS_MODE_JACOBI, g, nf, rdc, x0, nullptr, wr, fl(), a, c, b, false, sweep_xmode(ext, it, iters));
            for (int f = 0; f < nf; f++) std::swap(rd[f], wr[f]);
        }
        ex.relax_end();
        for (int f = 0; f < nf; f++) { *x[f] = rd[f]; *scratch[f] = wr[f]; }
    }
    void diffuse(int b, float *&x, const float *x0, float diff, float dt) {
        float a, c;
        coeffs(g.nx, diff, dt, &a, &c);
        smooth(b, x, x0, a, c, prm.iters_diffuse);
        lin_solve(b, x, x0, a, c, prm.iters_diffuse, false);
    }

    // ---- ProjectWithJobs (FluidSim.cs:1417-1521) ------------------------------------------------
    void project(float *ux, float *uy, float *uz) {
        ex.divergence(g, div, ux, uy, uz);
        // slabs + fused or extended sweeps: those are evaluated one plane into the ghost zone and read the right-hand
        // side there (a plain single sweep only reads div on owned planes)
        if (pair_ok(0, 6.0f, prm.solver_kind == FS_RED_BLACK ? FS_PAIR_RED_BLACK : FS_PAIR_JACOBI) ||
            (prm.solver_kind != FS_RED_BLACK && extendable(0)))
            ex.halo(g, div);
        if (prm.solver_kind == FS_RED_BLACK)
            lin_solve_rb(0, pressure, div, 1.0f, 6.0f, prm.iters_pressure, true);
        else
            lin_solve(0, pressure, div, 1.0f, 6.0f, prm.iters_pressure, true); // :1581-1582, p starts 0
        ex.gradient(g, ux, uy, uz, pressure, fl());
        finish_velocity(ux, uy, uz);
    }
    // BoundaryJob's obstacle pass on all velocity components (one launch) and their halos (one operation).
    void finish_velocity(float *ux, float *uy, float *uz) {
        if (g.hz && g_interior_obstacle) ex.halo(g, uz); // slabs: the z mirror reads the neighbours' new boundary planes
        if (n_obst) ex.mirror3(g, ux, uy, g.hz ? uz : nullptr, flags, obst_list, n_obst);
        float *fields[3] = {ux, uy, uz};
        ex.halo_n(g, fields, g.hz ? 3 : 2);
    }

    // ---- the step (Simulate, FluidSim.cs:551-570) -----------------------------------------------
    void step_body(float dt, float visc, float diff) {
        const float dt0 = dt * (float)(g.nx - 2); // :1526
        // VelocityStep :703-714
        diffuse_velocity(visc, dt);
        project(vx0, vy0, vz0);
        ex.halo_fence(); // slabs: the back-trace may gather from a neighbour slab, whose fields must be complete
        ex.advect_velocity(g, vx, vy, vz, vx0, vy0, vz0, fl(), dt0); // :710-711
        finish_velocity(vx, vy, vz);
        project(vx, vy, vz);
        // DensityStep :716-721
        diffuse(0, dens0, density, diff, dt);
        ex.halo_fence();
        ex.advect(g, density, dens0, vx, vy, vz, fl(), dt0, 0);
        ex.halo(g, density);
        if (prm.enable_obstacle && g_any_obstacle) { // :567-570 (global decision; the kernel is a no-op where flags are 0)
            ex.enforce(g, vx, vy, vz, flags, prm.cell_size, prm.raw_viscosity);
            float *fields[3] = {vx, vy, vz};
            ex.halo_n(g, fields, g.hz ? 3 : 2);
        }
        ex.halo_commit();
    }

    int step(float dt, float visc, float diff) {
        // The launch sequence is static for given (dt, visc, diff): the executor may capture it once
        // (CUDA graph) and replay it.  Pointer roles rotate inside step_body, so a replay is only valid
        // if the rotation is the identity over one step -- checked by the executor via the roles hash.
        float *before[11] = {density, dens0, vx, vy, vz, vx0, vy0, vz0, pressure, div, tmp};
        if (!ex.replay_step(dt, visc, diff, before)) {
            ex.begin_step(dt, visc, diff, before);
            step_body(dt, visc, diff);
            float *after[11] = {density, dens0, vx, vy, vz, vx0, vy0, vz0, pressure, div, tmp};
            ex.end_step(after);
        } else {
            float **roles[11] = {&density, &dens0, &vx, &vy, &vz, &vx0, &vy0, &vz0, &pressure, &div, &tmp};
            ex.roles_after_replay(roles);
        }
        return check();
    }

    // ---- visualisation (next row N2) -----------------------------------------------------------------------
    int render(const fs_vis_params *vp, float *out, long long n) {
        if (!vp || !out || n != g.sy * g.ny * 4) return fail(FS_ERR_BAD_ARGUMENT, "out_rgba must hold nx*ny*4 floats");
        if (vp->gradient_key_count < 0 || vp->gradient_key_count > 8) return fail(FS_ERR_BAD_ARGUMENT, "gradient_key_count must be 0..8");
        int kl = 0;
        if (g.hz) {
            if (vp->z_slice < zb || vp->z_slice >= ze) return fail(FS_ERR_BAD_ARGUMENT, "z_slice is not owned by this handle");
            kl = vp->z_slice - g.zoff;
        }
        float *dev = (float *)ex.render_buffer(sizeof(float) * (size_t)n);
        if (!dev) return fail(FS_ERR_OUT_OF_MEMORY, ex.error());
        ex.visualize(g, *vp, density + g.sz * kl, pressure + g.sz * kl, mask + g.sz * kl, dev);
        ex.download(out, dev, sizeof(float) * (size_t)n);
        return check();
    }

    // ---- streamline glyphs (next row N3) -------------------------------------------------------------------
    int streamlines(int skip, float scale, int z_slice, float *out, long long count) {
        if (skip < 1 || !out) return fail(FS_ERR_BAD_ARGUMENT, "skip >= 1 and a destination are required");
        const long long want = (long long)(g.nx / skip) * (g.ny / skip);
        if (count != want) return fail(FS_ERR_BAD_ARGUMENT, "count must equal (nx/skip)*(ny/skip)");
        int kl = 0;
        if (g.hz) {
            if (z_slice < zb || z_slice >= ze) return fail(FS_ERR_BAD_ARGUMENT, "z_slice is not owned by this handle");
            kl = z_slice - g.zoff;
        }
        if (count == 0) return FS_OK;
        float *dev = (float *)ex.render_buffer(sizeof(float) * 4 * (size_t)count);
        if (!dev) return fail(FS_ERR_OUT_OF_MEMORY, ex.error());
        ex.streamlines(g, skip, scale, vx + g.sz * kl, vy + g.sz * kl, mask + g.sz * kl, dev, count);
        ex.download(out, dev, sizeof(float) * 4 * (size_t)count);
        return check();
    }

    // ---- field access -----------------------------------------------------------------------------
    float *field_ptr(int f) {
        switch (f) {
        case FS_DENSITY: return density;
        case FS_VX: return vx;
        case FS_VY: return vy;
        case FS_VZ: return vz;
        case FS_VX0: return vx0;
        case FS_VY0: return vy0;
        case FS_VZ0: return vz0;
        case FS_PRESSURE: return pressure;
        case FS_DIVERGENCE: return div;
        default: return nullptr;
        }
    }
    float **field_slot(int f) {
        switch (f) {
        case FS_DENSITY: return &density;
        case FS_VX: return &vx;
        case FS_VY: return &vy;
        case FS_VZ: return &vz;
        case FS_VX0: return &vx0;
        case FS_VY0: return &vy0;
        case FS_VZ0: return &vz0;
        case FS_PRESSURE: return &pressure;
        case FS_DIVERGENCE: return &div;
        default: return nullptr;
        }
    }
    int get_field(int f, float *out, long long n) {
        float *p = field_ptr(f);
        if (!p) return fail(FS_ERR_BAD_ARGUMENT, "unknown or unallocated field");
        if (!out || n != nowned) return fail(FS_ERR_BAD_ARGUMENT, "n must equal the owned voxel count");
        ex.download(out, p + g.sz * g.kb, sizeof(float) * nowned);
        return check();
    }
    int get_field_async(int f, float *out, long long n) {
        float *p = field_ptr(f);
        if (!p) return fail(FS_ERR_BAD_ARGUMENT, "unknown or unallocated field");
        if (!out || n != nowned) return fail(FS_ERR_BAD_ARGUMENT, "n must equal the owned voxel count");
        ex.download_async(out, p + g.sz * g.kb, sizeof(float) * nowned, f);
        return check();
    }
    int set_field(int f, const float *in, long long n) {
        float *p = field_ptr(f);
        if (!p) return fail(FS_ERR_BAD_ARGUMENT, "unknown or unallocated field");
        if (!in || n != nowned) return fail(FS_ERR_BAD_ARGUMENT, "n must equal the owned voxel count");
        ex.upload(p + g.sz * g.kb, in, sizeof(float) * nowned);
        ex.halo(g, p);
        return check();
    }
};
