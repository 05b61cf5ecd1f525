// fs_abi.inl -- the extern "C" surface declared in include/fluidsolver.h, implemented over
// SolverCore<FS_EXEC>.  Included exactly once by fluidsolver.cu (FS_EXEC = CudaExec: the product) and
// by tests/host_emul/host_emul.cpp (FS_EXEC = HostExec: CPU-only test scaffolding).
#ifndef FS_EXEC
#error "define FS_EXEC before including fs_abi.inl"
#endif

#include <new>

struct fs_solver {
    SolverCore<FS_EXEC> core;
};

static std::string g_create_error;

#define FS_GUARD(s)                                                                                                    \
    if (!(s)) return FS_ERR_BAD_ARGUMENT;                                                                              \
    SolverCore<FS_EXEC> &c = (s)->core;                                                                                \
    c.ex.make_current();

extern "C" {

int fs_abi_version(void) { return FS_ABI_VERSION; }

int fs_create(const fs_params *params, fs_solver **out) {
    if (!params || !out) { g_create_error = "null argument"; return FS_ERR_BAD_ARGUMENT; }
    *out = nullptr;
    fs_solver *s = new (std::nothrow) fs_solver();
    if (!s) { g_create_error = "host allocation failed"; return FS_ERR_OUT_OF_MEMORY; }
    const int rc = s->core.init(*params);
    if (rc != FS_OK) {
        g_create_error = s->core.err;
        s->core.destroy();
        delete s;
        return rc;
    }
    *out = s;
    return FS_OK;
}

void fs_destroy(fs_solver *s) {
    if (!s) return;
    s->core.ex.make_current();
    s->core.destroy();
    delete s;
}

int fs_reset(fs_solver *s) { FS_GUARD(s); return c.reset(); }

const char *fs_last_error(const fs_solver *s) { return s ? s->core.err.c_str() : g_create_error.c_str(); }

int fs_slab_range(const fs_solver *s, int32_t *z_begin, int32_t *z_end, int64_t *owned_voxels) {
    if (!s) return FS_ERR_BAD_ARGUMENT;
    if (z_begin) *z_begin = s->core.zb;
    if (z_end) *z_end = s->core.ze;
    if (owned_voxels) *owned_voxels = s->core.nowned;
    return FS_OK;
}

int fs_set_obstacles(fs_solver *s, const uint8_t *mask, int64_t n) { FS_GUARD(s); return c.set_obstacles(mask, n); }

int fs_slab_halo_range(const fs_solver *s, int32_t *z_begin, int32_t *z_end) {
    if (!s) return FS_ERR_BAD_ARGUMENT;
    if (z_begin) *z_begin = s->core.g.zoff;
    if (z_end) *z_end = s->core.g.zoff + s->core.g.nzl;
    return FS_OK;
}

int fs_set_obstacles_slab(fs_solver *s, const uint8_t *mask, int64_t n, int32_t global_any, int32_t global_interior) {
    FS_GUARD(s);
    return c.set_obstacles_slab(mask, n, global_any != 0, global_interior != 0);
}

int fs_build_obstacles(fs_solver *s, const fs_obstacle_shape *shape, int64_t *obstacle_cells) { FS_GUARD(s); return c.build_obstacles(shape, obstacle_cells); }

int fs_get_obstacles(fs_solver *s, uint8_t *out, int64_t n) { FS_GUARD(s); return c.get_obstacles(out, n); }

int fs_add_density(fs_solver *s, float x, float y, float z, float amount) {
    FS_GUARD(s);
    return c.add_cells(1, &x, &y, &z, &amount, nullptr, nullptr, nullptr);
}

int fs_add_velocity(fs_solver *s, float x, float y, float z, float ax, float ay, float az) {
    FS_GUARD(s);
    return c.add_cells(1, &x, &y, &z, nullptr, &ax, &ay, &az);
}

int fs_add_source_cells(fs_solver *s, int64_t count, const float *x, const float *y, const float *z,
                        const float *density, const float *ax, const float *ay, const float *az) {
    FS_GUARD(s);
    return c.add_cells(count, x, y, z, density, ax, ay, az);
}

int fs_add_sources(fs_solver *s, const float *density, const float *vx, const float *vy, const float *vz) {
    FS_GUARD(s);
    return c.add_dense(density, vx, vy, vz);
}

int fs_step(fs_solver *s, float dt, float visc, float diff) { FS_GUARD(s); return c.step(dt, visc, diff); }

int fs_sync(fs_solver *s) { FS_GUARD(s); c.ex.sync(); return c.check(); }

int fs_get_field(fs_solver *s, int32_t field, float *out, int64_t n) { FS_GUARD(s); return c.get_field(field, out, n); }

int fs_set_field(fs_solver *s, int32_t field, const float *in, int64_t n) { FS_GUARD(s); return c.set_field(field, in, n); }

int fs_get_field_async(fs_solver *s, int32_t field, float *out, int64_t n) { FS_GUARD(s); return c.get_field_async(field, out, n); }

int fs_wait_transfers(fs_solver *s) { FS_GUARD(s); c.ex.wait_transfers(); return c.check(); }

int fs_render_rgba(fs_solver *s, const fs_vis_params *vp, float *out_rgba, int64_t n) { FS_GUARD(s); return c.render(vp, out_rgba, n); }

int fs_streamlines(fs_solver *s, int32_t skip, float scale, int32_t z_slice, float *out_segments, int64_t count) {
    FS_GUARD(s);
    return c.streamlines(skip, scale, z_slice, out_segments, count);
}

int fs_get_metrics(fs_solver *s, float *mean_density, float *max_speed, double *sum_density) {
    FS_GUARD(s);
    double sum = 0.0;
    float mx = 0.0f;
    c.ex.metrics(c.g, c.density, c.vx, c.vy, c.vz, &sum, &mx);
    if (sum_density) *sum_density = sum;
    if (mean_density) *mean_density = (float)(sum / (double)c.nowned);
    if (max_speed) *max_speed = mx;
    return c.check();
}

// ---- operator entry points ---------------------------------------------------------------------------
static bool fs_valid_b(const SolverCore<FS_EXEC> &c, int b) { return b >= 0 && b <= (c.g.hz ? 3 : 2); }

int fs_op_set_bnd(fs_solver *s, int32_t field, int32_t b) {
    FS_GUARD(s);
    float *p = c.field_ptr(field);
    if (!p || !fs_valid_b(c, b)) return c.fail(FS_ERR_BAD_ARGUMENT, "bad field or b");
    c.ex.bnd(c.g, p, b);
    c.mirror(p, b);
    c.ex.halo(c.g, p);
    return c.check();
}

int fs_op_smooth(fs_solver *s, int32_t dst, int32_t src, int32_t b, float a, float cc, int32_t iters) {
    FS_GUARD(s);
    float **d = c.field_slot(dst);
    float *x0 = c.field_ptr(src);
    if (!d || !*d || !x0 || dst == src || !fs_valid_b(c, b) || iters < 0) return c.fail(FS_ERR_BAD_ARGUMENT, "bad argument");
    c.smooth(b, *d, x0, a, cc, iters);
    return c.check();
}

int fs_op_lin_solve(fs_solver *s, int32_t dst, int32_t rhs, int32_t b, float a, float cc, int32_t iters,
                    int32_t solver_kind) {
    FS_GUARD(s);
    float **d = c.field_slot(dst);
    float *r = c.field_ptr(rhs);
    if (!d || !*d || !r || dst == rhs || !fs_valid_b(c, b) || iters < 0) return c.fail(FS_ERR_BAD_ARGUMENT, "bad argument");
    c.ex.halo(c.g, r); // slabs: fused / extended sweeps read the right-hand side one plane into the ghost zone
    if (solver_kind == FS_RED_BLACK)
        c.lin_solve_rb(b, *d, r, a, cc, iters, false);
    else
        c.lin_solve(b, *d, r, a, cc, iters, false);
    return c.check();
}

int fs_op_diffuse(fs_solver *s, int32_t dst, int32_t src, int32_t b, float diff, float dt) {
    FS_GUARD(s);
    float **d = c.field_slot(dst);
    float *x0 = c.field_ptr(src);
    if (!d || !*d || !x0 || dst == src || !fs_valid_b(c, b)) return c.fail(FS_ERR_BAD_ARGUMENT, "bad argument");
    c.diffuse(b, *d, x0, diff, dt);
    return c.check();
}

int fs_op_project(fs_solver *s, int32_t use_v0_fields) {
    FS_GUARD(s);
    if (use_v0_fields) c.project(c.vx0, c.vy0, c.vz0);
    else c.project(c.vx, c.vy, c.vz);
    return c.check();
}

int fs_op_advect(fs_solver *s, int32_t dst, int32_t src, int32_t b, int32_t use_v0_fields, float dt) {
    FS_GUARD(s);
    float *d = c.field_ptr(dst), *d0 = c.field_ptr(src);
    if (!d || !d0 || dst == src || !fs_valid_b(c, b)) return c.fail(FS_ERR_BAD_ARGUMENT, "bad argument");
    float *ux = use_v0_fields ? c.vx0 : c.vx, *uy = use_v0_fields ? c.vy0 : c.vy, *uz = use_v0_fields ? c.vz0 : c.vz;
    if (d == ux || d == uy || d == uz) return c.fail(FS_ERR_BAD_ARGUMENT, "dst aliases the carrier velocity");
    const float dt0 = dt * (float)(c.g.nx - 2);
    c.ex.halo_fence();
    c.ex.advect(c.g, d, d0, ux, uy, uz, c.fl(), dt0, b);
    if (b == 3 && c.g_interior_obstacle) c.ex.halo(c.g, d);
    c.mirror(d, b);
    c.ex.halo(c.g, d);
    return c.check();
}

int fs_op_advect_velocity(fs_solver *s, float dt) {
    FS_GUARD(s);
    const float dt0 = dt * (float)(c.g.nx - 2);
    c.ex.halo_fence();
    c.ex.advect_velocity(c.g, c.vx, c.vy, c.vz, c.vx0, c.vy0, c.vz0, c.fl(), dt0);
    c.finish_velocity(c.vx, c.vy, c.vz);
    return c.check();
}

int fs_op_enforce_obstacles(fs_solver *s) {
    FS_GUARD(s);
    if (c.g_any_obstacle) {
        c.ex.enforce(c.g, c.vx, c.vy, c.vz, c.flags, c.prm.cell_size, c.prm.raw_viscosity);
        float *fields[3] = {c.vx, c.vy, c.vz};
        c.ex.halo_n(c.g, fields, c.g.hz ? 3 : 2);
    }
    return c.check();
}

// ---- measurement ---------------------------------------------------------------------------------------
int fs_timer_start(fs_solver *s) { FS_GUARD(s); c.ex.timer_start(); return c.check(); }

int fs_timer_stop(fs_solver *s, float *elapsed_ms) {
    FS_GUARD(s);
    const float ms = c.ex.timer_stop();
    if (elapsed_ms) *elapsed_ms = ms;
    return c.check();
}

int64_t fs_launch_count(const fs_solver *s) { return s ? s->core.ex.launches : 0; }

int fs_bench_sweep(fs_solver *s, int32_t kind_and_fill, int32_t b, int32_t reps, float *avg_ms, double *algo_bytes) {
    FS_GUARD(s);
    const int kind = kind_and_fill & 15, fill = kind_and_fill >> 4; // fill: 0 as is, 1 random normals, 2 zeros
    if (reps < 1 || kind < 0 || kind > FS_BENCH_KIND_MAX || fill < 0 || fill > 2 || !fs_valid_b(c, b)) return c.fail(FS_ERR_BAD_ARGUMENT, "bad argument");
    // scratch operands: in = vx0, rhs = vy0, out = tmp (by default whatever the last step left there)
    if (fill == 1) { c.ex.fill_random(c.vx0, c.nloc, 1u); c.ex.fill_random(c.vy0, c.nloc, 2u); }
    if (fill == 2) { c.ex.zero(c.vx0, sizeof(float) * c.nloc); c.ex.zero(c.vy0, sizeof(float) * c.nloc); }
    const long long interior = (long long)(c.g.nx) * c.g.ny * (c.ze - c.zb);
    const double fl1 = c.fl() ? 1.0 : 0.0;
    // SURVEY.md section 8(d), algorithmic bytes per voxel per launch (3D): smoother 4 R + 4 W, Jacobi 8 R + 4 W,
    // red-black full sweep = one Jacobi sweep's worth, divergence 12 R + 4 W + 4 W (the p = 0 store the reference
    // makes; this build drops it), gradient 4 R + 24 RW, advect of one scalar 12 R + 4 R + 4 W, fused advect of the
    // three velocity components 12 R + 12 W (+ the flag byte wherever the grid has obstacles).
    // Fused pairs: TWO sweeps' worth of algorithmic bytes per launch.
    double per_voxel = 0.0;
    switch (kind) {
    case 0: per_voxel = 8.0 + fl1; break;
    case 1: case 2: per_voxel = 12.0 + fl1; break;
    case 3: per_voxel = (c.g.hz ? 20.0 : 16.0) + fl1; break;
    case 4: per_voxel = (c.g.hz ? 24.0 : 16.0) + fl1; break;
    case 5: per_voxel = c.g.hz ? 20.0 : 16.0; break;
    case 6: per_voxel = (c.g.hz ? 28.0 : 20.0) + fl1; break;
    case 7: per_voxel = 2.0 * (12.0 + fl1); break;
    case 8: per_voxel = 2.0 * (8.0 + fl1); break;
    case 9: per_voxel = 12.0 + fl1; break;
    }
    float a, cc;
    SolverCore<FS_EXEC>::coeffs(c.g.nx, 1e-4f, 0.1f, &a, &cc);
    const float dt0 = 0.1f * 128.0f / (float)c.g.nx * (float)(c.g.nx - 2);
    bool ok = true;
    auto once = [&]() {
        switch (kind) {
        case 0: c.ex.relax(FS_MODE_SMOOTH, c.g, c.vx0, c.vy0, c.vx0, c.tmp, c.fl(), a, cc, b, false, true); break;
        case 1: c.ex.relax(FS_MODE_JACOBI, c.g, c.vx0, c.vy0, nullptr, c.tmp, c.fl(), a, cc, b, false, true); break;
        case 2: c.ex.rb_half(c.g, c.tmp, c.vy0, c.fl(), a, cc, 0, b); c.ex.rb_half(c.g, c.tmp, c.vy0, c.fl(), a, cc, 1, b); break;
        // once-per-step kernels on the live fields (no halo ops: the callers run these on every rank or on none)
        case 3: c.ex.advect(c.g, c.tmp, c.density, c.vx, c.vy, c.vz, c.fl(), dt0, 0); break;
        case 4: c.ex.advect_velocity(c.g, c.vx0, c.vy0, c.vz0, c.vx, c.vy, c.vz, c.fl(), dt0); break;
        case 5: c.ex.divergence(c.g, c.div, c.vx, c.vy, c.vz); break;
        case 6: c.ex.gradient(c.g, c.vx0, c.vy0, c.vz0, c.pressure, c.fl()); break;
        case 7: ok = c.ex.relax_pair(FS_PAIR_JACOBI, c.g, c.vx0, c.vy0, c.tmp, c.fl(), a, cc, b, false, true); break;
        case 8: ok = c.ex.relax_pair(FS_PAIR_SMOOTH, c.g, c.vx0, nullptr, c.tmp, c.fl(), a, cc, b, false, true); break;
        case 9: ok = c.ex.relax_pair(FS_PAIR_RED_BLACK, c.g, c.vx0, c.vy0, c.tmp, c.fl(), a, cc, b, false, true); break;
        }
    };
    c.ex.pair_force = true;
    for (int w = 0; w < 2; w++) once(); // warm-up
    if (!ok) { c.ex.pair_force = false; return c.fail(FS_ERR_UNSUPPORTED, "the fused pair kernel does not support this grid"); }
    c.ex.timer_start();
    for (int r = 0; r < reps; r++) once();
    const float ms = c.ex.timer_stop();
    c.ex.pair_force = false;
    if (avg_ms) *avg_ms = ms / (float)reps;
    if (algo_bytes) *algo_bytes = per_voxel * (double)interior;
    return c.check();
}

int fs_selftest_division(fs_solver *s, float divisor, uint64_t first_bits, uint64_t count, uint64_t *mismatches) {
    FS_GUARD(s);
    if (!mismatches) return c.fail(FS_ERR_BAD_ARGUMENT, "null argument");
    *mismatches = c.ex.division_selftest(divisor, first_bits, count);
    return c.check();
}

// ---- multi-GPU wiring ------------------------------------------------------------------------------------
int fs_halo_export(fs_solver *s, void *blob, int64_t blob_bytes) {
    FS_GUARD(s);
    if (!blob || blob_bytes < FS_IPC_BLOB_BYTES) return c.fail(FS_ERR_BAD_ARGUMENT, "blob too small");
    memset(blob, 0, (size_t)blob_bytes);
    const int rc = c.ex.halo_export(c, blob);
    return rc ? c.fail(rc, c.ex.error()) : FS_OK;
}

int fs_halo_connect(fs_solver *s, const void *lower_blob, const void *upper_blob, int32_t same_process) {
    FS_GUARD(s);
    const int rc = c.ex.halo_connect(c, lower_blob, upper_blob, same_process);
    return rc ? c.fail(rc, c.ex.error()) : FS_OK;
}

} // extern "C"
