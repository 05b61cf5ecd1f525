// host_emul.cpp -- TEST SCAFFOLDING ONLY (never shipped, never loaded by the product package).
//
// Builds tests/host_emul/libfluidsolver_hostemul.so: the SAME orchestration (csrc/fs_core.h), the SAME
// per-cell functions (csrc/fs_cellops.cuh) and the SAME C ABI (csrc/fs_abi.inl) as libfluidsolver.so,
// but with a plain-loop executor instead of CUDA kernels.  Purpose: the CPU-only test tier can check
// the ring-scatter / obstacle / buffer-rotation logic and the ABI argument handling against the
// oracle without a GPU.  It does NOT cover relax_vec4 or any launch geometry -- those are covered by
// the -m gpu tests, which call the real library.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../3dfluidsimulation_b200/csrc/fs_core.h"

struct HostExec {
    bool use_graph = false;
    int64_t launches = 0;
    std::string msg;
    std::chrono::steady_clock::time_point t0;

    const std::string &error() const { return msg; }
    bool failed() { return false; }
    void make_current() {}
    int open(int) { return 0; }
    void close() {}
    void *alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
    void free(void *p) { ::free(p); }
    void zero(void *p, size_t bytes) { memset(p, 0, bytes); }
    void copy(void *d, const void *s, size_t bytes) { memcpy(d, s, bytes); }
    void upload(void *d, const void *s, size_t bytes) { memcpy(d, s, bytes); }
    void download(void *d, const void *s, size_t bytes) { memcpy(d, s, bytes); }
    void sync() {}

    template <class F>
    void cells(const FsGrid &g, F f) {
        int k0 = 0, k1 = 1;
        if (g.hz) {
            const int zb = g.zoff + g.kb, ze = g.zoff + g.ke;
            k0 = (zb < 1 ? 1 : zb) - g.zoff;
            k1 = (ze > g.nz - 1 ? g.nz - 1 : ze) - g.zoff;
        }
        // reversed traversal on purpose: results must not depend on the order cells are visited in
        for (int kl = k1 - 1; kl >= k0; kl--)
            for (int j = g.ny - 2; j >= 1; j--)
                for (int i = g.nx - 2; i >= 1; i--) f(i, j, kl);
        launches++;
    }

    void relax(int mode, const FsGrid &g, const float *in, const float *rhs, const float *stale, float *out,
               const uint8_t *flags, float a, float c, int b, bool in_zero) {
        if (mode == FS_MODE_SMOOTH)
            cells(g, [&](int i, int j, int kl) { fs_relax_cell<FS_MODE_SMOOTH>(g, in, rhs, stale, out, flags, a, c, b, in_zero, i, j, kl); });
        else
            cells(g, [&](int i, int j, int kl) { fs_relax_cell<FS_MODE_JACOBI>(g, in, rhs, stale, out, flags, a, c, b, in_zero, i, j, kl); });
    }
    void rb_half(const FsGrid &g, float *x, const float *rhs, const uint8_t *flags, float a, float c, int colour) {
        cells(g, [&](int i, int j, int kl) {
            if (((i + j + kl + g.zoff) & 1) == colour) fs_rb_cell(g, x, rhs, flags, a, c, i, j, kl);
        });
    }
    void bnd(const FsGrid &g, float *x, int b) {
        cells(g, [&](int i, int j, int kl) { fs_bnd_cell(g, x, b, i, j, kl); });
    }
    void mirror(const FsGrid &g, float *x, const uint8_t *flags, const long long *list, long long n, int b) {
        for (long long t = n - 1; t >= 0; t--) fs_mirror_cell(g, x, flags, b, list[t]);
        launches++;
    }
    void divergence(const FsGrid &g, float *div, const float *ux, const float *uy, const float *uz) {
        cells(g, [&](int i, int j, int kl) { fs_divergence_cell(g, div, ux, uy, uz, i, j, kl); });
    }
    void gradient(const FsGrid &g, float *ux, float *uy, float *uz, const float *p, const uint8_t *flags) {
        cells(g, [&](int i, int j, int kl) { fs_gradient_cell(g, ux, uy, uz, p, flags, i, j, kl); });
    }
    void advect(const FsGrid &g, float *d, const float *d0, const float *ux, const float *uy, const float *uz,
                const uint8_t *flags, float dt0, int b) {
        cells(g, [&](int i, int j, int kl) {
            auto samp = [&](int ii, int jj, int kk) { return d0[fs_idx(g, ii, jj, kk - g.zoff)]; };
            fs_advect_cell(g, d, samp, ux, uy, uz, flags, dt0, b, i, j, kl);
        });
    }
    void advect_velocity(const FsGrid &g, float *dx, float *dy, float *dz, const float *sx, const float *sy,
                         const float *sz, const uint8_t *flags, float dt0) {
        cells(g, [&](int i, int j, int kl) {
            auto px = [&](int ii, int jj, int kk) { return sx[fs_idx(g, ii, jj, kk - g.zoff)]; };
            auto py = [&](int ii, int jj, int kk) { return sy[fs_idx(g, ii, jj, kk - g.zoff)]; };
            auto pz = [&](int ii, int jj, int kk) { return sz[fs_idx(g, ii, jj, kk - g.zoff)]; };
            fs_advect_velocity_cell(g, dx, dy, dz, px, py, pz, sx, sy, sz, flags, dt0, i, j, kl);
        });
    }
    void enforce(const FsGrid &g, float *ux, float *uy, float *uz, const uint8_t *flags, float cell, float rawvisc) {
        cells(g, [&](int i, int j, int kl) { fs_enforce_cell(g, ux, uy, uz, flags, cell, rawvisc, i, j, kl); });
    }
    void build_flags(const FsGrid &g, const uint8_t *mask, uint8_t *flags) {
        for (int kl = 0; kl < g.nzl; kl++)
            for (int j = 0; j < g.ny; j++)
                for (int i = 0; i < g.nx; i++) flags[fs_idx(g, i, j, kl)] = fs_flags_cell(g, mask, i, j, kl);
    }
    void fill_random(float *dst, long long n, unsigned seed) { for (long long t = 0; t < n; t++) dst[t] = (float)((t * 2654435761u + seed) % 2001) / 1000.0f - 1.0f; }
    void axpy(float *dst, const float *src, long long n) {
        for (long long t = 0; t < n; t++) dst[t] += src[t];
    }
    void scatter_add(float *dst[4], const long long *idx, const float *src[4], long long n) {
        for (long long t = 0; t < n; t++)
            for (int f = 0; f < 4; f++)
                if (dst[f]) dst[f][idx[t]] += src[f][t];
    }
    void metrics(const FsGrid &g, const float *d, const float *ux, const float *uy, const float *uz, double *sum, float *mx) {
        double acc = 0;
        float m = 0;
        for (long long t = g.sz * g.kb; t < g.sz * g.ke; t++) {
            acc += d[t];
            float q = ux[t] * ux[t] + uy[t] * uy[t];
            if (g.hz) q = q + uz[t] * uz[t];
            m = fmaxf(m, sqrtf(q));
        }
        *sum = acc;
        *mx = m;
    }
    unsigned long long division_selftest(float, unsigned long long, unsigned long long) { return 0; } // plain `/` here
    void halo(const FsGrid &, float *) {}
    template <class Core> int halo_export(Core &, void *) { msg = "host emulation: no multi-GPU"; return FS_ERR_UNSUPPORTED; }
    template <class Core> int halo_connect(Core &, const void *, const void *, int) { msg = "host emulation: no multi-GPU"; return FS_ERR_UNSUPPORTED; }
    void invalidate_graph() {}
    bool replay_step(float, float, float, float *const[11]) { return false; }
    void roles_after_replay(float **[11]) {}
    void begin_step(float, float, float, float *const[11]) {}
    void end_step(float *const[11]) {}
    void timer_start() { t0 = std::chrono::steady_clock::now(); }
    float timer_stop() { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

#define FS_EXEC HostExec
#include "../../3dfluidsimulation_b200/csrc/fs_abi.inl"
