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

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <string>
#include <vector>

#include "../../3dfluidsimulation_b200/csrc/fs_core.h"

struct HostExec {
    bool use_graph = false;
    bool pair_force = false;
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
    void download_async(void *d, const void *s, size_t bytes, int) { memcpy(d, s, bytes); }
    void wait_transfers() {}
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

    // xmode: FS_X_* (fs_cellops.cuh).  An extended sweep also computes the first ghost plane on each side that has a
    // neighbour, from local data only, and acknowledges the ghost planes it read (see halo()).
    bool can_extend(const FsGrid &g) const {
        const char *e = getenv("FS_EXTEND"); // opt-in, as in the CUDA executor
        return halo_on && g.hz && e && e[0] == '1';
    }
    void relax(int mode, const FsGrid &g, const float *in, const float *rhs, const float *stale, float *out,
               const uint8_t *flags, float a, float c, int b, bool in_zero, int xmode) {
        const int zb = g.zoff + g.kb, ze = g.zoff + g.ke;
        int k0 = g.hz ? (zb < 1 ? 1 : zb) - g.zoff : 0, k1 = g.hz ? (ze > g.nz - 1 ? g.nz - 1 : ze) - g.zoff : 1;
        const bool extend = halo_on && xmode == FS_X_EXTEND;
        if (extend) {
            if (lo.present) k0 -= 1;
            if (hi.present) k1 += 1;
        }
        if (mode == FS_MODE_SMOOTH)
            cells_range(g, k0, k1, [&](int i, int j, int kl) { fs_relax_cell<FS_MODE_SMOOTH>(g, in, rhs, stale, out, flags, a, c, b, in_zero, i, j, kl); });
        else
            cells_range(g, k0, k1, [&](int i, int j, int kl) { fs_relax_cell<FS_MODE_JACOBI>(g, in, rhs, stale, out, flags, a, c, b, in_zero, i, j, kl); });
        launches++;
        if (xmode == FS_X_EXCHANGE || xmode == FS_X_EXCHANGE_OPEN) halo(g, out);
        if (extend) ack();
    }
    void relax_n(int mode, const FsGrid &g, int nf, const float *const *in, const float *const *rhs, const float *const *stale,
                 float *const *out, const uint8_t *flags, float a, float c, const int *b, bool in_zero, int xmode) {
        const bool exch = xmode == FS_X_EXCHANGE || xmode == FS_X_EXCHANGE_OPEN;
        for (int f = 0; f < nf; f++)
            relax(mode, g, in[f], rhs ? rhs[f] : nullptr, stale ? stale[f] : nullptr, out[f], flags, a, c, b[f], in_zero,
                  exch ? FS_X_NONE : xmode);
        if (exch) halo_n(g, out, nf);
    }
    void halo_n(const FsGrid &g, float *const *fields, int nf) {
        for (int f = 0; f < nf; f++) halo(g, fields[f]);
    }
    // Fused two-stage sweep, emulated with the per-cell functions and LOCAL data only (so that the CPU tier checks the
    // two-plane halo logic): stage 1 on the owned interior planes plus one plane each side into a scratch copy of
    // `in`, stage 2 on the owned interior planes.  FS_EMUL_NO_PAIR=1 makes the orchestration fall back to single sweeps.
    bool pair_supported(const FsGrid &g, float, int /*kind*/) const {
        const char *e = getenv("FS_EMUL_NO_PAIR");
        return g.hz && !(e && e[0] == '1');
    }
    template <class F>
    void cells_range(const FsGrid &g, int klo, int khi, F f) { // interior rows / columns of local planes [klo, khi)
        for (int kl = khi - 1; kl >= klo; kl--)
            for (int j = g.ny - 2; j >= 1; j--)
                for (int i = g.nx - 2; i >= 1; i--) f(i, j, kl);
    }
    bool relax_pair(int kind, const FsGrid &g, const float *in, const float *rhs, float *out, const uint8_t *flags,
                    float a, float c, int b, bool in_zero, bool fuse_halo) {
        if (!pair_supported(g, c, kind)) return false;
        const int zb = g.zoff + g.kb, ze = g.zoff + g.ke;
        const int k0 = (zb < 1 ? 1 : zb) - g.zoff, k1 = (ze > g.nz - 1 ? g.nz - 1 : ze) - g.zoff; // owned interior [k0, k1)
        if (k1 <= k0) return true;
        const long long n = g.sz * g.nzl;
        std::vector<float> y(n);
        if (in_zero) std::fill(y.begin(), y.end(), 0.0f); else memcpy(y.data(), in, sizeof(float) * n);
        const float *src = in_zero ? nullptr : in;
        std::vector<float> zeros;
        if (in_zero) { zeros.assign(n, 0.0f); src = zeros.data(); }
        // stage 1 range: one plane beyond the owned interior planes where that plane is itself interior
        const int s0 = (k0 - 1 + g.zoff >= 1) ? k0 - 1 : k0, s1 = (k1 + g.zoff <= g.nz - 2) ? k1 + 1 : k1;
        if (kind == FS_PAIR_RED_BLACK) {
            cells_range(g, s0, s1, [&](int i, int j, int kl) {
                if (((i + j + kl + g.zoff) & 1) == 0) fs_rb_cell(g, y.data(), rhs, flags, a, c, i, j, kl);
            });
            // owned planes only: a faster neighbour may already be storing its planes of this op into out's ghosts.
            // A colour-1 cell has colour-0 neighbours only, which stage 2 does not change: read them from y.
            memcpy(out + g.sz * g.kb, y.data() + g.sz * g.kb, sizeof(float) * g.sz * (g.ke - g.kb));
            cells_range(g, k0, k1, [&](int i, int j, int kl) {
                const long long idx = fs_idx(g, i, j, kl);
                if (((i + j + kl + g.zoff) & 1) != 1 || (flags && (flags[idx] & FS_OB_SELF))) return;
                const float *x = y.data();
                float s = ((x[idx + 1] + x[idx - 1]) + x[idx + g.sy]) + x[idx - g.sy];
                s = (s + x[idx + g.sz]) + x[idx - g.sz];
                out[idx] = (rhs[idx] + a * s) / c;
            });
            cells_range(g, k0, k1, [&](int i, int j, int kl) { fs_bnd_cell(g, out, b, i, j, kl); });
        } else if (kind == FS_PAIR_JACOBI) {
            cells_range(g, s0, s1, [&](int i, int j, int kl) { fs_relax_cell<FS_MODE_JACOBI>(g, src, rhs, nullptr, y.data(), flags, a, c, b, false, i, j, kl); });
            cells_range(g, k0, k1, [&](int i, int j, int kl) { fs_relax_cell<FS_MODE_JACOBI>(g, y.data(), rhs, nullptr, out, flags, a, c, b, false, i, j, kl); });
        } else {
            cells_range(g, s0, s1, [&](int i, int j, int kl) { fs_relax_cell<FS_MODE_SMOOTH>(g, src, nullptr, src, y.data(), flags, a, c, b, false, i, j, kl); });
            std::vector<float> y2(y); // obstacle cells copy the stage's input
            cells_range(g, k0, k1, [&](int i, int j, int kl) { fs_relax_cell<FS_MODE_SMOOTH>(g, y.data(), nullptr, y2.data(), out, flags, a, c, b, false, i, j, kl); });
        }
        launches++;
        if (fuse_halo) halo(g, out);
        return true;
    }
    bool rb_half(const FsGrid &g, float *x, const float *rhs, const uint8_t *flags, float a, float c, int colour, int /*b*/) {
        cells(g, [&](int i, int j, int kl) {
            if (((i + j + kl + g.zoff) & 1) == colour) fs_rb_cell(g, x, rhs, flags, a, c, i, j, kl);
        });
        return false;
    }
    void bnd(const FsGrid &g, float *x, int b) {
        cells(g, [&](int i, int j, int kl) { fs_bnd_cell(g, x, b, i, j, kl); });
    }
    void mirror(const FsGrid &g, float *x, const uint8_t *flags, const long long *list, long long n, int b) {
        for (long long t = n - 1; t >= 0; t--) fs_mirror_cell(g, x, flags, b, list[t]);
        launches++;
    }
    void mirror3(const FsGrid &g, float *ux, float *uy, float *uz, const uint8_t *flags, const long long *list, long long n) {
        mirror(g, ux, flags, list, n, 1);
        mirror(g, uy, flags, list, n, 2);
        if (uz) mirror(g, uz, flags, list, n, 3);
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
            auto samp = [&](int kk) { int l; const float *b = resolve(g, d0, kk, &l); return b + l * g.sz; };
            fs_advect_cell(g, d, samp, ux, uy, uz, flags, dt0, b, i, j, kl);
        });
    }
    void advect_velocity(const FsGrid &g, float *dx, float *dy, float *dz, const float *sx, const float *sy,
                         const float *sz, const uint8_t *flags, float dt0) {
        cells(g, [&](int i, int j, int kl) {
            auto px = [&](int kk) { int l; const float *b = resolve(g, sx, kk, &l); return b + l * g.sz; };
            auto py = [&](int kk) { int l; const float *b = resolve(g, sy, kk, &l); return b + l * g.sz; };
            auto pz = [&](int kk) { int l; const float *b = resolve(g, sz, kk, &l); return b + l * g.sz; };
            fs_advect_velocity_cell(g, dx, dy, dz, px, py, pz, sx, sy, sz, flags, dt0, i, j, kl);
        });
    }
    void enforce(const FsGrid &g, float *ux, float *uy, float *uz, const uint8_t *flags, float cell, float rawvisc) {
        cells(g, [&](int i, int j, int kl) { fs_enforce_cell(g, ux, uy, uz, flags, cell, rawvisc, i, j, kl); });
    }
    std::vector<float> render_buf;
    void *render_buffer(size_t bytes) { render_buf.resize(bytes / sizeof(float)); return render_buf.data(); }
    void visualize(const FsGrid &g, const fs_vis_params &vp, const float *d, const float *p, const uint8_t *mask, float *rgba) {
        for (long long t = 0; t < g.sz; t++) {
            const FsColor c = fs_visualize_cell(vp, d[t], p[t], mask[t] != 0, (int)(t % g.nx), (int)(t / g.nx));
            rgba[4 * t] = c.r; rgba[4 * t + 1] = c.g; rgba[4 * t + 2] = c.b; rgba[4 * t + 3] = c.a;
        }
    }
    void streamlines(const FsGrid &g, int skip, float scale, const float *ux, const float *uy, const uint8_t *mask, float *out,
                     long long count) {
        for (long long t = 0; t < count; t++) fs_streamline_glyph(g.nx, g.ny, skip, scale, ux, uy, mask, (int)t, out + 4 * t);
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
    std::vector<long long> stage_idx;
    std::vector<float> stage_amt[4];
    bool source_stage(long long cap, long long **idx, float *amt[4]) {
        stage_idx.resize((size_t)cap);
        *idx = stage_idx.data();
        for (int f = 0; f < 4; f++) { stage_amt[f].resize((size_t)cap); amt[f] = stage_amt[f].data(); }
        return true;
    }
    void scatter_add_staged(float *dst[4], long long, long long n) {
        for (long long t = 0; t < n; t++)
            for (int f = 0; f < 4; f++)
                if (dst[f]) dst[f][stage_idx[t]] += stage_amt[f][t];
    }
    bool scan_obstacles(const FsGrid &g, const uint8_t *mask, int kl0, int kl1, bool *any_local, long long **list, long long *count) {
        std::vector<long long> v;
        bool any = false;
        for (long long t = g.sz * g.nzl - 1; t >= 0; t--) { // reversed on purpose: the list order must not matter
            if (!mask[t]) continue;
            any = true;
            const int i = (int)(t % g.nx), j = (int)((t / g.nx) % g.ny), kl = (int)(t / g.sz);
            if (i >= 1 && i <= g.nx - 2 && j >= 1 && j <= g.ny - 2 && kl >= kl0 && kl < kl1) v.push_back(t);
        }
        *any_local = any;
        *count = (long long)v.size();
        *list = nullptr;
        if (!v.empty()) {
            *list = (long long *)malloc(sizeof(long long) * v.size());
            memcpy(*list, v.data(), sizeof(long long) * v.size());
        }
        return true;
    }
    bool build_shape(const FsGrid &g, const fs_obstacle_shape &sh, uint8_t *mask, long long *total, long long *interior) {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        const bool hz = g.hz != 0;
        std::vector<uint8_t> inside((size_t)g.sz, 0), reach((size_t)g.sz, 0);
        const bool need_fill = !(sh.kind == 0 && hz);
        const bool seed_in_grid = sh.seed_x >= 0 && sh.seed_x < nx && sh.seed_y >= 0 && sh.seed_y < ny && (!hz || (sh.seed_z >= 0 && sh.seed_z < nz));
        bool seed_ok;
        if (need_fill) {
            for (int y = 0; y < ny; y++)
                for (int x = 0; x < nx; x++)
                    inside[x + (size_t)y * nx] = sh.kind == 0 ? fs_shape_inside_circle(sh, false, x, y, 0) : fs_shape_inside_xy(sh, x, y);
            if (seed_in_grid) { // RecursiveFloodFill :329-351 with an explicit stack
                std::vector<std::pair<int, int>> todo{{sh.seed_x, sh.seed_y}};
                while (!todo.empty()) {
                    const auto [x, y] = todo.back();
                    todo.pop_back();
                    if (x < 0 || x >= nx || y < 0 || y >= ny || reach[x + (size_t)y * nx] || !inside[x + (size_t)y * nx]) continue;
                    reach[x + (size_t)y * nx] = 1;
                    todo.emplace_back(x + 1, y); todo.emplace_back(x - 1, y); todo.emplace_back(x, y + 1); todo.emplace_back(x, y - 1);
                }
            }
            seed_ok = seed_in_grid && fs_shape_in_span(sh, hz && sh.kind != 0, sh.seed_z);
        } else {
            seed_ok = seed_in_grid && fs_shape_inside_circle(sh, true, sh.seed_x, sh.seed_y, sh.seed_z);
        }
        for (int kl = 0; kl < g.nzl; kl++)
            for (int y = 0; y < ny; y++)
                for (int x = 0; x < nx; x++)
                    mask[fs_idx(g, x, y, kl)] = seed_ok ? fs_shape_mask(sh, hz, nx, reach.data(), seed_ok, x, y, kl + g.zoff) : 0;
        long long tot = 0, inter = 0;
        for (int z = 0; z < nz; z++)
            for (int y = 0; y < ny; y++)
                for (int x = 0; x < nx; x++) {
                    if (!seed_ok || !fs_shape_mask(sh, hz, nx, reach.data(), seed_ok, x, y, z)) continue;
                    tot++;
                    if (x >= 1 && x <= nx - 2 && y >= 1 && y <= ny - 2 && (!hz || (z >= 1 && z <= nz - 2))) inter++;
                }
        *total = tot;
        *interior = inter;
        return true;
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
    // ---- multi-slab emulation: one host thread per slab handle, same lock-step protocol as the CUDA
    // executor (wait for the neighbours' op seq-1, store boundary planes into their ghosts, publish seq,
    // wait for the neighbours' seq) with std::atomic instead of device flags.
    struct HostBlob {
        uint32_t magic;
        int32_t rank, nzl, kb, ke, zoff, nbuf;
        void *raw_field[11];
        std::atomic<unsigned> *seq, *ack;
    };
    struct Peer {
        bool present = false;
        float *base[11] = {};
        std::atomic<unsigned> *seq = nullptr, *ack = nullptr;
        int nzl = 0, zoff = 0;
    };
    bool halo_on = false;
    Peer lo, hi;
    std::atomic<unsigned> *my_seq = nullptr, *my_ack = nullptr;
    bool ack_since_last_op = false;
    unsigned ops = 0;
    std::vector<float *> bufs;
    int buf_index(const float *p) const {
        for (size_t i = 0; i < bufs.size(); i++)
            if (bufs[i] == p) return (int)i;
        return -1;
    }
    static void spin(std::atomic<unsigned> *a, unsigned target) {
        while ((int)(a->load(std::memory_order_acquire) - target) < 0) std::this_thread::yield();
    }
    // after an extended sweep: the ghost planes the neighbours' last operation filled have been read
    void ack() {
        if (!halo_on) return;
        my_ack->store(ops, std::memory_order_release);
        ack_since_last_op = true;
    }
    void halo(const FsGrid &g, float *field) {
        if (!halo_on) return;
        const unsigned op = ++ops;
        if (lo.present) spin(lo.seq, op - 1);
        if (hi.present) spin(hi.seq, op - 1);
        if (ack_since_last_op) { // the neighbours ran the same extended sweep: wait until they have read their ghosts
            if (lo.present) spin(lo.ack, op - 1);
            if (hi.present) spin(hi.ack, op - 1);
            ack_since_last_op = false;
        }
        const int bi = field ? buf_index(field) : -1;
        if (bi >= 0) {
            if (lo.present) memcpy(lo.base[bi] + g.sz * (lo.nzl - FS_GHOST), field + g.sz * g.kb, sizeof(float) * g.sz * FS_GHOST);
            if (hi.present) memcpy(hi.base[bi], field + g.sz * (g.ke - FS_GHOST), sizeof(float) * g.sz * FS_GHOST);
        }
        my_seq->store(op, std::memory_order_release);
        if (lo.present) spin(lo.seq, op);
        if (hi.present) spin(hi.seq, op);
    }
    void halo_fence() { halo(FsGrid{}, nullptr); }
    void relax_end() {}
    void halo_commit() {}
    template <class Core> int halo_export(Core &c, void *blob) {
        if (!my_seq) my_seq = new std::atomic<unsigned>(0);
        if (!my_ack) my_ack = new std::atomic<unsigned>(0);
        bufs = c.allocated;
        HostBlob b{};
        b.magic = 0x48454d31u; b.rank = c.prm.slab_rank; b.nzl = c.g.nzl; b.kb = c.g.kb; b.ke = c.g.ke; b.zoff = c.g.zoff;
        b.nbuf = (int)bufs.size();
        for (size_t i = 0; i < bufs.size(); i++) b.raw_field[i] = bufs[i];
        b.seq = my_seq;
        b.ack = my_ack;
        static_assert(sizeof(HostBlob) <= FS_IPC_BLOB_BYTES, "blob too large");
        memcpy(blob, &b, sizeof(b));
        return FS_OK;
    }
    int connect_one(Peer &p, const void *blob, int expect_rank) {
        HostBlob b;
        memcpy(&b, blob, sizeof(b));
        if (b.magic != 0x48454d31u || b.rank != expect_rank) { msg = "blob mismatch"; return FS_ERR_BAD_ARGUMENT; }
        for (int i = 0; i < b.nbuf; i++) p.base[i] = (float *)b.raw_field[i];
        p.seq = b.seq; p.ack = b.ack; p.nzl = b.nzl; p.zoff = b.zoff; p.present = true;
        return FS_OK;
    }
    template <class Core> int halo_connect(Core &c, const void *lower_blob, const void *upper_blob, int same_process) {
        if (!same_process || !my_seq) { msg = "host emulation: same_process only, export first"; return FS_ERR_UNSUPPORTED; }
        int rc = FS_OK;
        if (lower_blob) rc = connect_one(lo, lower_blob, c.prm.slab_rank - 1);
        if (rc == FS_OK && upper_blob) rc = connect_one(hi, upper_blob, c.prm.slab_rank + 1);
        halo_on = rc == FS_OK;
        return rc;
    }
    const float *resolve(const FsGrid &g, const float *field, int kk, int *kl) const { // slab view for advect
        int l = kk - g.zoff;
        if (l >= 0 && l < g.nzl) { *kl = l; return field; }
        const int bi = buf_index(field);
        if (l < 0 && lo.present) { *kl = kk - lo.zoff; return lo.base[bi]; }
        if (l >= g.nzl && hi.present) { *kl = kk - hi.zoff; return hi.base[bi]; }
        *kl = 0;
        return nullptr;
    }
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

// marks this library as the CPU test scaffold (3dfluidsimulation_b200/slab.py uses it to allow several slab
// handles without distinct CUDA devices)
extern "C" int fs_host_emulation(void) { return 1; }
