// fluidsolver.cu -- libfluidsolver.so: CUDA executor (sm_100a) + the C ABI of include/fluidsolver.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false --extended-lambda
//        -shared -Xcompiler -fPIC   (see 3dfluidsimulation_b200/build.py)
// -fmad=false keeps the reference's unfused fp32 arithmetic (results are bit-identical to the oracle
// for every kernel except the obstacle drag, whose exp() differs by at most 1 ulp).
//
// There is no CPU path in this library: every operator launches a kernel on the handle's stream.
#include <cuda_runtime.h>
#include <stdio.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "fs_core.h"
#include "fs_kernels.cuh"

#define FS_CUDA(call)                                                                                                  \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) set_error(#call, e_);                                                                   \
    } while (0)

struct CudaExec {
    int dev = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st_halo = nullptr;                  // side stream: the halo push runs beside the interior launch
    cudaStream_t st_copy = nullptr;                  // copy stream: pipelined device->host readback (fs_get_field_async)
    float *stage[FS_FIELD_COUNT] = {};               // per-field device snapshots the copy stream reads from
    size_t stage_bytes[FS_FIELD_COUNT] = {};
    cudaEvent_t ev_snap[FS_FIELD_COUNT] = {}, ev_sent[FS_FIELD_COUNT] = {};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_ends = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool bad = false;
    std::string msg;
    int64_t launches = 0;
    bool force_generic = false; // FS_FORCE_GENERIC=1: scalar per-cell kernels for the sweeps (tests)
    int tune_zchunk = 0, tune_by = 0; // FS_ZCHUNK / FS_BLOCK_Y: override the sweep's z chunk length / CTA rows (experiments)
    int pair_rows = FS_PAIR_MAX_ROWS, pair_zchunk = 0; // FS_PAIR_ROWS / FS_PAIR_ZCHUNK (experiments)
    int pair_mode = 2;          // FS_PAIR: 0 never use the fused two-stage sweep, 1 always, 2 auto (see pair_supported)
    bool pair_force = false;
    bool no_advect_vec4 = false; // FS_NO_ADVECT_VEC4=1: per-cell advect kernels (tests compare both)
    bool pair_slabs = false;    // FS_PAIR_SLABS=1: auto mode also fuses Jacobi / smoother sweeps on z-slabs (measured slower at N = 2: 56.9 vs 47.3 ms/step)
    int l2_ahead = 2;           // prefetch.global.L2 of the planes n steps ahead in the sweep (FS_L2_AHEAD overrides; 0 = off).
                                // measured 512^3: Jacobi 297 -> 246 us, smoother 256 -> 224 us (profiles/r01f_l2_prefetch.md)
    bool use_graph = false;
    int sm_count = 148;

    // device scratch for metrics / scatter_add
    double *d_sum = nullptr;
    unsigned int *d_max = nullptr;
    void *scratch = nullptr;
    size_t scratch_bytes = 0;

    // CUDA graph cache: one executable graph per (dt, visc, diff, role configuration)
    struct GraphEntry {
        float dt, visc, diff;
        float *before[11], *after[11];
        cudaGraphExec_t exec;
        int64_t launches;
    };
    std::vector<GraphEntry> graphs;
    bool capturing = false;
    GraphEntry pending;
    int64_t launches_at_begin = 0;

    void set_error(const char *what, cudaError_t e) {
        if (!bad) msg = std::string(what) + ": " + cudaGetErrorString(e);
        bad = true;
    }
    const std::string &error() const { return msg; }
    bool failed() {
        if (!bad && !capturing) {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) set_error("kernel launch", e);
        }
        return bad;
    }
    void make_current() { cudaSetDevice(dev); }
    // Launch with an explicit priority attribute when the target is the halo stream: a stream's priority is honoured for
    // direct launches but NOT carried into the kernel nodes of a captured graph (measured: the priority stream alone
    // changed the back-to-back sweeps, 60.6 -> 52.5 us at N = 8, and left the graph-replayed step where it was); the
    // attribute is captured with the node.
    int halo_priority = 0;
    bool halo_prio_attr = false; // FS_HALO_PRIORITY=1: highest stream priority + launch-priority attribute for the side stream
    template <class... KArgs, class... Args>
    void launch_on(cudaStream_t stream, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args &&...args) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        int n = 0;
        if (stream == st_halo && halo_prio_attr) { attr[0].id = cudaLaunchAttributePriority; attr[0].val.priority = halo_priority; n = 1; }
        cfg.attrs = attr; cfg.numAttrs = n;
        FS_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
        launches++;
    }

    int open(int device) {
        dev = device;
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            msg = "no CUDA device available (libfluidsolver has no CPU path)";
            bad = true;
            return 1;
        }
        if (device < 0 || device >= count) { msg = "device_id out of range"; bad = true; return 1; }
        FS_CUDA(cudaSetDevice(dev));
        FS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        {   // the halo stream runs the boundary chunks of a sweep and the push beside the interior launch: its CTAs must be
            // dispatched FIRST (highest priority), otherwise they queue behind the interior CTAs and the exchange is serialised
            // after the sweep instead of hidden behind it
            int prio_lo = 0, prio_hi = 0;
            FS_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            const char *np = getenv("FS_HALO_PRIORITY"); // off by default: no effect on graph replay, slightly slower (profiles/r02e_exchange.md)
            halo_prio_attr = np && np[0] == '1';
            halo_priority = prio_hi;
            FS_CUDA(cudaStreamCreateWithPriority(&st_halo, cudaStreamNonBlocking, halo_prio_attr ? prio_hi : prio_lo));
        }
        FS_CUDA(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
        FS_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        FS_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        FS_CUDA(cudaEventCreateWithFlags(&ev_ends, cudaEventDisableTiming));
        FS_CUDA(cudaEventCreate(&ev0));
        FS_CUDA(cudaEventCreate(&ev1));
        FS_CUDA(cudaMalloc(&d_sum, sizeof(double)));
        FS_CUDA(cudaMalloc(&d_max, sizeof(unsigned int)));
        FS_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        const char *fg = getenv("FS_FORCE_GENERIC");
        force_generic = fg && fg[0] == '1';
        if (const char *e = getenv("FS_ZCHUNK")) tune_zchunk = atoi(e);
        if (const char *e = getenv("FS_BLOCK_Y")) tune_by = atoi(e);
        if (const char *e = getenv("FS_PAIR_ROWS")) pair_rows = atoi(e);
        if (const char *e = getenv("FS_PAIR_ZCHUNK")) pair_zchunk = atoi(e);
        if (const char *e = getenv("FS_NO_ADVECT_VEC4")) no_advect_vec4 = e[0] == '1';
        if (const char *e = getenv("FS_NO_TILEMAP")) no_tilemap = e[0] == '1';
        if (const char *e = getenv("FS_EXTEND")) extend_sweeps = e[0] != '0';
        if (const char *e = getenv("FS_XCHG_IN_SWEEP")) xchg_in_sweep = e[0] != '0';
        if (const char *e = getenv("FS_PUSH_CTAS")) push_ctas = atoi(e);
        if (const char *e = getenv("FS_PAIR")) pair_mode = atoi(e);
        if (const char *e = getenv("FS_NO_PAIR")) { if (e[0] == '1') pair_mode = 0; }
        if (const char *e = getenv("FS_PAIR_SLABS")) pair_slabs = e[0] != '0';
        if (pair_rows < 3) pair_rows = 3;
        if (pair_rows > FS_PAIR_MAX_ROWS) pair_rows = FS_PAIR_MAX_ROWS;
        const char *la = getenv("FS_L2_AHEAD");
        if (la) l2_ahead = atoi(la);
        return bad ? 1 : 0;
    }
    void close() {
        invalidate_graph();
        halo_close();
        source_close();
        if (d_sum) cudaFree(d_sum);
        if (d_max) cudaFree(d_max);
        if (scratch) cudaFree(scratch);
        if (render_buf) cudaFree(render_buf);
        render_buf = nullptr; render_bytes = 0;
        if (tile_cum) cudaFree(tile_cum);
        tile_cum = nullptr;
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_ends) cudaEventDestroy(ev_ends);
        if (st_copy) { cudaStreamSynchronize(st_copy); cudaStreamDestroy(st_copy); }
        for (int f = 0; f < FS_FIELD_COUNT; f++) {
            if (stage[f]) cudaFree(stage[f]);
            if (ev_snap[f]) cudaEventDestroy(ev_snap[f]);
            if (ev_sent[f]) cudaEventDestroy(ev_sent[f]);
            stage[f] = nullptr; ev_snap[f] = ev_sent[f] = nullptr; stage_bytes[f] = 0;
        }
        st_copy = nullptr;
        if (st_halo) cudaStreamDestroy(st_halo);
        if (st) cudaStreamDestroy(st);
        d_sum = nullptr; d_max = nullptr; scratch = nullptr; ev0 = ev1 = ev_fork = ev_join = ev_ends = nullptr; st = st_halo = nullptr;
    }

    // ---- memory -------------------------------------------------------------------------------
    void *alloc(size_t bytes) {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { set_error("cudaMalloc", e); return nullptr; }
        return p;
    }
    void free(void *p) { if (p) cudaFree(p); }
    void zero(void *p, size_t bytes) { FS_CUDA(cudaMemsetAsync(p, 0, bytes, st)); }
    void copy(void *dst, const void *src, size_t bytes) { FS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st)); }
    void upload(void *dst, const void *src, size_t bytes) {
        FS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        FS_CUDA(cudaStreamSynchronize(st)); // the caller's buffer is only valid during the call
    }
    void download(void *dst, const void *src, size_t bytes) {
        FS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaStreamSynchronize(st));
    }
    // Pipelined readback: snapshot the field on the compute stream (device->device, ~0.2 ms per 512 MB), then copy
    // the snapshot to the host on the copy stream while the next step computes.  `slot` = field id.
    void download_async(void *dst, const void *src, size_t bytes, int slot) {
        if (stage_bytes[slot] < bytes) {
            if (stage[slot]) { cudaStreamSynchronize(st_copy); cudaFree(stage[slot]); }
            stage[slot] = (float *)alloc(bytes);
            stage_bytes[slot] = stage[slot] ? bytes : 0;
            if (!stage[slot]) return;
        }
        if (!ev_snap[slot]) {
            FS_CUDA(cudaEventCreateWithFlags(&ev_snap[slot], cudaEventDisableTiming));
            FS_CUDA(cudaEventCreateWithFlags(&ev_sent[slot], cudaEventDisableTiming));
        } else {
            FS_CUDA(cudaStreamWaitEvent(st, ev_sent[slot], 0)); // the previous transfer out of this snapshot is done
        }
        FS_CUDA(cudaMemcpyAsync(stage[slot], src, bytes, cudaMemcpyDeviceToDevice, st));
        FS_CUDA(cudaEventRecord(ev_snap[slot], st));
        FS_CUDA(cudaStreamWaitEvent(st_copy, ev_snap[slot], 0));
        FS_CUDA(cudaMemcpyAsync(dst, stage[slot], bytes, cudaMemcpyDeviceToHost, st_copy));
        FS_CUDA(cudaEventRecord(ev_sent[slot], st_copy));
    }
    void wait_transfers() { FS_CUDA(cudaStreamSynchronize(st_copy)); }
    void sync() {
        FS_CUDA(cudaStreamSynchronize(st));
        if (halo_on && my_flags) {
            unsigned e = 0;
            FS_CUDA(cudaMemcpy(&e, my_flags + FS_HF_ERROR, sizeof(e), cudaMemcpyDeviceToHost));
            if (e && !bad) {
                bad = true;
                msg = e == 2 ? "halo exchange timed out: a neighbouring slab did not publish its planes (rank missing or failed)"
                             : "advection back-trace left the neighbouring slab (slab too thin for this CFL)";
            }
        }
    }
    void *get_scratch(size_t bytes) {
        if (bytes > scratch_bytes) {
            if (scratch) { cudaStreamSynchronize(st); cudaFree(scratch); }
            scratch_bytes = bytes * 2;
            scratch = alloc(scratch_bytes);
        }
        return scratch;
    }

    // ---- launch helpers --------------------------------------------------------------------------
    static void interior_planes(const FsGrid &g, int *kl0, int *count) {
        if (!g.hz) { *kl0 = 0; *count = 1; return; }
        const int zb = g.zoff + g.kb, ze = g.zoff + g.ke;
        const int k0 = zb < 1 ? 1 : zb, k1 = ze > g.nz - 1 ? g.nz - 1 : ze;
        *kl0 = k0 - g.zoff;
        *count = k1 > k0 ? k1 - k0 : 0;
    }
    template <class F>
    void cells(const FsGrid &g, F f) {
        int kl0, cnt;
        interior_planes(g, &kl0, &cnt);
        cells_range(g, kl0, cnt, f);
    }
    template <class F>
    void cells_range(const FsGrid &g, int kl0, int cnt, F f) { // interior rows / columns of local planes [kl0, kl0+cnt)
        if (cnt <= 0) return;
        const dim3 block(64, 4, 1);
        const dim3 grid((g.nx - 2 + 63) / 64, (g.ny - 2 + 3) / 4, cnt);
        cells_kernel<<<grid, block, 0, st>>>(g, kl0, f);
        launches++;
    }
    template <class F>
    void linear(long long n, F f) {
        if (n <= 0) return;
        linear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, f);
        launches++;
    }

    // ---- sweeps ----------------------------------------------------------------------------------
    // xmode (fs_cellops.cuh, FS_X_*): how a sweep on z-slabs relates to the halo exchange.
    //   FS_X_NONE      owned planes, no exchange
    //   FS_X_EXCHANGE  owned planes, then the FS_GHOST boundary planes of `out` go to the neighbours (one halo operation)
    //   FS_X_EXTEND    owned planes PLUS the first ghost plane on each internal side, computed redundantly from the
    //                  two-deep ghost zone of `in`; no exchange.  An FS_X_EXCHANGE* sweep must follow: it finds `in` valid
    //                  one plane into the ghost zone, which is all a single sweep reads, and its push restores both planes.
    //                  Two sweeps per halo operation instead of one.
    //   FS_X_EXCHANGE_OPEN  as FS_X_EXCHANGE, but the caller promises that the next sweep is an FS_X_EXTEND one: the fork is
    //                  left open, so that only the boundary chunks of the next sweep wait for the neighbours' planes.
    void relax(int mode, const FsGrid &g, const float *in, const float *rhs, const float *stale, float *out,
               const uint8_t *flags, float a, float c, int b, bool in_zero, int xmode) {
        const float *ins[1] = {in}, *rhss[1] = {rhs}, *stales[1] = {stale};
        float *outs[1] = {out};
        relax_n(mode, g, 1, ins, rhss, stales, outs, flags, a, c, &b, in_zero, xmode);
    }
    bool can_extend(const FsGrid &g) const { return halo_on && g.hz && extend_sweeps; }
    // One sweep over nf <= FS_BATCH fields that share a, c and the flags (the velocity components of a diffusion sweep),
    // in ONE launch and, on z-slabs, ONE halo operation.
    void relax_n(int mode, const FsGrid &g, int nf, const float *const *in, const float *const *rhs, const float *const *stale,
                 float *const *out, const uint8_t *flags, float a, float c, const int *b, bool in_zero, int xmode) {
        int kl0, cnt;
        interior_planes(g, &kl0, &cnt);
        if (cnt <= 0) return;
        const bool exchange = halo_on && (xmode == FS_X_EXCHANGE || xmode == FS_X_EXCHANGE_OPEN);
        const bool extend = halo_on && xmode == FS_X_EXTEND;
        if (extend) { // one plane into the ghost zone on each side that has a neighbour
            if (lo.present) { kl0 -= 1; cnt += 1; }
            if (hi.present) cnt += 1;
        }
        const bool c_ok = c != 0.0f && c == c && c - c == 0.0f; // finite, non-zero: fs_div's precondition
        if (g.nx % 4 == 0 && !force_generic && c_ok) {
            FsRelaxBatch batch{};
            batch.nf = nf;
            FsTileMap tiles{};
            if (flags && tile_cum && !no_tilemap) { tiles.cum = tile_cum; tiles.tx_count = tile_tx; tiles.ty_count = tile_ty; }
            for (int f = 0; f < nf; f++) {
                batch.in[f] = in[f]; batch.rhs[f] = rhs ? rhs[f] : nullptr; batch.stale[f] = stale ? stale[f] : nullptr;
                batch.out[f] = out[f]; batch.b[f] = b[f];
            }
            const int groups = g.nx / 4;
            // small grids: smaller CTAs so that there are enough of them to occupy 148 SMs
            const long long thread_planes = (long long)groups * (g.ny - 2) * cnt * nf;
            const int threads = thread_planes >= (long long)sm_count * 256 * 16 ? 256 : 64;
            int bx = 32;
            while (bx / 2 >= groups && bx > 1) bx /= 2;
            const int by = tune_by > 0 ? tune_by : threads / bx; // FS_BLOCK_Y (experiments)
            const int gxn = (groups + bx - 1) / bx, gyn = (g.ny - 2 + by - 1) / by;
            const long long blocks_xy = (long long)gxn * gyn;
            const long long target = (long long)sm_count * 16;
            // z chunk per CTA for `planes` planes: short enough for >= ~4 waves of 4 resident CTAs/SM (tail effect), long
            // enough that re-reading the two halo planes per chunk stays <= 2/16 of one field
            auto chunk_for = [&](int planes) {
                long long z = (long long)planes * blocks_xy * nf / target;
                z = z < 4 ? 4 : (z > 16 ? 16 : z);
                if (tune_zchunk > 0) z = tune_zchunk; // FS_ZCHUNK (experiments)
                if (z > planes) z = planes;
                return (int)z;
            };
            const int iz = in_zero ? 1 : 0;
            const dim3 block(bx, by, 1);
            cudaStream_t ls = st;                                  // stream of the next launch
            int kl_b = kl0, kl_e = kl0 + cnt, kl_alt = 0, zc = 1;  // plane range / chunk length of the next launch
            FsSweepXchg xc{};
#define FS_LAUNCH_RELAX(MODE_, HZ_, NZ_, BASE_, STRIDE_) \
    do { const dim3 grid(gxn, gyn, (NZ_) * nf); \
         launch_on(ls, relax_vec4<MODE_, HZ_, false>, grid, block, g, batch, flags, tiles, a, c, iz, kl_b, kl_e, zc, BASE_, STRIDE_, l2_ahead, kl_alt, xc); } while (0)
#define FS_LAUNCH_RELAX_XCHG(MODE_, NZ_) \
    do { const dim3 grid(gxn, gyn, (NZ_) * nf + 1); \
         launch_on(ls, relax_vec4<MODE_, true, true>, grid, block, g, batch, flags, tiles, a, c, iz, kl_b, kl_e, zc, 0, 1, l2_ahead, kl_alt, xc); } while (0)
#define FS_LAUNCH_RELAX_MODE(NZ_, BASE_, STRIDE_) \
    do { if (mode == FS_MODE_SMOOTH) { if (g.hz) FS_LAUNCH_RELAX(FS_MODE_SMOOTH, true, NZ_, BASE_, STRIDE_); else FS_LAUNCH_RELAX(FS_MODE_SMOOTH, false, NZ_, BASE_, STRIDE_); } \
         else { if (g.hz) FS_LAUNCH_RELAX(FS_MODE_JACOBI, true, NZ_, BASE_, STRIDE_); else FS_LAUNCH_RELAX(FS_MODE_JACOBI, false, NZ_, BASE_, STRIDE_); } } while (0)
            // the whole range [kl0, kl0+cnt) as one launch on stream ls
            auto launch_all = [&]() {
                kl_b = kl0; kl_e = kl0 + cnt; zc = chunk_for(cnt);
                FS_LAUNCH_RELAX_MODE((cnt + zc - 1) / zc, 0, 1);
            };
            // [kl0, kl0+ends) and [kl0+cnt-ends, kl0+cnt) as one two-chunk launch; everything between as another
            auto launch_ends = [&](int ends) {
                kl_b = kl0; kl_e = kl0 + cnt; kl_alt = kl0 + cnt - ends; zc = ends;
                FS_LAUNCH_RELAX_MODE(2, 0, 0);
            };
            auto launch_middle = [&](int ends) {
                const int mid = cnt - 2 * ends;
                kl_b = kl0 + ends; kl_e = kl0 + cnt - ends; zc = chunk_for(mid);
                FS_LAUNCH_RELAX_MODE((mid + zc - 1) / zc, 0, 1);
            };
            const int be = 4, ee = 3; // planes per end of an exchange sweep's / an extended sweep's boundary launch (see below)
            if (exchange && xchg_in_sweep && g.hz && cnt >= 2 * be + 4) {
                // ONE launch: ends, push CTAs, middle (relax_vec4<.., XCHG = true>)
                if (fork_open) join_fork();
                const unsigned op = ++ops_since_commit;
                xc.h = halo_args(g, out, nf, op);
                xc.h.need_ack = ack_since_last_op ? 1 : 0;
                ack_since_last_op = false;
                xc.plane_elems = g.sz * FS_GHOST;
                xc.ends = be;
                xc.ends_ctas = (unsigned)(2 * nf * gxn * gyn);
                int pc = push_ctas > 0 ? push_ctas : 16 * nf;
                if (pc > gxn * gyn) pc = gxn * gyn;
                xc.push_ctas = (unsigned)pc;
                const int mid = cnt - 2 * be;
                kl_b = kl0; kl_e = kl0 + cnt; kl_alt = kl0 + cnt - be; zc = chunk_for(mid);
                const int nz = 2 + (mid + zc - 1) / zc;
                if (mode == FS_MODE_SMOOTH) FS_LAUNCH_RELAX_XCHG(FS_MODE_SMOOTH, nz);
                else FS_LAUNCH_RELAX_XCHG(FS_MODE_JACOBI, nz);
            } else if (exchange && cnt >= 2 * be + 4) {
                // fork: side stream = the two ends of the slab, then the P2P push kernel (stores them into the neighbours'
                // ghosts, signals, awaits theirs); main stream = everything between, concurrently; join
                if (fork_open) join_fork();                        // (callers pair OPEN with EXTEND; defensive)
                FS_CUDA(cudaEventRecord(ev_fork, st));
                FS_CUDA(cudaStreamWaitEvent(st_halo, ev_fork, 0));
                ls = st_halo;
                launch_ends(be);
                FS_CUDA(cudaEventRecord(ev_ends, st_halo));
                halo_n_on_stream(g, out, nf, st_halo);
                ls = st;
                launch_middle(be);
                if (xmode == FS_X_EXCHANGE_OPEN) fork_open = true;
                else join_fork();
            } else if (extend && fork_open && cnt >= 2 * ee + 4) {
                // The previous sweep's fork is still open: its push (side stream) may be in flight.  Only the ends of THIS
                // sweep read ghost planes, so they follow the push on the side stream, while the middle starts as soon as
                // the previous sweep's ends are done.  Data flow: the ends ([kl0, kl0+3) and its mirror image) read the
                // previous output on planes kl0-1 .. kl0+3, all of which the previous sweep's 4-plane ends or the push wrote
                // (same stream); they overwrite planes <= kl0+2 of the previous INPUT buffer, which the previous middle
                // launch (planes >= kl0+be, reading from kl0+be-1 = kl0+3 at the owned range; one less here) no longer reads.
                ls = st_halo;
                launch_ends(ee);
                halo_ack_on_stream(st_halo);
                ls = st;
                FS_CUDA(cudaStreamWaitEvent(st, ev_ends, 0));
                launch_middle(ee);
                join_fork();
            } else {
                if (fork_open) join_fork();
                launch_all();
                if (exchange) halo_n_on_stream(g, out, nf, st);
                if (extend) halo_ack_on_stream(st);
            }
#undef FS_LAUNCH_RELAX_MODE
#undef FS_LAUNCH_RELAX_XCHG
#undef FS_LAUNCH_RELAX
            return;
        }
        if (fork_open) join_fork();
        for (int f = 0; f < nf; f++) {
            const float *in_f = in[f], *rhs_f = rhs ? rhs[f] : nullptr, *stale_f = stale ? stale[f] : nullptr;
            float *out_f = out[f];
            const int b_f = b[f];
            if (mode == FS_MODE_SMOOTH)
                cells_range(g, kl0, cnt, [=] __device__(int i, int j, int kl) { fs_relax_cell<FS_MODE_SMOOTH>(g, in_f, rhs_f, stale_f, out_f, flags, a, c, b_f, in_zero, i, j, kl); });
            else
                cells_range(g, kl0, cnt, [=] __device__(int i, int j, int kl) { fs_relax_cell<FS_MODE_JACOBI>(g, in_f, rhs_f, stale_f, out_f, flags, a, c, b_f, in_zero, i, j, kl); });
        }
        if (exchange) halo_n_on_stream(g, out, nf, st); // per-cell fallback: push after the whole sweep
        if (extend) halo_ack_on_stream(st);
    }
    // FS_EXTEND=1: two sweeps per halo operation (extended sweeps).  Built and bit exact; measured neutral on NVSwitch
    // (N = 8, 512^3: 18.0 vs 17.7 ms per step) because the single-launch exchange sweep already hides the operation, so the
    // default is one operation per sweep and no acknowledgement launches.  For links slower than NVLink.
    bool extend_sweeps = false;
    bool xchg_in_sweep = true;    // FS_XCHG_IN_SWEEP=0: fork / join with a separate push kernel instead of the single launch
    int push_ctas = 0;            // FS_PUSH_CTAS: CTAs of the single launch that carry the halo operation (default 16 per field)
    bool fork_open = false;       // an FS_X_EXCHANGE_OPEN sweep has left the side stream un-joined
    void join_fork() {
        FS_CUDA(cudaEventRecord(ev_join, st_halo));
        FS_CUDA(cudaStreamWaitEvent(st, ev_join, 0));  // (also required before a graph capture ends)
        fork_open = false;
    }
    // end of a run of sweeps: nothing may stay forked
    void relax_end() {
        if (fork_open) join_fork();
    }
    // Fused two-stage sweep (fs_kernels.cuh relax_pair): out = S2(S1(in)).  Returns false when this grid / field
    // cannot take it (the caller then issues two single sweeps): 2D, nx % 4 != 0, unusable divisor, FS_NO_PAIR.
    // Policy (FS_PAIR = 0 never / 1 always / unset: auto).  Measured on B200 (profiles/r02b_pair_kernel.md): the fused
    // kernel moves 13 B/voxel per two stages but is issue bound -- 0.64 ms per Jacobi pair at 512^3 against 2 x 0.26 ms
    // for two single sweeps that already run at the HBM roofline, and on two z-slabs 56.9 vs 47.3 ms per step; the
    // red-black form beats its two-launch version (0.66 vs 0.75 ms).  So auto = red-black only (FS_PAIR_SLABS=1 adds
    // Jacobi / smoother pairs on z-slabs for experiments).
    bool pair_supported(const FsGrid &g, float c, int kind) const {
        const bool c_ok = c != 0.0f && c == c && c - c == 0.0f;
        if (!(g.hz && g.nx % 4 == 0 && !force_generic && c_ok)) return false;
        if (pair_force) return true;                     // fs_bench_sweep times the kernel whatever the policy says
        if (pair_mode == 0) return false;
        if (pair_mode == 1) return true;
        return kind == FS_PAIR_RED_BLACK || (halo_on && pair_slabs);
    }
    bool relax_pair(int kind, const FsGrid &g, const float *in, const float *rhs, float *out, const uint8_t *flags,
                    float a, float c, int b, bool in_zero, bool fuse_halo) {
        if (!pair_supported(g, c, kind)) return false;
        int kl0, cnt;
        interior_planes(g, &kl0, &cnt);
        if (cnt <= 0) return true;
        const int ncols = g.nx / 4;
        const int wcols = ncols < FS_PAIR_W ? ncols : FS_PAIR_W, cols = wcols + 2;
        int rows = pair_rows;
        if (rows > g.ny) rows = g.ny;                    // ny - 2 inner rows + the two ring rows
        const int threads = ((cols * rows + 31) / 32) * 32;
        const int gxn = (ncols + wcols - 1) / wcols, gyn = (g.ny - 2 + rows - 3) / (rows - 2);
        // z chunk per CTA: stage 1 is evaluated on zchunk + 2 planes, so long chunks; but enough CTAs for ~2 waves
        long long zchunk = (long long)cnt * gxn * gyn / ((long long)sm_count * 2);
        if (zchunk < 8) zchunk = 8;
        if (zchunk > 64) zchunk = 64;
        if (halo_on && zchunk > cnt / 4) zchunk = cnt / 4 < 4 ? 4 : cnt / 4; // slabs: keep interior chunks to overlap the push
        if (pair_zchunk > 0) zchunk = pair_zchunk;
        if (zchunk > cnt) zchunk = cnt;
        const int nchunks = (int)((cnt + zchunk - 1) / zchunk);
        const int iz = in_zero ? 1 : 0, zc = (int)zchunk;
#define FS_LAUNCH_PAIR(NZ_, BASE_, STRIDE_) \
    do { const dim3 grid(gxn, gyn, NZ_); \
         const dim3 blk(threads, 1, 1); const float *no_rhs = nullptr; \
         if (kind == FS_PAIR_JACOBI) launch_on(st, relax_pair_kernel<FS_PAIR_JACOBI>, grid, blk, g, in, rhs, out, flags, a, c, b, iz, kl0, kl0 + cnt, zc, BASE_, STRIDE_, l2_ahead, wcols, rows); \
         else if (kind == FS_PAIR_SMOOTH) launch_on(st, relax_pair_kernel<FS_PAIR_SMOOTH>, grid, blk, g, in, no_rhs, out, flags, a, c, b, iz, kl0, kl0 + cnt, zc, BASE_, STRIDE_, l2_ahead, wcols, rows); \
         else launch_on(st, relax_pair_kernel<FS_PAIR_RED_BLACK>, grid, blk, g, in, rhs, out, flags, a, c, b, iz, kl0, kl0 + cnt, zc, BASE_, STRIDE_, l2_ahead, wcols, rows); } while (0)
        if (halo_on && fuse_halo && nchunks > 2) {
            FS_CUDA(cudaEventRecord(ev_fork, st));
            FS_CUDA(cudaStreamWaitEvent(st_halo, ev_fork, 0));
            { cudaStream_t main_st = st; st = st_halo;
              FS_LAUNCH_PAIR(2, 0, nchunks - 1);           // the two chunks holding the slab's boundary planes
              st = main_st; }
            halo_on_stream(g, out, st_halo);
            FS_CUDA(cudaEventRecord(ev_join, st_halo));
            FS_LAUNCH_PAIR(nchunks - 2, 1, 1);             // interior chunks, concurrently
            FS_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
        } else {
            FS_LAUNCH_PAIR(nchunks, 0, 1);
            if (halo_on && fuse_halo) halo(g, out);
        }
#undef FS_LAUNCH_PAIR
        return true;
    }
    // returns true when the launch also performed set_bnd (colour 1 of the float4 kernel)
    bool rb_half(const FsGrid &g, float *x, const float *rhs, const uint8_t *flags, float a, float c, int colour, int b) {
        dim3 grid, block;
        int kl0;
        const bool c_ok = c != 0.0f && c == c && c - c == 0.0f;
        if (c_ok && vec4_geometry(g, &grid, &block, &kl0)) {
            const int ring = colour == 1 ? 1 : 0;
            if (g.hz) rb_vec4<true><<<grid, block, 0, st>>>(g, x, rhs, flags, a, c, colour, b, ring, kl0);
            else rb_vec4<false><<<grid, block, 0, st>>>(g, x, rhs, flags, a, c, colour, b, ring, kl0);
            launches++;
            return ring != 0;
        }
        cells(g, [=] __device__(int i, int j, int kl) {
            if (((i + j + kl + g.zoff) & 1) == colour) fs_rb_cell(g, x, rhs, flags, a, c, i, j, kl);
        });
        return false;
    }
    void bnd(const FsGrid &g, float *x, int b) {
        cells(g, [=] __device__(int i, int j, int kl) {
            const bool edge = i == 1 || i == g.nx - 2 || j == 1 || j == g.ny - 2 ||
                              (g.hz && (kl + g.zoff == 1 || kl + g.zoff == g.nz - 2));
            if (edge) fs_bnd_cell(g, x, b, i, j, kl);
        });
    }
    void mirror(const FsGrid &g, float *x, const uint8_t *flags, const long long *list, long long n, int b) {
        linear(n, [=] __device__(long long t) { fs_mirror_cell(g, x, flags, b, list[t]); });
    }
    // the obstacle pass on every velocity component in one launch (each component mirrors along its own axis)
    void mirror3(const FsGrid &g, float *ux, float *uy, float *uz, const uint8_t *flags, const long long *list, long long n) {
        const int nf = uz ? 3 : 2;
        linear(n * nf, [=] __device__(long long t) {
            const int f = (int)(t / n);
            fs_mirror_cell(g, f == 0 ? ux : (f == 1 ? uy : uz), flags, f + 1, list[t - f * n]);
        });
    }
    // launch geometry of the one-plane-per-thread float4 kernels
    bool vec4_geometry(const FsGrid &g, dim3 *grid, dim3 *block, int *kl0) {
        int cnt;
        interior_planes(g, kl0, &cnt);
        if (cnt <= 0 || g.nx % 4 != 0 || force_generic) return false;
        const int groups = g.nx / 4;
        int bx = 32;
        while (bx / 2 >= groups && bx > 1) bx /= 2;
        const int by = 256 / bx;
        *block = dim3(bx, by, 1);
        *grid = dim3((groups + bx - 1) / bx, (g.ny - 2 + by - 1) / by, cnt);
        return true;
    }
    void divergence(const FsGrid &g, float *div, const float *ux, const float *uy, const float *uz) {
        dim3 grid, block;
        int kl0;
        if (vec4_geometry(g, &grid, &block, &kl0)) {
            if (g.hz) divergence_vec4<true><<<grid, block, 0, st>>>(g, div, ux, uy, uz, kl0);
            else divergence_vec4<false><<<grid, block, 0, st>>>(g, div, ux, uy, uz, kl0);
            launches++;
            return;
        }
        cells(g, [=] __device__(int i, int j, int kl) { fs_divergence_cell(g, div, ux, uy, uz, i, j, kl); });
    }
    void gradient(const FsGrid &g, float *ux, float *uy, float *uz, const float *p, const uint8_t *flags) {
        dim3 grid, block;
        int kl0;
        if (vec4_geometry(g, &grid, &block, &kl0)) {
            if (g.hz) gradient_vec4<true><<<grid, block, 0, st>>>(g, ux, uy, uz, p, flags, kl0);
            else gradient_vec4<false><<<grid, block, 0, st>>>(g, ux, uy, uz, p, flags, kl0);
            launches++;
            return;
        }
        cells(g, [=] __device__(int i, int j, int kl) { fs_gradient_cell(g, ux, uy, uz, p, flags, i, j, kl); });
    }
    FsSlabView slab_view(const FsGrid &g, const float *field) const {
        FsSlabView v{};
        v.loc = field; v.zoff = g.zoff; v.nzl = g.nzl;
        if (halo_on) {
            const int bi = buf_index(field);
            if (lo.present && bi >= 0) { v.lo = lo.base[bi]; v.lo_zoff = lo.zoff; v.lo_nzl = lo.nzl; }
            if (hi.present && bi >= 0) { v.hi = hi.base[bi]; v.hi_zoff = hi.zoff; v.hi_nzl = hi.nzl; }
            v.err = my_flags + FS_HF_ERROR;
        }
        return v;
    }
    void advect(const FsGrid &g, float *d, const float *d0, const float *ux, const float *uy, const float *uz,
                const uint8_t *flags, float dt0, int b) {
        const FsSlabView v = slab_view(g, d0);
        dim3 grid, block;
        int kl0;
        if (!no_advect_vec4 && vec4_geometry(g, &grid, &block, &kl0)) {
            FsAdvectBatch ab{};
            ab.dst[0] = d; ab.src[0] = v; ab.b[0] = b;
            if (g.hz) advect_vec4<1, true><<<grid, block, 0, st>>>(g, ab, ux, uy, uz, flags, dt0, kl0);
            else advect_vec4<1, false><<<grid, block, 0, st>>>(g, ab, ux, uy, uz, flags, dt0, kl0);
            launches++;
            return;
        }
        cells(g, [=] __device__(int i, int j, int kl) {
            auto samp = [&](int kk) { return fs_slab_plane(v, g, kk); };
            fs_advect_cell(g, d, samp, ux, uy, uz, flags, dt0, b, i, j, kl);
        });
    }
    void advect_velocity(const FsGrid &g, float *dx, float *dy, float *dz, const float *sx, const float *sy,
                         const float *sz, const uint8_t *flags, float dt0) {
        const FsSlabView vx_ = slab_view(g, sx), vy_ = slab_view(g, sy), vz_ = g.hz ? slab_view(g, sz) : FsSlabView{};
        dim3 grid, block;
        int kl0;
        if (!no_advect_vec4 && vec4_geometry(g, &grid, &block, &kl0)) {
            FsAdvectBatch ab{};
            ab.dst[0] = dx; ab.dst[1] = dy; ab.dst[2] = dz;
            ab.src[0] = vx_; ab.src[1] = vy_; ab.src[2] = vz_;
            ab.b[0] = 1; ab.b[1] = 2; ab.b[2] = 3;
            if (g.hz) advect_vec4<3, true><<<grid, block, 0, st>>>(g, ab, sx, sy, sz, flags, dt0, kl0);
            else advect_vec4<2, false><<<grid, block, 0, st>>>(g, ab, sx, sy, sz, flags, dt0, kl0);
            launches++;
            return;
        }
        cells(g, [=] __device__(int i, int j, int kl) {
            auto px = [&](int kk) { return fs_slab_plane(vx_, g, kk); };
            auto py = [&](int kk) { return fs_slab_plane(vy_, g, kk); };
            auto pz = [&](int kk) { return fs_slab_plane(vz_, g, kk); };
            fs_advect_velocity_cell(g, dx, dy, dz, px, py, pz, sx, sy, sz, flags, dt0, i, j, kl);
        });
    }
    void enforce(const FsGrid &g, float *ux, float *uy, float *uz, const uint8_t *flags, float cell, float rawvisc) {
        cells(g, [=] __device__(int i, int j, int kl) { fs_enforce_cell(g, ux, uy, uz, flags, cell, rawvisc, i, j, kl); });
    }
    float *render_buf = nullptr;
    size_t render_bytes = 0;
    void *render_buffer(size_t bytes) {
        if (bytes > render_bytes) {
            if (render_buf) { cudaStreamSynchronize(st); cudaFree(render_buf); }
            render_buf = (float *)alloc(bytes);
            render_bytes = render_buf ? bytes : 0;
        }
        return render_buf;
    }
    void visualize(const FsGrid &g, const fs_vis_params &vp, const float *d, const float *p, const uint8_t *mask, float *rgba) {
        const int nx = g.nx;
        linear(g.sz, [=] __device__(long long t) {
            const FsColor c = fs_visualize_cell(vp, d[t], p[t], mask[t] != 0, (int)(t % nx), (int)(t / nx));
            reinterpret_cast<float4 *>(rgba)[t] = make_float4(c.r, c.g, c.b, c.a);
        });
    }
    void streamlines(const FsGrid &g, int skip, float scale, const float *ux, const float *uy, const uint8_t *mask, float *out,
                     long long count) {
        const int nx = g.nx, ny = g.ny;
        linear(count, [=] __device__(long long t) { fs_streamline_glyph(nx, ny, skip, scale, ux, uy, mask, (int)t, out + 4 * t); });
    }
    // coarse obstacle map for the sweeps (fs_kernels.cuh FsTileMap), rebuilt with the flags
    int *tile_cum = nullptr;
    int tile_tx = 0, tile_ty = 0;
    bool no_tilemap = false; // FS_NO_TILEMAP=1 (experiments)
    void build_flags(const FsGrid &g, const uint8_t *mask, uint8_t *flags) {
        if (tile_cum) { cudaStreamSynchronize(st); cudaFree(tile_cum); tile_cum = nullptr; }
        if (g.nx % 4 == 0) {
            tile_tx = (g.nx + FS_TILE_X - 1) / FS_TILE_X;
            tile_ty = (g.ny - 2 + FS_TILE_Y - 1) / FS_TILE_Y;
            const long long ntiles = (long long)tile_tx * tile_ty;
            tile_cum = (int *)alloc(sizeof(int) * (size_t)(ntiles * (g.nzl + 1)));
            if (tile_cum) {
                tilemap_mark_kernel<<<(unsigned)((ntiles * g.nzl + 255) / 256), 256, 0, st>>>(g, mask, tile_cum, tile_tx, tile_ty);
                tilemap_scan_kernel<<<(unsigned)((ntiles + 255) / 256), 256, 0, st>>>(tile_cum, (int)ntiles, g.nzl);
                launches += 2;
            }
        }
        const long long n = g.sz * g.nzl;
        linear(n, [=] __device__(long long t) {
            const int i = (int)(t % g.nx), j = (int)((t / g.nx) % g.ny), kl = (int)(t / g.sz);
            flags[t] = fs_flags_cell(g, mask, i, j, kl);
        });
    }
    void fill_random(float *dst, long long n, unsigned seed) { // uniform in [-1, 1), hash based (bench only)
        linear(n, [=] __device__(long long t) {
            unsigned h = (unsigned)t * 2654435761u ^ (unsigned)(t >> 32) ^ (seed * 0x9E3779B9u);
            h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
            dst[t] = (float)(h >> 8) * (1.0f / 8388608.0f) - 1.0f;
        });
    }
    void axpy(float *dst, const float *src, long long n) {
        linear(n, [=] __device__(long long t) { dst[t] += src[t]; });
    }
    // ---- source staging ring: pinned host slot -> device slot -> scatter kernel, no per-call synchronisation ----
    struct SourceSlot {
        char *host = nullptr, *dev = nullptr;
        size_t bytes = 0;
        cudaEvent_t done = nullptr;
        bool pending = false;
    };
    static const int kSourceSlots = 4;
    SourceSlot src_slots[kSourceSlots];
    int src_next = 0, src_cur = 0;
    bool source_stage(long long cap, long long **idx, float *amt[4]) {
        SourceSlot &sl = src_slots[src_cur = src_next];
        src_next = (src_next + 1) % kSourceSlots;
        if (sl.pending) { FS_CUDA(cudaEventSynchronize(sl.done)); sl.pending = false; } // only if the GPU is a whole ring behind
        const size_t need = (sizeof(long long) + 4 * sizeof(float)) * (size_t)cap;
        if (need > sl.bytes) {
            if (sl.host) cudaFreeHost(sl.host);
            if (sl.dev) cudaFree(sl.dev);
            sl.host = sl.dev = nullptr;
            sl.bytes = need * 2;
            if (cudaMallocHost(&sl.host, sl.bytes) != cudaSuccess || cudaMalloc(&sl.dev, sl.bytes) != cudaSuccess) {
                msg = "source staging allocation failed"; bad = true; sl.bytes = 0;
                return false;
            }
            if (!sl.done) FS_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        }
        *idx = (long long *)sl.host;
        float *f = (float *)(sl.host + sizeof(long long) * (size_t)cap);
        for (int k = 0; k < 4; k++) amt[k] = f + (size_t)cap * k;
        return true;
    }
    void scatter_add_staged(float *dst[4], long long cap, long long n) {
        if (n <= 0) return;
        SourceSlot &sl = src_slots[src_cur];
        const size_t used = (sizeof(long long) + 4 * sizeof(float)) * (size_t)cap;
        FS_CUDA(cudaMemcpyAsync(sl.dev, sl.host, used, cudaMemcpyHostToDevice, st));
        const long long *di = (const long long *)sl.dev;
        const float *a0 = (const float *)(sl.dev + sizeof(long long) * (size_t)cap), *a1 = a0 + cap, *a2 = a1 + cap, *a3 = a2 + cap;
        float *d0 = dst[0], *d1 = dst[1], *d2 = dst[2], *d3 = dst[3];
        // duplicates in idx are legal (two source cells clamped to one voxel): use atomics
        linear(n, [=] __device__(long long t) {
            const long long c = di[t];
            if (d0) atomicAdd(d0 + c, a0[t]);
            if (d1) atomicAdd(d1 + c, a1[t]);
            if (d2) atomicAdd(d2 + c, a2[t]);
            if (d3) atomicAdd(d3 + c, a3[t]);
        });
        FS_CUDA(cudaEventRecord(sl.done, st));
        sl.pending = true;
    }
    void source_close() {
        for (SourceSlot &sl : src_slots) {
            if (sl.host) cudaFreeHost(sl.host);
            if (sl.dev) cudaFree(sl.dev);
            if (sl.done) cudaEventDestroy(sl.done);
            sl = SourceSlot{};
        }
    }

    // ---- obstacle bookkeeping on the device ------------------------------------------------------------------------
    // any_local: some cell of the local planes (ghosts included) is an obstacle; list: local indices of the obstacle cells
    // on interior rows / columns of local planes [kl0, kl1) (the owned interior planes), in no particular order (the mirror
    // pass that consumes it is order independent).
    bool scan_obstacles(const FsGrid &g, const uint8_t *mask, int kl0, int kl1, bool *any_local, long long **list, long long *count) {
        unsigned long long *ctr = (unsigned long long *)get_scratch(3 * sizeof(unsigned long long));
        if (!ctr) return false;
        FS_CUDA(cudaMemsetAsync(ctr, 0, 3 * sizeof(unsigned long long), st));
        const long long n = g.sz * g.nzl;
        const int nx = g.nx, ny = g.ny;
        const long long sz = g.sz;
        auto interior_owned = [=] __device__(long long t) {
            const int i = (int)(t % nx), j = (int)((t / nx) % ny), kl = (int)(t / sz);
            return i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2 && kl >= kl0 && kl < kl1;
        };
        linear(n, [=] __device__(long long t) {
            if (!mask[t]) return;
            atomicAdd(ctr + 0, 1ull); // (rare cells: contention is irrelevant next to the plane-sized reads)
            if (interior_owned(t)) atomicAdd(ctr + 1, 1ull);
        });
        unsigned long long h[3] = {0, 0, 0};
        FS_CUDA(cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaStreamSynchronize(st));
        *any_local = h[0] != 0;
        *count = (long long)h[1];
        *list = nullptr;
        if (h[1] == 0) return !bad;
        long long *out = (long long *)alloc(sizeof(long long) * h[1]);
        if (!out) return false;
        linear(n, [=] __device__(long long t) {
            if (mask[t] && interior_owned(t)) out[atomicAdd(ctr + 2, 1ull)] = t;
        });
        *list = out;
        return !bad;
    }
    // SetupObstacles on the device (next row N4): inside test, the reference's 4-neighbour flood fill on the xy cross
    // section, the local planes of the mask, and the GLOBAL obstacle counts (evaluated over the whole grid analytically,
    // so that every slab takes the same decisions without any exchange).
    bool build_shape(const FsGrid &g, const fs_obstacle_shape &sh, uint8_t *mask, long long *total, long long *interior) {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        const bool hz = g.hz != 0;
        const long long plane = g.sz;
        uint8_t *inside = (uint8_t *)alloc(2 * plane + 2 * sizeof(unsigned long long) + 16);
        if (!inside) return false;
        uint8_t *reach = inside + plane;
        unsigned long long *ctr = (unsigned long long *)(inside + ((2 * plane + 15) / 16) * 16);
        FS_CUDA(cudaMemsetAsync(inside, 0, 2 * plane + 2 * sizeof(unsigned long long) + 16, st));
        const bool need_fill = !(sh.kind == 0 && hz);
        const bool seed_in_grid = sh.seed_x >= 0 && sh.seed_x < nx && sh.seed_y >= 0 && sh.seed_y < ny && (!hz || (sh.seed_z >= 0 && sh.seed_z < nz));
        bool seed_ok = false;
        const fs_obstacle_shape shape = sh;
        if (need_fill) {
            linear(plane, [=] __device__(long long t) {
                const int x = (int)(t % nx), y = (int)(t / nx);
                inside[t] = shape.kind == 0 ? fs_shape_inside_circle(shape, false, x, y, 0) : fs_shape_inside_xy(shape, x, y);
            });
            if (seed_in_grid) {
                flood_fill_kernel<<<1, 1024, 0, st>>>(nx, ny, inside, reach, shape.seed_x, shape.seed_y);
                launches++;
            }
            seed_ok = seed_in_grid && fs_shape_in_span(sh, hz && sh.kind != 0, sh.seed_z);
        } else {
            seed_ok = seed_in_grid && fs_shape_inside_circle(sh, true, sh.seed_x, sh.seed_y, sh.seed_z);
        }
        const int zoff = g.zoff;
        const long long nloc = plane * g.nzl;
        const bool ok = seed_ok;
        linear(nloc, [=] __device__(long long t) {
            const int x = (int)(t % nx), y = (int)((t / nx) % ny), kl = (int)(t / plane);
            mask[t] = ok ? fs_shape_mask(shape, hz, nx, reach, ok, x, y, kl + zoff) : 0;
        });
        const long long ncell = plane * nz;
        linear(ncell, [=] __device__(long long t) {
            const int x = (int)(t % nx), y = (int)((t / nx) % ny), z = (int)(t / plane);
            if (!ok || !fs_shape_mask(shape, hz, nx, reach, ok, x, y, z)) return;
            atomicAdd(ctr + 0, 1ull);
            if (x >= 1 && x <= nx - 2 && y >= 1 && y <= ny - 2 && (!hz || (z >= 1 && z <= nz - 2))) atomicAdd(ctr + 1, 1ull);
        });
        unsigned long long h[2] = {0, 0};
        FS_CUDA(cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaStreamSynchronize(st));
        cudaFree(inside);
        *total = (long long)h[0];
        *interior = (long long)h[1];
        return !bad;
    }
    void metrics(const FsGrid &g, const float *d, const float *ux, const float *uy, const float *uz, double *sum, float *mx) {
        FS_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), st));
        FS_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned int), st));
        const long long n = g.sz * (g.ke - g.kb);
        const long long off = g.sz * g.kb;
        double *ds = d_sum;
        unsigned int *dm = d_max;
        const int hz = g.hz;
        const int blocks = sm_count * 8;
        const long long per = (n + (long long)blocks * 256 - 1) / ((long long)blocks * 256);
        metrics_kernel<<<blocks, 256, 0, st>>>(d + off, ux + off, uy + off, hz ? uz + off : nullptr, n, per, ds, dm);
        launches++;
        unsigned int bits = 0;
        FS_CUDA(cudaMemcpyAsync(sum, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaMemcpyAsync(&bits, d_max, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaStreamSynchronize(st));
        memcpy(mx, &bits, sizeof(float));
    }

    unsigned long long division_selftest(float c, unsigned long long first, unsigned long long count) {
        unsigned long long *d = (unsigned long long *)get_scratch(sizeof(unsigned long long));
        unsigned long long h = 0;
        FS_CUDA(cudaMemsetAsync(d, 0, sizeof(h), st));
        division_selftest_kernel<<<sm_count * 8, 256, 0, st>>>(c, first, count, d);
        launches++;
        FS_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaStreamSynchronize(st));
        return h;
    }

    // ---- halo exchange (z-slabs) ---------------------------------------------------------------------
    // See fs_kernels.cuh ("z-slab halo exchange over peer memory") for the protocol.
    struct HaloBlob {              // what fs_halo_export hands to the neighbours (<= FS_IPC_BLOB_BYTES)
        uint32_t magic;
        int32_t rank, world, dev, nzl, kb, ke, zoff, nbuf;
        uint64_t pid;
        cudaIpcMemHandle_t field[11];
        cudaIpcMemHandle_t flags;
        void *raw_field[11];
        void *raw_flags;
    };
    struct Peer {
        bool present = false, ipc = false;
        float *base[11] = {};
        unsigned *flags = nullptr;
        int nzl = 0, kb = 0, ke = 0, zoff = 0, dev = -1;
    };
    bool halo_on = false;
    // FS_HALO_TRACE=<path prefix>: every halo operation records four %globaltimer stamps (start, previous op of the
    // neighbours seen, planes stored, neighbours' planes landed); halo_close() writes <prefix>.rank<r>.csv
    static const unsigned kHaloTraceCap = 1u << 16;
    unsigned long long *halo_trace = nullptr;
    std::string halo_trace_path;
    int halo_rank = 0;
    Peer lo, hi;
    unsigned *my_flags = nullptr;       // FS_HF_WORDS words, device
    std::vector<float *> bufs;          // this slab's field allocations, same order on every rank
    unsigned ops_since_commit = 0;      // halo ops enqueued since the last halo_commit

    int buf_index(const float *p) const {
        for (size_t i = 0; i < bufs.size(); i++)
            if (bufs[i] == p) return (int)i;
        return -1;
    }
    FsHaloArgs halo_args(const FsGrid &g, float *const *fields, int nf, unsigned op_offset) const {
        FsHaloArgs h{};
        if (!halo_on) return h;
        h.my_flags = my_flags;
        h.op_offset = op_offset;
        h.nf = nf;
        h.trace = halo_trace;
        h.trace_cap = kHaloTraceCap;
        if (lo.present) h.lo_flags = lo.flags;
        if (hi.present) h.hi_flags = hi.flags;
        for (int f = 0; f < nf; f++) {
            const int bi = buf_index(fields[f]);
            // FS_GHOST planes each way: my lowest owned planes -> the lower neighbour's top ghosts, my highest -> the upper's bottom ghosts
            h.lo_src[f] = fields[f] + g.sz * g.kb;
            h.hi_src[f] = fields[f] + g.sz * (g.ke - FS_GHOST);
            if (lo.present && bi >= 0) h.lo_plane[f] = lo.base[bi] + g.sz * (lo.nzl - FS_GHOST);
            if (hi.present && bi >= 0) h.hi_plane[f] = hi.base[bi];
        }
        return h;
    }
    // Every halo_push_kernel retires only after the neighbours' planes of the same op have landed, so kernels
    // ordered after it can read the ghost planes without any further wait.
    void halo(const FsGrid &g, float *field) { halo_on_stream(g, field, st); }
    void halo_on_stream(const FsGrid &g, float *field, cudaStream_t stream) {
        float *fields[1] = {field};
        halo_n_on_stream(g, fields, field ? 1 : 0, stream);
    }
    void halo_n(const FsGrid &g, float *const *fields, int nf) { halo_n_on_stream(g, fields, nf, st); }
    void halo_n_on_stream(const FsGrid &g, float *const *fields, int nf, cudaStream_t stream) {
        if (!halo_on) return;
        const unsigned op = ++ops_since_commit;
        FsHaloArgs h = halo_args(g, fields, nf, op);
        h.need_ack = ack_since_last_op ? 1 : 0;
        ack_since_last_op = false;
        const long long plane = g.sz * FS_GHOST;
        int blocks = (int)((plane / 4 + 255) / 256) * (nf > 1 ? nf : 1);
        if (blocks > sm_count * 2) blocks = sm_count * 2;
        if (blocks < 1 || nf == 0) blocks = 1;
        launch_on(stream, halo_push_kernel, dim3(blocks), dim3(256), h, nf ? plane : 0LL);
    }
    // After an FS_X_EXTEND sweep (which read ghost planes WITHOUT a halo operation of its own): tell the neighbours that the
    // ghost planes their last push filled have been consumed.  Their next push -- the first operation that overwrites
    // those planes -- waits for this (FsHaloArgs::need_ack), see the protocol notes in fs_kernels.cuh.
    bool ack_since_last_op = false;
    void halo_ack_on_stream(cudaStream_t stream) {
        if (!halo_on) return;
        launch_on(stream, halo_ack_kernel, dim3(1), dim3(32), my_flags, lo.present ? lo.flags : nullptr,
                  hi.present ? hi.flags : nullptr, ops_since_commit);
        ack_since_last_op = true;
    }
    void halo_fence() { // neighbours have finished everything enqueued before this point, and vice versa
        if (!halo_on) return;
        halo(FsGrid{}, nullptr);
    }
    void halo_commit() {
        if (!halo_on || !ops_since_commit) return;
        halo_commit_kernel<<<1, 1, 0, st>>>(my_flags, ops_since_commit);
        launches++;
        ops_since_commit = 0;
    }
    template <class Core>
    int halo_export(Core &c, void *blob) {
        static_assert(sizeof(HaloBlob) <= FS_IPC_BLOB_BYTES, "blob too large");
        if (c.prm.slab_count < 2) { msg = "halo export needs slab_count >= 2"; return FS_ERR_BAD_ARGUMENT; }
        if (!my_flags) {
            my_flags = (unsigned *)alloc(sizeof(unsigned) * FS_HF_WORDS);
            if (!my_flags) return FS_ERR_OUT_OF_MEMORY;
            FS_CUDA(cudaMemset(my_flags, 0, sizeof(unsigned) * FS_HF_WORDS));
        }
        bufs = c.allocated;
        HaloBlob b{};
        b.magic = 0x46534831u;
        b.rank = c.prm.slab_rank; b.world = c.prm.slab_count; b.dev = dev;
        b.nzl = c.g.nzl; b.kb = c.g.kb; b.ke = c.g.ke; b.zoff = c.g.zoff; b.nbuf = (int)bufs.size();
        b.pid = (uint64_t)getpid();
        for (size_t i = 0; i < bufs.size(); i++) {
            FS_CUDA(cudaIpcGetMemHandle(&b.field[i], bufs[i]));
            b.raw_field[i] = bufs[i];
        }
        FS_CUDA(cudaIpcGetMemHandle(&b.flags, my_flags));
        b.raw_flags = my_flags;
        memcpy(blob, &b, sizeof(b));
        return bad ? FS_ERR_CUDA : FS_OK;
    }
    int connect_one(Peer &p, const void *blob, int expect_rank, int same_process) {
        HaloBlob b;
        memcpy(&b, blob, sizeof(b));
        if (b.magic != 0x46534831u || b.rank != expect_rank || b.nbuf != (int)bufs.size()) { msg = "halo blob does not match the expected neighbour"; return FS_ERR_BAD_ARGUMENT; }
        p.nzl = b.nzl; p.kb = b.kb; p.ke = b.ke; p.zoff = b.zoff; p.dev = b.dev;
        if (same_process) {
            if (b.dev != dev) {
                int can = 0;
                FS_CUDA(cudaDeviceCanAccessPeer(&can, dev, b.dev));
                if (!can) { msg = "no peer access between the two devices"; return FS_ERR_COMM; }
                cudaError_t e = cudaDeviceEnablePeerAccess(b.dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess", e); return FS_ERR_COMM; }
                cudaGetLastError();
            }
            for (int i = 0; i < b.nbuf; i++) p.base[i] = (float *)b.raw_field[i];
            p.flags = (unsigned *)b.raw_flags;
        } else {
            for (int i = 0; i < b.nbuf; i++) {
                void *ptr = nullptr;
                cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.field[i], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(field)", e); return FS_ERR_COMM; }
                p.base[i] = (float *)ptr;
            }
            void *ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.flags, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(flags)", e); return FS_ERR_COMM; }
            p.flags = (unsigned *)ptr;
            p.ipc = true;
        }
        p.present = true;
        return FS_OK;
    }
    template <class Core>
    int halo_connect(Core &c, const void *lower_blob, const void *upper_blob, int same_process) {
        if (!my_flags) { msg = "call fs_halo_export first"; return FS_ERR_BAD_ARGUMENT; }
        const int r = c.prm.slab_rank, P = c.prm.slab_count;
        if ((r > 0) != (lower_blob != nullptr) || (r < P - 1) != (upper_blob != nullptr)) { msg = "a blob is needed for exactly the existing neighbours"; return FS_ERR_BAD_ARGUMENT; }
        int rc = FS_OK;
        if (lower_blob) rc = connect_one(lo, lower_blob, r - 1, same_process);
        if (rc == FS_OK && upper_blob) rc = connect_one(hi, upper_blob, r + 1, same_process);
        if (rc != FS_OK) return rc;
        halo_on = true;
        halo_rank = r;
        if (const char *tp = getenv("FS_HALO_TRACE")) {
            halo_trace_path = tp;
            halo_trace = (unsigned long long *)alloc(sizeof(unsigned long long) * 4 * kHaloTraceCap);
            if (halo_trace) FS_CUDA(cudaMemset(halo_trace, 0, sizeof(unsigned long long) * 4 * kHaloTraceCap));
        }
        invalidate_graph();
        return FS_OK;
    }
    void halo_trace_dump() {
        if (!halo_trace) return;
        std::vector<unsigned long long> h(4 * (size_t)kHaloTraceCap);
        if (cudaMemcpy(h.data(), halo_trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const std::string path = halo_trace_path + ".rank" + std::to_string(halo_rank) + ".csv";
            if (FILE *f = fopen(path.c_str(), "w")) {
                fprintf(f, "slot,start_ns,prev_seen_ns,stored_ns,landed_ns\n");
                for (unsigned i = 0; i < kHaloTraceCap; i++)
                    if (h[4 * i]) fprintf(f, "%u,%llu,%llu,%llu,%llu\n", i, h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                fclose(f);
            }
        }
        cudaFree(halo_trace);
        halo_trace = nullptr;
    }
    void halo_close() {
        halo_trace_dump();
        for (Peer *p : {&lo, &hi}) {
            if (p->present && p->ipc) {
                for (float *b : p->base) if (b) cudaIpcCloseMemHandle(b);
                if (p->flags) cudaIpcCloseMemHandle(p->flags);
            }
            *p = Peer{};
        }
        if (my_flags) cudaFree(my_flags);
        my_flags = nullptr;
        halo_on = false;
    }

    // ---- CUDA graph capture / replay of one step --------------------------------------------------------
    void invalidate_graph() {
        for (auto &e : graphs) cudaGraphExecDestroy(e.exec);
        graphs.clear();
    }
    bool replay_step(float dt, float visc, float diff, float *const before[11]) {
        if (!use_graph) return false;
        for (auto &e : graphs) {
            if (e.dt == dt && e.visc == visc && e.diff == diff && memcmp(e.before, before, sizeof(e.before)) == 0) {
                FS_CUDA(cudaGraphLaunch(e.exec, st));
                launches += e.launches;
                replayed = &e;
                return true;
            }
        }
        return false;
    }
    GraphEntry *replayed = nullptr;
    void roles_after_replay(float **roles[11]) {
        for (int i = 0; i < 11; i++) *roles[i] = replayed->after[i];
    }
    void begin_step(float dt, float visc, float diff, float *const before[11]) {
        if (!use_graph) return;
        pending.dt = dt; pending.visc = visc; pending.diff = diff;
        memcpy(pending.before, before, sizeof(pending.before));
        launches_at_begin = launches;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) { set_error("cudaStreamBeginCapture", e); return; }
        capturing = true;
    }
    void end_step(float *const after[11]) {
        if (!capturing) return;
        capturing = false;
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (e != cudaSuccess) { set_error("cudaStreamEndCapture", e); return; }
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { set_error("cudaGraphInstantiate", e); return; }
        memcpy(pending.after, after, sizeof(pending.after));
        pending.exec = exec;
        pending.launches = launches - launches_at_begin;
        if (graphs.size() >= 16) invalidate_graph();
        graphs.push_back(pending);
        FS_CUDA(cudaGraphLaunch(exec, st)); // the captured work has not run yet
    }

    // ---- timing ------------------------------------------------------------------------------------------
    void timer_start() { FS_CUDA(cudaEventRecord(ev0, st)); }
    float timer_stop() {
        float ms = 0.f;
        FS_CUDA(cudaEventRecord(ev1, st));
        FS_CUDA(cudaEventSynchronize(ev1));
        FS_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        return ms;
    }
};

#define FS_EXEC CudaExec
#include "fs_abi.inl"
