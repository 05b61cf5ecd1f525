// fs_kernels.cuh -- sm_100a kernels of the stable-fluids step.
//
// Two families:
//  * relax_vec4: the hot sweep (pass-1 smoother / Jacobi; 320 of the 328 sweeps of a 512^3 K_p=80
//    step).  One thread owns a float4 of x for one row and marches in z, keeping the z-1/z/z+1
//    centre values in registers; y and x neighbours come through L1 (read-only path), rhs/flags are
//    streamed with L1::no_allocate.  set_bnd faces/edges/corners are written by the owning thread as
//    whole float4 rows ("ring scatter"), so no separate boundary pass exists.  HBM-bound: 13 B/voxel
//    (Jacobi) or 9 B/voxel (smoother); no tensor cores (no contraction anywhere on this path).
//  * cells_kernel<F>: one thread per interior cell running a fs_cellops.cuh functor.  Used for the
//    once-per-step kernels (divergence, gradient, advect, obstacle pass) and as the scalar fallback
//    of the sweeps when nx % 4 != 0.
#pragma once
#include <cuda_runtime.h>

#include "fs_cellops.cuh"

// ---- generic per-cell launch ------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) cells_kernel(FsGrid g, int kl0, F f) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = kl0 + blockIdx.z;
    if (i <= g.nx - 2 && j <= g.ny - 2) f(i, j, kl);
}

template <class F>
__global__ void __launch_bounds__(256) linear_kernel(long long n, F f) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) f(t);
}

// ---- vector loads/stores ---------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 ld4_stream(const float *p) { // read once: do not allocate in L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ld4_plain(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ uint32_t ld_flags4(const uint8_t *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st4(float *p, const float v[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// ---- correctly rounded division by a launch-constant divisor ------------------------------------------
// `x / c` compiles to div.rn.f32, whose FCHK guard sends zero / denormal numerators to a ~40 instruction
// slow path -- and a smoke-plume field is exactly zero, denormal or tiny almost everywhere (18% of the cells
// of a 512^3 plume are in 1e-45..1e-30 after 8 steps).  profiles/r01a_*: with IEEE `/` every division took
// the slow path and the sweep was issue bound at 41% of DRAM peak; a float FMA sequence with an IEEE
// fallback for tiny numerators (r01b) still lost 40% on real plume data.  The version below has no data
// dependent path at all and is exact for every finite numerator, denormal results included:
//   in double, with C = (double)c and Y = RN(1/C) computed once:
//     q = RN(n*Y); r = n - C*q (exact: FMA); q2 = RN(q + r*Y)          => q2 = (n/c)(1 + d), |d| <= 2^-105
//   result = (float)q2.
// Why the final rounding is always the IEEE one: a quotient of two 24-bit floats that is not itself a
// float rounding boundary (a 25-bit midpoint, or a midpoint of the denormal grid) is at least 2^-49
// (relative) away from the nearest one, far more than 2^-53 + 2^-105; and if it IS a boundary it is a double,
// so q2 equals it exactly and cvt.rn resolves the tie to even like div.rn does.  NaN propagates.  The one
// deviation: a +-inf numerator (a field that has already blown up) gives NaN instead of +-inf.  A divisor
// that is zero, infinite or NaN never reaches this code (the host launches the per-cell kernel instead).
// fs_selftest_division() checks fs_div against __fdiv_rn bit for bit over any range of numerator bit
// patterns (the GPU tests run all 2^32 of them for the divisors the step uses).
struct FsDivisor {
    double c, rc;
    float cf;
    int safe;
};
__device__ __forceinline__ FsDivisor fs_make_divisor(float c) {
    FsDivisor d;
    d.cf = c;
    d.c = (double)c;
    d.rc = 1.0 / (double)c;
    d.safe = (c != 0.0f && fabsf(c) <= 3.402823466e38f) ? 1 : 0; // finite and non-zero (NaN fails the compare)
    return d;
}
__device__ __forceinline__ float fs_div(float n, const FsDivisor &d) {
    const double nd = (double)n;
    const double q = __dmul_rn(nd, d.rc);
    // residual as -(C*q - n) rather than (n - C*q): the same value, but for n = -0 it yields -0, so that the
    // quotient keeps the sign IEEE division gives a zero numerator
    const double t = __fma_rn(d.c, q, -nd);
    const double q2 = __fma_rn(-t, d.rc, q);
    return (float)q2;
}

__global__ void __launch_bounds__(256)
division_selftest_kernel(float c, unsigned long long first, unsigned long long count, unsigned long long *mismatches) {
    const FsDivisor d = fs_make_divisor(c);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const float n = __uint_as_float((unsigned)(first + t));
        const float want = __fdiv_rn(n, c);
        const float got = d.safe ? fs_div(n, d) : want;
        const bool both_nan = (want != want) && (got != got);
        const bool inf_in = fabsf(n) > 3.402823466e38f; // documented deviation: inf numerator -> NaN
        if (!both_nan && !inf_in && __float_as_uint(want) != __float_as_uint(got)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ---- z-slab halo exchange over peer memory (NVLink P2P) ----------------------------------------------
// One process (or handle) per GPU.  A slab's neighbours map its field buffers and its flag block (CUDA IPC
// or plain peer access) and STORE boundary planes straight into its ghost planes; there is no receive
// side copy and no collective.  Every halo-producing operation has a sequence number
//     seq = flags[FS_HF_BASE] + op_offset
// (FS_HF_BASE lives in device memory and is advanced by halo_commit_kernel, so a CUDA graph replays with
// fresh numbers).  Protocol of operation `seq` on every rank ("strict lock step"):
//   wait   flags[FROM_LO] >= seq-1 and flags[FROM_HI] >= seq-1   (the neighbours' stores of op seq-1 have landed here;
//          redundant after a push of op seq-1, which already blocked on the same words -- it only matters for the
//          first op after fs_halo_connect)
//   store  my lowest owned planes -> lower neighbour's top ghost planes, my highest -> upper's bottom ghost planes
//   signal __threadfence_system(); neighbour.flags[FROM_HI or FROM_LO] = seq
//   wait   flags[FROM_*] >= seq   (the neighbours' planes of THIS op have landed: the kernel does not retire earlier)
// What the signal does NOT say: that the neighbour has finished READING the ghost planes a later push overwrites.
// Write-after-read safety of the ghosts comes from the solver's buffer discipline instead, and every new halo
// producing op has to respect it:
//   * ping-pong sweeps (smoother / Jacobi / fused pair): op seq+1 writes the ghosts of the OTHER buffer; the ghosts of
//     a buffer are rewritten two ops later, and the neighbour cannot have signalled op seq+1 before the kernels of
//     op seq that read them were complete (its push of seq+1 is stream-ordered after them);
//   * in-place kernels (red-black colour passes, mirror, gradient, obstacle pass) re-push planes whose ghost copies
//     are only read by kernels ordered BEFORE the neighbour's previous signal, or carry identical values.
// An in-place op followed by a halo of a field whose ghost a neighbour kernel may still be reading would race; such
// an op needs a halo_fence() (consumer-done) first.
//   * extended sweeps (FS_X_EXTEND: two sweeps per operation -- the first also computes the first ghost plane from the
//     two-deep ghost zone, the second exchanges) read ghost planes in a kernel that belongs to NO operation, and the very
//     next operation overwrites them (same ping-pong buffer).  Here the discipline above does not hold, so the consumer
//     says so explicitly: halo_ack_kernel stores seq into the neighbours' ACK words once the reading launch is complete,
//     and the next push (FsHaloArgs::need_ack) also waits for ACK >= seq-1 before it stores.
// Schedule of the relaxation sweeps on a slab (CudaExec::relax_n): side stream = the two ends of the slab (a two-chunk
// launch of relax_vec4) -> halo_push_kernel (wait seq-1, store, signal seq, wait for the incoming seq); main stream = the
// planes between, concurrently; join.  With extended sweeps the join is deferred: only the ENDS of the extended sweep
// (side stream, after the push) read ghost planes, its middle starts as soon as the previous sweep's ends are done, so
// one exchange hides behind two interior launches.  (Fusing wait / store / signal INTO relax_vec4 was built twice --
// round 1 inlined, round 2 as a separate instantiation for the ends -- and measured slower both times: register
// pressure in the 64-register sweep, profiles/r01d_halo_variants.md, profiles/r02e_exchange.md.)
enum { FS_HF_FROM_LO = 0, FS_HF_FROM_HI = 1, FS_HF_BASE = 2, FS_HF_CNT_LO = 3, FS_HF_ENDS = 4, FS_HF_ERROR = 5, FS_HF_ACK_FROM_LO = 6, FS_HF_ACK_FROM_HI = 7, FS_HF_WORDS = 8 };

#define FS_BATCH 3 // fields per batched launch / halo operation (the velocity components)
struct FsHaloArgs {
    float *lo_plane[FS_BATCH];       // lower neighbour's top ghost planes of each field (nullptr: no neighbour / fence)
    float *hi_plane[FS_BATCH];       // upper neighbour's bottom ghost planes
    const float *lo_src[FS_BATCH];   // this slab's lowest / highest owned planes of each field
    const float *hi_src[FS_BATCH];
    int nf;                          // number of fields (0: pure fence)
    unsigned *my_flags;              // this slab's flag block
    unsigned *lo_flags;              // neighbours' flag blocks (peer memory)
    unsigned *hi_flags;
    unsigned op_offset;
    int need_ack;                    // an extended sweep read the ghost planes since the last operation: wait for the neighbours' acknowledgement too
    unsigned long long *trace;       // optional (FS_HALO_TRACE): 4 %globaltimer stamps per operation, ring of trace_cap entries
    unsigned trace_cap;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long fs_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Bounded: a neighbour that died or never launched must not leave this GPU spinning for ever.  After
// FS_HALO_TIMEOUT_NS the waiter records FS_HF_ERROR = 2 in its own flag block and carries on (the fields are then
// wrong; fs_sync / fs_get_* report the error); later waits of the same slab give up at once.
#ifndef FS_HALO_TIMEOUT_NS
#define FS_HALO_TIMEOUT_NS 30000000000ull
#endif
__device__ __forceinline__ void halo_spin_until(const unsigned *flag, unsigned target, unsigned *err_word) {
    if ((int)(ld_acquire_sys(flag) - target) >= 0) return;
    if (*(volatile unsigned *)err_word == 2u) return; // an earlier wait has already timed out: do not add 30 s per wait
    const unsigned long long t0 = fs_globaltimer_ns();
    while ((int)(ld_acquire_sys(flag) - target) < 0) {
        __nanosleep(64);
        if (fs_globaltimer_ns() - t0 > FS_HALO_TIMEOUT_NS) {
            *(volatile unsigned *)err_word = 2u;
            return;
        }
    }
}
__device__ __forceinline__ unsigned halo_seq(const FsHaloArgs &h) {
    return *(volatile const unsigned *)(h.my_flags + FS_HF_BASE) + h.op_offset;
}

// One CTA's share of halo operation h: copies the FS_GHOST boundary planes of up to FS_BATCH fields into the neighbours'
// ghost planes (nf = 0: pure fence) after waiting for seq-1, then signals seq and (last CTA to finish) waits for the
// neighbours' seq.  plane_elems = FS_GHOST*nx*ny.  cta / ncta: this CTA's index among the CTAs working on the operation;
// tid / nthr: linear thread index / CTA size.  ends_target > 0: the planes are being produced by other CTAs of the SAME
// launch (relax_vec4<.., XCHG>), which count themselves in FS_HF_ENDS when their stores are done -- wait for that many.
__device__ __forceinline__ void halo_push_cta(const FsHaloArgs &h, const long long plane_elems, const unsigned cta, const unsigned ncta,
                                           const unsigned tid, const unsigned nthr, const unsigned ends_target) {
    __shared__ unsigned s_seq;
    if (tid == 0) {
        const unsigned seq = halo_seq(h);
        unsigned long long *tr = (h.trace && cta == 0) ? h.trace + 4ull * (seq % h.trace_cap) : nullptr;
        if (tr) tr[0] = fs_globaltimer_ns();             // start
        if (h.lo_flags) halo_spin_until(h.my_flags + FS_HF_FROM_LO, seq - 1, h.my_flags + FS_HF_ERROR);
        if (h.hi_flags) halo_spin_until(h.my_flags + FS_HF_FROM_HI, seq - 1, h.my_flags + FS_HF_ERROR);
        if (h.need_ack) { // ... and they have finished reading the ghost planes that operation filled (extended sweeps)
            if (h.lo_flags) halo_spin_until(h.my_flags + FS_HF_ACK_FROM_LO, seq - 1, h.my_flags + FS_HF_ERROR);
            if (h.hi_flags) halo_spin_until(h.my_flags + FS_HF_ACK_FROM_HI, seq - 1, h.my_flags + FS_HF_ERROR);
        }
        if (ends_target) halo_spin_until(h.my_flags + FS_HF_ENDS, ends_target, h.my_flags + FS_HF_ERROR);
        if (tr) tr[1] = fs_globaltimer_ns();             // neighbours' previous operation seen (and my planes complete)
        s_seq = seq;
    }
    __syncthreads();
    const long long n4 = plane_elems / 4, stride = (long long)ncta * nthr;
#pragma unroll
    for (int f = 0; f < FS_BATCH; f++) {
        if (f >= h.nf) break;
        float *lo_dst = h.lo_plane[f], *hi_dst = h.hi_plane[f];
        const float *lo_src = h.lo_src[f], *hi_src = h.hi_src[f];
        // (ld.cg: the planes may have been written by other SMs during this very launch)
        for (long long t = (long long)cta * nthr + tid; t < n4; t += stride) {
            if (lo_dst) reinterpret_cast<float4 *>(lo_dst)[t] = __ldcg(reinterpret_cast<const float4 *>(lo_src) + t);
            if (hi_dst) reinterpret_cast<float4 *>(hi_dst)[t] = __ldcg(reinterpret_cast<const float4 *>(hi_src) + t);
        }
        for (long long t = n4 * 4 + (long long)cta * nthr + tid; t < plane_elems; t += stride) {
            if (lo_dst) lo_dst[t] = __ldcg(lo_src + t);
            if (hi_dst) hi_dst[t] = __ldcg(hi_src + t);
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        if (atomicAdd(h.my_flags + FS_HF_CNT_LO, 1u) == ncta - 1) {
            h.my_flags[FS_HF_CNT_LO] = 0;
            if (ends_target) h.my_flags[FS_HF_ENDS] = 0;  // every CTA of the operation has seen the count: re-arm it
            __threadfence_system();
            unsigned long long *tr = h.trace ? h.trace + 4ull * (s_seq % h.trace_cap) : nullptr;
            if (tr) tr[2] = fs_globaltimer_ns();         // all planes stored
            if (h.lo_flags) st_release_sys(h.lo_flags + FS_HF_FROM_HI, s_seq);
            if (h.hi_flags) st_release_sys(h.hi_flags + FS_HF_FROM_LO, s_seq);
            // ... and do not retire before the neighbours' planes of the same op have landed here: whatever
            // is ordered after this kernel may read the ghost planes (no separate wait launch needed)
            if (h.lo_flags) halo_spin_until(h.my_flags + FS_HF_FROM_LO, s_seq, h.my_flags + FS_HF_ERROR);
            if (h.hi_flags) halo_spin_until(h.my_flags + FS_HF_FROM_HI, s_seq, h.my_flags + FS_HF_ERROR);
            if (tr) tr[3] = fs_globaltimer_ns();         // neighbours' planes of this operation have landed
        }
    }
}
__global__ void __launch_bounds__(256)
halo_push_kernel(const FsHaloArgs h, long long plane_elems) {
    halo_push_cta(h, plane_elems, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x, 0u);
}

__global__ void halo_commit_kernel(unsigned *flags, unsigned ops) { flags[FS_HF_BASE] += ops; }

// Published after an extended sweep (one that read ghost planes without a halo operation of its own): "I have consumed
// the ghost planes your operation base + op_offset filled".  The neighbours' next push waits for it.
__global__ void halo_ack_kernel(const unsigned *my_flags, unsigned *lo_flags, unsigned *hi_flags, const unsigned op_offset) {
    if (threadIdx.x == 0) {
        const unsigned seq = *(volatile const unsigned *)(my_flags + FS_HF_BASE) + op_offset;
        if (lo_flags) st_release_sys(lo_flags + FS_HF_ACK_FROM_HI, seq);
        if (hi_flags) st_release_sys(hi_flags + FS_HF_ACK_FROM_LO, seq);
    }
}

// Gather source for the semi-Lagrangian back-trace: a field as seen from one slab -- its own planes
// (ghosts included) plus the two neighbour slabs' copies through peer memory (NVLink loads).  A back-trace
// that leaves even the neighbour slabs sets FS_HF_ERROR (reported by fs_sync / fs_get_field).
struct FsSlabView {
    const float *loc, *lo, *hi;
    int zoff, nzl, lo_zoff, lo_nzl, hi_zoff, hi_nzl;
    unsigned *err;
};
__device__ __forceinline__ const float *fs_slab_plane(const FsSlabView &v, const FsGrid &g, int kk) {
    int kl = kk - v.zoff;
    if (kl >= 0 && kl < v.nzl) return v.loc + kl * g.sz;
    const float *base = kl < 0 ? v.lo : v.hi;
    const int pl = kl < 0 ? kk - v.lo_zoff : kk - v.hi_zoff;
    const int pn = kl < 0 ? v.lo_nzl : v.hi_nzl;
    if (!base || pl < 0 || pl >= pn) { // the back-trace left even the neighbour slab: flag it, read something valid
        if (v.err) *v.err = 1u;
        return v.loc;
    }
    return base + pl * g.sz;
}

// ---- the hot sweep ---------------------------------------------------------------------------------
// Requirements: nx % 4 == 0 (so every row start is 16-byte aligned in a cudaMalloc'd array).
// Latency hiding: prefetch.global.L2 of the planes l2_ahead iterations ahead (holds no registers; 297 -> 246 us at
// 512^3).  A register prefetch of the next plane was measured and rejected (78 registers -> 3 CTAs/SM, 336 us).
// Grid: x = ceil(nx/4 / blockDim.x), y = ceil((ny-2) / blockDim.y), z = number of z chunks.
// kl_begin/kl_end: owned interior local planes [kl_begin, kl_end); each block marches zchunk of them.
// Up to FS_BATCH fields per launch (the velocity components of one diffusion sweep share a, c and the flags): blockIdx.z
// enumerates (z chunk, field), each CTA works on one field, so batching costs the inner loop nothing.
struct FsRelaxBatch {
    const float *in[FS_BATCH], *rhs[FS_BATCH], *stale[FS_BATCH];
    float *out[FS_BATCH];
    int b[FS_BATCH];
    int nf;
};
// Coarse obstacle map: for every tile of FS_TILE_X x FS_TILE_Y cells (interior rows counted from j = 1) the running count
// along z of planes in which the tile holds an obstacle cell: cum[kl][ty][tx] = number of such planes below local plane
// kl (nzl + 1 entries per tile).  A sweep CTA whose tile is obstacle free over all its planes (two loads, one compare
// before the z loop) skips the flag stream altogether: 1 of the 13 B/voxel of a Jacobi sweep, 1 of 9 for the smoother.
#define FS_TILE_X 128
#define FS_TILE_Y 8
struct FsTileMap {
    const int *cum; // nullptr: no map, always stream the flags
    int tx_count, ty_count;
};
// step 1: cum[kl + 1][tile] = 1 if the tile holds an obstacle cell in local plane kl (one thread per tile and plane)
__global__ void __launch_bounds__(256)
tilemap_mark_kernel(const FsGrid g, const uint8_t *__restrict__ mask, int *cum, int tx_count, int ty_count) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ntiles = (long long)tx_count * ty_count;
    if (t >= ntiles * g.nzl) return;
    const int kl = (int)(t / ntiles), tile = (int)(t % ntiles);
    const int tx = tile % tx_count, ty = tile / tx_count;
    const int x_lo = tx * FS_TILE_X, x_hi = min(x_lo + FS_TILE_X, g.nx);
    const int j_lo = 1 + ty * FS_TILE_Y, j_hi = min(j_lo + FS_TILE_Y, g.ny - 1);
    int any = 0;
    for (int j = j_lo; j < j_hi && !any; j++) {
        const uint32_t *row = reinterpret_cast<const uint32_t *>(mask + fs_idx(g, 0, j, kl)); // nx % 4 == 0 here
        for (int x = x_lo; x < x_hi; x += 4)
            if (row[x >> 2]) { any = 1; break; }
    }
    cum[(kl + 1) * ntiles + tile] = any;
}
// step 2: running count along z, in place (one thread per tile)
__global__ void __launch_bounds__(256)
tilemap_scan_kernel(int *cum, int ntiles, int nzl) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    int count = 0;
    cum[t] = 0;
    for (int kl = 1; kl <= nzl; kl++) {
        count += cum[(long long)kl * ntiles + t];
        cum[(long long)kl * ntiles + t] = count;
    }
}
// XCHG = true: the sweep of a z-slab that also carries its halo operation, as ONE launch.  Block order (blockIdx.z is the
// slowest index, CTAs are dispatched in linear order): first the two ENDS of the slab (x.ends planes each, every field),
// then one z-slice whose first x.push_ctas CTAs are not sweep CTAs at all but run halo_push_cta -- they wait until the ends
// CTAs have counted themselves done (FS_HF_ENDS), store the boundary planes into the neighbours' ghost planes, signal, and
// the last of them waits for the neighbours' planes -- then the MIDDLE chunks, which hide all of that.  Because the push
// CTAs are dispatched BEFORE the middle CTAs they hold their SM slots from the start; a separate push kernel only gets
// slots when the middle launch (64 registers x 1024 threads = the whole register file of every SM) has no CTA left to
// dispatch, which serialised the exchange behind the sweep (profiles/r02e_exchange.md).  The hot loop is untouched: the
// ends CTAs add one atomic after it, the push CTAs leave before it.
struct FsSweepXchg {
    FsHaloArgs h;
    long long plane_elems;   // FS_GHOST * nx * ny
    int ends;                // planes per end
    unsigned push_ctas;      // CTAs of the push slice that work on the operation
    unsigned ends_ctas;      // CTAs of the two ends (all fields): the count that completes FS_HF_ENDS
};
template <int MODE, bool HZ, bool XCHG>
__global__ void __launch_bounds__(256, 4)
relax_vec4(const FsGrid g, const FsRelaxBatch batch, const uint8_t *__restrict__ flags, const FsTileMap tiles, const float a,
           const float c, const int in_zero, const int kl_begin, const int kl_end, const int zchunk, const int zc_base,
           const int zc_stride, const int l2_ahead, const int kl_alt, const FsSweepXchg x) {
    unsigned bz = blockIdx.z;
    if (XCHG) {
        const unsigned push_slice = 2u * (unsigned)batch.nf;
        if (bz == push_slice) {
            const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
            if (cta < x.push_ctas)
                halo_push_cta(x.h, x.plane_elems, cta, x.push_ctas, threadIdx.y * blockDim.x + threadIdx.x,
                              blockDim.x * blockDim.y, x.ends_ctas);
            return;
        }
        if (bz > push_slice) bz -= 1;
    }
    const int fld = batch.nf > 1 ? (int)(bz % (unsigned)batch.nf) : 0;
    const int zblk = batch.nf > 1 ? (int)(bz / (unsigned)batch.nf) : (int)bz;
    const float *__restrict__ in = fld == 0 ? batch.in[0] : (fld == 1 ? batch.in[1] : batch.in[2]);
    const float *__restrict__ rhs = fld == 0 ? batch.rhs[0] : (fld == 1 ? batch.rhs[1] : batch.rhs[2]);
    const float *stale = fld == 0 ? batch.stale[0] : (fld == 1 ? batch.stale[1] : batch.stale[2]);
    float *out = fld == 0 ? batch.out[0] : (fld == 1 ? batch.out[1] : batch.out[2]);
    const int b = fld == 0 ? batch.b[0] : (fld == 1 ? batch.b[1] : batch.b[2]);
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int x0 = gx * 4;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    // z chunk of this CTA: zc_base + blockIdx.z * zc_stride.  One launch (0, 1) covers every chunk; with z-slabs the
    // sweep is issued as two launches of this same kernel -- the two chunks that hold the slab's boundary planes
    // first (base 0, stride nchunks-1), then the interior chunks (base 1, stride 1) -- with the P2P halo push
    // between them, so the exchange overlaps the interior (fluidsolver.cu, CudaExec::relax).
    // (zc_stride == 0: a two-chunk launch for the two ends of a slab -- chunk 0 starts at kl_begin, chunk 1 at kl_alt)
    const int zc = zc_base + zblk * zc_stride;
    int k_lo = zc_stride == 0 ? (zblk == 0 ? kl_begin : kl_alt) : kl_begin + zc * zchunk;
    int k_hi = min(k_lo + zchunk, kl_end);
    if (XCHG) { // ends: [kl_begin, +ends) and [kl_alt = kl_end - ends, kl_end); middle chunks between
        k_lo = zblk == 0 ? kl_begin : (zblk == 1 ? kl_alt : kl_begin + x.ends + (zblk - 2) * zchunk);
        k_hi = zblk < 2 ? k_lo + x.ends : min(k_lo + zchunk, kl_alt);
    }
    const bool active = x0 < g.nx && j <= g.ny - 2 && k_lo < k_hi;
    if (active) {
    const FsDivisor dv = fs_make_divisor(c);
    const bool first_x = x0 == 0, last_x = x0 + 4 == g.nx;
    const int fxl[4] = {first_x ? 1 : 0, 0, 0, last_x ? 1 : 0};
    const float *stale_src = stale ? stale : out;
    // extra ring rows this thread owns in y (set_bnd faces/edges), -1 = none
    const int jr = j == 1 ? 0 : (j == g.ny - 2 ? g.ny - 1 : -1);
    const int jr2 = (j == 1 && j == g.ny - 2) ? g.ny - 1 : -1; // ny == 3: both

    const long long idx0 = fs_idx(g, x0, j, k_lo);
    // per-thread running pointers (one 64-bit add each per plane instead of re-deriving every address)
    const float *pin = in + idx0;
    const float *prh = rhs ? rhs + idx0 : nullptr;
    const uint8_t *pfl = flags ? flags + idx0 : nullptr;
    bool use_fl = flags != nullptr;                       // (pfl is advanced every plane, so it cannot be the test)
    if (use_fl && tiles.cum) { // this tile is obstacle free over all of this CTA's planes: no flag stream
        const long long tstep = (long long)tiles.ty_count * tiles.tx_count;
        const int *tc = tiles.cum + ((j - 1) / FS_TILE_Y) * tiles.tx_count + x0 / FS_TILE_X;
        use_fl = __ldg(tc + k_hi * tstep) != __ldg(tc + k_lo * tstep);
    }
    float *pout = out + idx0;
    const long long sy = g.sy, sz = g.sz;
    float4 prev = make_float4(0.f, 0.f, 0.f, 0.f), cur = prev, next = prev;
    if (!in_zero) {
        cur = ld4(pin);
        if (HZ) prev = ld4(pin - sz);
    }

    for (int kl = k_lo; kl < k_hi; kl++, pin += sz, prh += sz, pfl += sz, pout += sz) {
        float4 up = make_float4(0.f, 0.f, 0.f, 0.f), dn = up;
        float left = 0.f, right = 0.f;
        if (!in_zero) {
            if (HZ) next = ld4(pin + sz);
            up = ld4(pin + sy);
            dn = ld4(pin - sy);
            if (!first_x) left = __ldg(pin - 1);
            if (!last_x) right = __ldg(pin + 4);
        }
        if (l2_ahead > 0 && kl + l2_ahead < k_hi) { // planes of iteration kl + l2_ahead (all inside this chunk's range)
            // start the DRAM->L2 transfers of a later iteration now; unlike a register prefetch this holds no registers
            if (!in_zero) prefetch_l2(pin + (long long)(l2_ahead + 1) * sz);
            if (MODE == FS_MODE_JACOBI) prefetch_l2(prh + (long long)l2_ahead * sz);
            if (use_fl) prefetch_l2(pfl + (long long)l2_ahead * sz);
        }
        float4 r4 = cur;
        if (MODE == FS_MODE_JACOBI) r4 = ld4_stream(prh);
        const uint32_t fl = use_fl ? ld_flags4(pfl) : 0u;

        const float cv[6] = {left, cur.x, cur.y, cur.z, cur.w, right};
        const float upv[4] = {up.x, up.y, up.z, up.w}, dnv[4] = {dn.x, dn.y, dn.z, dn.w};
        const float nxv[4] = {next.x, next.y, next.z, next.w}, pvv[4] = {prev.x, prev.y, prev.z, prev.w};
        const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
        float v[4];
#pragma unroll
        for (int l = 0; l < 4; l++) {
            float s = ((cv[l + 2] + cv[l]) + upv[l]) + dnv[l];
            if (HZ) s = (s + nxv[l]) + pvv[l];
            const float num = rv[l] + a * s;
            v[l] = fs_div(num, dv); // the host only launches this kernel for finite, non-zero c
        }
        if (fl & 0x01010101u) { // some lane is an obstacle cell (rare): copy / stale value instead
            float st[4] = {cv[1], cv[2], cv[3], cv[4]};
            if (MODE == FS_MODE_SMOOTH) {
                const float4 s4 = ld4_plain(stale_src + (pout - out));
                st[0] = s4.x; st[1] = s4.y; st[2] = s4.z; st[3] = s4.w;
            }
#pragma unroll
            for (int l = 0; l < 4; l++)
                if ((fl >> (8 * l)) & 1u) v[l] = st[l];
        }

        const int k = kl + g.zoff;
        const int kr = HZ ? (k == 1 ? kl - 1 : (k == g.nz - 2 ? kl + 1 : -1)) : -1;
        if (jr < 0 && kr < 0) { // no y/z ring row derives from this row: one store.  The x-face cells (x = 0, nx-1) are
            // lanes of the same float4 and are fixed up in place (fs_ring_value with fx only); sending every x-edge
            // WARP through the generic ring path below cost half the warps of a 512-wide grid ~100 issue slots per plane
            if (first_x) v[0] = b == 1 ? -v[1] : v[1];
            if (last_x) v[3] = b == 1 ? -v[2] : v[2];
            st4(pout, v);
        } else {
            // ring lanes take their nearest interior lane's value; fs_ring_value applies the face rules
            if (first_x) v[0] = v[1];
            if (last_x) v[3] = v[2];
            const int kr2 = (HZ && k == 1 && k == g.nz - 2) ? kl + 1 : -1; // nz == 3: both
            const int zs[3] = {kl, kr, kr2}, fzs[3] = {0, 1, 1};
            const int ys[3] = {j, jr, jr2}, fys[3] = {0, 1, 1};
#pragma unroll
            for (int zi = 0; zi < 3; zi++) {
                if (zi > 0 && zs[zi] < 0) continue;
#pragma unroll
                for (int yi = 0; yi < 3; yi++) {
                    if (yi > 0 && ys[yi] < 0) continue;
                    float o[4];
#pragma unroll
                    for (int l = 0; l < 4; l++) o[l] = fs_ring_value(v[l], fxl[l], fys[yi], fzs[zi], b);
                    float *po = out + fs_idx(g, x0, ys[yi], zs[zi]);
                    st4(po, o);
                }
            }
        }
        // BoundaryJob's obstacle pass, fused: a velocity component's obstacle cells take the mirrored mean of their
        // fluid neighbours' NEW values along axis b (recomputed, see fs_mirror_fused); the stores above wrote the
        // pre-mirror value, this thread's later store wins
        if ((fl & 0x01010101u) && b != 0) {
            if (b != 3 || HZ) {
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    if (!((fl >> (8 * l)) & 1u) || (l == 0 && first_x) || (l == 3 && last_x)) continue;
                    pout[l] = fs_mirror_fused<MODE>(g, in, rhs, (uint8_t)(fl >> (8 * l)), a, c, b, in_zero != 0, v[l], x0 + l, j, kl);
                }
            }
        }
        prev = cur;
        cur = next;
    }
    } // active
    if (XCHG && zblk < 2) { // an ends CTA: its planes are complete
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            __threadfence();
            atomicAdd(x.h.my_flags + FS_HF_ENDS, 1u);
        }
    }
}

// ---- float4 versions of the once-per-step stencils --------------------------------------------------------
// Same thread mapping as relax_vec4 (a float4 of x per thread, nx % 4 == 0), one plane per thread.
struct FsVec4Pos {
    int x0, j, kl, jr, jr2, kr, kr2;
    bool first_x, last_x;
};
__device__ __forceinline__ FsVec4Pos fs_vec4_pos(const FsGrid &g, int x0, int j, int kl) {
    FsVec4Pos p;
    p.x0 = x0; p.j = j; p.kl = kl;
    p.first_x = x0 == 0; p.last_x = x0 + 4 == g.nx;
    p.jr = j == 1 ? 0 : (j == g.ny - 2 ? g.ny - 1 : -1);
    p.jr2 = (j == 1 && j == g.ny - 2) ? g.ny - 1 : -1;
    const int k = kl + g.zoff;
    p.kr = g.hz ? (k == 1 ? kl - 1 : (k == g.nz - 2 ? kl + 1 : -1)) : -1;
    p.kr2 = (g.hz && k == 1 && k == g.nz - 2) ? kl + 1 : -1;
    return p;
}
// Stores the row and every set_bnd ring row / lane that derives from it (ring scatter, b = field kind).
__device__ __forceinline__ void fs_vec4_store_ring(float *out, const FsGrid &g, const FsVec4Pos &p, float v[4], int b) {
    if (p.jr < 0 && p.kr < 0) { // x-face lanes are fixed up in place, see relax_vec4
        if (p.first_x) v[0] = b == 1 ? -v[1] : v[1];
        if (p.last_x) v[3] = b == 1 ? -v[2] : v[2];
        st4(out + fs_idx(g, p.x0, p.j, p.kl), v);
        return;
    }
    if (p.first_x) v[0] = v[1];
    if (p.last_x) v[3] = v[2];
    const int fxl[4] = {p.first_x ? 1 : 0, 0, 0, p.last_x ? 1 : 0};
    const int zs[3] = {p.kl, p.kr, p.kr2}, ys[3] = {p.j, p.jr, p.jr2};
#pragma unroll
    for (int zi = 0; zi < 3; zi++) {
        if (zi > 0 && zs[zi] < 0) continue;
#pragma unroll
        for (int yi = 0; yi < 3; yi++) {
            if (yi > 0 && ys[yi] < 0) continue;
            float o[4];
#pragma unroll
            for (int l = 0; l < 4; l++) o[l] = fs_ring_value(v[l], fxl[l], yi > 0, zi > 0, b);
            st4(out + fs_idx(g, p.x0, ys[yi], zs[zi]), o);
        }
    }
}

// ---- fused two-stage sweep: two relaxation stages per pass over HBM ---------------------------------------------
// out = S2(S1(in)) in ONE launch, the intermediate field y1 = S1(in) never touching HBM:
//   FS_PAIR_JACOBI     S1 = S2 = LinearSolveIterationJob + BoundaryJob   (two Jacobi iterations, FluidSim.cs:1188-1233)
//   FS_PAIR_SMOOTH     S1 = S2 = DiffuseJob + BoundaryJob                (two pass-1 iterations, :1034-1069)
//   FS_PAIR_RED_BLACK  S1 = colour-0 half sweep, S2 = colour-1 half sweep + BoundaryJob (one red-black iteration, config 5)
// HBM traffic: in 4 + rhs 4 + flags 1 + out 4 = 13 B/voxel per TWO stages (26 B as two launches) plus the halo re-reads,
// which mostly hit L2 because the tiles that share them run in the same wave.
// Tile: a CTA owns W <= 32 float4 columns x (rows - 2) rows of OUTPUT and marches in z.  Stage 1 is evaluated on the
// tile plus a one-cell ring -- (W + 2) float4 columns (only one lane of each outer column is needed, the other three are
// free SIMD width) x rows -- by a FLATTENED thread index t -> (row = t / (W+2), column = t % (W+2)), so that no lane idles
// whatever nx is; the same thread then owns the same column in stage 2 (threads of the outer ring idle there).  Per
// z step: the global loads of stage 1 at plane k are issued (z-1/z/z+1 of `in` in registers, x/y neighbours through
// L1 as in relax_vec4); while they are in flight stage 2 runs at plane k-2 (y1's z neighbours from registers, its
// x/y neighbours from shared memory); then the stage-1 arithmetic, whose result goes to shared memory (three
// rotating slots: one barrier per step) and stays in registers as the z column of y1.
// Ring cells of the intermediate field: Jacobi / smoother stages end with set_bnd, so stage 2 must see
// y1(ring) = +-y1(nearest interior cell): x faces are fixed up inside the float4 before it is stored, y / z faces
// are substituted on the fly for the rows / planes next to them (edges and corners are never read by a stencil);
// a red-black sweep applies set_bnd only after both colours, so there the ring simply passes `in` through.
// Obstacle cells copy (requires b == 0 or no interior obstacle: no mirror pass between the stages -- the host
// checks).  The final set_bnd ring is written exactly as relax_vec4 does it.  Bit exact against two single sweeps.
#define FS_PAIR_W 32
#define FS_PAIR_MAX_ROWS 16
#define FS_PAIR_MAX_THREADS ((FS_PAIR_W + 2) * FS_PAIR_MAX_ROWS)

template <int KIND>
__global__ void __launch_bounds__(FS_PAIR_MAX_THREADS, 1)
relax_pair_kernel(const FsGrid g, const float *__restrict__ in, const float *__restrict__ rhs, float *out,
                  const uint8_t *__restrict__ flags, const float a, const float c, const int b, const int in_zero,
                  const int kl_begin, const int kl_end, const int zchunk, const int zc_base, const int zc_stride,
                  const int l2_ahead, const int wcols, const int rows) {
    __shared__ float4 s_y1[3][FS_PAIR_MAX_THREADS];
    const int cols = wcols + 2;
    const int t = threadIdx.x;
    const int r = t / cols, cidx = t - r * cols;
    const int ncols = g.nx >> 2;
    const int gx = (int)blockIdx.x * wcols + cidx - 1;
    const int j = (int)blockIdx.y * (rows - 2) + r;
    const bool live = r < rows && gx >= 0 && gx < ncols && j <= g.ny - 1;   // owns a real float4 column
    const bool row_in = j >= 1 && j <= g.ny - 2;
    const bool inner = live && row_in && cidx >= 1 && cidx <= wcols && r >= 1 && r <= rows - 2; // stage-2 owner
    const int zc = zc_base + (int)blockIdx.z * zc_stride;
    const int k_lo = kl_begin + zc * zchunk;            // stage 2 (output) planes [k_lo, k_hi), local indices
    const int k_hi = min(k_lo + zchunk, kl_end);
    if (k_lo >= k_hi) return;                            // uniform per CTA
    const int x0 = gx * 4;
    const bool first_x = gx == 0, last_x = gx == ncols - 1;
    const long long sy = g.sy, sz = g.sz;
    const FsDivisor dv = fs_make_divisor(c);
    const float sgn_x = b == 1 ? -1.0f : 1.0f, sgn_y = b == 2 ? -1.0f : 1.0f, sgn_z = b == 3 ? -1.0f : 1.0f;
    const bool yface = j == 1 || j == g.ny - 2;

    // running pointers at plane kl = k_lo - 1 (the first stage-1 plane)
    const long long idx0 = live ? fs_idx(g, x0, j, k_lo - 1) : 0;
    const float *pin = in + idx0;
    const float *prh = rhs ? rhs + idx0 : nullptr;
    const uint8_t *pfl = flags ? flags + idx0 : nullptr;
    float *pout = out + idx0 - 2 * sz;                   // stage 2 trails two planes behind stage 1

    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a_prev = zero4, a_cur = zero4, a_next = zero4;    // z column of `in`: planes kl-1, kl, kl+1
    float4 y_a = zero4, y_b = zero4, y_c = zero4;            // z column of y1: planes kl-3, kl-2, kl-1
    float4 r_p = zero4, r_pp = zero4;                        // rhs of planes kl-1, kl-2 (stage 2 uses kl-2)
    uint32_t f_p = 0u, f_pp = 0u;
    if (live && !in_zero) {
        a_cur = ld4(pin);
        // plane k_lo - 2 exists unless plane k_lo - 1 is the global ring plane (then stage 1 only passes it through)
        if ((k_lo - 1) + g.zoff >= 1) a_prev = ld4(pin - sz);
    }

    // shared-memory slots of y1 planes kl (written), kl-1, kl-2 (read by stage 2); rotated instead of indexed
    float4 *s_w = &s_y1[0][t], *s_m = &s_y1[1][t], *s_r = &s_y1[2][t];
    // Per step kl: (1) issue the global loads of stage 1 at plane kl; (2) stage 2 at plane kl-2 from registers and
    // shared memory while those loads are in flight; (3) stage 1 arithmetic at plane kl; (4) publish y1(kl), one barrier.
    for (int kl = k_lo - 1; kl <= k_hi + 1; kl++, pin += sz, prh += sz, pfl += sz, pout += sz) {
        const int k = kl + g.zoff;                       // global plane of stage 1
        const bool s1 = kl <= k_hi;                      // (the last step only drains stage 2)
        const bool plane_in = s1 && k >= 1 && k <= g.nz - 2; // interior plane: stage 1 updates; ring plane: passes through
        const bool upd = live && plane_in && row_in;
        // ---------------- (1) loads of stage 1 at plane kl ----------------
        float4 up = zero4, dn = zero4, r4 = zero4;
        float left = 0.f, right = 0.f;
        uint32_t fl = 0u;
        if (live && s1 && !in_zero && (plane_in || kl < k_hi)) a_next = ld4(pin + sz);
        if (upd) {
            if (!in_zero) {
                up = ld4(pin + sy);
                dn = ld4(pin - sy);
                if (!first_x) left = __ldg(pin - 1);
                if (!last_x) right = __ldg(pin + 4);
            }
            if (KIND != FS_PAIR_SMOOTH) r4 = ld4_stream(prh);
            fl = flags ? ld_flags4(pfl) : 0u;
            // (no prefetch.global.L2 here: measured without effect on this kernel, profiles/r02b_pair_kernel.md)
        }
        // ---------------- (2) stage 2 at plane kl - 2 ----------------
        if (inner && kl - 2 >= k_lo && kl - 2 < k_hi) {
            const int k2 = k - 2;
            float4 up2 = s_r[cols], dn2 = s_r[-cols];
            const float left2 = s_r[-1].w, right2 = s_r[1].x;
            float4 zp = y_a, zn = y_c;
            const bool zface = k2 == 1 || k2 == g.nz - 2;
            if (KIND != FS_PAIR_RED_BLACK && (yface || zface)) { // set_bnd y / z faces of the intermediate field, on the fly
                // (a real branch: as predicated multiplies these 16 instructions cost every warp their issue slots)
                const float4 ys = make_float4(sgn_y * y_b.x, sgn_y * y_b.y, sgn_y * y_b.z, sgn_y * y_b.w);
                const float4 zs4 = make_float4(sgn_z * y_b.x, sgn_z * y_b.y, sgn_z * y_b.z, sgn_z * y_b.w);
                if (j == 1) dn2 = ys;
                if (j == g.ny - 2) up2 = ys;
                if (k2 == 1) zp = zs4;
                if (k2 == g.nz - 2) zn = zs4;
            }
            const float4 rr = KIND == FS_PAIR_SMOOTH ? y_b : r_pp;
            const float cv[6] = {left2, y_b.x, y_b.y, y_b.z, y_b.w, right2};
            const float upv[4] = {up2.x, up2.y, up2.z, up2.w}, dnv[4] = {dn2.x, dn2.y, dn2.z, dn2.w};
            const float nxv[4] = {zn.x, zn.y, zn.z, zn.w}, pvv[4] = {zp.x, zp.y, zp.z, zp.w};
            const float rv[4] = {rr.x, rr.y, rr.z, rr.w};
            const int par = (j + k2) & 1;                // x0 % 4 == 0: lane l has colour (l + j + k) & 1
            float v[4];
#pragma unroll
            for (int l = 0; l < 4; l++) {
                float s = ((cv[l + 2] + cv[l]) + upv[l]) + dnv[l];
                s = (s + nxv[l]) + pvv[l];
                const float val = fs_div(rv[l] + a * s, dv);
                bool keep = ((f_pp >> (8 * l)) & 1u) != 0u || (l == 0 && first_x) || (l == 3 && last_x);
                if (KIND == FS_PAIR_RED_BLACK) keep = keep || (((l + par) & 1) == 0);
                v[l] = keep ? cv[l + 1] : val;
            }
            if (!yface && !zface) {
                if (first_x) v[0] = sgn_x * v[1];
                if (last_x) v[3] = sgn_x * v[2];
                st4(pout, v);
            } else {
                const FsVec4Pos pos = fs_vec4_pos(g, x0, j, kl - 2);
                fs_vec4_store_ring(out, g, pos, v, b);
            }
        }
        // ---------------- (3) stage 1 arithmetic at plane kl ----------------
        float4 v1 = a_cur;                               // ring rows / planes / dead lanes pass `in` through
        if (upd) {
            const float4 rs = KIND == FS_PAIR_SMOOTH ? a_cur : r4;
            const float cv[6] = {left, a_cur.x, a_cur.y, a_cur.z, a_cur.w, right};
            const float upv[4] = {up.x, up.y, up.z, up.w}, dnv[4] = {dn.x, dn.y, dn.z, dn.w};
            const float nxv[4] = {a_next.x, a_next.y, a_next.z, a_next.w};
            const float pvv[4] = {a_prev.x, a_prev.y, a_prev.z, a_prev.w};
            const float rv[4] = {rs.x, rs.y, rs.z, rs.w};
            const int par = (j + k) & 1;
            float v[4];
#pragma unroll
            for (int l = 0; l < 4; l++) {
                float s = ((cv[l + 2] + cv[l]) + upv[l]) + dnv[l];
                s = (s + nxv[l]) + pvv[l];
                const float val = fs_div(rv[l] + a * s, dv);
                bool keep = ((fl >> (8 * l)) & 1u) != 0u || (l == 0 && first_x) || (l == 3 && last_x);
                if (KIND == FS_PAIR_RED_BLACK) keep = keep || (((l + par) & 1) != 0);
                v[l] = keep ? cv[l + 1] : val;
            }
            if (KIND != FS_PAIR_RED_BLACK) {             // set_bnd x faces of the intermediate field
                if (first_x) v[0] = sgn_x * v[1];
                if (last_x) v[3] = sgn_x * v[2];
            }
            v1 = make_float4(v[0], v[1], v[2], v[3]);
        }
        // ---------------- (4) publish y1(kl) ----------------
        *s_w = v1;
        __syncthreads();
        a_prev = a_cur; a_cur = a_next;
        y_a = y_b; y_b = y_c; y_c = v1;
        r_pp = r_p; r_p = r4; f_pp = f_p; f_p = fl;
        float4 *tmp_s = s_r; s_r = s_m; s_m = s_w; s_w = tmp_s;
    }
}

// ProjectDivergenceJob :1080-1095 + BoundaryJob(b = 0); 20 B/voxel in 3D (p = 0 is not stored).
template <bool HZ>
__global__ void __launch_bounds__(256)
divergence_vec4(const FsGrid g, float *__restrict__ div, const float *__restrict__ vx, const float *__restrict__ vy,
                const float *__restrict__ vz, const int kl0) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = kl0 + blockIdx.z;
    if (x0 >= g.nx || j > g.ny - 2) return;
    const FsVec4Pos pos = fs_vec4_pos(g, x0, j, kl);
    const long long idx = fs_idx(g, x0, j, kl);
    const float4 c = ld4(vx + idx), yt = ld4(vy + idx + g.sy), yb = ld4(vy + idx - g.sy);
    const float left = pos.first_x ? 0.f : __ldg(vx + idx - 1), right = pos.last_x ? 0.f : __ldg(vx + idx + 4);
    float4 zu = make_float4(0.f, 0.f, 0.f, 0.f), zd = zu;
    if (HZ) { zu = ld4(vz + idx + g.sz); zd = ld4(vz + idx - g.sz); }
    const float xs[6] = {left, c.x, c.y, c.z, c.w, right};
    const float ytv[4] = {yt.x, yt.y, yt.z, yt.w}, ybv[4] = {yb.x, yb.y, yb.z, yb.w};
    const float zuv[4] = {zu.x, zu.y, zu.z, zu.w}, zdv[4] = {zd.x, zd.y, zd.z, zd.w};
    const FsDivisor dn = fs_make_divisor((float)g.nx);
    float v[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        float s = ((xs[l + 2] - xs[l]) + ytv[l]) - ybv[l];
        if (HZ) s = (s + zuv[l]) - zdv[l];
        v[l] = fs_div(-0.5f * s, dn);
    }
    fs_vec4_store_ring(div, g, pos, v, 0);
}

// ProjectVelocityAdjustJob :1107-1122 + BoundaryJob(b = 1/2/3) faces, in place; 29 B/voxel in 3D.
template <bool HZ>
__global__ void __launch_bounds__(256)
gradient_vec4(const FsGrid g, float *vx, float *vy, float *vz, const float *__restrict__ p,
              const uint8_t *__restrict__ flags, const int kl0) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = kl0 + blockIdx.z;
    if (x0 >= g.nx || j > g.ny - 2) return;
    const FsVec4Pos pos = fs_vec4_pos(g, x0, j, kl);
    const long long idx = fs_idx(g, x0, j, kl);
    const float nf = (float)g.nx;
    const float4 pc = ld4(p + idx), pt = ld4(p + idx + g.sy), pb = ld4(p + idx - g.sy);
    const float pl = pos.first_x ? 0.f : __ldg(p + idx - 1), pr = pos.last_x ? 0.f : __ldg(p + idx + 4);
    const uint32_t fl = flags ? ld_flags4(flags + idx) : 0u;
    const float ps[6] = {pl, pc.x, pc.y, pc.z, pc.w, pr};
    const float ptv[4] = {pt.x, pt.y, pt.z, pt.w}, pbv[4] = {pb.x, pb.y, pb.z, pb.w};
    {
        const float4 u = ld4_plain(vx + idx);
        float v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int l = 0; l < 4; l++)
            if (!((fl >> (8 * l)) & 1u)) v[l] = v[l] - 0.5f * (ps[l + 2] - ps[l]) * nf;
        fs_vec4_store_ring(vx, g, pos, v, 1);
    }
    {
        const float4 u = ld4_plain(vy + idx);
        float v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int l = 0; l < 4; l++)
            if (!((fl >> (8 * l)) & 1u)) v[l] = v[l] - 0.5f * (ptv[l] - pbv[l]) * nf;
        fs_vec4_store_ring(vy, g, pos, v, 2);
    }
    if (HZ) {
        const float4 pu = ld4(p + idx + g.sz), pd = ld4(p + idx - g.sz);
        const float puv[4] = {pu.x, pu.y, pu.z, pu.w}, pdv[4] = {pd.x, pd.y, pd.z, pd.w};
        const float4 u = ld4_plain(vz + idx);
        float v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int l = 0; l < 4; l++)
            if (!((fl >> (8 * l)) & 1u)) v[l] = v[l] - 0.5f * (puv[l] - pdv[l]) * nf;
        fs_vec4_store_ring(vz, g, pos, v, 3);
    }
}

// AdvectJob :1138-1185 + BoundaryJob, float4 per thread, NF fields along one shared back-trace (NF = 3: the velocity
// components of VelocityStep :710-711; NF = 1: density or a single component).  The per-cell form issues 8 scalar gathers
// per cell and field (24 + 3 loads per cell for the velocity) and is bound by load-instruction / L1 wavefront throughput
// (1.3 TB/s at 512^3).  Here a thread owns 4 consecutive cells; wherever they share one integer displacement
// (i0 - i, j0, k0 equal, i0 - i in {-1, 0}: everywhere the flow moves less than a cell per step, i.e. the whole grid
// outside the plume core) the 5 source values a row contributes come from ONE 16-byte load plus one scalar, 8 loads per
// field for 4 cells instead of 32.  Other threads fall back to the per-cell gathers.  Same arithmetic, bit for bit.
struct FsAdvectBatch {
    float *dst[FS_BATCH];
    FsSlabView src[FS_BATCH];
    int b[FS_BATCH];
};
__device__ __forceinline__ void fs_row5(const float *row, int x0, int di, bool first_x, bool last_x, float r[5]) {
    const float4 m = ld4(row + x0);                      // cells x0 .. x0+3
    if (di == 0) {
        r[0] = m.x; r[1] = m.y; r[2] = m.z; r[3] = m.w;
        r[4] = last_x ? 0.0f : __ldg(row + x0 + 4);       // only the ring lane of the last float4 would use it
    } else {
        r[1] = m.x; r[2] = m.y; r[3] = m.z; r[4] = m.w;
        r[0] = first_x ? 0.0f : __ldg(row + x0 - 1);
    }
}
template <int NF, bool HZ>
__global__ void __launch_bounds__(256)
advect_vec4(const FsGrid g, const FsAdvectBatch ab, const float *__restrict__ velx, const float *__restrict__ vely,
            const float *__restrict__ velz, const uint8_t *__restrict__ flags, const float dt0, const int kl0) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = kl0 + blockIdx.z;
    if (x0 >= g.nx || j > g.ny - 2) return;
    const FsVec4Pos pos = fs_vec4_pos(g, x0, j, kl);
    const long long idx = fs_idx(g, x0, j, kl);
    const float4 ux = ld4(velx + idx), uy = ld4(vely + idx);
    float4 uz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HZ) uz = ld4(velz + idx);
    const uint32_t fl = flags ? ld_flags4(flags + idx) : 0u;
    const float uxv[4] = {ux.x, ux.y, ux.z, ux.w}, uyv[4] = {uy.x, uy.y, uy.z, uy.w}, uzv[4] = {uz.x, uz.y, uz.z, uz.w};
    FsAdvectWeights w[4];
#pragma unroll
    for (int l = 0; l < 4; l++) w[l] = fs_advect_weights(g, dt0, uxv[l], uyv[l], uzv[l], x0 + l, j, kl + g.zoff);
    // shared displacement of the INTERIOR lanes (the x-face lanes x = 0 / nx-1 are ring cells: not advected, their velocity
    // carries set_bnd's sign flip, and letting them veto the fast path sent half the warps of a 512-wide grid down the
    // gather path)
    const int di = pos.first_x ? w[1].i0 - (x0 + 1) : w[0].i0 - x0;   // (selects, not w[lr]: a dynamic index would put w in local memory)
    const int j0 = pos.first_x ? w[1].j0 : w[0].j0, k0 = pos.first_x ? w[1].k0 : w[0].k0;
    bool uniform = di == 0 || di == -1;
#pragma unroll
    for (int l = 0; l < 4; l++) {
        const bool ring_lane = (l == 0 && pos.first_x) || (l == 3 && pos.last_x);
        uniform = uniform && (ring_lane || ((w[l].i0 - (x0 + l) == di) && w[l].j0 == j0 && w[l].k0 == k0));
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
        const FsSlabView &sv = ab.src[f];
        float v[4];
        if (uniform) {
            const float *p0 = fs_slab_plane(sv, g, k0) + j0 * g.sy;
            float A[5], B[5];
            fs_row5(p0, x0, di, pos.first_x, pos.last_x, A);
            fs_row5(p0 + g.sy, x0, di, pos.first_x, pos.last_x, B);
            float lo[4];
#pragma unroll
            for (int l = 0; l < 4; l++)
                lo[l] = w[l].s0 * (w[l].t0 * A[l] + w[l].t1 * B[l]) + w[l].s1 * (w[l].t0 * A[l + 1] + w[l].t1 * B[l + 1]);
            if (HZ) {
                const float *p1 = fs_slab_plane(sv, g, k0 + 1) + j0 * g.sy;
                float C[5], D[5];
                fs_row5(p1, x0, di, pos.first_x, pos.last_x, C);
                fs_row5(p1 + g.sy, x0, di, pos.first_x, pos.last_x, D);
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    const float hi = w[l].s0 * (w[l].t0 * C[l] + w[l].t1 * D[l]) + w[l].s1 * (w[l].t0 * C[l + 1] + w[l].t1 * D[l + 1]);
                    v[l] = w[l].u0 * lo[l] + w[l].u1 * hi;
                }
            } else {
#pragma unroll
                for (int l = 0; l < 4; l++) v[l] = lo[l];
            }
        } else {
#pragma unroll
            for (int l = 0; l < 4; l++) v[l] = fs_advect_interp(g, w[l], [&](int kk) { return fs_slab_plane(sv, g, kk); });
        }
#pragma unroll
        for (int l = 0; l < 4; l++)
            if ((fl >> (8 * l)) & 1u) v[l] = 0.0f;      // the reference's output array is fresh: obstacle cells get 0 (:1529, :1148-1156)
        fs_vec4_store_ring(ab.dst[f], g, pos, v, ab.b[f]);
    }
}

// Red-black Gauss-Seidel half sweep (BASELINE config 5; oracle fo_lin_solve_rb), float4 per thread, in place.
// A cell of colour (i+j+k)&1 reads only cells of the other colour, which no thread writes in this launch, so
// the update order is free.  The colour-1 launch also performs set_bnd (ring scatter from the now final row),
// so a full sweep is two launches and 26 B/voxel (a fused single-pass form is the next step).
template <bool HZ>
__global__ void __launch_bounds__(256)
rb_vec4(const FsGrid g, float *x, const float *__restrict__ rhs, const uint8_t *__restrict__ flags, const float a,
        const float c, const int colour, const int b, const int write_ring, const int kl0) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = kl0 + blockIdx.z;
    if (x0 >= g.nx || j > g.ny - 2) return;
    FsVec4Pos pos = fs_vec4_pos(g, x0, j, kl);
    const long long idx = fs_idx(g, x0, j, kl);
    const float4 cur = ld4_plain(x + idx), up = ld4_plain(x + idx + g.sy), dn = ld4_plain(x + idx - g.sy);
    const float left = pos.first_x ? 0.f : x[idx - 1], right = pos.last_x ? 0.f : x[idx + 4];
    float4 zu = make_float4(0.f, 0.f, 0.f, 0.f), zd = zu;
    if (HZ) { zu = ld4_plain(x + idx + g.sz); zd = ld4_plain(x + idx - g.sz); }
    const float4 r4 = ld4_stream(rhs + idx);
    const uint32_t fl = flags ? ld_flags4(flags + idx) : 0u;
    const float cv[6] = {left, cur.x, cur.y, cur.z, cur.w, right};
    const float upv[4] = {up.x, up.y, up.z, up.w}, dnv[4] = {dn.x, dn.y, dn.z, dn.w};
    const float zuv[4] = {zu.x, zu.y, zu.z, zu.w}, zdv[4] = {zd.x, zd.y, zd.z, zd.w};
    const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
    const FsDivisor dv = fs_make_divisor(c);
    const int par = (j + kl + g.zoff) & 1; // x0 is a multiple of 4: lane l has colour (l + j + k) & 1
    float v[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        float s = ((cv[l + 2] + cv[l]) + upv[l]) + dnv[l];
        if (HZ) s = (s + zuv[l]) + zdv[l];
        const float val = fs_div(rv[l] + a * s, dv);
        const bool ring_lane = (l == 0 && pos.first_x) || (l == 3 && pos.last_x);
        const bool upd = (((l + par) & 1) == colour) && !((fl >> (8 * l)) & 1u) && !ring_lane;
        v[l] = upd ? val : cv[l + 1];
    }
    if (write_ring) {
        fs_vec4_store_ring(x, g, pos, v, b);
    } else {
        st4(x + idx, v);
    }
}

// ---- RecursiveFloodFill, FluidSim.cs:329-351, as a fixed point: a cell of the shape becomes an obstacle once one of
// its 4 neighbours is one, starting from the seed.  One CTA; `reach` only ever goes 0 -> 1, so the concurrent
// reads are benign and the fixed point is the reference's fill.  The cross-section is at most 1024^2 cells.
__global__ void __launch_bounds__(1024)
flood_fill_kernel(const int nx, const int ny, const uint8_t *__restrict__ inside, uint8_t *reach, const int seed_x, const int seed_y) {
    __shared__ int changed;
    volatile uint8_t *r = reach;
    const long long n = (long long)nx * ny;
    if (threadIdx.x == 0 && inside[seed_x + (long long)seed_y * nx]) r[seed_x + (long long)seed_y * nx] = 1;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        bool mine = false;
        for (long long c = threadIdx.x; c < n; c += blockDim.x) {
            if (!inside[c] || r[c]) continue;
            const int x = (int)(c % nx), y = (int)(c / nx);
            if ((x > 0 && r[c - 1]) || (x < nx - 1 && r[c + 1]) || (y > 0 && r[c - nx]) || (y < ny - 1 && r[c + nx])) {
                r[c] = 1;
                mine = true;
            }
        }
        if (mine) changed = 1;
        __syncthreads();
        if (!changed) break;
    }
}

// ---- metrics (LogCurrentMetrics, FluidSim.cs:582-594): sum of density, max |V| -------------------------
__global__ void __launch_bounds__(256)
metrics_kernel(const float *__restrict__ d, const float *__restrict__ ux, const float *__restrict__ uy,
               const float *__restrict__ uz, long long n, long long per, double *sum, unsigned int *mx) {
    double acc = 0.0;
    float m = 0.0f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long r = 0; r < per && t < n; r++, t += stride) {
        acc += (double)d[t];
        float q = ux[t] * ux[t] + uy[t] * uy[t];
        if (uz) q = q + uz[t] * uz[t];
        m = fmaxf(m, sqrtf(q));
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    __shared__ double s_acc[8];
    __shared__ float s_m[8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_acc[w] = acc; s_m[w] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { acc += s_acc[i]; m = fmaxf(m, s_m[i]); }
        atomicAdd(sum, acc);
        atomicMax(mx, __float_as_uint(m)); // non-negative floats order like their bit patterns
    }
}
