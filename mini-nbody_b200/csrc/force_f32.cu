// K1: all-pairs softened-gravity force, FP32, hand-written for sm_100a.
//
// Computes, for every i-body of this rank's slice and a range of j-bodies,
//     a_i += (r_j - r_i) * (|r_j - r_i|^2 + 1e-9)^(-3/2)
// which is the per-pair dataflow of the reference pipeline (dxy.vhd:94-122, dzsoft.vhd:177-202,
// dxyz_soft.vhd:149-150, fxyz.vhd:101-127, cube.vhd:66-70): d = target - this, softening added to
// dist^2, rsqrt, cube, three accumulating FMAs; unit masses, self-pair included.
//
// Mapping (the reference streams one j per clock past 12 i-pipelines, top_level.vhd:44,233-249):
//   * each thread keeps I i-bodies in registers (register blocking); a CTA covers I*THREADS
//     i-bodies (blockIdx.x) and one j-split (blockIdx.y); partial sums go to slot
//     slot0 + blockIdx.y and are added by the integrate kernel (the reference likewise keeps 16
//     interleaved partial sums per (body, dim) and tree-adds them, fxyz.vhd:120-145,
//     final_adder.vhd:88-104);
//   * j-bodies arrive by 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) of STAGE_BLOCKS layout
//     blocks per stage into an NSTAGES-deep shared-memory ring, full/empty mbarriers, producer =
//     thread 0 with a two-stage look-ahead and one stage of slack;
//   * the inner loop reads four consecutive j-bodies per row with one broadcast LDS.128 and runs
//     the arithmetic as packed f32x2 pairs over j (FADD2 / FFMA2 / FMUL2, the i-operand entering as
//     a scalar broadcast, the softening as an immediate), two MUFU.RSQ per pair: 11 packed ops + 2
//     MUFU per 2 interactions.  Measured on B200 (profiles/r01_microbench.md): FFMA2 sustains 128
//     lane-FMA/clk/SM in half the issue slots of scalar FFMA, beside which a MUFU costs ~4 issue
//     cycles (scalar loop: 15.6 cycles per interaction, 60 % of peak, kept as variant 5); a packed op
//     takes max(2, distinct register pairs read) cycles, so the floor of this loop is 11.5 cycles per
//     interaction and the kernel runs at 12.8 (78 % of the 20-flop FP32 peak).
#include "nbody_internal.cuh"
#include "force_f32_inner.cuh"
#include "stream.cuh"
#include <cooperative_groups.h>

namespace nb {

constexpr int RING_PAD = 64;      // shared memory between the j-ring and the mbarriers (see force_unit)

// Second-level accumulators (shared memory, one private f2 per thread, accumulator and dimension):
// after every j-stage (SB blocks = 512 j = 256 adds per register accumulator) the register sums are
// folded into them and cleared.  Without this a body with one very close neighbour (pair term ~1e7 at
// N = 1M) keeps that term in a register accumulator whose ulp is then ~1, and every later term
// smaller than half an ulp is absorbed: measured 1.9e-4 relative error on such a body, reproduced
// digit for digit by a NumPy emulation of the summation order.  Each thread
// touches only its own words, so no barrier is needed; cost 3*I (LDS.64 + FADD2 + STS.64) per stage
// (folding after every block instead was measured 4 % slower, per stage ~1 %).
template <int I, int THREADS>
__device__ __forceinline__ void fold_accumulators(IState<I>& s, f2* acc2, int tid) {
#pragma unroll
    for (int q = 0; q < I; q++) {
        f2* p = acc2 + (size_t)(q * 3) * THREADS + tid;
        p[0] = add2(p[0], s.ax[q]); p[THREADS] = add2(p[THREADS], s.ay[q]); p[2 * THREADS] = add2(p[2 * THREADS], s.az[q]);
        s.ax[q] = s.ay[q] = s.az[q] = pk(0.f, 0.f);
    }
}

// One work unit = (i-tile, j-split): the partial accelerations of I*THREADS i-bodies over one j-range.
// COHERENT: the positions may have been written by other CTAs of the SAME launch (fused multi-step kernel
// below), so the i-bodies are loaded past L1 (ld.global.cg) and the mbarriers are re-initialised per unit.
template <int I, int THREADS, int SB, int NS, bool PACKED, int PIPE, bool FOLD, int UNROLL, bool EPS_RT, bool COHERENT>
__device__ __forceinline__ void force_unit(const ForceArgs& a, const int tile, const int split, unsigned char* smem_raw, const bool reinit = false) {
    static_assert(THREADS % BLK == 0 || BLK % THREADS == 0, "thread/block mapping");
    static_assert(NS >= 3, "need >= 3 stages for the look-ahead scheme");
    static_assert(PIPE != 2 || SB == 4, "the row-major stage path is validated for 4-block stages only (an 8-block probe gave wrong sums and 0.4 % speed: not pursued)");
    constexpr int STAGE_FLOATS = SB * 3 * BLK;
    constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
    constexpr int NWARPS = THREADS / 32;
    constexpr int LOOKAHEAD = NS - 2;              // tiles in flight beyond the one being consumed

    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    // RING_PAD bytes behind the last stage: the rotated loop's final prefetch reads up to 16*UNROLL bytes past its rows, which
    // must never be mbarrier storage (non-mbarrier access to an mbarrier object is undefined)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_BYTES + RING_PAD);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NS);
    f2* acc2 = reinterpret_cast<f2*>(smem_raw + (size_t)NS * STAGE_BYTES + RING_PAD + 2 * NS * 8);

    const int tid = threadIdx.x;
    const float* __restrict__ pos = static_cast<const float*>(a.pos);
    const float eps = EPS_RT ? a.eps32 : EPS_F32;        // compile-time immediate in the default instantiations
    if (FOLD) {
#pragma unroll
        for (int q = 0; q < 3 * I; q++) acc2[(size_t)q * THREADS + tid] = pk(0.f, 0.f);
    }

    // j-range of this split, in rotated block coordinates
    const int jb0 = (int)(((long long)split * a.j_len) / a.nsplit);
    const int jb1 = (int)(((long long)(split + 1) * a.j_len) / a.nsplit);
    const int ntiles = (jb1 - jb0 + SB - 1) / SB;

    if (tid == 0) {
        if (COHERENT && reinit)                     // a later unit of a persistent CTA: the objects of the previous unit are idle
            for (int s = 0; s < 2 * NS; s++) mbar_inval(full0 + 8 * s);
        for (int s = 0; s < NS; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NWARPS); }
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int k) {                       // thread 0 only: bring tile k into stage k % NS
        const int st = k % NS;
        const int rb = jb0 + k * SB;
        const int cnt = min(SB, jb1 - rb);
        int p = a.j_rot0 + rb; if (p >= a.total_blocks) p -= a.total_blocks;
        const uint32_t bar = full0 + 8 * st;
        const uint32_t dst = smem_u32(stage_buf + (size_t)st * STAGE_FLOATS);
        mbar_expect_tx(bar, (uint32_t)cnt * 3 * BLK * 4);
        if (PIPE == 2) {
            // row-major stage [X: SB*BLK][Y: SB*BLK][Z: SB*BLK]: one 512 B bulk copy per row of every block,
            // so the consumer walks one flat x-row per stage (loads for the next iteration issued a whole
            // iteration ahead, see the loop below)
            for (int c = 0; c < cnt; c++) {
                const float* src = pos + (size_t)p * 3 * BLK;
#pragma unroll
                for (int d = 0; d < 3; d++)
                    bulk_g2s(dst + (uint32_t)((d * SB + c) * BLK * 4), src + d * BLK, BLK * 4, bar);
                if (++p == a.total_blocks) p = 0;
            }
            return;
        }
        const int first = min(cnt, a.total_blocks - p);
        bulk_g2s(dst, pos + (size_t)p * 3 * BLK, (uint32_t)first * 3 * BLK * 4, bar);
        if (first < cnt)                            // range wraps past the end of the array
            bulk_g2s(dst + (uint32_t)first * 3 * BLK * 4, pos, (uint32_t)(cnt - first) * 3 * BLK * 4, bar);
    };
    if (tid == 0) {
        // fused multi-GPU pass: a CTA whose j-range reaches past the rank's own slice acquires the peers' step flags first
        if (a.fuse && a.wait.flags != nullptr && jb1 > a.local_len) wait_for_peers(a.wait);
        for (int k = 0; k < LOOKAHEAD && k < ntiles; k++) issue(k);
    }

    // i-bodies of this thread: local i-block (tile*IB + q*(THREADS/BLK) + tid/BLK), lane tid % BLK
    constexpr int IB = I * THREADS / BLK;           // i-blocks per CTA
    constexpr int TB = THREADS / BLK;               // i-blocks covered by one "row" of threads
    const int lane_in_blk = tid % BLK;
    IState<I> s;
    int iblk[I];
#pragma unroll
    for (int q = 0; q < I; q++) {
        iblk[q] = tile * IB + q * TB + tid / BLK;
        const int ib = min(iblk[q], a.n_iblk - 1);  // idle threads of a ragged last tile alias a valid block
        const float* pi = pos + ((size_t)(a.i_blk0 + ib) * 3) * BLK + lane_in_blk;
        if (COHERENT) { s.nx[q] = -__ldcg(pi); s.ny[q] = -__ldcg(pi + BLK); s.nz[q] = -__ldcg(pi + 2 * BLK); }
        else { s.nx[q] = -pi[0]; s.ny[q] = -pi[BLK]; s.nz[q] = -pi[2 * BLK]; }
        s.ax[q] = s.ay[q] = s.az[q] = pk(0.f, 0.f);
    }

    for (int k = 0; k < ntiles; k++) {
        const int st = k % NS;
        if (tid == 0 && k + LOOKAHEAD < ntiles) {
            const int kn = k + LOOKAHEAD;
            if (kn >= NS) mbar_wait(empty0 + 8 * (kn % NS), (uint32_t)((kn / NS) - 1) & 1u);
            issue(kn);
        }
        mbar_wait(full0 + 8 * st, (uint32_t)(k / NS) & 1u);
        const int cnt = min(SB, jb1 - (jb0 + k * SB));
        const float* sb = stage_buf + (size_t)st * STAGE_FLOATS;
        if (PIPE == 2) {
            // rotated loop over the flat x/y/z rows of the stage: the UNROLL j-groups of iteration it+1 are
            // loaded while iteration it computes (on its last iteration the loop reads <= 16*UNROLL bytes
            // past the rows it owns: still inside the shared-memory allocation, never used).  This is the
            // form tools/sass_sched.py re-schedules: no LDS latency at the top of the body.
            constexpr int ROW4 = SB * BLK / 4;
            const float4* sx = reinterpret_cast<const float4*>(sb);
            float4 X[UNROLL], Y[UNROLL], Z[UNROLL];
#pragma unroll
            for (int g = 0; g < UNROLL; g++) { X[g] = sx[g]; Y[g] = sx[g + ROW4]; Z[g] = sx[g + 2 * ROW4]; }
            const int niter = cnt * (BLK / (4 * UNROLL));
#pragma unroll 1
            for (int it = 0; it < niter; it++) {
#pragma unroll
                for (int g = 0; g < UNROLL; g++) interact4<I>(s, X[g], Y[g], Z[g], eps);
                sx += UNROLL;
#pragma unroll
                for (int g = 0; g < UNROLL; g++) { X[g] = sx[g]; Y[g] = sx[g + ROW4]; Z[g] = sx[g + 2 * ROW4]; }
            }
        } else if (PIPE == 3) {
            // measured alternative (north_star: "warp-shuffle broadcast of j-tiles"): every lane keeps one j-body of a
            // 32-body group in registers and the pair operands are broadcast with SHFL instead of a broadcast LDS.128.
            // 6 SHFL per pair of j (each reads and writes the register file, which is the bottleneck of this loop)
            // against 1.5 LDS.128: see DESIGN.md section 4 for the numbers; not the shipped path.
            const int lane = tid & 31;
            for (int b = 0; b < cnt; b++) {
                const float* row = sb + b * 3 * BLK;
                for (int g = 0; g < BLK / 32; g++) {
                    const float xj = row[g * 32 + lane], yj = row[BLK + g * 32 + lane], zj = row[2 * BLK + g * 32 + lane];
#pragma unroll 2
                    for (int jj = 0; jj < 32; jj += 4) {
                        float4 X, Y, Z;
                        X.x = __shfl_sync(0xffffffffu, xj, jj); X.y = __shfl_sync(0xffffffffu, xj, jj + 1);
                        X.z = __shfl_sync(0xffffffffu, xj, jj + 2); X.w = __shfl_sync(0xffffffffu, xj, jj + 3);
                        Y.x = __shfl_sync(0xffffffffu, yj, jj); Y.y = __shfl_sync(0xffffffffu, yj, jj + 1);
                        Y.z = __shfl_sync(0xffffffffu, yj, jj + 2); Y.w = __shfl_sync(0xffffffffu, yj, jj + 3);
                        Z.x = __shfl_sync(0xffffffffu, zj, jj); Z.y = __shfl_sync(0xffffffffu, zj, jj + 1);
                        Z.z = __shfl_sync(0xffffffffu, zj, jj + 2); Z.w = __shfl_sync(0xffffffffu, zj, jj + 3);
                        interact4<I>(s, X, Y, Z, eps);
                    }
                }
            }
        } else if (PIPE == 1) {
            // software-pipelined: the next group's three LDS.128 are in flight while the current
            // group is being computed (ping-pong register sets, rows of a block are 512 B apart)
            const float4* sx = reinterpret_cast<const float4*>(sb);
            float4 X0 = sx[0], Y0 = sx[BLK / 4], Z0 = sx[2 * (BLK / 4)];
            const int ngroups = cnt * (BLK / 4);
            for (int q = 0; q < ngroups; q += 2) {
                // group q+1 lives in the same block (BLK/4 = 32 groups per block, q even)
                const float4 X1 = sx[1], Y1 = sx[1 + BLK / 4], Z1 = sx[1 + 2 * (BLK / 4)];
                if (PACKED) interact4<I>(s, X0, Y0, Z0, eps); else interact4_scalar<I>(s, X0, Y0, Z0, eps);
                // group q+2: next block when q+2 crosses a multiple of 32 (skip the y and z rows)
                sx += 2;
                if (((q + 2) & (BLK / 4 - 1)) == 0) sx += 2 * (BLK / 4);
                if (q + 2 < ngroups) { X0 = sx[0]; Y0 = sx[BLK / 4]; Z0 = sx[2 * (BLK / 4)]; }
                if (PACKED) interact4<I>(s, X1, Y1, Z1, eps); else interact4_scalar<I>(s, X1, Y1, Z1, eps);
            }
        } else {
            for (int b = 0; b < cnt; b++) {
                const float4* sx = reinterpret_cast<const float4*>(sb + b * 3 * BLK);
#pragma unroll UNROLL
                for (int g = 0; g < BLK / 4; g++) {
                    const float4 X = sx[g], Y = sx[g + BLK / 4], Z = sx[g + 2 * (BLK / 4)];
                    if (PACKED) interact4<I>(s, X, Y, Z, eps); else interact4_scalar<I>(s, X, Y, Z, eps);
                }
            }
        }
        if (FOLD) fold_accumulators<I, THREADS>(s, acc2, tid);     // once per stage: chains of <= SB*BLK/2 adds
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * st);
    }

    if (a.fuse) {
        // fused mode: the tile's slot for this split in the L2-resident ring, thread-private columns [q*3+d][tid]
        float* __restrict__ w = static_cast<float*>(a.ws) + ((size_t)(tile % a.ring) * a.nsplit + split) * (I * 3 * THREADS) + tid;
#pragma unroll
        for (int q = 0; q < I; q++) {
            float lo, hi;
            const f2* p2 = acc2 + (size_t)(q * 3) * THREADS + tid;
            upk(FOLD ? p2[0] : s.ax[q], lo, hi); w[(size_t)(q * 3) * THREADS] = lo + hi;
            upk(FOLD ? p2[THREADS] : s.ay[q], lo, hi); w[(size_t)(q * 3 + 1) * THREADS] = lo + hi;
            upk(FOLD ? p2[2 * THREADS] : s.az[q], lo, hi); w[(size_t)(q * 3 + 2) * THREADS] = lo + hi;
        }
        return;
    }
    float* __restrict__ part = static_cast<float*>(a.part) + (size_t)(a.slot0 + split) * a.n_iblk * 3 * BLK;
#pragma unroll
    for (int q = 0; q < I; q++) {
        if (iblk[q] < a.n_iblk) {
            float lo, hi;
            float* o = part + (size_t)iblk[q] * 3 * BLK + lane_in_blk;
            const f2* p2 = acc2 + (size_t)(q * 3) * THREADS + tid;      // register sums are zero after the last fold
            upk(FOLD ? p2[0] : s.ax[q], lo, hi); o[0] = lo + hi;
            upk(FOLD ? p2[THREADS] : s.ay[q], lo, hi); o[BLK] = lo + hi;
            upk(FOLD ? p2[2 * THREADS] : s.az[q], lo, hi); o[2 * BLK] = lo + hi;
        }
    }
}

template <int I, int THREADS, int SB, int NS, int MINB, bool PACKED, int PIPE, bool FOLD, int UNROLL, bool EPS_RT>
__global__ void __launch_bounds__(THREADS, MINB) force_f32_kernel(const ForceArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // (tile, split) of this CTA: 2-D grid (tile, split), or a 1-D grid -- split-major over ALL tiles (order 0: whole waves of
    // equal CTAs; every tile keeps its own slots), or split-major inside GROUPS of ring/2 tiles taken one after the other
    // (order 1: the slots of a group are read back from L2 while the next group fills the other half of the ring, so partial
    // sums never travel to HBM however many tiles there are).  ONE call site of force_unit: the post-ptxas re-scheduler
    // patches one hot loop per kernel.
    __shared__ int s_last;
    int tile = blockIdx.x, split = blockIdx.y;
    if (a.fuse || a.order == 1) fused_cta_of((int)blockIdx.x, (int)gridDim.x / a.nsplit, a.nsplit, a.ring, a.order, &tile, &split);
    const int tid = threadIdx.x;
    if (a.fuse && tile >= a.ring) {
        // ring position reuse: tile - ring must have been reduced (always long true: CTAs are dispatched in order; the wait
        // only makes the reuse safe, with the usual time-out instead of a hang)
        if (tid == 0) {
            const unsigned long long t0 = globaltimer_ns();
            while (*(volatile unsigned int*)(a.tile_done + (tile - a.ring)) != a.epoch) {
                if (globaltimer_ns() - t0 > 20000000000ull) { if (a.wait.err) atomicExch(a.wait.err, 2); break; }
                __nanosleep(200);
            }
            __threadfence();
        }
        __syncthreads();
    }
    force_unit<I, THREADS, SB, NS, PACKED, PIPE, FOLD, UNROLL, EPS_RT, false>(a, tile, split, smem_raw);
    if (!a.fuse) return;
    __threadfence();                                           // this split's sums are visible device-wide ...
    __syncthreads();                                           // ... before the CTA counts itself
    if (tid == 0) {
        const bool last = atomicAdd(a.tile_counter + tile, 1u) == (unsigned)a.nsplit - 1u;
        if (last) a.tile_counter[tile] = 0u;
        s_last = last;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float acc[I][3];
#pragma unroll
    for (int q = 0; q < I; q++) acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
    const float* __restrict__ w = static_cast<const float*>(a.ws) + (size_t)(tile % a.ring) * a.nsplit * (I * 3 * THREADS) + tid;
#pragma unroll 2
    for (int sl = 0; sl < a.nsplit; sl++) {                   // fixed order => deterministic, same sum as integrate_kernel
#pragma unroll
        for (int q = 0; q < I; q++)
#pragma unroll
            for (int d = 0; d < 3; d++) acc[q][d] += __ldcg(w + ((size_t)sl * (I * 3) + q * 3 + d) * THREADS);
    }
    tile_epilogue<float, I, THREADS>(a.ep, tile, acc);
    __syncthreads();                                           // every thread has read the tile's slots ...
    if (tid == 0) { __threadfence(); *(volatile unsigned int*)(a.tile_done + tile) = a.epoch; }   // ... before the ring position is released
}

// ---- fused multi-step kernel for launch-bound sizes ------------------------------------------------------
// One cooperative launch runs nsteps whole time steps (bodyForce + integrate) on one GPU: persistent CTAs take
// the (i-tile, j-split) units of the step exactly as the two-kernel path would launch them (same instantiation,
// same splits => the same partial sums, bit for bit); the CTA that finishes the LAST split of an i-tile adds the
// tile's partials in slot order, updates v and x and writes the tile's slice of pos[next] (what integrate_kernel
// does), and one grid barrier per step separates readers of pos[cur] from writers of the buffer it becomes.
// At N = 4096 this replaces 2 launches x 10 steps by one launch (launch-bound sizes).
template <int I, int THREADS, int SB, int NS, int MINB, bool EPS_RT>
__global__ void __launch_bounds__(THREADS, MINB) step_fused_f32_kernel(const FusedStepArgs fa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_last;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    constexpr int IB = I * THREADS / BLK, TB = THREADS / BLK;
    const int tid = threadIdx.x, lane_in_blk = tid % BLK;
    const int units = fa.i_tiles * fa.nsplit;
    int cur = fa.cur;
    for (int step = 0; step < fa.nsteps; step++, cur ^= 1) {
        ForceArgs a{};
        a.pos = fa.pos[cur]; a.part = fa.part; a.total_blocks = fa.n_iblk; a.i_blk0 = 0; a.n_iblk = fa.n_iblk;
        a.j_rot0 = 0; a.j_len = fa.n_iblk; a.nsplit = fa.nsplit; a.slot0 = 0; a.eps32 = fa.eps32;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int tile = unit % fa.i_tiles, split = unit / fa.i_tiles;
            force_unit<I, THREADS, SB, NS, true, 0, true, 2, EPS_RT, true>(a, tile, split, smem_raw, step > 0 || unit != (int)blockIdx.x);
            __threadfence();                                   // this CTA's partial sums are visible device-wide ...
            __syncthreads();                                   // ... before it counts itself (also: mbarriers idle again)
            if (tid == 0) s_last = (atomicAdd(fa.tile_counter + tile, 1u) == (unsigned)fa.nsplit - 1u);
            __syncthreads();
            if (s_last) {
                __threadfence();
                const float* __restrict__ part = static_cast<const float*>(fa.part);
                const float* __restrict__ pc = static_cast<const float*>(fa.pos[cur]);
                float* __restrict__ pn = static_cast<float*>(fa.pos[cur ^ 1]);
                float* __restrict__ vel = static_cast<float*>(fa.vel);
                const size_t slot_stride = (size_t)fa.n_iblk * 3 * BLK;
#pragma unroll
                for (int q = 0; q < I; q++) {
                    const int ib = tile * IB + q * TB + tid / BLK;
                    if (ib >= fa.n_iblk) continue;
                    const size_t loc = (size_t)ib * 3 * BLK + lane_in_blk;
                    float vx = __ldcg(vel + loc), vy = __ldcg(vel + loc + BLK), vz = __ldcg(vel + loc + 2 * BLK);
                    float x = __ldcg(pc + loc), y = __ldcg(pc + loc + BLK), z = __ldcg(pc + loc + 2 * BLK);
                    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll 8
                    for (int sl = 0; sl < fa.nsplit; sl++) {    // fixed order => deterministic, same as integrate_kernel
                        const float* p = part + (size_t)sl * slot_stride + loc;
                        ax += __ldcg(p); ay += __ldcg(p + BLK); az += __ldcg(p + 2 * BLK);
                    }
                    if ((long long)ib * BLK + lane_in_blk < fa.n) {      // padding bodies never move
                        vx = fmaf(fa.dt_v, ax, vx); vy = fmaf(fa.dt_v, ay, vy); vz = fmaf(fa.dt_v, az, vz);
                        x = fmaf(vx, fa.dt_x, x); y = fmaf(vy, fa.dt_x, y); z = fmaf(vz, fa.dt_x, z);
                    }
                    vel[loc] = vx; vel[loc + BLK] = vy; vel[loc + 2 * BLK] = vz;
                    pn[loc] = x; pn[loc + BLK] = y; pn[loc + 2 * BLK] = z;
                }
                if (tid == 0) fa.tile_counter[tile] = 0u;       // ready for the next step (ordered by the grid barrier)
            }
        }
        // pos[next] complete and visible (generic proxy), also to the bulk copies (async proxy) of the next step
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        grid.sync();                                       // measured: faster here than a hand-rolled count+generation barrier
    }
}

// ---- stream-K force pass (default from 6144 bodies per GPU) ------------------------------------------------
// One segment = the tile's I*THREADS i-bodies against granules [ja, jb) of a phase's rotated j-range.  Same
// TMA/mbarrier ring, same inner loop (interact4) and the same two accumulation levels as force_unit above; what
// differs:
//   * the j-range is cut at GRAN = 16 bodies, not at layout blocks: a stage is SG granules laid out row-major
//     [X: SG*16][Y][Z]; thread 0 brings it in with one bulk copy per (layout block touched, row), <= 64 B .. 512 B
//     each, so a stage may start and end anywhere inside a block (this is what lets every CTA of the pass get the
//     same amount of work whatever N is);
//   * the ring runs on through the segments of a persistent CTA (kbase = stages consumed so far): no barrier
//     re-initialisation, no drain beyond the one the data dependence forces;
//   * a third accumulation level: every CHAIN_STAGES stages (65 536 j) the shared-memory f32x2 sums are folded into
//     `res` (plain floats), so a segment that spans a million j keeps every chain as short as the (tile, split)
//     kernel did with its 65 536-body splits;
//   * the result stays in `res` for the stream driver (stream.cuh), which stores / reduces / integrates it.
template <int I, int THREADS, int SG, int NS, int LOOP, int UNROLL, bool EPS_RT>
__device__ __forceinline__ void force_segment_f32(const StreamArgs& a, const int tile, const int rot0, const int ja, const int jb,
                                                  unsigned char* smem_raw, int& kbase) {
    static_assert(NS >= 3, "need >= 3 stages for the look-ahead scheme");
    static_assert(LOOP != 2 || GRAN % (4 * UNROLL) == 0, "the rotated loop consumes whole granules per iteration");
    constexpr int ROWF = SG * GRAN;                  // floats per row of a stage
    constexpr int STAGE_FLOATS = 3 * ROWF;
    constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
    constexpr int LOOKAHEAD = NS - 2;
    constexpr int CHAIN_STAGES = 65536 / ROWF;

    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    // 64 B of padding behind the last stage: the rotated loop's final prefetch reads up to 16*UNROLL bytes past its rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_BYTES + RING_PAD);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NS);
    f2* acc2 = reinterpret_cast<f2*>(smem_raw + (size_t)NS * STAGE_BYTES + RING_PAD + 2 * NS * 8);
    float* res = reinterpret_cast<float*>(acc2 + (size_t)3 * I * THREADS);

    const int tid = threadIdx.x;
    const float* __restrict__ pos = static_cast<const float*>(a.pos);
    const float eps = EPS_RT ? a.eps32 : EPS_F32;
#pragma unroll
    for (int q = 0; q < 3 * I; q++) { acc2[(size_t)q * THREADS + tid] = pk(0.f, 0.f); res[(size_t)q * THREADS + tid] = 0.f; }

    const int nst = (jb - ja + SG - 1) / SG;
    const int k0 = kbase;                            // global index of this segment's first stage
    constexpr int ptid = 0;                          // producer: thread 0 (a short first stage that block-aligns the later ones, and a
                                                     // producer warp that differs between the CTAs of an SM, were measured: no gain)
    auto stage_g0 = [&](int k) { return ja + k * SG; };
    auto stage_cnt = [&](int k) { return min(SG, jb - (ja + k * SG)); };
    auto issue = [&](int k) {                        // producer thread only: bring stage k of the segment into ring slot (k0+k) % NS
        const int gk = k0 + k, slot = gk % NS;
        if (gk >= NS) mbar_wait(empty0 + 8 * slot, (uint32_t)((gk / NS) - 1) & 1u);
        int g = stage_g0(k);
        int cnt = stage_cnt(k);
        const uint32_t bar = full0 + 8 * slot;
        uint32_t dst = smem_u32(stage_buf + (size_t)slot * STAGE_FLOATS);
        mbar_expect_tx(bar, (uint32_t)cnt * GRAN * 3 * 4);
        int p = rot0 + g / GPB; if (p >= a.total_blocks) p -= a.total_blocks;
        int off = g % GPB;
        while (cnt > 0) {
            const int take = min(GPB - off, cnt);
            const float* src = pos + (size_t)p * 3 * BLK + off * GRAN;
#pragma unroll
            for (int d = 0; d < 3; d++) bulk_g2s(dst + (uint32_t)(d * ROWF * 4), src + d * BLK, (uint32_t)take * GRAN * 4, bar);
            dst += (uint32_t)take * GRAN * 4; cnt -= take; off = 0;
            if (++p == a.total_blocks) p = 0;
        }
    };
    if (tid == ptid)
        for (int k = 0; k < LOOKAHEAD && k < nst; k++) issue(k);

    constexpr int IB = I * THREADS / BLK, TB = THREADS / BLK;
    const int lane_in_blk = tid % BLK;
    IState<I> s;
#pragma unroll
    for (int q = 0; q < I; q++) {
        const int ib = min(tile * IB + q * TB + tid / BLK, a.n_iblk - 1);   // idle threads of a ragged last tile alias a valid block
        const float* pi = pos + ((size_t)(a.i_blk0 + ib) * 3) * BLK + lane_in_blk;
        s.nx[q] = -pi[0]; s.ny[q] = -pi[BLK]; s.nz[q] = -pi[2 * BLK];
        s.ax[q] = s.ay[q] = s.az[q] = pk(0.f, 0.f);
    }

    for (int k = 0; k < nst; k++) {
        const int gk = k0 + k, slot = gk % NS;
        if (tid == ptid && k + LOOKAHEAD < nst) issue(k + LOOKAHEAD);
        mbar_wait(full0 + 8 * slot, (uint32_t)(gk / NS) & 1u);
        const int cnt = stage_cnt(k);
        const float* sb = stage_buf + (size_t)slot * STAGE_FLOATS;
        constexpr int ROW4 = ROWF / 4;
        const float4* sx = reinterpret_cast<const float4*>(sb);
        if (LOOP == 2) {
            // rotated loop over the flat rows: the UNROLL j-groups of iteration it+1 are loaded while iteration it
            // computes; this is the form sass_sched.py re-schedules
            float4 X[UNROLL], Y[UNROLL], Z[UNROLL];
#pragma unroll
            for (int g = 0; g < UNROLL; g++) { X[g] = sx[g]; Y[g] = sx[g + ROW4]; Z[g] = sx[g + 2 * ROW4]; }
            const int niter = cnt * (GRAN / (4 * UNROLL));
#pragma unroll 1
            for (int it = 0; it < niter; it++) {
#pragma unroll
                for (int g = 0; g < UNROLL; g++) interact4<I>(s, X[g], Y[g], Z[g], eps);
                sx += UNROLL;
#pragma unroll
                for (int g = 0; g < UNROLL; g++) { X[g] = sx[g]; Y[g] = sx[g + ROW4]; Z[g] = sx[g + 2 * ROW4]; }
            }
        } else {
            const int ng = cnt * (GRAN / 4);
#pragma unroll UNROLL
            for (int g = 0; g < ng; g++) {
                const float4 X = sx[g], Y = sx[g + ROW4], Z = sx[g + 2 * ROW4];
                interact4<I>(s, X, Y, Z, eps);
            }
        }
        fold_accumulators<I, THREADS>(s, acc2, tid);            // level 2: once per stage
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * slot);
        if ((k % CHAIN_STAGES) == CHAIN_STAGES - 1 || k == nst - 1) {   // level 3: every 65 536 j and at the end
#pragma unroll
            for (int q = 0; q < 3 * I; q++) {
                float lo, hi;
                upk(acc2[(size_t)q * THREADS + tid], lo, hi);
                res[(size_t)q * THREADS + tid] += lo + hi;
                acc2[(size_t)q * THREADS + tid] = pk(0.f, 0.f);
            }
        }
    }
    kbase = k0 + nst;
}

template <int I, int THREADS, int SG, int NS, int MINB, int LOOP, int UNROLL, bool EPS_RT>
__global__ void __launch_bounds__(THREADS, MINB) force_stream_f32_kernel(const StreamArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int STAGE_BYTES = 3 * SG * GRAN * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_BYTES + RING_PAD);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + NS + s), THREADS / 32); }
        fence_mbar_init();
    }
    __syncthreads();
    float* res = reinterpret_cast<float*>(smem_raw + (size_t)NS * STAGE_BYTES + RING_PAD + 2 * NS * 8 + (size_t)3 * I * THREADS * 8);
    int kbase = 0;
    stream_run<float, I, THREADS>(a, res, [&](int tile, int phase, int ja, int jb) {
        force_segment_f32<I, THREADS, SG, NS, LOOP, UNROLL, EPS_RT>(a, tile, a.ph_rot0[phase], ja, jb, smem_raw, kbase);
    });
}

// ---- variant table --------------------------------------------------------------------------------
//        id  name                  I  THREADS SB NS MINB packed pipe  fold unroll ctas/SM (hint for host-only planning)  run-time softening
#define NB_F32_VARIANTS(X)                                                      \
    X(0, "p_i4_t256",           4, 256, 4, 4, 1, true,  0, true, 2,  1, false)        \
    X(1, "p_i4_t128",           4, 128, 4, 4, 2, true,  0, true, 2,  2, false)        \
    X(2, "p_i2_t256",           2, 256, 4, 4, 2, true,  0, true, 2,  2, false)        \
    X(3, "p_i8_t128",           8, 128, 4, 4, 1, true,  0, true, 2,  2, false)        \
    X(4, "p_i2_t128",           2, 128, 4, 4, 4, true,  0, true, 2,  4, false)        \
    X(5, "scalar_i4_t256",      4, 256, 4, 4, 1, false, 0, true, 2,  1, false)        \
    X(6, "p_i1_t128",           1, 128, 2, 4, 4, true,  0, true, 2,  7, false)        \
    X(7, "p_i8_t128_nofold",    8, 128, 4, 4, 1, true,  0, false, 2, 2, false)        \
    X(8, "p_i12_t128",         12, 128, 4, 4, 1, true,  0, true, 2,  2, false)        \
    X(9, "p_i8_t128_pipe",      8, 128, 4, 4, 1, true,  1, true, 2,  2, false)        \
    X(10, "p_i6_t128",          6, 128, 4, 4, 2, true,  0, true, 2,  2, false)        \
    X(11, "p_i8_t128_sb8",      8, 128, 8, 4, 1, true,  0, true, 2,  2, false)        \
    X(12, "p_i8_t128_u1",       8, 128, 4, 4, 1, true,  0, true, 1,  2, false)        \
    X(13, "p_i8_t128_rot",      8, 128, 4, 4, 1, true,  2, true, 2,  2, false)        \
    X(14, "p_i8_t128_rot_u4",   8, 128, 4, 4, 1, true,  2, true, 4,  2, false)        \
    X(15, "p_i8_t128_rot_eps",  8, 128, 4, 4, 1, true,  2, true, 4,  2, true)         \
    X(16, "p_i2_t128_eps",      2, 128, 4, 4, 4, true,  0, true, 2,  4, true)         \
    X(17, "p_i1_t128_eps",      1, 128, 2, 4, 4, true,  0, true, 2,  7, true)         \
    X(18, "p_i8_t128_shfl",     8, 128, 4, 4, 1, true,  3, true, 2,  2, false) 

// stream-K instantiations (force_stream_f32_kernel); ids continue the table above
//        id  name                    I  THREADS SG NS MINB loop unroll ctas/SM  run-time softening
#define NB_F32_STREAM_VARIANTS(X)                                            \
    X(19, "s_i8_t128_rot_u4",     8, 128, 32, 4, 1, 2, 4,  2, false)            \
    X(20, "s_i8_t128_rot_eps",    8, 128, 32, 4, 1, 2, 4,  2, true)             \
    X(21, "s_i2_t128",            2, 128, 32, 4, 4, 0, 2,  4, false)            \
    X(22, "s_i2_t128_eps",        2, 128, 32, 4, 4, 0, 2,  4, true)             \
    X(23, "s_i4_t128",            4, 128, 32, 4, 2, 0, 2,  2, false)            \
    X(24, "s_i8_t128_u1",         8, 128, 32, 4, 1, 0, 1,  2, false)

static const ForceVariant g_variants[] = {
#define X(id, name, I, T, SB, NS, MINB, P, PIPE, FOLD, UNR, OCC, EPS) {name, I, T, SB, NS, P ? 1 : 0, OCC, FOLD ? 1 : 0, EPS ? 1 : 0, 0},
    NB_F32_VARIANTS(X)
#undef X
#define X(id, name, I, T, SG, NS, MINB, LOOP, UNR, OCC, EPS) {name, I, T, SG / GPB, NS, 1, OCC, 1, EPS ? 1 : 0, 1},
    NB_F32_STREAM_VARIANTS(X)
#undef X
};

int force_f32_num_variants() { return (int)(sizeof(g_variants) / sizeof(g_variants[0])); }
const ForceVariant& force_f32_variant(int v) { return g_variants[v]; }

static size_t smem_bytes(const ForceVariant& v) {
    if (v.stream)       // ring + pad + barriers + level-2 f32x2 sums + level-3 / result floats
        return (size_t)v.stages * v.stage_blocks * 3 * BLK * 4 + RING_PAD + 2 * v.stages * 8 + (size_t)v.i_per_thread * 3 * v.threads * 12;
    return (size_t)v.stages * v.stage_blocks * 3 * BLK * 4 + RING_PAD + 2 * v.stages * 8 + (v.fold ? (size_t)v.i_per_thread * 3 * v.threads * 8 : 0);
}

cudaError_t force_f32_setup(int variant) {
    cudaError_t e = cudaErrorInvalidValue;
    switch (variant) {
#define X(id, name, I, T, SB, NS, MINB, P, PIPE, FOLD, UNR, OCC, EPS) \
    case id: e = cudaFuncSetAttribute(force_f32_kernel<I, T, SB, NS, MINB, P, PIPE, FOLD, UNR, EPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(g_variants[id])); break;
        NB_F32_VARIANTS(X)
#undef X
#define X(id, name, I, T, SG, NS, MINB, LOOP, UNR, OCC, EPS) \
    case id: e = cudaFuncSetAttribute(force_stream_f32_kernel<I, T, SG, NS, MINB, LOOP, UNR, EPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(g_variants[id])); break;
        NB_F32_STREAM_VARIANTS(X)
#undef X
    }
    return e;
}

int force_f32_occupancy(int variant) {
    int nblk = 0;
    const size_t sm = smem_bytes(g_variants[variant]);
    cudaError_t e = cudaErrorInvalidValue;
    switch (variant) {
#define X(id, name, I, T, SB, NS, MINB, P, PIPE, FOLD, UNR, OCC, EPS) \
    case id: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, force_f32_kernel<I, T, SB, NS, MINB, P, PIPE, FOLD, UNR, EPS>, T, sm); break;
        NB_F32_VARIANTS(X)
#undef X
#define X(id, name, I, T, SG, NS, MINB, LOOP, UNR, OCC, EPS) \
    case id: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, force_stream_f32_kernel<I, T, SG, NS, MINB, LOOP, UNR, EPS>, T, sm); break;
        NB_F32_STREAM_VARIANTS(X)
#undef X
    }
    return e == cudaSuccess && nblk > 0 ? nblk : g_variants[variant].ctas_per_sm_hint;
}

cudaError_t force_f32_launch(int variant, const ForceArgs& a, cudaStream_t st) {
    if (variant < 0 || variant >= force_f32_num_variants()) return cudaErrorInvalidValue;
    const ForceVariant& v = g_variants[variant];
    if (v.stream) return cudaErrorInvalidValue;         // stream-K instantiations take StreamArgs (force_f32_stream_launch)
    const int ib = v.tile_bodies() / BLK;
    dim3 grid((a.n_iblk + ib - 1) / ib, a.nsplit, 1);
    if (grid.x == 0 || grid.y == 0 || a.j_len <= 0) return cudaSuccess;
    if (a.fuse || a.order == 1) grid = dim3(grid.x * grid.y, 1, 1);         // 1-D grid; order 1 = tile-major: blockIdx.x = tile * nsplit + split
    const size_t sm = smem_bytes(v);
    switch (variant) {
#define X(id, name, I, T, SB, NS, MINB, P, PIPE, FOLD, UNR, OCC, EPS) \
    case id: force_f32_kernel<I, T, SB, NS, MINB, P, PIPE, FOLD, UNR, EPS><<<grid, T, sm, st>>>(a); break;
        NB_F32_VARIANTS(X)
#undef X
    }
    return cudaGetLastError();
}

cudaError_t force_f32_stream_launch(int variant, const StreamArgs& a, cudaStream_t st) {
    if (variant < 0 || variant >= force_f32_num_variants() || !g_variants[variant].stream) return cudaErrorInvalidValue;
    if (a.grid <= 0 || a.i_tiles <= 0 || a.ph_begin >= a.ph_end) return cudaSuccess;
    const size_t sm = smem_bytes(g_variants[variant]);
    switch (variant) {
#define X(id, name, I, T, SG, NS, MINB, LOOP, UNR, OCC, EPS) \
    case id: force_stream_f32_kernel<I, T, SG, NS, MINB, LOOP, UNR, EPS><<<a.grid, T, sm, st>>>(a); break;
        NB_F32_STREAM_VARIANTS(X)
#undef X
    }
    return cudaGetLastError();
}

// twin of the in-kernel reduction (store_all passes): one CTA per tile, segments added in slot order
cudaError_t force_f32_stream_reduce_launch(int variant, const StreamArgs& a, cudaStream_t st) {
    if (variant < 0 || variant >= force_f32_num_variants() || !g_variants[variant].stream) return cudaErrorInvalidValue;
    if (a.i_tiles <= 0) return cudaSuccess;
    switch (variant) {
#define X(id, name, I, T, SG, NS, MINB, LOOP, UNR, OCC, EPS) \
    case id: stream_reduce_kernel<float, I, T><<<a.i_tiles, T, 0, st>>>(a); break;
        NB_F32_STREAM_VARIANTS(X)
#undef X
    }
    return cudaGetLastError();
}

// fused step kernel: instantiations mirror the narrow variants (the planner picks variant 6 / 17 below 6144 bodies)
//   variant 6 / 17 (I=1, SB=2)   variant 4 / 16 (I=2, SB=4)
template <int I, int SB, int MINB, bool EPS>
static cudaError_t fused_launch_t(const FusedStepArgs& fa, int sms, cudaStream_t st, int* grid_out) {
    auto kern = step_fused_f32_kernel<I, 128, SB, 4, MINB, EPS>;
    const size_t sm = (size_t)4 * SB * 3 * BLK * 4 + RING_PAD + 2 * 4 * 8 + (size_t)I * 3 * 128 * 8;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, sm);
    if (e != cudaSuccess) return e;
    const int units = fa.i_tiles * fa.nsplit;
    const int grid = units < occ * sms ? units : occ * sms;
    if (grid_out) *grid_out = grid;
    if (grid <= 0) return cudaErrorInvalidValue;
    void* args[] = {const_cast<FusedStepArgs*>(&fa)};
    return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(128), args, sm, st);
}

bool force_f32_fused_supported(int variant) { return variant == 6 || variant == 17 || variant == 4 || variant == 16; }

cudaError_t force_f32_fused_launch(int variant, const FusedStepArgs& fa, int sms, cudaStream_t st, int* grid_out) {
    switch (variant) {
        case 6: return fused_launch_t<1, 2, 4, false>(fa, sms, st, grid_out);
        case 17: return fused_launch_t<1, 2, 4, true>(fa, sms, st, grid_out);
        case 4: return fused_launch_t<2, 4, 4, false>(fa, sms, st, grid_out);
        case 16: return fused_launch_t<2, 4, 4, true>(fa, sms, st, grid_out);
    }
    return cudaErrorInvalidValue;
}

}  // namespace nb
