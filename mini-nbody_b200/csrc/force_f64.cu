// K2: all-pairs softened-gravity force, FP64 accuracy path (DFMA-bound), sm_100a.
//
// Same decomposition and TMA/mbarrier j-ring as K1 (force_f32.cu); per-pair dataflow of the
// reference pipeline (dxy.vhd:94-122, dzsoft.vhd:177-202, dxyz_soft.vhd:149-150, fxyz.vhd:101-127,
// cube.vhd:66-70) in binary64.  rsqrt and cube are one step: y0 = MUFU.RSQ64H seed (rel. error d ~ 2^-22),
// e = 1 - s*y0^2, and s^(-3/2) = y0^3 (1 - e)^(-3/2) = y0^3 (1 + 3e/2 + 15e^2/8 + 35e^3/16 ...), cut after
// the e^2 term (|e| < 2^-20 => truncation < 3e-18, far below one ulp): u = y0*y0, e = fma(-s,u,1), c = u*y0,
// p = fma(e,15/8,3/2), w = c*e, r3 = fma(w,p,c) = 6 DP ops for rsqrt AND cube (refining y first and cubing it
// afterwards took 7; 1.0/sqrt() ~25).  Per interaction: 3 DADD + 3 DFMA (dist^2) + 6 + 3 DFMA = 15 FP64-pipe
// ops + 1 MUFU; B200 sustains ~59 DFMA lane-ops/clk/SM (profiles/r01_microbench.md) => ceiling ~1.14e12
// interactions/s.
#include "nbody_internal.cuh"
#include "stream.cuh"

namespace nb {

template <int I, int THREADS, int SB, int NS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) force_f64_kernel(const ForceArgs a) {
    static_assert(NS >= 3, "need >= 3 stages");
    constexpr int STAGE_ELEMS = SB * 3 * BLK;
    constexpr int STAGE_BYTES = STAGE_ELEMS * 8;
    constexpr int NWARPS = THREADS / 32;
    constexpr int LOOKAHEAD = NS - 2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_buf = reinterpret_cast<double*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NS);

    const int tid = threadIdx.x;
    const double* __restrict__ pos = static_cast<const double*>(a.pos);
    const double eps = a.eps64;                       // 1e-9 unless nbody_set_softening changed it
    const int split = blockIdx.y;
    const int jb0 = (int)(((long long)split * a.j_len) / a.nsplit);
    const int jb1 = (int)(((long long)(split + 1) * a.j_len) / a.nsplit);
    const int ntiles = (jb1 - jb0 + SB - 1) / SB;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NWARPS); }
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int k) {
        const int st = k % NS;
        const int rb = jb0 + k * SB;
        const int cnt = min(SB, jb1 - rb);
        int p = a.j_rot0 + rb; if (p >= a.total_blocks) p -= a.total_blocks;
        const uint32_t bar = full0 + 8 * st;
        const uint32_t dst = smem_u32(stage_buf + (size_t)st * STAGE_ELEMS);
        mbar_expect_tx(bar, (uint32_t)cnt * 3 * BLK * 8);
        const int first = min(cnt, a.total_blocks - p);
        bulk_g2s(dst, pos + (size_t)p * 3 * BLK, (uint32_t)first * 3 * BLK * 8, bar);
        if (first < cnt)
            bulk_g2s(dst + (uint32_t)first * 3 * BLK * 8, pos, (uint32_t)(cnt - first) * 3 * BLK * 8, bar);
    };
    if (tid == 0)
        for (int k = 0; k < LOOKAHEAD && k < ntiles; k++) issue(k);

    constexpr int IB = I * THREADS / BLK;
    constexpr int TB = THREADS / BLK;
    const int lane_in_blk = tid % BLK;
    double xi[I], yi[I], zi[I], ax[I], ay[I], az[I];
    int iblk[I];
#pragma unroll
    for (int q = 0; q < I; q++) {
        iblk[q] = blockIdx.x * IB + q * TB + tid / BLK;
        const int ib = min(iblk[q], a.n_iblk - 1);
        const double* pi = pos + ((size_t)(a.i_blk0 + ib) * 3) * BLK + lane_in_blk;
        xi[q] = pi[0]; yi[q] = pi[BLK]; zi[q] = pi[2 * BLK];
        ax[q] = ay[q] = az[q] = 0.0;
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
        const double* sb = stage_buf + (size_t)st * STAGE_ELEMS;
        for (int b = 0; b < cnt; b++) {
            const double2* sx = reinterpret_cast<const double2*>(sb + b * 3 * BLK);
#pragma unroll 2
            for (int g = 0; g < BLK / 2; g++) {
                const double2 X = sx[g], Y = sx[g + BLK / 2], Z = sx[g + 2 * (BLK / 2)];
                const double xs[2] = {X.x, X.y}, ys[2] = {Y.x, Y.y}, zs[2] = {Z.x, Z.y};
#pragma unroll
                for (int q = 0; q < I; q++) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const double dx = xs[h] - xi[q], dy = ys[h] - yi[q], dz = zs[h] - zi[q];
                        double s = fma(dx, dx, eps); s = fma(dy, dy, s); s = fma(dz, dz, s);
                        const double y0 = rsqrt_approx64(s);          // MUFU.RSQ64H: s^(-1/2) (1 + d), |d| ~ 2^-22
                        const double u = y0 * y0;
                        const double e = fma(-s, u, 1.0);             // e = 1 - s*y0^2 = -2d - d^2
                        const double c = u * y0;                      // y0^3
                        const double p = fma(e, 1.875, 1.5);
                        const double w = c * e;
                        const double r3 = fma(w, p, c);               // y0^3 (1 + 3e/2 + 15e^2/8) = s^(-3/2) (1 + O(e^3))
                        ax[q] = fma(dx, r3, ax[q]); ay[q] = fma(dy, r3, ay[q]); az[q] = fma(dz, r3, az[q]);
                    }
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * st);
    }

    double* __restrict__ part = static_cast<double*>(a.part) + (size_t)(a.slot0 + split) * a.n_iblk * 3 * BLK;
#pragma unroll
    for (int q = 0; q < I; q++) {
        if (iblk[q] < a.n_iblk) {
            double* o = part + (size_t)iblk[q] * 3 * BLK + lane_in_blk;
            o[0] = ax[q]; o[BLK] = ay[q]; o[2 * BLK] = az[q];
        }
    }
}

// ---- stream-K force pass, FP64 (see StreamArgs / stream.cuh and force_segment_f32) -------------------------
// Stage = SG granules of 16 j, row-major [X: SG*16][Y][Z] doubles; register accumulators over the whole segment
// (binary64 chains need no second level), handed to the stream driver through `res` in shared memory.
template <int I, int THREADS, int SG, int NS>
__device__ __forceinline__ void force_segment_f64(const StreamArgs& a, const int tile, const int rot0, const int ja, const int jb,
                                                  unsigned char* smem_raw, int& kbase) {
    static_assert(NS >= 3, "need >= 3 stages");
    constexpr int ROWE = SG * GRAN;
    constexpr int STAGE_ELEMS = 3 * ROWE;
    constexpr int STAGE_BYTES = STAGE_ELEMS * 8;
    constexpr int LOOKAHEAD = NS - 2;
    double* stage_buf = reinterpret_cast<double*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NS);
    double* res = reinterpret_cast<double*>(smem_raw + (size_t)NS * STAGE_BYTES + 2 * NS * 8);

    const int tid = threadIdx.x;
    const double* __restrict__ pos = static_cast<const double*>(a.pos);
    const double eps = a.eps64;
    const int nst = (jb - ja + SG - 1) / SG;
    const int k0 = kbase;
    auto issue = [&](int k) {
        const int gk = k0 + k, slot = gk % NS;
        if (gk >= NS) mbar_wait(empty0 + 8 * slot, (uint32_t)((gk / NS) - 1) & 1u);
        int g = ja + k * SG;
        int cnt = min(SG, jb - g);
        const uint32_t bar = full0 + 8 * slot;
        uint32_t dst = smem_u32(stage_buf + (size_t)slot * STAGE_ELEMS);
        mbar_expect_tx(bar, (uint32_t)cnt * GRAN * 3 * 8);
        int p = rot0 + g / GPB; if (p >= a.total_blocks) p -= a.total_blocks;
        int off = g % GPB;
        while (cnt > 0) {
            const int take = min(GPB - off, cnt);
            const double* src = pos + (size_t)p * 3 * BLK + off * GRAN;
#pragma unroll
            for (int d = 0; d < 3; d++) bulk_g2s(dst + (uint32_t)(d * ROWE * 8), src + d * BLK, (uint32_t)take * GRAN * 8, bar);
            dst += (uint32_t)take * GRAN * 8; cnt -= take; off = 0;
            if (++p == a.total_blocks) p = 0;
        }
    };
    if (tid == 0)
        for (int k = 0; k < LOOKAHEAD && k < nst; k++) issue(k);

    constexpr int IB = I * THREADS / BLK, TB = THREADS / BLK;
    const int lane_in_blk = tid % BLK;
    double xi[I], yi[I], zi[I], ax[I], ay[I], az[I];
#pragma unroll
    for (int q = 0; q < I; q++) {
        const int ib = min(tile * IB + q * TB + tid / BLK, a.n_iblk - 1);
        const double* pi = pos + ((size_t)(a.i_blk0 + ib) * 3) * BLK + lane_in_blk;
        xi[q] = pi[0]; yi[q] = pi[BLK]; zi[q] = pi[2 * BLK];
        ax[q] = ay[q] = az[q] = 0.0;
    }

    for (int k = 0; k < nst; k++) {
        const int gk = k0 + k, slot = gk % NS;
        if (tid == 0 && k + LOOKAHEAD < nst) issue(k + LOOKAHEAD);
        mbar_wait(full0 + 8 * slot, (uint32_t)(gk / NS) & 1u);
        const int cnt = min(SG, jb - (ja + k * SG));
        constexpr int ROW2 = ROWE / 2;
        const double2* sx = reinterpret_cast<const double2*>(stage_buf + (size_t)slot * STAGE_ELEMS);
        const int ng = cnt * (GRAN / 2);
#pragma unroll 2
        for (int g = 0; g < ng; g++) {
            const double2 X = sx[g], Y = sx[g + ROW2], Z = sx[g + 2 * ROW2];
            const double xs[2] = {X.x, X.y}, ys[2] = {Y.x, Y.y}, zs[2] = {Z.x, Z.y};
#pragma unroll
            for (int q = 0; q < I; q++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const double dx = xs[h] - xi[q], dy = ys[h] - yi[q], dz = zs[h] - zi[q];
                    double s = fma(dx, dx, eps); s = fma(dy, dy, s); s = fma(dz, dz, s);
                    const double y0 = rsqrt_approx64(s);          // MUFU.RSQ64H: s^(-1/2) (1 + d), |d| ~ 2^-22
                    const double u = y0 * y0;
                    const double e = fma(-s, u, 1.0);
                    const double c = u * y0;
                    const double p = fma(e, 1.875, 1.5);
                    const double w = c * e;
                    const double r3 = fma(w, p, c);               // s^(-3/2) (1 + O(e^3)), see the header of this file
                    ax[q] = fma(dx, r3, ax[q]); ay[q] = fma(dy, r3, ay[q]); az[q] = fma(dz, r3, az[q]);
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * slot);
    }
#pragma unroll
    for (int q = 0; q < I; q++) {
        res[(size_t)(q * 3 + 0) * THREADS + tid] = ax[q];
        res[(size_t)(q * 3 + 1) * THREADS + tid] = ay[q];
        res[(size_t)(q * 3 + 2) * THREADS + tid] = az[q];
    }
    kbase = k0 + nst;
}

template <int I, int THREADS, int SG, int NS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) force_stream_f64_kernel(const StreamArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int STAGE_BYTES = 3 * SG * GRAN * 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_BYTES);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + NS + s), THREADS / 32); }
        fence_mbar_init();
    }
    __syncthreads();
    double* res = reinterpret_cast<double*>(smem_raw + (size_t)NS * STAGE_BYTES + 2 * NS * 8);
    int kbase = 0;
    stream_run<double, I, THREADS>(a, res, [&](int tile, int phase, int ja, int jb) {
        force_segment_f64<I, THREADS, SG, NS>(a, tile, a.ph_rot0[phase], ja, jb, smem_raw, kbase);
    });
}

// variant sweep at C3 (profiles/r01_f64_variant_sweep.jsonl): every shape lands within 964-1021 G inter/s -- the
// FP64 pipe, not the schedule, is the limit; I=4 x 256 threads (one CTA per SM) is 1.6 % ahead at N = 65 536
#define NB_F64_VARIANTS(X)                        \
    X(0, "d_i2_t256_s2x4", 2, 256, 2, 4, 2, 2)       \
    X(1, "d_i4_t128_s2x4", 4, 128, 2, 4, 2, 2)       \
    X(2, "d_i1_t128_s2x4", 1, 128, 2, 4, 4, 4)       \
    X(3, "d_i2_t128_s2x4", 2, 128, 2, 4, 4, 4)       \
    X(4, "d_i4_t256_s2x4", 4, 256, 2, 4, 1, 1)

// stream-K instantiations (force_stream_f64_kernel); ids continue the table above
//        id  name             I  THREADS SG NS MINB ctas/SM
#define NB_F64_STREAM_VARIANTS(X)                  \
    X(5, "ds_i4_t256",   4, 256, 16, 4, 1, 1)        \
    X(6, "ds_i4_t128",   4, 128, 16, 4, 2, 2)        \
    X(7, "ds_i2_t128",   2, 128, 16, 4, 4, 4)

static const ForceVariant g_variants64[] = {
#define X(id, name, I, T, SB, NS, MINB, OCC) {name, I, T, SB, NS, 0, OCC, 0, 1, 0},
    NB_F64_VARIANTS(X)
#undef X
#define X(id, name, I, T, SG, NS, MINB, OCC) {name, I, T, SG / GPB, NS, 0, OCC, 0, 1, 1},
    NB_F64_STREAM_VARIANTS(X)
#undef X
};
int force_f64_num_variants() { return (int)(sizeof(g_variants64) / sizeof(g_variants64[0])); }
const ForceVariant& force_f64_variant(int v) { return g_variants64[v]; }
static size_t smem_bytes64(const ForceVariant& v) {
    if (v.stream) return (size_t)v.stages * v.stage_blocks * 3 * BLK * 8 + 2 * v.stages * 8 + (size_t)v.i_per_thread * 3 * v.threads * 8;
    return (size_t)v.stages * v.stage_blocks * 3 * BLK * 8 + 2 * v.stages * 8;
}

cudaError_t force_f64_setup(int variant) {
    cudaError_t e = cudaErrorInvalidValue;
    switch (variant) {
#define X(id, name, I, T, SB, NS, MINB, OCC) \
    case id: e = cudaFuncSetAttribute(force_f64_kernel<I, T, SB, NS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes64(g_variants64[id])); break;
        NB_F64_VARIANTS(X)
#undef X
#define X(id, name, I, T, SG, NS, MINB, OCC) \
    case id: e = cudaFuncSetAttribute(force_stream_f64_kernel<I, T, SG, NS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes64(g_variants64[id])); break;
        NB_F64_STREAM_VARIANTS(X)
#undef X
    }
    return e;
}

int force_f64_occupancy(int variant) {
    int nblk = 0;
    const size_t sm = smem_bytes64(g_variants64[variant]);
    cudaError_t e = cudaErrorInvalidValue;
    switch (variant) {
#define X(id, name, I, T, SB, NS, MINB, OCC) \
    case id: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, force_f64_kernel<I, T, SB, NS, MINB>, T, sm); break;
        NB_F64_VARIANTS(X)
#undef X
#define X(id, name, I, T, SG, NS, MINB, OCC) \
    case id: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, force_stream_f64_kernel<I, T, SG, NS, MINB>, T, sm); break;
        NB_F64_STREAM_VARIANTS(X)
#undef X
    }
    return e == cudaSuccess && nblk > 0 ? nblk : g_variants64[variant].ctas_per_sm_hint;
}

cudaError_t force_f64_launch(int variant, const ForceArgs& a, cudaStream_t st) {
    if (variant < 0 || variant >= force_f64_num_variants()) return cudaErrorInvalidValue;
    const ForceVariant& v = g_variants64[variant];
    if (v.stream) return cudaErrorInvalidValue;         // stream-K instantiations take StreamArgs (force_f64_stream_launch)
    const int ib = v.tile_bodies() / BLK;
    dim3 grid((a.n_iblk + ib - 1) / ib, a.nsplit, 1);
    if (grid.x == 0 || grid.y == 0 || a.j_len <= 0) return cudaSuccess;
    const size_t sm = smem_bytes64(v);
    switch (variant) {
#define X(id, name, I, T, SB, NS, MINB, OCC) \
    case id: force_f64_kernel<I, T, SB, NS, MINB><<<grid, T, sm, st>>>(a); break;
        NB_F64_VARIANTS(X)
#undef X
    }
    return cudaGetLastError();
}

cudaError_t force_f64_stream_launch(int variant, const StreamArgs& a, cudaStream_t st) {
    if (variant < 0 || variant >= force_f64_num_variants() || !g_variants64[variant].stream) return cudaErrorInvalidValue;
    if (a.grid <= 0 || a.i_tiles <= 0 || a.ph_begin >= a.ph_end) return cudaSuccess;
    const size_t sm = smem_bytes64(g_variants64[variant]);
    switch (variant) {
#define X(id, name, I, T, SG, NS, MINB, OCC) \
    case id: force_stream_f64_kernel<I, T, SG, NS, MINB><<<a.grid, T, sm, st>>>(a); break;
        NB_F64_STREAM_VARIANTS(X)
#undef X
    }
    return cudaGetLastError();
}

cudaError_t force_f64_stream_reduce_launch(int variant, const StreamArgs& a, cudaStream_t st) {
    if (variant < 0 || variant >= force_f64_num_variants() || !g_variants64[variant].stream) return cudaErrorInvalidValue;
    if (a.i_tiles <= 0) return cudaSuccess;
    switch (variant) {
#define X(id, name, I, T, SG, NS, MINB, OCC) \
    case id: stream_reduce_kernel<double, I, T><<<a.i_tiles, T, 0, st>>>(a); break;
        NB_F64_STREAM_VARIANTS(X)
#undef X
    }
    return cudaGetLastError();
}

}  // namespace nb
