// K6: all time steps of a SMALL system in one cooperative launch, no partial sums between CTAs (FP32, one GPU).
//
// For N up to a few thousand bodies a step of the tiled kernels is a chain of dependent L2 round trips (bulk copy, partial
// store, tile counter, partial load, state store, grid barrier: ~6 us at N = 1024, 13.5 us at C1's N = 4096 for 5.4 us of
// arithmetic).  Here the decomposition is turned round: a CTA owns IPC = ceil(N / CTAs) i-bodies -- one per lane (IL per
// lane from 4737 bodies) -- and ALL of their interactions, so nothing is reduced across CTAs:
//   * the whole position array (12 B per body: 48 KB at C1) is brought into shared memory at the top of every step, one
//     1536 B bulk copy (TMA, UBLKCP) and one mbarrier per layout block, issued back to back by one thread;
//   * the 16 warps of the CTA share the same i-bodies (lane = body) and split the j-sweep: warp w takes the 4-body groups
//     [w*G/16, (w+1)*G/16) and starts as soon as ITS blocks have landed; same inner loop as K1 (interact4: packed f32x2 over
//     pairs of j, one broadcast LDS.128 per row and 4 j), register accumulators only (a slice is <= 512 j: chains of <= 256
//     adds, the length K1 folds at);
//   * the 16 partial sums per body meet in shared memory and are added in warp order (fixed order: deterministic), the owner
//     thread integrates -- velocities and the CTA's own positions live in registers for the whole call -- and stores the new
//     position; one grid barrier per step separates the writers of pos[next] from the bulk copies that read it.
// Per step: barrier + one L2 round trip for the first block + the arithmetic.  Reference analogue: the FPGA holds its 12
// i-bodies in registers and streams every j past them (S/top_level.vhd:44,233-249); here a CTA does that with 28-128.
#include <cooperative_groups.h>

#include "force_f32_inner.cuh"
#include "nbody_internal.cuh"

namespace nb {

constexpr int SMALL_WARPS = 16;

template <int IL, bool EPS_RT>
__global__ void __launch_bounds__(SMALL_WARPS * 32, 1) step_small_f32_kernel(const SmallStepArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int nblk = a.n_iblk;
    float* pos_s = reinterpret_cast<float*>(smem_raw);                                  // [nblk][3][BLK], the layout of HBM
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nblk * 3 * BLK * 4);          // one per block
    float* red = reinterpret_cast<float*>(bars + nblk);                                  // [WARPS][IL][3][32]
    float* ipos = red + SMALL_WARPS * IL * 3 * 32;                                       // [IL][3][32]: the CTA's own bodies
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const float eps = EPS_RT ? a.eps32 : EPS_F32;

    if (tid == 0) {
        for (int b = 0; b < nblk; b++) mbar_init(smem_u32(bars + b), 1);
        fence_mbar_init();
    }
    // bodies of this CTA: [i0, i0 + ipc) ; owner thread of body (q, lane) is thread q*32 + lane (warps 0 .. IL-1)
    const int i0 = blockIdx.x * a.ipc;
    const bool owner = warp < IL;
    const int my = i0 + warp * 32 + lane;                         // meaningful for owner threads only
    const bool live = owner && (warp * 32 + lane) < a.ipc && my < a.n;
    const size_t my_off = (size_t)(my / BLK) * 3 * BLK + (my % BLK);
    float x = PAD_F32, y = PAD_F32, z = PAD_F32, vx = 0.f, vy = 0.f, vz = 0.f;
    int cur = a.cur;
    if (live) {
        const float* p = static_cast<const float*>(a.pos[cur]) + my_off;
        const float* v = static_cast<const float*>(a.vel) + my_off;
        x = p[0]; y = p[BLK]; z = p[2 * BLK]; vx = v[0]; vy = v[BLK]; vz = v[2 * BLK];
    }
    if (owner) { ipos[(warp * 3 + 0) * 32 + lane] = x; ipos[(warp * 3 + 1) * 32 + lane] = y; ipos[(warp * 3 + 2) * 32 + lane] = z; }
    __syncthreads();

    // this warp's share of the j-sweep, in groups of 4 consecutive bodies
    const int groups = nblk * (BLK / 4);
    const int g0 = (int)(((long long)warp * groups) / SMALL_WARPS), g1 = (int)(((long long)(warp + 1) * groups) / SMALL_WARPS);

    for (int step = 0; step < a.nsteps; step++, cur ^= 1) {
        if (tid == 0) {
            const float* src = static_cast<const float*>(a.pos[cur]);
            for (int b = 0; b < nblk; b++) {
                const uint32_t bar = smem_u32(bars + b);
                mbar_expect_tx(bar, 3 * BLK * 4);
                bulk_g2s(smem_u32(pos_s + (size_t)b * 3 * BLK), src + (size_t)b * 3 * BLK, 3 * BLK * 4, bar);
            }
        }
        IState<IL> s;
#pragma unroll
        for (int q = 0; q < IL; q++) {
            s.nx[q] = -ipos[(q * 3 + 0) * 32 + lane]; s.ny[q] = -ipos[(q * 3 + 1) * 32 + lane]; s.nz[q] = -ipos[(q * 3 + 2) * 32 + lane];
            s.ax[q] = s.ay[q] = s.az[q] = pk(0.f, 0.f);
        }
        int g = g0;
        while (g < g1) {
            const int b = g / (BLK / 4), ge = min(g1, (b + 1) * (BLK / 4));
            mbar_wait(smem_u32(bars + b), (uint32_t)step & 1u);
            const float4* sx = reinterpret_cast<const float4*>(pos_s + (size_t)b * 3 * BLK) + (g - b * (BLK / 4));
            const int cnt = ge - g;
#pragma unroll 4
            for (int k = 0; k < cnt; k++) interact4<IL>(s, sx[k], sx[k + BLK / 4], sx[k + 2 * (BLK / 4)], eps);
            g = ge;
        }
#pragma unroll
        for (int q = 0; q < IL; q++) {
            float lo, hi;
            float* r = red + ((size_t)(warp * IL + q) * 3) * 32 + lane;
            upk(s.ax[q], lo, hi); r[0] = lo + hi;
            upk(s.ay[q], lo, hi); r[32] = lo + hi;
            upk(s.az[q], lo, hi); r[64] = lo + hi;
        }
        __syncthreads();                                          // all partial sums in place; every warp is done with pos_s and ipos
        if (owner) {
            float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
            for (int w = 0; w < SMALL_WARPS; w++) {               // fixed order => deterministic
                const float* r = red + ((size_t)(w * IL + warp) * 3) * 32 + lane;
                ax += r[0]; ay += r[32]; az += r[64];
            }
            if (live) {
                vx = fmaf(a.dt_v, ax, vx); vy = fmaf(a.dt_v, ay, vy); vz = fmaf(a.dt_v, az, vz);
                if (a.write_pos) {                                 // 0: kick only (bodyForce: v += dt * F, positions stay)
                    x = fmaf(vx, a.dt_x, x); y = fmaf(vy, a.dt_x, y); z = fmaf(vz, a.dt_x, z);
                    float* pn = static_cast<float*>(a.pos[cur ^ 1]) + my_off;
                    pn[0] = x; pn[BLK] = y; pn[2 * BLK] = z;
                }
            }
            ipos[(warp * 3 + 0) * 32 + lane] = x; ipos[(warp * 3 + 1) * 32 + lane] = y; ipos[(warp * 3 + 2) * 32 + lane] = z;
        }
        // pos[next] complete and visible device-wide, also to the bulk copies (async proxy) of the next step
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        grid.sync();
    }
    if (live) {
        float* v = static_cast<float*>(a.vel) + my_off;
        v[0] = vx; v[BLK] = vy; v[2 * BLK] = vz;
    }
}

static size_t small_smem_bytes(int nblk, int il) {
    return (size_t)nblk * 3 * BLK * 4 + (size_t)nblk * 8 + (size_t)(SMALL_WARPS + 1) * il * 3 * 32 * 4;
}

// bodies per CTA and CTAs for n bodies on `sms` SMs: one CTA per SM, <= 32 * SMALL_MAX_IL bodies each
bool step_small_plan(int n, int sms, int* ipc, int* ctas, int* il) {
    if (n <= 0 || sms <= 0) return false;
    const int per = (n + sms - 1) / sms;
    const int l = (per + 31) / 32;
    if (l > 4) return false;
    *il = l == 3 ? 4 : l;
    *ipc = per; *ctas = (n + per - 1) / per;
    const int nblk = (n + BLK - 1) / BLK;
    return small_smem_bytes(nblk, *il) <= (size_t)200 * 1024;
}

template <int IL, bool EPS>
static cudaError_t small_launch_t(const SmallStepArgs& a, int ctas, cudaStream_t st) {
    auto kern = step_small_f32_kernel<IL, EPS>;
    const size_t sm = small_smem_bytes(a.n_iblk, IL);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    void* args[] = {const_cast<SmallStepArgs*>(&a)};
    return cudaLaunchCooperativeKernel((const void*)kern, dim3(ctas), dim3(SMALL_WARPS * 32), args, sm, st);
}

cudaError_t step_small_launch(const SmallStepArgs& a, int ctas, int il, bool eps_rt, cudaStream_t st) {
    switch (il * 2 + (eps_rt ? 1 : 0)) {
        case 2: return small_launch_t<1, false>(a, ctas, st);
        case 3: return small_launch_t<1, true>(a, ctas, st);
        case 4: return small_launch_t<2, false>(a, ctas, st);
        case 5: return small_launch_t<2, true>(a, ctas, st);
        case 8: return small_launch_t<4, false>(a, ctas, st);
        case 9: return small_launch_t<4, true>(a, ctas, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace nb
