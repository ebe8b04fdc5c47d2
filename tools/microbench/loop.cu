// Third microbenchmark: the force kernel's own inner loop (force_f32_inner.cuh) over a resident
// shared-memory tile, no TMA / mbarrier / epilogue: the ceiling of the instruction mix itself as a
// function of I (i-bodies per thread) and warps per SM sub-partition.  Evidence only.
#include <cstdio>
#include <cuda_runtime.h>
#include "force_f32_sched.cuh"
using namespace nb;

// experiment modes: 0 = product loop; 1 = no MUFU (r := d2); 2 = one MUFU per pair (hi half reuses lo)
template <int I, int MODE>
__device__ __forceinline__ void interact4x(IState<I>& s, const float4 X, const float4 Y, const float4 Z) {
    const f2 eps2 = pk(EPS_F32, EPS_F32);
    const f2 xs[2] = {pk(X.x, X.y), pk(X.z, X.w)}, ys[2] = {pk(Y.x, Y.y), pk(Y.z, Y.w)}, zs[2] = {pk(Z.x, Z.y), pk(Z.z, Z.w)};
#pragma unroll
    for (int i = 0; i < I; i++) {
        const f2 nx2 = pk(s.nx[i], s.nx[i]), ny2 = pk(s.ny[i], s.ny[i]), nz2 = pk(s.nz[i], s.nz[i]);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const f2 dx = add2(xs[h], nx2), dy = add2(ys[h], ny2), dz = add2(zs[h], nz2);
            f2 d2 = fma2(dx, dx, eps2); d2 = fma2(dy, dy, d2); d2 = fma2(dz, dz, d2);
            float d2lo, d2hi; upk(d2, d2lo, d2hi);
            f2 r;
            if (MODE == 1) r = d2;
            else if (MODE == 2) { const float q = rsqrt_approx(d2lo); r = pk(q, d2hi); }
            else r = pk(rsqrt_approx(d2lo), rsqrt_approx(d2hi));
            const f2 r3 = mul2(mul2(r, r), r);
            s.ax[i] = fma2(dx, r3, s.ax[i]); s.ay[i] = fma2(dy, r3, s.ay[i]); s.az[i] = fma2(dz, r3, s.az[i]);
        }
    }
}

// MODE 4: pair-major order (all i for j-pair A, then all i for pair B)
// MODE 5: stage-major: all distances+dist^2 first, then all rsqrt, then all cubes+accumulates
// MODE 6: as 5 but per j-pair (half the live registers)
template <int I, int MODE>
__device__ __forceinline__ void interact4y(IState<I>& s, const float4 X, const float4 Y, const float4 Z) {
    const f2 eps2 = pk(EPS_F32, EPS_F32);
    const f2 xs[2] = {pk(X.x, X.y), pk(X.z, X.w)}, ys[2] = {pk(Y.x, Y.y), pk(Y.z, Y.w)}, zs[2] = {pk(Z.x, Z.y), pk(Z.z, Z.w)};
    if (MODE == 4) {
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int i = 0; i < I; i++) {
                const f2 dx = add2(xs[h], pk(s.nx[i], s.nx[i])), dy = add2(ys[h], pk(s.ny[i], s.ny[i])), dz = add2(zs[h], pk(s.nz[i], s.nz[i]));
                f2 d2 = fma2(dx, dx, eps2); d2 = fma2(dy, dy, d2); d2 = fma2(dz, dz, d2);
                float lo, hi; upk(d2, lo, hi);
                const f2 r = pk(rsqrt_approx(lo), rsqrt_approx(hi));
                const f2 r3 = mul2(mul2(r, r), r);
                s.ax[i] = fma2(dx, r3, s.ax[i]); s.ay[i] = fma2(dy, r3, s.ay[i]); s.az[i] = fma2(dz, r3, s.az[i]);
            }
    } else {
#pragma unroll
        for (int h0 = 0; h0 < 2; h0 += (MODE == 5 ? 2 : 1)) {
            constexpr int NH = MODE == 5 ? 2 : 1;
            f2 dx[NH][I], dy[NH][I], dz[NH][I], q[NH][I];
#pragma unroll
            for (int hh = 0; hh < NH; hh++)
#pragma unroll
                for (int i = 0; i < I; i++) {
                    const int h = h0 + hh;
                    dx[hh][i] = add2(xs[h], pk(s.nx[i], s.nx[i])); dy[hh][i] = add2(ys[h], pk(s.ny[i], s.ny[i])); dz[hh][i] = add2(zs[h], pk(s.nz[i], s.nz[i]));
                    f2 d2 = fma2(dx[hh][i], dx[hh][i], eps2); d2 = fma2(dy[hh][i], dy[hh][i], d2); q[hh][i] = fma2(dz[hh][i], dz[hh][i], d2);
                }
#pragma unroll
            for (int hh = 0; hh < NH; hh++)
#pragma unroll
                for (int i = 0; i < I; i++) { float lo, hi; upk(q[hh][i], lo, hi); q[hh][i] = pk(rsqrt_approx(lo), rsqrt_approx(hi)); }
#pragma unroll
            for (int hh = 0; hh < NH; hh++)
#pragma unroll
                for (int i = 0; i < I; i++) {
                    const f2 r3 = mul2(mul2(q[hh][i], q[hh][i]), q[hh][i]);
                    s.ax[i] = fma2(dx[hh][i], r3, s.ax[i]); s.ay[i] = fma2(dy[hh][i], r3, s.ay[i]); s.az[i] = fma2(dz[hh][i], r3, s.az[i]);
                }
        }
    }
}

template <int I, int THREADS, int MINB, int MODE>
__global__ void __launch_bounds__(THREADS, MINB) k_loop(float* out, int reps, int blocks) {
    extern __shared__ __align__(128) float tile[];
    for (int t = threadIdx.x; t < blocks * 3 * BLK; t += THREADS) tile[t] = (float)((t * 2654435761u) >> 8) * (1.f / 16777216.f);
    __syncthreads();
    IState<I> s;
#pragma unroll
    for (int q = 0; q < I; q++) { s.nx[q] = -0.01f * (threadIdx.x + q); s.ny[q] = tile[(threadIdx.x * 7 + q) % (blocks * 3 * BLK)]; s.nz[q] = tile[(threadIdx.x * 13 + q * 5) % (blocks * 3 * BLK)]; s.ax[q] = s.ay[q] = s.az[q] = pk(0.f, 0.f); }
    for (int r = 0; r < reps; r++) {
        if (MODE == 3) { sched_tile<I>(s, smem_u32(tile), blocks); continue; }
        for (int b = 0; b < blocks; b++) {
            const float4* sx = reinterpret_cast<const float4*>(tile + b * 3 * BLK);
#pragma unroll 2
            for (int g = 0; g < BLK / 4; g++) {
                const float4 X = sx[g], Y = sx[g + BLK / 4], Z = sx[g + 2 * (BLK / 4)];
                if (MODE >= 4) interact4y<I, MODE>(s, X, Y, Z); else interact4x<I, MODE>(s, X, Y, Z);
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int q = 0; q < I; q++) { float lo, hi; upk(s.ax[q], lo, hi); acc += lo + hi; upk(s.ay[q], lo, hi); acc += lo + hi; upk(s.az[q], lo, hi); acc += lo + hi; }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

static int g_sms; static float* g_out;
template <int I, int THREADS, int MINB, int MODE>
static void run(int ctas_per_sm) {
    const int blocks = 4, reps = 400;
    const size_t sm = blocks * 3 * BLK * 4;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_loop<I, THREADS, MINB, MODE>, THREADS, sm);
    if (ctas_per_sm > occ) return;
    const int grid = g_sms * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int t = 0; t < 4; t++) {
        cudaEventRecord(e0); k_loop<I, THREADS, MINB, MODE><<<grid, THREADS, sm>>>(g_out, reps, blocks); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (t && ms < best) best = ms;
    }
    const double inter = (double)grid * THREADS * I * (double)reps * blocks * BLK;
    const double rate = inter / (best * 1e-3);
    printf("{\"test\": \"loop\", \"mode\": %d, \"I\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"warps_per_smsp\": %.1f, \"occ_max\": %d, \"ms\": %.3f, \"G_inter_s\": %.1f, \"cyc_per_inter@1965\": %.3f}\n",
           MODE, I, THREADS, ctas_per_sm, ctas_per_sm * THREADS / 128.0, occ, best, rate / 1e9, g_sms * 128 * 1.965e9 / rate);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); g_sms = p.multiProcessorCount;
    cudaMalloc(&g_out, sizeof(float) * g_sms * 16 * 256);
    for (int c = 1; c <= 2; c++) { run<8, 128, 1, 0>(c); }
    for (int c = 1; c <= 2; c++) { run<8, 128, 1, 1>(c); }
    for (int c = 2; c <= 3; c++) { run<4, 128, 1, 0>(c); }
    for (int c = 2; c <= 2; c++) { run<12, 128, 1, 0>(c); }
    printf("{\"done\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
