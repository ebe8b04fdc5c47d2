// Fifth microbenchmark: do the FP32 pipe and the FP64 pipe of a B200 SM sub-partition run side by side?
// K1 is bound by the FP32 pipe / register file (12.0 cycles per interaction) while the half-rate FP64 pipe
// (58.7 DFMA lane-ops/clk/SM, pipes.cu) idles.  If warps that evaluate interactions in FP64 arithmetic
// (3 DADD + 3 DFMA + MUFU.RSQ64H + 2 DMUL + 3 DFMA, no refinement: FP32-level accuracy) can be resident
// beside the FP32 warps without slowing them down, a hybrid kernel gains their throughput on top.
//   test "pure":  FFMA2 stream warps + DFMA stream warps in one CTA (register-file / dispatch sharing)
//   test "loop":  the product FP32 inner loop (I=4 / I=8) beside an FP64 interaction loop over the same tile
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../mini-nbody_b200/csrc -o dual dual.cu
// Evidence only (output committed under profiles/).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "force_f32_inner.cuh"
using namespace nb;

__device__ __forceinline__ f2 fma2v(f2 a, f2 b, f2 c) { f2 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ double dfmav(double a, double b, double c) { double r; asm volatile("fma.rn.f64 %0,%1,%2,%3;" : "=d"(r) : "d"(a), "d"(b), "d"(c)); return r; }

// ---- pure streams: warps [0, w32) run FFMA2 chains, warps [w32, w32 + w64) run DFMA chains -----------------
__global__ void __launch_bounds__(512) k_pure(float* out, long long* clk, int w32, int it32, int it64, float b, float c) {
    const int warp = threadIdx.x >> 5;
    const long long t0 = clock64();
    float res = 0.f;
    if (warp < w32) {
        f2 a[8]; const f2 b2 = pk(b, b), c2 = pk(c, c);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = pk(threadIdx.x * 1e-3f + i, 1.f + i);
        for (int it = 0; it < it32; it++) {
#pragma unroll
            for (int u = 0; u < 32; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = fma2v(a[i], b2, c2);
        }
        for (int i = 0; i < 8; i++) { float x, y; upk(a[i], x, y); res += x + y; }
    } else {
        double a[8]; const double bd = b, cd = c;
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < it64; it++) {
#pragma unroll
            for (int u = 0; u < 32; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = dfmav(a[i], bd, cd);
        }
        for (int i = 0; i < 8; i++) res += (float)a[i];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)(clk + (warp < w32 ? 0 : 1)), (unsigned long long)(t1 - t0));
}

// ---- loops: FP32 product loop (I32 i-bodies per thread) beside an FP64 loop (I64 i-bodies per thread) --------
template <int I64>
__device__ __forceinline__ void interact2_f64(double (&xi)[I64], double (&yi)[I64], double (&zi)[I64],
                                              double (&ax)[I64], double (&ay)[I64], double (&az)[I64],
                                              const double2 X, const double2 Y, const double2 Z) {
    const double xs[2] = {X.x, X.y}, ys[2] = {Y.x, Y.y}, zs[2] = {Z.x, Z.y};
#pragma unroll
    for (int q = 0; q < I64; q++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double dx = xs[h] - xi[q], dy = ys[h] - yi[q], dz = zs[h] - zi[q];
            double s = fma(dx, dx, 1e-9); s = fma(dy, dy, s); s = fma(dz, dz, s);
            const double r = rsqrt_approx64(s);              // MUFU.RSQ64H, rel. error ~2^-22: FP32-level accuracy
            const double r3 = (r * r) * r;
            ax[q] = fma(dx, r3, ax[q]); ay[q] = fma(dy, r3, ay[q]); az[q] = fma(dz, r3, az[q]);
        }
}

template <int I32, int I64, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_loop2(float* out, long long* clk, int w32, int reps32, int reps64, int blocks) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* tile = reinterpret_cast<float*>(smem);
    double* dtile = reinterpret_cast<double*>(smem + (size_t)blocks * 3 * BLK * 4);
    for (int t = threadIdx.x; t < blocks * 3 * BLK; t += THREADS) {
        tile[t] = (float)((t * 2654435761u) >> 8) * (1.f / 16777216.f);
        dtile[t] = (double)tile[t];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const long long t0 = clock64();
    float res = 0.f;
    if (warp < w32) {
        IState<I32> s;
#pragma unroll
        for (int q = 0; q < I32; q++) { s.nx[q] = -0.01f * (threadIdx.x + q); s.ny[q] = tile[(threadIdx.x * 7 + q) % (blocks * 3 * BLK)]; s.nz[q] = tile[(threadIdx.x * 13 + q * 5) % (blocks * 3 * BLK)]; s.ax[q] = s.ay[q] = s.az[q] = pk(0.f, 0.f); }
        for (int r = 0; r < reps32; r++)
            for (int b = 0; b < blocks; b++) {
                const float4* sx = reinterpret_cast<const float4*>(tile + b * 3 * BLK);
#pragma unroll 2
                for (int g = 0; g < BLK / 4; g++) interact4<I32>(s, sx[g], sx[g + BLK / 4], sx[g + 2 * (BLK / 4)]);
            }
#pragma unroll
        for (int q = 0; q < I32; q++) { float lo, hi; upk(s.ax[q], lo, hi); res += lo + hi; upk(s.ay[q], lo, hi); res += lo + hi; upk(s.az[q], lo, hi); res += lo + hi; }
    } else {
        double xi[I64], yi[I64], zi[I64], ax[I64], ay[I64], az[I64];
#pragma unroll
        for (int q = 0; q < I64; q++) { xi[q] = 0.01 * (threadIdx.x + q); yi[q] = dtile[(threadIdx.x * 7 + q) % (blocks * 3 * BLK)]; zi[q] = dtile[(threadIdx.x * 13 + q * 5) % (blocks * 3 * BLK)]; ax[q] = ay[q] = az[q] = 0.0; }
        for (int r = 0; r < reps64; r++)
            for (int b = 0; b < blocks; b++) {
                const double2* sx = reinterpret_cast<const double2*>(dtile + b * 3 * BLK);
#pragma unroll 2
                for (int g = 0; g < BLK / 2; g++) interact2_f64<I64>(xi, yi, zi, ax, ay, az, sx[g], sx[g + BLK / 2], sx[g + 2 * (BLK / 2)]);
            }
#pragma unroll
        for (int q = 0; q < I64; q++) res += (float)(ax[q] + ay[q] + az[q]);
    }
    const long long t1 = clock64();
    out[blockIdx.x * THREADS + threadIdx.x] = res;
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)(clk + (warp < w32 ? 0 : 1)), (unsigned long long)(t1 - t0));
}

static int g_sms; static float* g_out; static long long* g_clk;

static void pure(int w32, int w64, int it32, int it64) {
    const int threads = 32 * (w32 + w64);
    long long h[2] = {0, 0};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int t = 0; t < 3; t++) {
        cudaMemset(g_clk, 0, 16);
        cudaEventRecord(e0); k_pure<<<g_sms, threads>>>(g_out, g_clk, w32, it32, it64, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (t && ms < best) best = ms;
    }
    cudaMemcpy(h, g_clk, 16, cudaMemcpyDeviceToHost);
    const double f32_ops = (double)w32 * 32 * it32 * 256 * 2, f64_ops = (double)w64 * 32 * it64 * 256;    // lane-FMAs per SM
    printf("{\"test\": \"pure\", \"w32\": %d, \"w64\": %d, \"it32\": %d, \"it64\": %d, \"ms\": %.4f, \"clk32\": %lld, \"clk64\": %lld, "
           "\"f32_lane_fma_per_clk_sm\": %.2f, \"f64_lane_fma_per_clk_sm\": %.2f}\n",
           w32, w64, it32, it64, best, h[0], h[1], h[0] ? f32_ops / h[0] : 0.0, h[1] ? f64_ops / h[1] : 0.0);
}

template <int I32, int I64, int THREADS>
static void loop2(int w32, int reps32, int reps64) {
    const int blocks = 4;
    const size_t sm = (size_t)blocks * 3 * BLK * 12;
    cudaFuncSetAttribute(k_loop2<I32, I64, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_loop2<I32, I64, THREADS>);
    long long h[2] = {0, 0};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int t = 0; t < 3; t++) {
        cudaMemset(g_clk, 0, 16);
        cudaEventRecord(e0); k_loop2<I32, I64, THREADS><<<g_sms, THREADS, sm>>>(g_out, g_clk, w32, reps32, reps64, blocks); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (t && ms < best) best = ms;
    }
    cudaMemcpy(h, g_clk, 16, cudaMemcpyDeviceToHost);
    const int w64 = THREADS / 32 - w32;
    const double i32 = (double)w32 * 32 * I32 * reps32 * blocks * BLK, i64 = (double)w64 * 32 * I64 * reps64 * blocks * BLK;   // interactions per SM
    const double tot = (i32 + i64) * g_sms / (best * 1e-3);
    printf("{\"test\": \"loop2\", \"I32\": %d, \"I64\": %d, \"threads\": %d, \"regs\": %d, \"w32\": %d, \"w64\": %d, \"reps32\": %d, \"reps64\": %d, \"ms\": %.4f, "
           "\"clk32\": %lld, \"clk64\": %lld, \"cyc_per_inter_f32_warps\": %.3f, \"cyc_per_inter_f64_warps\": %.3f, \"G_inter_s_total\": %.1f, \"err\": \"%s\"}\n",
           I32, I64, THREADS, fa.numRegs, w32, w64, reps32, reps64, best, h[0], h[1],
           i32 ? h[0] * 128.0 / i32 : 0.0, i64 ? h[1] * 128.0 / i64 : 0.0, tot / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); g_sms = p.multiProcessorCount;
    cudaMalloc(&g_out, sizeof(float) * g_sms * 1024); cudaMalloc(&g_clk, 16);
    // pure streams: alone, then together (per SMSP: 2 FFMA2 warps, 1 or 2 DFMA warps)
    pure(8, 0, 2000, 0); pure(0, 4, 0, 1000); pure(0, 8, 0, 1000);
    pure(8, 4, 2000, 1000); pure(8, 8, 2000, 500); pure(8, 4, 2000, 2000); pure(4, 4, 2000, 1000);
    // loops: 8 FP32 warps alone, 4 FP64 warps alone, then both; reps chosen so both sides run about equally long
    loop2<4, 2, 384>(8, 400, 0);   loop2<4, 2, 384>(8, 0, 200);   loop2<4, 2, 384>(8, 400, 200);  loop2<4, 2, 384>(8, 400, 400);
    loop2<4, 4, 384>(8, 400, 0);   loop2<4, 4, 384>(8, 0, 100);   loop2<4, 4, 384>(8, 400, 100);  loop2<4, 4, 384>(8, 400, 200);
    loop2<8, 2, 384>(8, 200, 0);   loop2<8, 2, 384>(8, 0, 200);   loop2<8, 2, 384>(8, 200, 200);  loop2<8, 2, 384>(8, 200, 400);
    loop2<4, 2, 256>(4, 400, 200); loop2<4, 2, 512>(8, 400, 200);
    printf("{\"done\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
