// Pipe-rate microbenchmarks for B200 (sm_100a): fixes the roofline denominators used in
// DESIGN.md / bench.py.  Measures warp-instruction issue rates of FFMA, FFMA2 (fma.rn.f32x2),
// MUFU.RSQ, the 11:1 force-loop mix (scalar and packed), DFMA and MUFU.RSQ64H.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Not part of the product path; evidence only (output committed under profiles/).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ double dfma(double a, double b, double c) { double r; asm volatile("fma.rn.f64 %0,%1,%2,%3;" : "=d"(r) : "d"(a), "d"(b), "d"(c)); return r; }
__device__ __forceinline__ double rsq64(double x) { double r; asm volatile("rsqrt.approx.ftz.f64 %0,%1;" : "=d"(r) : "d"(x)); return r; }

constexpr int NCH = 8;     // independent dependency chains per thread
constexpr int INNER = 64;  // unrolled ops per chain per outer iteration

// mode 0: scalar FFMA   (ops counted: NCH*INNER warp-instr / iter, 1 FMA lane-op each)
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float b, float c) {
    float a[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < INNER; u++)
#pragma unroll
            for (int i = 0; i < NCH; i++) a[i] = ffma(a[i], b, c);
    }
    float s = 0; for (int i = 0; i < NCH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mode 1: packed FFMA2
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float b, float c) {
    uint64_t a[NCH]; uint64_t b2 = pk(b, b), c2 = pk(c, c);
#pragma unroll
    for (int i = 0; i < NCH; i++) a[i] = pk(threadIdx.x * 1e-3f + i, 1.f + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < INNER; u++)
#pragma unroll
            for (int i = 0; i < NCH; i++) a[i] = fma2(a[i], b2, c2);
    }
    float s = 0; for (int i = 0; i < NCH; i++) { float x, y; upk(a[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mode 2: MUFU.RSQ only
__global__ void __launch_bounds__(256) k_rsq(float* out, int iters) {
    float a[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) a[i] = 1.0f + threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < INNER; u++)
#pragma unroll
            for (int i = 0; i < NCH; i++) a[i] = rsq(a[i]);
    }
    float s = 0; for (int i = 0; i < NCH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mode 3: scalar force-loop mix: 3 FADD + 3 FFMA + RSQ + 2 FMUL + 3 FFMA per "interaction"
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float xj, float yj, float zj) {
    float xi[NCH], ax[NCH], ay[NCH], az[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { xi[i] = threadIdx.x * 1e-3f + i; ax[i] = ay[i] = az[i] = 0.f; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                float dx = fadd(xj, -xi[i]), dy = fadd(yj, -xi[i]), dz = fadd(zj, -xi[i]);
                float s = ffma(dx, dx, 1e-9f); s = ffma(dy, dy, s); s = ffma(dz, dz, s);
                float r = rsq(s);
                float r3 = fmul(fmul(r, r), r);
                ax[i] = ffma(dx, r3, ax[i]); ay[i] = ffma(dy, r3, ay[i]); az[i] = ffma(dz, r3, az[i]);
            }
            xj += 0.25f; yj += 0.5f; zj += 0.125f;   // uniform-datapath noise, 3 extra FADD per 8 interactions
        }
    }
    float s = 0; for (int i = 0; i < NCH; i++) s += ax[i] + ay[i] + az[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mode 4: packed force-loop mix: per 2 interactions 3 FADD2 + 3 FFMA2 + 2 RSQ + 2 FMUL2 + 3 FFMA2
__global__ void __launch_bounds__(256) k_mix2(float* out, int iters, float xj, float yj, float zj) {
    float xi[NCH]; uint64_t ax[NCH], ay[NCH], az[NCH];
    const uint64_t eps = pk(1e-9f, 1e-9f);
#pragma unroll
    for (int i = 0; i < NCH; i++) { xi[i] = threadIdx.x * 1e-3f + i; ax[i] = ay[i] = az[i] = pk(0.f, 0.f); }
    uint64_t xj2 = pk(xj, xj + 1.f), yj2 = pk(yj, yj + 1.f), zj2 = pk(zj, zj + 1.f);
    const uint64_t inc = pk(0.25f, 0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                uint64_t nx = pk(-xi[i], -xi[i]);
                uint64_t dx = add2(xj2, nx), dy = add2(yj2, nx), dz = add2(zj2, nx);
                uint64_t s = fma2(dx, dx, eps); s = fma2(dy, dy, s); s = fma2(dz, dz, s);
                float s0, s1; upk(s, s0, s1);
                uint64_t r = pk(rsq(s0), rsq(s1));
                uint64_t r3 = mul2(mul2(r, r), r);
                ax[i] = fma2(dx, r3, ax[i]); ay[i] = fma2(dy, r3, ay[i]); az[i] = fma2(dz, r3, az[i]);
            }
            xj2 = add2(xj2, inc); yj2 = add2(yj2, inc); zj2 = add2(zj2, inc);
        }
    }
    float s = 0; for (int i = 0; i < NCH; i++) { float a, b; upk(ax[i], a, b); s += a + b; upk(ay[i], a, b); s += a + b; upk(az[i], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mode 5: DFMA
__global__ void __launch_bounds__(256) k_dfma(float* out, int iters, double b, double c) {
    double a[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < INNER; u++)
#pragma unroll
            for (int i = 0; i < NCH; i++) a[i] = dfma(a[i], b, c);
    }
    double s = 0; for (int i = 0; i < NCH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
// mode 6: MUFU.RSQ64H
__global__ void __launch_bounds__(256) k_rsq64(float* out, int iters) {
    double a[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < INNER; u++)
#pragma unroll
            for (int i = 0; i < NCH; i++) a[i] = rsq64(a[i]);
    }
    double s = 0; for (int i = 0; i < NCH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
// SM clock estimate: cycles elapsed on one SM vs. wall time of the kernel
__global__ void k_clock(long long* out, int iters) {
    long long t0 = clock64();
    float a = threadIdx.x;
    for (int i = 0; i < iters; i++) a = ffma(a, 1.0001f, 0.5f);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)a; }
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"cc\": \"%d.%d\", \"sms\": %d, \"l2_bytes\": %d, \"smem_per_sm\": %zu, \"smem_optin\": %zu, \"regs_per_sm\": %d, \"clock_khz\": %d, \"global_mem\": %zu}\n",
           p.name, p.major, p.minor, p.multiProcessorCount, p.l2CacheSize, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.regsPerMultiprocessor, clk_khz, p.totalGlobalMem);
    const int sms = p.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256));
    long long* cout_; CK(cudaMalloc(&cout_, 16));
    // clock estimate under load
    {
        const int it = 4000000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_clock<<<sms, 128>>>(cout_, 1000); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_clock<<<sms, 128>>>(cout_, it); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[2]; cudaMemcpy(h, cout_, 16, cudaMemcpyDeviceToHost);
        printf("{\"test\": \"sm_clock\", \"cycles\": %lld, \"ms\": %.4f, \"mhz\": %.1f}\n", h[0], ms, h[0] / (ms * 1e3));
    }
    for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
        const int grid = sms * ctas_per_sm, thr = 256;
        const double warps = (double)grid * thr / 32;
        const int iters = 2000;
        struct R { const char* name; float ms; double winstr; double extra; };
        auto rep = [&](const char* name, float ms, double fp_winstr_per_warp, double xu_winstr_per_warp, double per_inter) {
            double t = ms * 1e-3;
            double fp = fp_winstr_per_warp * warps / t, xu = xu_winstr_per_warp * warps / t;
            printf("{\"test\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"fp_warp_instr_per_s\": %.4e, \"fp_lane_ops_per_clk_per_sm@1965\": %.2f, \"xu_lane_ops_per_clk_per_sm@1965\": %.2f",
                   name, ctas_per_sm, ms, fp, fp * 32 / sms / 1.965e9, xu * 32 / sms / 1.965e9);
            if (per_inter > 0) printf(", \"G_inter_per_s\": %.1f", per_inter * warps * 32 / t / 1e9);
            printf("}\n");
        };
        float ms;
        ms = time_ms([&] { k_ffma<<<grid, thr>>>(out, iters, 1.0001f, 0.5f); }, 5);
        rep("ffma", ms, (double)iters * INNER * NCH, 0, 0);
        ms = time_ms([&] { k_ffma2<<<grid, thr>>>(out, iters, 1.0001f, 0.5f); }, 5);
        rep("ffma2(lane-ops=2x instr)", ms, (double)iters * INNER * NCH * 2, 0, 0);
        ms = time_ms([&] { k_rsq<<<grid, thr>>>(out, iters / 4); }, 5);
        rep("mufu.rsq", ms, 0, (double)(iters / 4) * INNER * NCH, 0);
        ms = time_ms([&] { k_mix<<<grid, thr>>>(out, iters, 0.3f, 0.2f, 0.1f); }, 5);
        rep("mix_scalar(11fp+1xu)", ms, (double)iters * 8 * NCH * 11, (double)iters * 8 * NCH, (double)iters * 8 * NCH);
        ms = time_ms([&] { k_mix2<<<grid, thr>>>(out, iters, 0.3f, 0.2f, 0.1f); }, 5);
        rep("mix_packed(11fp2+2xu per 2)", ms, (double)iters * 4 * NCH * 22, (double)iters * 4 * NCH * 2, (double)iters * 4 * NCH * 2);
        ms = time_ms([&] { k_dfma<<<grid, thr>>>(out, iters / 2, 1.0001, 0.5); }, 5);
        rep("dfma", ms, (double)(iters / 2) * INNER * NCH, 0, 0);
        ms = time_ms([&] { k_rsq64<<<grid, thr>>>(out, iters / 8); }, 5);
        rep("mufu.rsq64h", ms, 0, (double)(iters / 8) * INNER * NCH, 0);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
