// Second microbenchmark: how FFMA/FFMA2 issue interacts with MUFU.RSQ on sm_100a.
// For each ratio NF:1 (NF fp32-pipe instructions per MUFU) reports SMSP cycles per group, so the
// cost of one MUFU in fp32-pipe cycles can be read off directly.  Evidence only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr int NCH = 12;

// PACKED=1: NF FFMA2 per MUFU ; PACKED=0: NF FFMA per MUFU.  NM MUFUs per group (NF*NM fp instrs).
template <int NF, int PACKED>
__global__ void __launch_bounds__(256) k_ratio(float* out, int iters, float b, float c) {
    uint64_t a2[NCH]; float a1[NCH]; float m[4];
    uint64_t b2 = pk(b, b), c2 = pk(c, c);
#pragma unroll
    for (int i = 0; i < NCH; i++) { a2[i] = pk(threadIdx.x * 1e-3f + i, 1.f + i); a1[i] = threadIdx.x * 1e-3f + i; }
#pragma unroll
    for (int i = 0; i < 4; i++) m[i] = 1.f + i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int g = 0; g < 12; g++) {           // 12 groups per iteration
            if (NF >= 0) m[g & 3] = rsq(m[g & 3]);
#pragma unroll
            for (int f = 0; f < (NF < 0 ? -NF : NF); f++) {
                const int ch = (g * (NF < 0 ? -NF : NF) + f) % NCH;
                if (PACKED) a2[ch] = fma2(a2[ch], b2, c2); else a1[ch] = ffma(a1[ch], b, c);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < NCH; i++) { float x, y; upk(a2[i], x, y); s += x + y + a1[i]; }
    for (int i = 0; i < 4; i++) s += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// single-op-type throughput for the packed forms the force loop uses
template <int OP>
__global__ void __launch_bounds__(256) k_op(float* out, int iters, float b, float c) {
    uint64_t a[NCH]; float sc[NCH];
    uint64_t b2 = pk(b, c);
#pragma unroll
    for (int i = 0; i < NCH; i++) { a[i] = pk(threadIdx.x * 1e-3f + i, 1.f + i); sc[i] = c + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 32; u++)
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (OP == 0) a[i] = add2(a[i], b2);                       // FADD2 reg,reg
                if (OP == 1) a[i] = add2(a[i], pk(sc[i], sc[i]));         // FADD2 reg, scalar-broadcast
                if (OP == 2) a[i] = mul2(a[i], b2);                       // FMUL2
                if (OP == 3) a[i] = fma2(a[i], a[i], pk(1e-9f, 1e-9f));   // FFMA2 a,a,imm
                if (OP == 4) a[i] = fma2(a[i], b2, a[(i + 1) % NCH]);     // FFMA2 3 distinct regs
                if (OP == 5) a[i] = mul2(a[i], a[i]);                     // FMUL2 a,a
            }
    }
    float s = 0;
    for (int i = 0; i < NCH; i++) { float x, y; upk(a[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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

static int g_sms; static float* g_out; static const double CLK = 1.9577e9;  // measured by pipes.cu k_clock

template <int NF, int PACKED>
static void run_ratio(int ctas_per_sm) {
    const int iters = 4000, grid = g_sms * ctas_per_sm;
    float ms = time_ms([&] { k_ratio<NF, PACKED><<<grid, 256>>>(g_out, iters, 1.0001f, 0.5f); }, 3);
    double warps_per_smsp = ctas_per_sm * 8 / 4.0;
    double groups = (double)iters * 12 * warps_per_smsp;
    printf("{\"test\": \"ratio\", \"packed\": %d, \"nf\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"smsp_cycles_per_group\": %.3f}\n",
           PACKED, NF, ctas_per_sm, ms, ms * 1e-3 * CLK / groups);
}
template <int OP>
static void run_op(const char* name) {
    const int iters = 2000, ctas_per_sm = 2, grid = g_sms * ctas_per_sm;
    float ms = time_ms([&] { k_op<OP><<<grid, 256>>>(g_out, iters, 1.0001f, 0.5f); }, 3);
    double instr = (double)iters * 32 * NCH * (ctas_per_sm * 8 / 4.0);
    printf("{\"test\": \"op\", \"name\": \"%s\", \"ms\": %.4f, \"smsp_cycles_per_instr\": %.3f}\n", name, ms, ms * 1e-3 * CLK / instr);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); g_sms = p.multiProcessorCount;
    cudaMalloc(&g_out, sizeof(float) * g_sms * 8 * 256);
    run_op<0>("FADD2 r,r"); run_op<1>("FADD2 r,bcast"); run_op<2>("FMUL2 r,r"); run_op<3>("FFMA2 a,a,imm");
    run_op<4>("FFMA2 r,r,r"); run_op<5>("FMUL2 a,a");
    for (int c = 1; c <= 4; c *= 2) {
        run_ratio<-11, 1>(c);   // no MUFU, 11 FFMA2 per group (baseline)
        run_ratio<2, 1>(c); run_ratio<4, 1>(c); run_ratio<6, 1>(c); run_ratio<8, 1>(c); run_ratio<11, 1>(c); run_ratio<16, 1>(c);
        run_ratio<-11, 0>(c);
        run_ratio<4, 0>(c); run_ratio<8, 0>(c); run_ratio<11, 0>(c); run_ratio<16, 0>(c); run_ratio<22, 0>(c);
    }
    cudaDeviceSynchronize();
    printf("{\"done\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
