// Offline probe: compiles the product inner loop (force_f32_inner.cuh) in a minimal kernel so that the
// SASS of the hot loop can be scored with the register-bank model (tools/sass_bank_model.py) without a GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DPI=8 -DPUNROLL=2 -cubin -o probe.cubin model_probe.cu
#include "../../mini-nbody_b200/csrc/force_f32_inner.cuh"
using namespace nb;
#ifndef PI
#define PI 8
#endif
#ifndef PUNROLL
#define PUNROLL 2
#endif
#ifndef PMINB
#define PMINB 1
#endif
constexpr int kUnroll = PUNROLL;
extern "C" __global__ void __launch_bounds__(128, PMINB) probe(const float* __restrict__ pos, float* out, int blocks, int reps) {
    extern __shared__ __align__(128) float tile[];
    for (int t = threadIdx.x; t < blocks * 3 * BLK; t += 128) tile[t] = pos[t];
    __syncthreads();
    IState<PI> s;
#pragma unroll
    for (int q = 0; q < PI; q++) {
        const float* pi = pos + (blockIdx.x * PI + q) * 3 * BLK + threadIdx.x;
        s.nx[q] = -pi[0]; s.ny[q] = -pi[BLK]; s.nz[q] = -pi[2 * BLK];
        s.ax[q] = s.ay[q] = s.az[q] = pk(0.f, 0.f);
    }
    for (int r = 0; r < reps; r++)
        for (int b = 0; b < blocks; b++) {
            const float4* sx = reinterpret_cast<const float4*>(tile + b * 3 * BLK);
#pragma unroll kUnroll
            for (int g = 0; g < BLK / 4; g++) {
                const float4 X = sx[g], Y = sx[g + BLK / 4], Z = sx[g + 2 * (BLK / 4)];
                interact4<PI>(s, X, Y, Z);
            }
        }
#pragma unroll
    for (int q = 0; q < PI; q++) {
        float lo, hi; float* o = out + (blockIdx.x * PI + q) * 3 * BLK + threadIdx.x;
        upk(s.ax[q], lo, hi); o[0] = lo + hi; upk(s.ay[q], lo, hi); o[BLK] = lo + hi; upk(s.az[q], lo, hi); o[2 * BLK] = lo + hi;
    }
}
