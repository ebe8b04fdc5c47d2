// Inner loop of K1 (force_f32.cu): the per-pair arithmetic of the reference pipeline
// (dxy.vhd:94-122, dzsoft.vhd:177-202, dxyz_soft.vhd:149-150, fxyz.vhd:101-127, cube.vhd:66-70) on
// four consecutive j-bodies (two f32x2 pairs) against I register-blocked i-bodies.
// Shared with tools/microbench/loop.cu so the microbenchmark times exactly this code.
#pragma once
#include "nbody_internal.cuh"

namespace nb {

template <int I>
struct IState {
    float nx[I], ny[I], nz[I];     // negated i-positions (scalar-broadcast operands of FADD2)
    f2 ax[I], ay[I], az[I];        // accumulators: .lo = even j, .hi = odd j
};

template <int I>
__device__ __forceinline__ void interact4(IState<I>& s, const float4 X, const float4 Y, const float4 Z, const float eps = EPS_F32) {
    const f2 eps2 = pk(eps, eps);
    const f2 xa = pk(X.x, X.y), xb = pk(X.z, X.w);
    const f2 ya = pk(Y.x, Y.y), yb = pk(Y.z, Y.w);
    const f2 za = pk(Z.x, Z.y), zb = pk(Z.z, Z.w);
#pragma unroll
    for (int i = 0; i < I; i++) {
        const f2 nx2 = pk(s.nx[i], s.nx[i]), ny2 = pk(s.ny[i], s.ny[i]), nz2 = pk(s.nz[i], s.nz[i]);
        {
            const f2 dx = add2(xa, nx2), dy = add2(ya, ny2), dz = add2(za, nz2);
            f2 d2 = fma2(dx, dx, eps2); d2 = fma2(dy, dy, d2); d2 = fma2(dz, dz, d2);
            float d2lo, d2hi; upk(d2, d2lo, d2hi);
            const f2 r = pk(rsqrt_approx(d2lo), rsqrt_approx(d2hi));
            const f2 r3 = mul2(mul2(r, r), r);
            s.ax[i] = fma2(dx, r3, s.ax[i]); s.ay[i] = fma2(dy, r3, s.ay[i]); s.az[i] = fma2(dz, r3, s.az[i]);
        }
        {
            const f2 dx = add2(xb, nx2), dy = add2(yb, ny2), dz = add2(zb, nz2);
            f2 d2 = fma2(dx, dx, eps2); d2 = fma2(dy, dy, d2); d2 = fma2(dz, dz, d2);
            float d2lo, d2hi; upk(d2, d2lo, d2hi);
            const f2 r = pk(rsqrt_approx(d2lo), rsqrt_approx(d2hi));
            const f2 r3 = mul2(mul2(r, r), r);
            s.ax[i] = fma2(dx, r3, s.ax[i]); s.ay[i] = fma2(dy, r3, s.ay[i]); s.az[i] = fma2(dz, r3, s.az[i]);
        }
    }
}

// scalar variant of the same loop (one FFMA per lane-op): kept as the measured baseline the
// packed loop is compared against.
template <int I>
__device__ __forceinline__ void interact4_scalar(IState<I>& s, const float4 X, const float4 Y, const float4 Z, const float eps = EPS_F32) {
    const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
    for (int i = 0; i < I; i++) {
        float alo, ahi, blo, bhi, clo, chi;
        upk(s.ax[i], alo, ahi); upk(s.ay[i], blo, bhi); upk(s.az[i], clo, chi);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float dx = xs[q] + s.nx[i], dy = ys[q] + s.ny[i], dz = zs[q] + s.nz[i];
            float d2 = fmaf(dx, dx, eps); d2 = fmaf(dy, dy, d2); d2 = fmaf(dz, dz, d2);
            const float r = rsqrt_approx(d2);
            const float r3 = (r * r) * r;
            if (q & 1) { ahi = fmaf(dx, r3, ahi); bhi = fmaf(dy, r3, bhi); chi = fmaf(dz, r3, chi); }
            else       { alo = fmaf(dx, r3, alo); blo = fmaf(dy, r3, blo); clo = fmaf(dz, r3, clo); }
        }
        s.ax[i] = pk(alo, ahi); s.ay[i] = pk(blo, bhi); s.az[i] = pk(clo, chi);
    }
}

}  // namespace nb
