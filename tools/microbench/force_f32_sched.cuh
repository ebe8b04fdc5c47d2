// Hand-scheduled inner loop of K1: the same per-pair arithmetic as force_f32_inner.cuh
// (dxy.vhd:94-122, dzsoft.vhd:177-202, dxyz_soft.vhd:149-150, fxyz.vhd:101-127, cube.vhd:66-70),
// software-pipelined by hand over the stream of (j-pair, i) interactions so that the instruction
// ORDER, not just the instruction mix, fits the sm_100a register file:
//
//   measured (tools/microbench/bank.cu, profiles/r01_microbench_bank.jsonl): a packed FP32 op takes
//   max(2, #distinct register pairs it reads) cycles (FFMA2 with three fresh pairs: 3.05), and a MUFU
//   costs ~0.7-1.0 extra cycle when it is issued behind an op that reads two pairs but ~0.2 behind an
//   op that reads one pair (its operand read then uses the bank slot the FP2 op leaves free).
//
// Per pair of j and per i-body the ops are
//   a1 a2 a3  FADD2  d = r_j - r_i                (pair + broadcast scalar)
//   f1        FFMA2  q = dx*dx + eps              (ONE pair)      <- host of MUFU #2
//   f2 f3     FFMA2  q += dy*dy ; q += dz*dz      (two pairs)
//   m1 m2     MUFU.RSQ on q.lo, q.hi
//   g1        FMUL2  r2 = r*r                     (ONE pair)      <- host of MUFU #1
//   g2        FMUL2  r3 = r2*r                    (two pairs)
//   h1 h2 h3  FFMA2  a += d*r3                    (three pairs, then r3 reused: 3+2+2 cycles)
// = 23 FP32-pipe cycles per 2 interactions if both MUFUs ride behind f1/g1.  A 4-slot ring keeps
// interactions k (being accumulated), k+1, k+2 (waiting for their rsqrt) and k+3 (distances being
// formed) in flight; every op is an `asm volatile` so ptxas keeps the order written here.
#pragma once
#include "../../mini-nbody_b200/csrc/force_f32_inner.cuh"

namespace nb {

__device__ __forceinline__ f2 vfma2(f2 a, f2 b, f2 c) { f2 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 vadd2(f2 a, f2 b) { f2 r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 vmul2(f2 a, f2 b) { f2 r; asm volatile("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float vrsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float4 vlds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

struct PSlot { f2 dx, dy, dz, q; };      // one in-flight pair-interaction: d and q = d2 -> r

struct JGroup { float4 X, Y, Z; };       // four consecutive j-bodies (two pairs)
__device__ __forceinline__ f2 jpair(const JGroup& g, int pair, int c) {
    const float4& v = c == 0 ? g.X : (c == 1 ? g.Y : g.Z);
    return pair == 0 ? pk(v.x, v.y) : pk(v.z, v.w);
}

// distances + dist^2 of interaction kk of group g into slot p (used by the prologue)
template <int I>
__device__ __forceinline__ void sched_s1(const IState<I>& s, PSlot& p, const JGroup& g, int kk) {
    const int pair = kk / I, i = kk % I;
    const f2 eps2 = pk(EPS_F32, EPS_F32);
    p.dx = vadd2(jpair(g, pair, 0), pk(s.nx[i], s.nx[i]));
    p.dy = vadd2(jpair(g, pair, 1), pk(s.ny[i], s.ny[i]));
    p.dz = vadd2(jpair(g, pair, 2), pk(s.nz[i], s.nz[i]));
    p.q = vfma2(p.dx, p.dx, eps2);
    p.q = vfma2(p.dy, p.dy, p.q);
    p.q = vfma2(p.dz, p.dz, p.q);
}
__device__ __forceinline__ void sched_rsq(PSlot& p) {
    float lo, hi; upk(p.q, lo, hi);
    lo = vrsq(lo); hi = vrsq(hi);
    p.q = pk(lo, hi);
}

// One group (4 j x I i = 2I pair-interactions) of the steady-state pipeline.
//   cur : the group whose interactions are being accumulated
//   nxt : the following group (its first three interactions enter the ring in the last three steps);
//         ignored when LAST (the ring then drains)
//   lds_addr != 0: issue the three LDS.128 of the group AFTER nxt into `pre` (next-next group's data)
template <int I, bool LAST>
__device__ __forceinline__ void sched_group(IState<I>& s, PSlot (&sl)[4], const JGroup& cur, const JGroup& nxt,
                                            JGroup& pre, uint32_t lds_addr) {
    constexpr int NK = 2 * I;
    const f2 eps2 = pk(EPS_F32, EPS_F32);
#pragma unroll
    for (int K = 0; K < NK; K++) {
        const int kin = K + 3;
        const bool from_next = kin >= NK;
        const bool have_in = !(LAST && from_next);
        const int kk = from_next ? kin - NK : kin;
        const int pair = kk / I, ii = kk % I;
        const JGroup& J = from_next ? nxt : cur;
        PSlot& in = sl[kin & 3];
        PSlot& out = sl[K & 3];
        PSlot& mid = sl[(K + 2) & 3];
        const bool have_mid = !(LAST && K + 2 >= NK);
        const int io = K % I;

        float qlo = 0.f, qhi = 0.f;
        if (have_mid) upk(mid.q, qlo, qhi);
        if (have_in) in.dx = vadd2(jpair(J, pair, 0), pk(s.nx[ii], s.nx[ii]));           // a1
        const f2 r2 = vmul2(out.q, out.q);                                                // g1 (one pair)
        if (have_mid) qlo = vrsq(qlo);                                                    // m1 behind g1
        if (have_in) in.dy = vadd2(jpair(J, pair, 1), pk(s.ny[ii], s.ny[ii]));           // a2
        if (have_in) in.q = vfma2(in.dx, in.dx, eps2);                                    // f1 (one pair)
        if (have_mid) { qhi = vrsq(qhi); mid.q = pk(qlo, qhi); }                          // m2 behind f1
        if (have_in) in.dz = vadd2(jpair(J, pair, 2), pk(s.nz[ii], s.nz[ii]));           // a3
        const f2 r3 = vmul2(r2, out.q);                                                   // g2
        if (have_in) in.q = vfma2(in.dy, in.dy, in.q);                                    // f2
        s.ax[io] = vfma2(out.dx, r3, s.ax[io]);                                           // h1
        s.ay[io] = vfma2(out.dy, r3, s.ay[io]);                                           // h2 (r3 reused)
        s.az[io] = vfma2(out.dz, r3, s.az[io]);                                           // h3 (r3 reused)
        if (have_in) in.q = vfma2(in.dz, in.dz, in.q);                                    // f3
        if (lds_addr != 0) {                                                              // prefetch, one LDS.128 per step
            if (K == 1) pre.X = vlds128(lds_addr);
            if (K == 3) pre.Y = vlds128(lds_addr + BLK * 4);
            if (K == 5) pre.Z = vlds128(lds_addr + 2 * BLK * 4);
        }
    }
}

// Prologue: fill the ring with interactions 0,1,2 of the first group (rsqrt of 0 and 1 issued;
// interaction 2's rsqrt is issued by step 0 of the first sched_group call).
template <int I>
__device__ __forceinline__ void sched_prologue(const IState<I>& s, PSlot (&sl)[4], const JGroup& g) {
    sched_s1<I>(s, sl[0], g, 0);
    sched_s1<I>(s, sl[1], g, 1);
    sched_rsq(sl[0]);
    sched_s1<I>(s, sl[2], g, 2);
    sched_rsq(sl[1]);
}

// All groups of `nblk` resident layout blocks starting at shared address `base` (bytes).
// Groups are consumed two per loop iteration with the roles of the three JGroup register sets
// rotating statically (A,B,C -> C,A,B ...) so no register moves are needed.
template <int I>
__device__ __forceinline__ void sched_tile(IState<I>& s, uint32_t base, int nblk) {
    constexpr int GPB = BLK / 4;                   // groups per block
    const int ngroups = nblk * GPB;                // multiple of 32
    auto gaddr = [&](int g) -> uint32_t {          // shared address of group g's x-row quad
        return base + (uint32_t)(g / GPB) * (3 * BLK * 4) + (uint32_t)(g % GPB) * 16;
    };
    PSlot sl[4];
    JGroup A, B, C;
    A.X = vlds128(gaddr(0)); A.Y = vlds128(gaddr(0) + BLK * 4); A.Z = vlds128(gaddr(0) + 2 * BLK * 4);
    B.X = vlds128(gaddr(1)); B.Y = vlds128(gaddr(1) + BLK * 4); B.Z = vlds128(gaddr(1) + 2 * BLK * 4);
    sched_prologue<I>(s, sl, A);
    // steady state: three groups per iteration (A cur, B nxt, C pre) -> (B, C, A) -> (C, A, B)
    int g = 0;
    for (; g + 3 < ngroups; g += 3) {
        sched_group<I, false>(s, sl, A, B, C, gaddr(g + 2));
        sched_group<I, false>(s, sl, B, C, A, gaddr(g + 3));
        sched_group<I, false>(s, sl, C, A, B, g + 4 < ngroups ? gaddr(g + 4) : 0);
    }
    // tail: ngroups % 3 is 2 for any multiple of 32 that is not a multiple of 3 ... handle 1, 2 or 3 left
    const int left = ngroups - g;                  // 1..3 groups left, A = group g, B = group g+1 (if any)
    if (left == 1) {
        sched_group<I, true>(s, sl, A, A, C, 0);
    } else if (left == 2) {
        sched_group<I, false>(s, sl, A, B, C, 0);
        sched_group<I, true>(s, sl, B, B, C, 0);
    } else {
        sched_group<I, false>(s, sl, A, B, C, gaddr(g + 2));
        sched_group<I, false>(s, sl, B, C, A, 0);
        sched_group<I, true>(s, sl, C, C, A, 0);
    }
}

}  // namespace nb
