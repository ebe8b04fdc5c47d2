// K3 integrate (+ reduction of the force kernel's partial sums), K4 layout conversion at the C-ABI
// boundary, K5 FP64 energy diagnostic, and the FFMA peak probe.  All HBM-bound streaming kernels:
// one thread per body, rows of 128 consecutive scalars => every access is a coalesced 512 B / 1 KiB
// row; grids are sized by body count (one CTA per layout block).
//
// Integrate semantics (BASELINE.json north_star; the host C code that would hold it is absent from
// the reference mount): bodyForce applies v += dt * F, integrate applies x += dt * v with the
// updated velocity.  The force result record of the reference is {Fx,Fy,Fz,0} per body
// (compute_store.vhd:203-242); here it stays in HBM as per-split partial sums that this kernel adds.
#include "nbody_internal.cuh"

namespace nb {

template <typename T>
__global__ void __launch_bounds__(BLK) integrate_kernel(const IntegrateArgs a) {
    const int ib = blockIdx.x;                 // local i-block
    const int lane = threadIdx.x;
    const size_t loc = (size_t)ib * 3 * BLK + lane;
    const size_t glb = (size_t)(a.i_blk0 + ib) * 3 * BLK + lane;
    const long long body = (long long)(a.i_blk0 + ib) * BLK + lane;

    // state loads first: they are in flight under the slot loop (at launch-bound sizes the kernel is a chain of
    // L2 round trips, not a stream)
    T* vel = a.vel ? static_cast<T*>(a.vel) + loc : nullptr;
    const T* __restrict__ pc = static_cast<const T*>(a.pos_cur) + glb;
    T vx = 0, vy = 0, vz = 0, x = 0, y = 0, z = 0;
    if (vel) { vx = vel[0]; vy = vel[BLK]; vz = vel[2 * BLK]; x = pc[0]; y = pc[BLK]; z = pc[2 * BLK]; }

    T ax = 0, ay = 0, az = 0;
    const T* __restrict__ part = static_cast<const T*>(a.part);
    const size_t slot_stride = (size_t)a.n_iblk * 3 * BLK;
#pragma unroll 8
    for (int s = 0; s < a.slots; s++) {        // fixed order => deterministic sum; unrolled so the loads overlap
        const T* p = part + (size_t)s * slot_stride + loc;
        ax += p[0]; ay += p[BLK]; az += p[2 * BLK];
    }
    if (a.acc_out) {
        T* o = static_cast<T*>(a.acc_out) + loc;
        o[0] = ax; o[BLK] = ay; o[2 * BLK] = az;
    }
    if (vel == nullptr) return;                // acceleration-only pass (nbody_accel)

    if (body < a.n) {                          // padding bodies never move
        const T dtv = (T)a.dt_v, dtx = (T)a.dt_x;
        vx = fma(dtv, ax, vx); vy = fma(dtv, ay, vy); vz = fma(dtv, az, vz);
        x = fma(vx, dtx, x); y = fma(vy, dtx, y); z = fma(vz, dtx, z);
    }
    vel[0] = vx; vel[BLK] = vy; vel[2 * BLK] = vz;
    if (a.pos_next) {
        T* pn = static_cast<T*>(a.pos_next) + glb;
        pn[0] = x; pn[BLK] = y; pn[2 * BLK] = z;
        // push exchange: the same slice goes straight into every peer's next-step buffer over NVLink
        for (int r = 0; r < a.n_peers; r++) {
            T* pp = static_cast<T*>(a.peer_pos_next[r]) + glb;
            pp[0] = x; pp[BLK] = y; pp[2 * BLK] = z;
        }
    }
    if (a.n_peers > 0 && a.peer_flags != nullptr) {
        // every CTA: make its peer stores visible system-wide, then count itself; the last one signals
        __threadfence_system();
        __syncthreads();
        if (lane == 0) {
            const unsigned int prev = atomicAdd(a.done_counter, 1u);
            if (prev == gridDim.x - 1) {
                *a.done_counter = 0u;
                __threadfence_system();
                for (int r = 0; r < a.n_peers; r++)
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.peer_flags[r] + a.flag_index), "l"(a.flag_value) : "memory");
            }
        }
    }
}

// ---- flag handshake of the push exchange ------------------------------------------------------------
__global__ void flag_signal_kernel(unsigned long long* const* peer_flags, int n_peers, int index, unsigned long long value) {
    const int p = threadIdx.x;
    if (p < n_peers) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flags[p] + index), "l"(value) : "memory");
    }
}
// waits until flags[i] >= value for every i != skip; gives up after ~20 s and raises *err instead of
// hanging the device (the waited-for writers run on OTHER GPUs, never on this one)
__global__ void flag_wait_kernel(const unsigned long long* flags, int count, int skip, unsigned long long value, int* err) {
    const int p = threadIdx.x;
    if (p >= count || p == skip) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + p) : "memory");
        if (v >= value) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) { atomicExch(err, 1); break; }
        __nanosleep(200);
    }
}
cudaError_t flag_signal_launch(unsigned long long* const* peer_flags, int n_peers, int index, unsigned long long value, cudaStream_t st) {
    if (n_peers <= 0) return cudaSuccess;
    flag_signal_kernel<<<1, 64, 0, st>>>(peer_flags, n_peers, index, value);      // one thread per peer, up to MAX_WORLD = 64 ranks
    return cudaGetLastError();
}
cudaError_t flag_wait_launch(const unsigned long long* flags, int count, int skip, unsigned long long value, int* err, cudaStream_t st) {
    flag_wait_kernel<<<1, 64, 0, st>>>(flags, count, skip, value, err);
    return cudaGetLastError();
}

// Drift only (integrate(): x += dt * v, no partial sums, no peers): a pure element-wise update of the rank's slice of the
// blocked arrays -- positions and velocities have the same [block][3][128] layout, padding bodies carry v = 0 and stay
// put -- so it runs as 16-byte vector loads/stores, one vector per thread: 36 B/body algorithmic (read pos + vel, write
// pos; the velocities are not rewritten), HBM-bound.  Same fma per scalar as integrate_kernel => bit-identical.
template <typename T, typename V>
__global__ void __launch_bounds__(256) drift_kernel(const V* __restrict__ pos_cur, const V* __restrict__ vel, V* __restrict__ pos_next,
                                                   size_t nvec, T dt) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= nvec) return;
    const V x = pos_cur[i], v = vel[i];
    V o;
    if constexpr (sizeof(V) == 16 && sizeof(T) == 4) { o.x = fma(v.x, dt, x.x); o.y = fma(v.y, dt, x.y); o.z = fma(v.z, dt, x.z); o.w = fma(v.w, dt, x.w); }
    else { o.x = fma(v.x, dt, x.x); o.y = fma(v.y, dt, x.y); }
    pos_next[i] = o;
}

cudaError_t integrate_launch(int precision, const IntegrateArgs& a, cudaStream_t st) {
    if (a.n_iblk <= 0) return cudaSuccess;
    if (a.slots == 0 && a.n_peers == 0 && a.acc_out == nullptr && a.vel != nullptr && a.pos_next != nullptr && a.dt_v == 0.0) {
        const size_t bytes = (size_t)a.n_iblk * 3 * BLK * (precision == 0 ? 4 : 8), nvec = bytes / 16;
        const size_t off = (size_t)a.i_blk0 * 3 * BLK * (precision == 0 ? 4 : 8);
        const unsigned grid = (unsigned)((nvec + 255) / 256);
        if (precision == 0)
            drift_kernel<float, float4><<<grid, 256, 0, st>>>((const float4*)((const char*)a.pos_cur + off), (const float4*)a.vel,
                                                             (float4*)((char*)a.pos_next + off), nvec, (float)a.dt_x);
        else
            drift_kernel<double, double2><<<grid, 256, 0, st>>>((const double2*)((const char*)a.pos_cur + off), (const double2*)a.vel,
                                                               (double2*)((char*)a.pos_next + off), nvec, (double)a.dt_x);
        return cudaGetLastError();
    }
    if (precision == 0) integrate_kernel<float><<<a.n_iblk, BLK, 0, st>>>(a);
    else integrate_kernel<double><<<a.n_iblk, BLK, 0, st>>>(a);
    return cudaGetLastError();
}

// ---- K4: Body{x,y,z,vx,vy,vz} AoS <-> tile-blocked SoA -------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BLK) aos_to_blocked_kernel(const T* __restrict__ aos, int n, int i_blk0, int n_iblk,
                                                            T* __restrict__ pos, T* __restrict__ vel, T pad,
                                                            int blk_first, long long aos_body0) {
    // converts layout blocks [blk_first, blk_first + gridDim.x); aos[0] is body aos_body0 (a rank's slice of the
    // caller's array when the bodies are sharded, the whole array otherwise)
    const int b = blk_first + blockIdx.x, lane = threadIdx.x;
    const long long body = (long long)b * BLK + lane;
    T x = pad, y = pad, z = pad, vx = 0, vy = 0, vz = 0;
    if (body < n) {
        const T* p = aos + (body - aos_body0) * 6;
        x = p[0]; y = p[1]; z = p[2]; vx = p[3]; vy = p[4]; vz = p[5];
    }
    T* o = pos + (size_t)b * 3 * BLK + lane;
    o[0] = x; o[BLK] = y; o[2 * BLK] = z;
    if (vel && b >= i_blk0 && b < i_blk0 + n_iblk) {
        T* v = vel + (size_t)(b - i_blk0) * 3 * BLK + lane;
        v[0] = vx; v[BLK] = vy; v[2 * BLK] = vz;
    }
}

cudaError_t aos_to_blocked_launch(int precision, const void* aos, int n, int i_blk0, int n_iblk, int total_blocks,
                                  void* pos, void* vel, cudaStream_t st, int blk_first, int n_blk, long long aos_body0) {
    if (n_blk < 0) { blk_first = 0; n_blk = total_blocks; aos_body0 = 0; }      // whole array
    if (n_blk == 0) return cudaSuccess;
    if (precision == 0)
        aos_to_blocked_kernel<float><<<n_blk, BLK, 0, st>>>((const float*)aos, n, i_blk0, n_iblk, (float*)pos, (float*)vel, PAD_F32, blk_first, aos_body0);
    else
        aos_to_blocked_kernel<double><<<n_blk, BLK, 0, st>>>((const double*)aos, n, i_blk0, n_iblk, (double*)pos, (double*)vel, PAD_F64, blk_first, aos_body0);
    return cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(BLK) blocked_to_aos_kernel(const T* __restrict__ pos, const T* __restrict__ vel, int n,
                                                            T* __restrict__ aos) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const long long body = (long long)b * BLK + lane;
    if (body >= n) return;
    const T* p = pos + (size_t)b * 3 * BLK + lane;
    const T* v = vel + (size_t)b * 3 * BLK + lane;
    T* o = aos + body * 6;
    o[0] = p[0]; o[1] = p[BLK]; o[2] = p[2 * BLK];
    o[3] = v[0]; o[4] = v[BLK]; o[5] = v[2 * BLK];
}

cudaError_t blocked_to_aos_launch(int precision, const void* pos, const void* vel, int n, void* aos, cudaStream_t st) {
    const int nb = (n + BLK - 1) / BLK;
    if (nb == 0) return cudaSuccess;
    if (precision == 0) blocked_to_aos_kernel<float><<<nb, BLK, 0, st>>>((const float*)pos, (const float*)vel, n, (float*)aos);
    else blocked_to_aos_kernel<double><<<nb, BLK, 0, st>>>((const double*)pos, (const double*)vel, n, (double*)aos);
    return cudaGetLastError();
}

// blocked [nb][3][BLK] -> packed {ax,ay,az} per body
template <typename T>
__global__ void __launch_bounds__(BLK) blocked_to_a3_kernel(const T* __restrict__ acc, int n, T* __restrict__ a3) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const long long body = (long long)b * BLK + lane;
    if (body >= n) return;
    const T* p = acc + (size_t)b * 3 * BLK + lane;
    T* o = a3 + body * 3;
    o[0] = p[0]; o[1] = p[BLK]; o[2] = p[2 * BLK];
}
cudaError_t blocked_to_a3_launch(int precision, const void* acc, int n, void* a3, cudaStream_t st) {
    const int nb = (n + BLK - 1) / BLK;
    if (nb == 0) return cudaSuccess;
    if (precision == 0) blocked_to_a3_kernel<float><<<nb, BLK, 0, st>>>((const float*)acc, n, (float*)a3);
    else blocked_to_a3_kernel<double><<<nb, BLK, 0, st>>>((const double*)acc, n, (double*)a3);
    return cudaGetLastError();
}

// FPGA mailbox image: 16-byte body words {x,y,z,pad} (top_level.vhd:206-208) -> blocked positions,
// and blocked accelerations -> 16-byte result words {Fx,Fy,Fz,0} (compute_store.vhd:242).
__global__ void __launch_bounds__(BLK) mailbox_to_blocked_kernel(const float4* __restrict__ words, int n, float* __restrict__ pos) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const long long body = (long long)b * BLK + lane;
    float x = PAD_F32, y = PAD_F32, z = PAD_F32;
    if (body < n) { const float4 w = words[body]; x = w.x; y = w.y; z = w.z; }
    float* o = pos + (size_t)b * 3 * BLK + lane;
    o[0] = x; o[BLK] = y; o[2 * BLK] = z;
}
__global__ void __launch_bounds__(BLK) blocked_to_mailbox_kernel(const float* __restrict__ acc, int n, float4* __restrict__ words) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const long long body = (long long)b * BLK + lane;
    if (body >= n) return;
    const float* p = acc + (size_t)b * 3 * BLK + lane;
    words[body] = make_float4(p[0], p[BLK], p[2 * BLK], 0.f);
}
cudaError_t mailbox_to_blocked_launch(const float* words, int n, int n_blocks, float* pos, cudaStream_t st) {
    mailbox_to_blocked_kernel<<<n_blocks, BLK, 0, st>>>((const float4*)words, n, pos);
    return cudaGetLastError();
}
cudaError_t blocked_to_mailbox_launch(const float* acc, int n, float* words, cudaStream_t st) {
    const int nb = (n + BLK - 1) / BLK;
    if (nb == 0) return cudaSuccess;
    blocked_to_mailbox_kernel<<<nb, BLK, 0, st>>>(acc, n, (float4*)words);
    return cudaGetLastError();
}

// ---- K5: total energy, FP64 arithmetic whatever the storage type ----------------------------------
// ke = 1/2 sum_i |v_i|^2 over the local slice; pe = -1/2 sum_{i local} sum_{j != i} (r_ij^2 + eps)^(-1/2).
// Diagnostic only (energy drift is reported, never gated on).
__device__ __forceinline__ double rsqrt_f64(double s) {
    // MUFU.RSQ64H seed (~2^-22) + one third-order step: y(1 + e/2 + 3e^2/8), e = 1 - s*y^2
    const double y = rsqrt_approx64(s);
    const double t = s * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double q = y * e;
    return fma(q, p, y);
}

template <typename T>
__global__ void __launch_bounds__(BLK) energy_kernel(const T* __restrict__ pos, const T* __restrict__ vel, int n, int i_blk0,
                                                    int total_blocks, double eps, double* __restrict__ out) {
    __shared__ double sj[3 * BLK];
    __shared__ double red[2 * (BLK / 32)];
    const int ib = blockIdx.x, lane = threadIdx.x;
    const long long body = (long long)(i_blk0 + ib) * BLK + lane;
    const T* pi = pos + (size_t)(i_blk0 + ib) * 3 * BLK + lane;
    const double xi = (double)pi[0], yi = (double)pi[BLK], zi = (double)pi[2 * BLK];
    double u = 0.0;
    const int jblocks = (n + BLK - 1) / BLK;
    for (int jb = 0; jb < jblocks; jb++) {
        const T* pj = pos + (size_t)jb * 3 * BLK + lane;
        __syncthreads();
        sj[lane] = (double)pj[0]; sj[BLK + lane] = (double)pj[BLK]; sj[2 * BLK + lane] = (double)pj[2 * BLK];
        __syncthreads();
        const long long j0 = (long long)jb * BLK;
        const int lim = (int)min((long long)BLK, (long long)n - j0);
#pragma unroll 4
        for (int j = 0; j < lim; j++) {
            const double dx = sj[j] - xi, dy = sj[BLK + j] - yi, dz = sj[2 * BLK + j] - zi;
            const double s = fma(dz, dz, fma(dy, dy, fma(dx, dx, eps)));
            const double r = rsqrt_f64(s);
            u += (j0 + j != body) ? r : 0.0;
        }
    }
    double k = 0.0;
    if (body < n) {
        const T* v = vel + (size_t)ib * 3 * BLK + lane;
        const double vx = (double)v[0], vy = (double)v[BLK], vz = (double)v[2 * BLK];
        k = 0.5 * (vx * vx + vy * vy + vz * vz);
    } else {
        u = 0.0;
    }
    for (int o = 16; o > 0; o >>= 1) { k += __shfl_down_sync(0xffffffffu, k, o); u += __shfl_down_sync(0xffffffffu, u, o); }
    if ((lane & 31) == 0) { red[lane / 32] = k; red[BLK / 32 + lane / 32] = u; }
    __syncthreads();
    if (lane == 0) {
        double ks = 0, us = 0;
        for (int w = 0; w < BLK / 32; w++) { ks += red[w]; us += red[BLK / 32 + w]; }
        atomicAdd(&out[0], ks);
        atomicAdd(&out[1], -0.5 * us);
    }
}

cudaError_t energy_launch(int precision, const void* pos, const void* vel, int n, int i_blk0, int n_iblk, int total_blocks,
                          double eps, double* out, cudaStream_t st) {
    if (n_iblk <= 0) return cudaSuccess;
    if (precision == 0) energy_kernel<float><<<n_iblk, BLK, 0, st>>>((const float*)pos, (const float*)vel, n, i_blk0, total_blocks, eps, out);
    else energy_kernel<double><<<n_iblk, BLK, 0, st>>>((const double*)pos, (const double*)vel, n, i_blk0, total_blocks, eps, out);
    return cudaGetLastError();
}

// ---- FFMA2 peak probe: measured FP32 roofline denominator and SM clock under load ------------------
__global__ void __launch_bounds__(256) ffma_probe_kernel(float* out, long long* cycles, int iters) {
    f2 acc[8];
    const f2 b = pk(1.0001f, 0.9999f), c = pk(0.5f, 0.25f);
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = pk(threadIdx.x * 1e-3f + i, 1.f + i);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = fma2(acc[i], b, c);
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; upk(acc[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
cudaError_t ffma_probe_launch(float* out, long long* cycles, int iters, int grid, cudaStream_t st) {
    ffma_probe_kernel<<<grid, 256, 0, st>>>(out, cycles, iters);
    return cudaGetLastError();
}

}  // namespace nb
