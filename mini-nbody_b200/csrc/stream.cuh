// Stream-K driver shared by the FP32 and FP64 force kernels: work decomposition, the fixed-order reduction of a
// tile's segments by the last-arriving CTA, and the integrate epilogue fused behind it (see StreamArgs in
// nbody_internal.cuh for the scheme).  The per-precision part -- one segment's sums, left in shared memory by the
// thread that owns the body -- is the SEG functor passed in by force_f32.cu / force_f64.cu.
#pragma once
#include "nbody_internal.cuh"

namespace nb {

// CTA whose unit range [c*U/G, (c+1)*U/G) holds unit u: the largest c with floor(c*U/G) <= u
__host__ __device__ __forceinline__ int stream_cta_of(long long u, long long U, int G) {
    return (int)(((u + 1) * G - 1) / U);
}
__host__ __device__ __forceinline__ long long stream_lo(int c, long long U, int G) { return ((long long)c * U) / G; }

// segments of tile t over all phases of the pass
__host__ __device__ __forceinline__ int stream_tile_segments(int t, int nphase, const int* ph_len, int T, int G) {
    int n = 0;
    for (int p = 0; p < nphase; p++) {
        const long long L = ph_len[p], U = (long long)T * L;
        n += stream_cta_of((long long)(t + 1) * L - 1, U, G) - stream_cta_of((long long)t * L, U, G) + 1;
    }
    return n;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

// spin until every peer's step flag reached w.value (one thread, in front of its first bulk copy that reads other
// ranks' slices); gives up after ~20 s and raises *err instead of hanging the device
__device__ __forceinline__ void wait_for_peers(const PeerWait& w) {
    if (w.seen) {
        unsigned long long s;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(s) : "l"(w.seen) : "memory");
        if (s >= w.value) { asm volatile("fence.proxy.async;" ::: "memory"); return; }
    }
    const unsigned long long t0 = globaltimer_ns();
    for (int p = 0; p < w.count; p++) {
        if (p == w.skip) continue;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w.flags + p) : "memory");
            if (v >= w.value) break;
            if (globaltimer_ns() - t0 > 20000000000ull) { atomicExch(w.err, 1); break; }
            __nanosleep(100);
        }
    }
    // later CTAs of this rank synchronise with this one (release/acquire at device scope is cumulative over the system-scope
    // acquires above); a racing CTA that stores the same value is harmless
    if (w.seen) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(w.seen), "l"(w.value) : "memory");
    // the peers' stores were made through the generic proxy; the bulk copies that read them use the async proxy
    asm volatile("fence.proxy.async;" ::: "memory");
}

// The integrate step for one finished tile (what integrate_kernel does per body): acc[q][d] is the total acceleration
// of the thread's q-th body.  Optional acceleration record, v += dt_v*a, x_next = x + dt_x*v, the push into the peers'
// next-step buffers, and the step flag once the last tile of the rank is through.  Called by all threads of the CTA.
template <typename T, int I, int THREADS>
__device__ __forceinline__ void tile_epilogue(const TileEpilogue& e, const int tile, const T (&acc)[I][3],
                                              const int col_lo = 0, const int col_hi = I * THREADS, const bool tile_complete = true) {
    // columns: body q of thread tid is column q*THREADS + tid of the tile; [col_lo, col_hi) is the share this call integrates
    // (the whole tile unless the tile's segments were reduced cooperatively); tile_complete: this call finishes the tile
    constexpr int IB = I * THREADS / BLK, TB = THREADS / BLK;
    const int tid = threadIdx.x, lane = tid % BLK;
    const T* __restrict__ pc = static_cast<const T*>(e.pos);
    T* __restrict__ pn = static_cast<T*>(e.pos_next);
    T* __restrict__ vel = static_cast<T*>(e.vel);
    T* __restrict__ ao = static_cast<T*>(e.acc_out);
    const T dtv = (T)e.dt_v, dtx = (T)e.dt_x;
#pragma unroll
    for (int q = 0; q < I; q++) {
        const int ib = tile * IB + q * TB + tid / BLK;
        if (ib >= e.n_iblk) continue;
        if (q * THREADS + tid < col_lo || q * THREADS + tid >= col_hi) continue;
        const size_t loc = (size_t)ib * 3 * BLK + lane;
        const size_t glb = (size_t)(e.i_blk0 + ib) * 3 * BLK + lane;
        if (ao) { ao[loc] = acc[q][0]; ao[loc + BLK] = acc[q][1]; ao[loc + 2 * BLK] = acc[q][2]; }
        if (!vel) continue;                                   // acceleration-only pass (nbody_accel)
        T vx = vel[loc], vy = vel[loc + BLK], vz = vel[loc + 2 * BLK];
        T x = pc[glb], y = pc[glb + BLK], z = pc[glb + 2 * BLK];
        if ((long long)(e.i_blk0 + ib) * BLK + lane < e.n) {  // padding bodies never move
            vx = fma(dtv, acc[q][0], vx); vy = fma(dtv, acc[q][1], vy); vz = fma(dtv, acc[q][2], vz);
            x = fma(vx, dtx, x); y = fma(vy, dtx, y); z = fma(vz, dtx, z);
        }
        vel[loc] = vx; vel[loc + BLK] = vy; vel[loc + 2 * BLK] = vz;
        if (pn) {
            pn[glb] = x; pn[glb + BLK] = y; pn[glb + 2 * BLK] = z;
            for (int r = 0; r < e.n_peers; r++) {             // push exchange: NVLink stores into every peer's pos[next]
                T* pp = static_cast<T*>(e.peer_pos_next[r]) + glb;
                pp[0] = x; pp[BLK] = y; pp[2 * BLK] = z;
            }
        }
    }
    if (e.n_peers > 0 && e.peer_flags != nullptr && pn != nullptr) {
        __threadfence_system();                               // this tile's peer stores are visible system-wide ...
        __syncthreads();
        if (tid == 0 && tile_complete) {                      // ... before the tile is counted; the last tile publishes the step
            if (atomicAdd(e.done_counter, 1u) == (unsigned)e.i_tiles - 1u) {
                *e.done_counter = 0u;
                __threadfence_system();
                for (int r = 0; r < e.n_peers; r++)
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(e.peer_flags[r] + e.flag_index), "l"(e.flag_value) : "memory");
            }
        }
    }
}

// Finish tile `tile` of a stream-K pass: total acceleration of each of its bodies = the tile's segment sums added in
// slot order (from_ws) or this CTA's own sums (`res`, shared memory, entry (q*3+d)*THREADS + tid), then the epilogue.
template <typename T, int I, int THREADS>
__device__ __forceinline__ void stream_finish_tile(const StreamArgs& a, const int tile, const T* res, const bool from_ws,
                                                   const int col_lo = 0, const int col_hi = I * THREADS, const bool tile_complete = true) {
    const int tid = threadIdx.x;
    T acc[I][3];
    if (!from_ws) {
#pragma unroll
        for (int q = 0; q < I; q++)
#pragma unroll
            for (int d = 0; d < 3; d++) acc[q][d] = res[(size_t)(q * 3 + d) * THREADS + tid];
    } else {
#pragma unroll
        for (int q = 0; q < I; q++) acc[q][0] = acc[q][1] = acc[q][2] = (T)0;
        const T* __restrict__ ws = static_cast<const T*>(a.ws);
        for (int p = 0; p < a.nphase; p++) {
            const long long L = a.ph_len[p], U = (long long)a.i_tiles * L;
            const int cf = stream_cta_of((long long)tile * L, U, a.grid), cl = stream_cta_of((long long)(tile + 1) * L - 1, U, a.grid);
            const T* w = ws + (size_t)(p * (a.i_tiles + a.grid) + tile + cf) * (I * 3 * THREADS) + tid;
            // fixed order => deterministic sum.  Two segments' loads are in flight at a time (the adds stay in slot order): the
            // tile's last arriver is alone on the critical path of the pass, and one segment per L2 round trip was 10 us for 19
            int c = cf;
            for (; c + 1 <= cl; c += 2, w += (size_t)2 * (I * 3 * THREADS)) {
                T v0[I * 3], v1[I * 3];
#pragma unroll
                for (int k = 0; k < I * 3; k++) {
                    const bool mine = (k / 3) * THREADS + tid >= col_lo && (k / 3) * THREADS + tid < col_hi;
                    v0[k] = mine ? __ldcg(w + (size_t)k * THREADS) : (T)0; v1[k] = mine ? __ldcg(w + (size_t)(I * 3 + k) * THREADS) : (T)0;
                }
#pragma unroll
                for (int k = 0; k < I * 3; k++) { acc[k / 3][k % 3] += v0[k]; acc[k / 3][k % 3] += v1[k]; }
            }
            if (c <= cl) {
#pragma unroll
                for (int k = 0; k < I * 3; k++)
                    if ((k / 3) * THREADS + tid >= col_lo && (k / 3) * THREADS + tid < col_hi) acc[k / 3][k % 3] += __ldcg(w + (size_t)k * THREADS);
            }
        }
    }
    tile_epilogue<T, I, THREADS>(a.ep, tile, acc, col_lo, col_hi, tile_complete);
}

// The persistent loop of one CTA.  seg(tile, phase, ja, jb) computes the sums of the tile's bodies over granules
// [ja, jb) of the phase's j-range and leaves them in `res` (shared memory, written and read by the owning thread).
// The CTA's unit range of every phase is worked out once (thread 0, the only 64-bit divisions on the way to the
// first bulk copy) and parked in shared memory: nothing of the decomposition stays in registers across the hot
// loop, where every register is spoken for.
template <typename T, int I, int THREADS, typename SEG>
__device__ __forceinline__ void stream_run(const StreamArgs& a, T* res, SEG&& seg) {
    __shared__ int s_last, s_waited;
    __shared__ long long s_u0[STREAM_MAX_PHASES], s_u1[STREAM_MAX_PHASES];
    __shared__ int s_t0[STREAM_MAX_PHASES], s_n[STREAM_MAX_PHASES];
    __shared__ unsigned long long s_prof[5];               // entry, last segment done, segments, reductions, ns in reductions
    const int tid = threadIdx.x;
    if (tid == 0) {
        if (a.prof) { s_prof[1] = s_prof[2] = s_prof[3] = s_prof[4] = 0; s_prof[0] = globaltimer_ns(); }
        s_waited = 0;
#pragma unroll 1
        for (int p = a.ph_begin; p < a.ph_end; p++) {
            const long long L = a.ph_len[p], U = (long long)a.i_tiles * L;
            const long long u0 = stream_lo(blockIdx.x, U, a.grid), u1 = stream_lo(blockIdx.x + 1, U, a.grid);
            const int t0 = (int)(u0 / L);
            s_u0[p] = u0; s_u1[p] = u1; s_t0[p] = t0;
            s_n[p] = u1 > u0 ? (int)((u1 - 1) / L) - t0 + 1 : 0;
        }
    }
    __syncthreads();
    for (int p = a.ph_begin; p < a.ph_end; p++) {
        for (int e = 0; e < *(volatile int*)&s_n[p]; e++) {
            const int t = *(volatile int*)&s_t0[p] + e;
            int ja, jb;
            {
                const long long L = a.ph_len[p], tl = (long long)t * L;
                const long long u0 = *(volatile long long*)&s_u0[p], u1 = *(volatile long long*)&s_u1[p];
                ja = e == 0 ? (int)(u0 - tl) : 0;
                jb = (int)min(u1 - tl, L);
            }
            if (p >= a.wait_from && a.wait.flags != nullptr && !*(volatile int*)&s_waited) {
                __syncthreads();                               // everybody has read s_waited == 0
                if (tid == 0) { wait_for_peers(a.wait); s_waited = 1; }
                __syncthreads();                               // the producer thread issues its bulk copies behind the acquire
            }
            seg(t, p, ja, jb);
            if (a.prof && tid == 0) { s_prof[1] = globaltimer_ns(); s_prof[2]++; }
            // the tile is this CTA's alone: finish it straight from shared memory
            if (!a.store_all && a.nphase == 1 && ja == 0 && jb == a.ph_len[p]) { stream_finish_tile<T, I, THREADS>(a, t, res, false); continue; }
            T* w = static_cast<T*>(a.ws) + (size_t)(p * (a.i_tiles + a.grid) + t + (int)blockIdx.x) * (I * 3 * THREADS) + tid;
#pragma unroll
            for (int k = 0; k < 3 * I; k++) w[(size_t)k * THREADS] = res[(size_t)k * THREADS + tid];
            if (a.store_all) continue;
            __threadfence();                                   // this segment's sums are visible device-wide ...
            __syncthreads();                                   // ... before the CTA counts itself
            if (a.coop) {                                      // cooperative reduction: just count; the shares follow below
                if (tid == 0) atomicAdd(a.tile_counter + t, 1u);
                continue;
            }
            if (tid == 0) {
                const int nseg = stream_tile_segments(t, a.nphase, a.ph_len, a.i_tiles, a.grid);
                const bool last = atomicAdd(a.tile_counter + t, 1u) == (unsigned)nseg - 1u;
                if (last) a.tile_counter[t] = 0u;              // ready for the next pass (stream-ordered)
                s_last = last;
            }
            __syncthreads();
            if (s_last) {
                unsigned long long t0 = 0;
                if (a.prof && tid == 0) t0 = globaltimer_ns();
                __threadfence();
                stream_finish_tile<T, I, THREADS>(a, t, res, true);
                if (a.prof && tid == 0) { s_prof[4] += globaltimer_ns() - t0; s_prof[3]++; }
            }
            __syncthreads();                                   // s_last may be rewritten by the next segment
        }
    }
    if (a.coop) {
        // Cooperative reduction (all CTAs of the pass are resident, so waiting for each other is safe): every CTA that holds a
        // segment of a cut tile waits until all of the tile's segments are in, then adds and integrates ITS share of the tile's
        // bodies -- the k-th of nseg contributors takes columns [k*C/nseg, (k+1)*C/nseg) -- with the segments in the same
        // slot order as ever, so the result does not depend on who reduces what.  The tile's last arriver no longer works
        // alone at the end of the pass (19 segments x 24 KB + epilogue: 10-17 us per tile at C3 x 8 against a 500 us pass).
        __shared__ int s_k, s_nseg;
        for (int p = a.ph_begin; p < a.ph_end; p++) {
            for (int e = 0; e < *(volatile int*)&s_n[p]; e++) {
                const int t = *(volatile int*)&s_t0[p] + e;
                {
                    const long long L = a.ph_len[p], tl = (long long)t * L;
                    const long long u0 = *(volatile long long*)&s_u0[p], u1 = *(volatile long long*)&s_u1[p];
                    const int ja = e == 0 ? (int)(u0 - tl) : 0, jb = (int)min(u1 - tl, L);
                    if (a.nphase == 1 && ja == 0 && jb == L) continue;          // finished alone above
                }
                __syncthreads();                               // s_k / s_nseg of the previous share are no longer read
                if (tid == 0) {
                    int k = 0, nseg = 0;
                    for (int q = 0; q < a.nphase; q++) {
                        const long long L = a.ph_len[q], U = (long long)a.i_tiles * L;
                        const int cf = stream_cta_of((long long)t * L, U, a.grid), cl = stream_cta_of((long long)(t + 1) * L - 1, U, a.grid);
                        if (q < p) k += cl - cf + 1;
                        if (q == p) k += (int)blockIdx.x - cf;
                        nseg += cl - cf + 1;
                    }
                    s_k = k; s_nseg = nseg;
                    const unsigned long long t0 = globaltimer_ns();
                    for (;;) {
                        unsigned int v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.tile_counter + t) : "memory");
                        if (v >= (unsigned)nseg) break;
                        if (globaltimer_ns() - t0 > 20000000000ull) { if (a.wait.err) atomicExch(a.wait.err, 3); break; }
                        __nanosleep(64);
                    }
                }
                __syncthreads();
                const int k = s_k, nseg = s_nseg;
                constexpr int C = I * THREADS;
                const int lo = (int)(((long long)k * C) / nseg), hi = (int)(((long long)(k + 1) * C) / nseg);
                unsigned long long tp = 0;
                if (a.prof && tid == 0) tp = globaltimer_ns();
                // the share that brings the tile's counter to 2 * nseg completes the tile (and resets the counter)
                __shared__ int s_complete;
                stream_finish_tile<T, I, THREADS>(a, t, res, true, lo, hi, false);
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    const bool last = atomicAdd(a.tile_counter + t, 1u) == 2u * (unsigned)nseg - 1u;
                    if (last) a.tile_counter[t] = 0u;
                    s_complete = last;
                }
                __syncthreads();
                if (s_complete && a.ep.n_peers > 0 && a.ep.peer_flags != nullptr && a.ep.pos_next != nullptr && tid == 0) {
                    // every share fenced its peer stores system-wide before it counted itself: the tile is out
                    if (atomicAdd(a.ep.done_counter, 1u) == (unsigned)a.ep.i_tiles - 1u) {
                        *a.ep.done_counter = 0u;
                        __threadfence_system();
                        for (int r = 0; r < a.ep.n_peers; r++)
                            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.ep.peer_flags[r] + a.ep.flag_index), "l"(a.ep.flag_value) : "memory");
                    }
                }
                if (a.prof && tid == 0) { s_prof[4] += globaltimer_ns() - tp; s_prof[3]++; }
            }
        }
    }
    if (a.prof && tid == 0) {
        unsigned long long* o = a.prof + (size_t)blockIdx.x * 8;
        o[0] = s_prof[0]; o[1] = s_prof[1]; o[2] = s_prof[2]; o[3] = s_prof[3]; o[4] = s_prof[4]; o[5] = globaltimer_ns();
        unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); o[6] = smid;
    }
}

// twin of the in-kernel reduction for store_all passes: one CTA per tile adds the segments in slot order
template <typename T, int I, int THREADS>
__global__ void __launch_bounds__(THREADS) stream_reduce_kernel(const StreamArgs a) {
    stream_finish_tile<T, I, THREADS>(a, blockIdx.x, nullptr, true);
}

}  // namespace nb
