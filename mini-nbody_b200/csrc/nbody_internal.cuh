// Internal declarations shared by the kernels and the C-ABI layer of libnbody_b200.so.
// sm_100a only.  Nothing here is exported.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nb {

// ---------------------------------------------------------------------------------------------
// Data layout in HBM ("tile-blocked SoA"): bodies are grouped in blocks of BLK = 128; block b of
// an array holds   [x_0..x_127][y_0..y_127][z_0..z_127]   (3*BLK scalars, 1536 B in FP32,
// 3072 B in FP64).  One j-stage of the force kernel is a contiguous run of blocks, so a single
// 1-D TMA bulk copy (cp.async.bulk) brings it into shared memory, and a float4 LDS on the x-row
// yields four consecutive j-bodies = two f32x2 operand pairs.  The reference's own body word is
// {x,y,z,pad} (top_level.vhd:206-208); positions only, no mass (fxyz.vhd:120-127).
// Rank r of W owns blocks [r*local_blocks, (r+1)*local_blocks); the array is padded to
// W*local_blocks blocks with bodies at PAD_COORD, whose contribution underflows to exactly 0
// (the analogue of the reference's write mask for padding slots, top_level.vhd:201-205).
// ---------------------------------------------------------------------------------------------
constexpr int BLK = 128;
constexpr float EPS_F32 = 1.0e-9f;     // 0x3089705F, dzsoft.vhd:177
constexpr double EPS_F64 = 1.0e-9;
constexpr float PAD_F32 = 1.0e18f;     // d^2 ~ 3e36 finite; rsqrt^3 ~ 2e-55 -> 0
constexpr double PAD_F64 = 1.0e150;    // d^2 ~ 3e300 finite; rsqrt^3 ~ 2e-451 -> 0
constexpr int MAX_SLOTS = 96;          // partial-acceleration slots (j-splits) per step

// What happens to a tile of i-bodies once its accelerations are complete: the integrate step (or parts of it), run by
// the CTA that completed the tile.  The same record serves the split-grid kernels' fused mode and the stream-K kernels.
struct TileEpilogue {
    const void* pos;           // positions the pass reads (pos[cur]); x_next = x + dt_x * v starts from them
    void* pos_next; void* vel; void* acc_out;          // each may be null
    double dt_v, dt_x;         // v += dt_v * a ; x_next = x + dt_x * v
    int i_blk0, n_iblk, n, i_tiles;
    // push exchange (optional): every finished tile also goes into each peer's pos[next] over NVLink and the CTA that
    // finishes the rank's LAST tile publishes flag_value in slot flag_index of every peer's flag array
    void* const* peer_pos_next; unsigned long long* const* peer_flags; unsigned int* done_counter;
    unsigned long long flag_value; int flag_index; int n_peers;
};
// other ranks' step flags a pass has to acquire before it reads their j-slices (push exchange; flags == null otherwise)
// `seen` is a device-local word: the first CTA that has acquired all peers' flags (7 system-scope loads, ~1 us each)
// records the step there, every later CTA of the rank needs one device-scope load instead
struct PeerWait { const unsigned long long* flags; int count, skip; unsigned long long value; int* err; unsigned long long* seen; };

struct ForceArgs {
    const void* pos;       // blocked SoA positions, total_blocks blocks (all ranks' bodies)
    void* part;            // partial accelerations [slot][n_iblk][3][BLK]
    int total_blocks;
    int i_blk0;            // first global block of this rank's i-slice
    int n_iblk;            // i-blocks of this rank
    int j_rot0;            // physical block at which this launch's (rotated) j-range starts
    int j_len;             // blocks in the j-range (wraps modulo total_blocks)
    int nsplit;            // j-splits of this launch == gridDim.y
    int slot0;             // first output slot
    float eps32;           // softening added to dist^2; read only by the run-time-softening instantiations
    double eps64;          //   (the default FP32 kernels carry 1e-9 as an immediate operand, dzsoft.vhd:177)
    // ---- fused mode (fuse != 0): 1-D grid, partial sums go to a ring of `ring` tiles x nsplit slots (L2-resident) instead
    // of one slot array per split, the CTA that completes a tile's last split adds the tile's slots in split order (fixed order: same sum as integrate_kernel, bit for bit)
    // and runs the epilogue; no integrate launch, no partial sums in HBM (reference analogue: the on-chip adder tree,
    // S/final_adder.vhd:88-104, S/compute_store.vhd:139-173)
    int fuse;
    int order;                       // 1-D grids: 0 = split-major over all tiles, 1 = split-major inside groups of ring/2 tiles
    int ring;                        // tiles whose slots are live at once; tile t uses ring position t % ring
    void* ws;                        // [ring][nsplit][I*3][THREADS]
    unsigned int* tile_counter;      // [i_tiles], zero between passes
    unsigned int* tile_done;         // [i_tiles], == epoch once the tile is reduced (its ring position may be reused)
    unsigned int epoch;
    int local_len;                   // j-blocks (rotated coordinates) that are this rank's own: CTAs reaching past them wait
    PeerWait wait;
    TileEpilogue ep;
};

// fused multi-step kernel (single GPU, launch-bound sizes): see step_fused_f32_kernel in force_f32.cu
struct FusedStepArgs {
    void* pos[2];          // double-buffered positions (blocked SoA, n_iblk blocks)
    void* vel;             // velocities
    void* part;            // partial accelerations [nsplit][n_iblk][3][BLK]
    unsigned int* tile_counter;   // i_tiles counters, zero between launches
    int n, n_iblk, i_tiles, nsplit;
    int cur;               // pos[cur] holds the state at entry
    int nsteps;
    float dt_v, dt_x;      // v += dt_v * a ; x += dt_x * v
    float eps32;
};

// whole-array-in-shared-memory multi-step kernel for small systems (step_small.cu)
struct SmallStepArgs {
    void* pos[2];          // double-buffered positions (blocked SoA, n_iblk blocks)
    void* vel;
    int n, n_iblk;
    int ipc;               // i-bodies per CTA
    int cur, nsteps;
    int write_pos;         // 1: full steps; 0: kick only (one step, positions and cur untouched)
    float dt_v, dt_x, eps32;
};
bool step_small_plan(int n, int sms, int* ipc, int* ctas, int* il);
cudaError_t step_small_launch(const SmallStepArgs& a, int ctas, int il, bool eps_rt, cudaStream_t st);

// ---- stream-K force pass (stream.cuh, force_f32.cu, force_f64.cu) -----------------------------------------
// One persistent launch per force pass.  The work of a pass is the linearised space
//     (i-tile t, j-granule g)  ->  u = t * L + g,      granule = GRAN consecutive j-bodies,
// of each PHASE (a rotated j-range; one phase on a single GPU, two when sharded: the rank's own j-slice first,
// the other ranks' slices second, so the exchange of the previous step hides under phase 0).  CTA c of G owns
// the contiguous unit range [c*U/G, (c+1)*U/G) of every phase: equal work to within one granule, whatever N.
// A range cuts at most two tiles per phase; a (tile, CTA) piece is a SEGMENT.  A tile whose only segment is one
// CTA's is finished by that CTA straight from shared memory; otherwise each segment's sums go to workspace slot
// phase*(T+G) + t + c (unique and monotone along the sweep), a per-tile counter finds the last arriver, and
// that CTA adds the tile's segments IN SLOT ORDER (fixed order: deterministic, independent of arrival order)
// and runs the integrate epilogue.  The workspace is (T+G) slots per phase and is rewritten every launch, so
// it lives in L2: partial sums no longer travel to HBM and back, and no separate integrate kernel runs.
// Reference analogue: the on-chip adder tree over the 16 interleaved partials, S/final_adder.vhd:88-104,
// S/compute_store.vhd:139-173.
constexpr int GRAN = 16;               // j-bodies per granule (one iteration of the unrolled inner loop)
constexpr int GPB = BLK / GRAN;        // granules per layout block
constexpr int STREAM_MAX_PHASES = 2;

struct StreamArgs {
    const void* pos;           // blocked SoA positions of all ranks (pos[cur])
    int total_blocks, i_blk0, n_iblk, n;
    int i_tiles, grid;         // T, G
    int nphase, ph_begin, ph_end;              // phases of the pass / phases THIS launch works on
    int ph_rot0[STREAM_MAX_PHASES];            // physical block at which the phase's rotated j-range starts
    int ph_len[STREAM_MAX_PHASES];             // length of the phase's j-range in granules
    float eps32; double eps64;
    void* ws;                  // segment sums [nphase*(T+G)][I*3][THREADS]
    unsigned int* tile_counter;                // T counters, zero between passes
    int store_all;             // 1: every segment goes to ws and nothing is reduced here (stream_reduce_kernel does it)
    int coop;                  // 1: cut tiles are reduced cooperatively by their contributors (needs all CTAs resident and all
                               //    phases in this launch); 0: by the tile's last arriver
    TileEpilogue ep;           // what integrate_kernel does for a finished tile (ep.pos == pos)
    PeerWait wait;             // phases >= wait_from read other ranks' positions: acquire their step flags first
    int wait_from;
    unsigned long long* prof;  // optional per-CTA timeline (globaltimer ns): [grid][8] = entry, first stage landed, loops done, segments, last-arriver reductions, exit
};

// (tile, split) of CTA `bid` of a fused split-grid pass (1-D grid of i_tiles * nsplit CTAs): split-major over all tiles
// (order 0), or split-major inside groups of ring/2 tiles taken one after the other (order 1).  Shared by the kernel and by
// nbody_fused_cta(), which lets the CPU tests check that the map is a bijection and that ring positions are reused safely.
__host__ __device__ __forceinline__ void fused_cta_of(int bid, int i_tiles, int nsplit, int ring, int order, int* tile, int* split) {
    if (order == 1) {
        const int gs = ring / 2 > 0 ? ring / 2 : 1, per_group = gs * nsplit;
        const int group = bid / per_group, r = bid % per_group;
        const int rest = i_tiles - group * gs, in_group = rest < gs ? rest : gs;
        *tile = group * gs + r % in_group; *split = r / in_group;
    } else {
        *tile = bid % i_tiles; *split = bid / i_tiles;
    }
}

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint32_t bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NB_DONE;\n"
        "bra NB_WAIT;\n"
        "NB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion reported on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// packed FP32 pairs (FFMA2 / FADD2 / FMUL2 on sm_100a)
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double rsqrt_approx64(double x) { double r; asm("rsqrt.approx.ftz.f64 %0,%1;" : "=d"(r) : "d"(x)); return r; }

// ---- kernel launchers (defined in the .cu files) -------------------------------------------------
struct ForceVariant {
    const char* name;
    int i_per_thread, threads, stage_blocks, stages, packed;
    int ctas_per_sm_hint;      // resident CTAs/SM expected from the register count (host-only planning)
    int fold;                  // two-level accumulation (second level in shared memory)
    int eps_rt;                // softening read from ForceArgs instead of the 1e-9 immediate
    int stream;                // stream-K persistent kernel (StreamArgs) instead of the (i-tile, j-split) grid
    int tile_bodies() const { return i_per_thread * threads; }
};
int force_f32_num_variants();
const ForceVariant& force_f32_variant(int v);
cudaError_t force_f32_launch(int variant, const ForceArgs& a, cudaStream_t st);
cudaError_t force_f32_setup(int variant);   // opt-in shared memory etc.; once per device
int force_f32_occupancy(int variant);       // resident CTAs/SM on the current device
bool force_f32_fused_supported(int variant);
cudaError_t force_f32_fused_launch(int variant, const FusedStepArgs& fa, int sms, cudaStream_t st, int* grid_out);

cudaError_t force_f32_stream_launch(int variant, const StreamArgs& a, cudaStream_t st);
cudaError_t force_f64_stream_launch(int variant, const StreamArgs& a, cudaStream_t st);
cudaError_t force_f32_stream_reduce_launch(int variant, const StreamArgs& a, cudaStream_t st);   // twin of the in-kernel reduction
cudaError_t force_f64_stream_reduce_launch(int variant, const StreamArgs& a, cudaStream_t st);

int force_f64_num_variants();
const ForceVariant& force_f64_variant(int v);
cudaError_t force_f64_launch(int variant, const ForceArgs& a, cudaStream_t st);
cudaError_t force_f64_setup(int variant);
int force_f64_occupancy(int variant);

// integrate / layout / energy kernels (integrate.cu)
struct IntegrateArgs {
    const void* part;      // [slots][n_iblk][3][BLK]
    int slots;
    int n_iblk;
    int i_blk0;            // global block offset of the local slice
    int n;                 // total bodies (global index >= n is padding and never moves)
    const void* pos_cur;   // full position array (current)
    void* pos_next;        // full position array (next); local slice is written
    void* vel;             // local velocities [n_iblk][3][BLK]
    void* acc_out;         // optional: summed accelerations [n_iblk][3][BLK] (may be null)
    double dt_v;           // v += dt_v * a
    double dt_x;           // x_next = x + dt_x * v
    // push exchange (optional): the kernel also stores its slice into every peer's pos_next through
    // peer-mapped memory and the last CTA to finish publishes flag_value in slot flag_index of every
    // peer's flag array (system-scope release after all CTAs' stores)
    void* const* peer_pos_next;  // device array of n_peers peer-mapped pos_next base pointers
    unsigned long long* const* peer_flags;   // device array of n_peers peer-mapped flag arrays
    unsigned int* done_counter;  // local, zero between launches
    unsigned long long flag_value;
    int flag_index;
    int n_peers;
};
cudaError_t flag_signal_launch(unsigned long long* const* peer_flags, int n_peers, int index, unsigned long long value, cudaStream_t st);
cudaError_t flag_wait_launch(const unsigned long long* flags, int count, int skip, unsigned long long value, int* err, cudaStream_t st);
cudaError_t integrate_launch(int precision, const IntegrateArgs& a, cudaStream_t st);
cudaError_t aos_to_blocked_launch(int precision, const void* aos, int n, int i_blk0, int n_iblk, int total_blocks,
                                  void* pos_blocks, void* vel_blocks, cudaStream_t st,
                                  int blk_first = 0, int n_blk = -1, long long aos_body0 = 0);
cudaError_t blocked_to_aos_launch(int precision, const void* pos_blocks, const void* vel_blocks, int n,
                                  void* aos, cudaStream_t st);
cudaError_t blocked_to_a3_launch(int precision, const void* acc_blocks, int n, void* a3, cudaStream_t st);
cudaError_t mailbox_to_blocked_launch(const float* words, int n, int n_blocks, float* pos_blocks, cudaStream_t st);
cudaError_t blocked_to_mailbox_launch(const float* acc_blocks, int n, float* words, cudaStream_t st);
cudaError_t energy_launch(int precision, const void* pos, const void* vel, int n, int i_blk0, int n_iblk,
                          int total_blocks, double eps, double* out_ke_pe, cudaStream_t st);
cudaError_t ffma_probe_launch(float* out, long long* cycles, int iters, int grid, cudaStream_t st);

}  // namespace nb
