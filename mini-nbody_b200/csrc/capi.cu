// The C-ABI layer of libnbody_b200.so (include/nbody.h): resident state, planning, streams,
// the per-step all-gather and the reference-shaped drop-in entry points.  Host code only; the
// kernels live in force_f32.cu / force_f64.cu / integrate.cu.
//
// Execution model per step and per rank (one CUDA device each):
//   compute stream:  [force pass A over the rank's own j-slice]           (needs nothing remote)
//                    wait(all-gather of this step's positions done)
//                    [force pass B over the other ranks' j-slices]
//                    [integrate: sum partial slots, v += dt*a, x_next = x + dt*v (own slice)]
//   comm stream:     wait(integrate) -> all-gather x_next slices (NCCL, in place) -> event
// Positions are double-buffered (pos[cur] is read, pos[cur^1] is written), so the exchange of
// step t overlaps pass A of step t+1.  With "exchange"=1 the integrate kernel itself stores its
// slice into every peer's pos[cur^1] through peer-mapped memory and the all-gather is replaced by
// a flag handshake (signal/wait kernels).
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is bound at run time (see NcclApi)

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nbody.h"
#include "nbody_internal.cuh"
#include "stream.cuh"

using namespace nb;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? -4 : -2, "CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #x); } while (0)
#define NC(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(-3, "NCCL error '%s' at %s:%d (%s)", g_nccl.GetErrorString(r_), __FILE__, __LINE__, #x); } while (0)
#define OK(x) do { int rc_ = (x); if (rc_ != 0) return rc_; } while (0)

// NCCL is bound lazily with dlopen so that (a) a single-GPU user needs no NCCL at all and (b) inside a
// process that already carries an NCCL (e.g. the one bundled with PyTorch) that copy is reused instead
// of a second, possibly older, libnccl.so.2 being pulled in ahead of it.
struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.so) return 0;
    const char* names[] = {getenv("NBODY_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* so = nullptr;
    for (const char* nm : names) { if (nm && *nm) { so = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (so) break; } }
    if (!so) return fail(-3, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define NB_SYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(so, name); if (!g_nccl.field) return fail(-3, "NCCL symbol %s missing", name); } while (0)
    NB_SYM(GetUniqueId, "ncclGetUniqueId"); NB_SYM(CommInitRank, "ncclCommInitRank"); NB_SYM(CommInitAll, "ncclCommInitAll");
    NB_SYM(CommDestroy, "ncclCommDestroy"); NB_SYM(AllGather, "ncclAllGather"); NB_SYM(AllReduce, "ncclAllReduce");
    NB_SYM(GroupStart, "ncclGroupStart"); NB_SYM(GroupEnd, "ncclGroupEnd"); NB_SYM(GetErrorString, "ncclGetErrorString");
#undef NB_SYM
    g_nccl.so = so;
    return 0;
}

// The library switches the calling thread's current device while it drives its ranks; every public entry
// point restores the caller's device on the way out.
struct DeviceGuard {
    int dev = -1;
    DeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
};

constexpr int MAX_WORLD = 64;
struct EvPair { cudaEvent_t a, b; int kind; };   // kind 0 = force, 1 = integrate

struct Rank {
    int device = 0, rank = 0;
    cudaStream_t st = nullptr, st_comm = nullptr;
    cudaEvent_t ev_local = nullptr, ev_gather = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    void* pos[2] = {nullptr, nullptr};
    void* vel = nullptr;
    void* part = nullptr;
    void* acc = nullptr;
    void* staging = nullptr;       // device AoS staging: the rank's slice (6 scalars per body); grown to all n bodies by the
    size_t staging_bytes = 0;      //   calls that move whole arrays (nbody_download, nbody_accel on the first rank)
    void* gather_tmp = nullptr;    // total_blocks*3*BLK scalars (velocity / acceleration gather)
    double* energy = nullptr;      // 2 doubles
    size_t part_bytes = 0;
    ncclComm_t comm = nullptr;
    std::vector<EvPair> evs; size_t ev_used = 0;
    // push exchange: peer-mapped views of the other ranks' buffers (IPC handles or peer access)
    void** peer_pos_dev[2] = {nullptr, nullptr};   // device arrays [n_peers] of peers' pos[b]
    unsigned long long* flags = nullptr;           // local: [2*MAX_WORLD] step flags, then epoch flags
    unsigned long long** peer_flags_dev = nullptr; // device array [n_peers]: peers' step-flag arrays
    unsigned long long** peer_epoch_dev = nullptr; // device array [n_peers]: peers' epoch-flag arrays
    unsigned int* done_counter = nullptr;
    unsigned int* tile_counter = nullptr;          // fused step kernel: one counter per i-tile, zero between launches
    int* err_flag = nullptr;
    unsigned long long* prof = nullptr;            // stream-K per-CTA timeline of the last pass (option "profile")
    std::vector<void*> ipc_opened;                 // pointers obtained with cudaIpcOpenMemHandle
    int n_peers = 0;
    bool push_ready = false;
    unsigned long long flag_waited = 0;            // step-flag value a wait kernel is already enqueued for on r.st
    bool owns_streams = true;                      // virtual ranks of one device share the first rank's streams
};
}  // namespace

struct nbody_ctx {
    int n = 0, precision = 0, world = 1;
    int esize = 4;
    int total_blocks = 0, local_blocks = 0;
    int cur = 0;
    bool have_state = false;
    bool single_process = true;
    bool virtual_ranks = false;      // several ranks of this process on ONE device (NBODY_VIRTUAL_RANKS=1): the sharded path on a one-GPU box
    int variant = 0, opt_splits = 0, opt_overlap = 1, opt_exchange = 0, opt_timing = 0;
    int opt_stream = -1;             // stream-K force pass: -1 auto (default_variant), 0 never pick a stream variant by default
    int opt_grid = 0;                // stream-K: CTAs of the persistent launch (0 = resident slots, sms * ctas_per_sm)
    int opt_fuse = -1;               // split-grid FP32 kernels: in-kernel last-arriver reduction + integrate (-1/1 on, 0 = slot array + integrate kernel)
    int fuse_ring = 1; unsigned int fuse_epoch = 0;
    int opt_order = -1;              // CTA order of the fused split-grid pass: 1 tile-major + ring, 0 split-major, -1 auto
    int fuse_order = 1;
    int opt_profile = 0;             // stream-K: record a per-CTA timeline of every pass (nbody_stream_profile reads the last one)
    int opt_coop = 0;                // stream-K: 1 = cut tiles reduced cooperatively by their contributors, 0 = by the tile's last arriver
                                     // (default: measured no faster -- the tail is a chain of dependent round trips, not bandwidth:
                                     // a share of 1/19 of a tile takes the same ~9 us as the whole tile)
    int opt_twin = 0;                // stream-K: 1 = every segment to the workspace + separate reduce launch (bit-identity twin)
    int opt_small = -1;              // whole-array-in-shared-memory multi-step kernel (step_small.cu): -1 auto, 0 off, 1 on where it fits
    long long small_launches = 0;
    int opt_fused = -1;              // fused multi-step kernel: -1 auto (single GPU, FP32, narrow variants), 0 off, 1 on
    double softening = 1.0e-9;       // added to dist^2 (S/dzsoft.vhd:177); nbody_set_softening changes it
    int sms = 148, ctas_per_sm = 0;
    nbody_plan_t plan{};
    std::vector<Rank> ranks;      // ranks driven by this process
    long long launches = 0, fused_launches = 0;
    unsigned long long step_counter = 0;
    double last_step_ms = 0;
    bool gather_pending = false;
    // CUDA graph of two consecutive steps (parity-neutral) for launch-bound sizes, single GPU only
    cudaGraphExec_t graph = nullptr; double graph_dt = 0; int graph_variant = -1, graph_slots = -1, graph_cur = -1;
    int opt_graph = -1;                       // -1 auto (bodies per GPU < 65536), 0 off, 1 on
    unsigned long long flag_pending = 0;   // push exchange: step-flag value the next remote-j pass must wait for
    unsigned long long epoch = 0;
    size_t block_bytes() const { return (size_t)3 * BLK * esize; }
};

namespace {

int variant_count(int precision) { return precision == NBODY_F32 ? force_f32_num_variants() : force_f64_num_variants(); }
const ForceVariant& variant_of(int precision, int v) { return precision == NBODY_F32 ? force_f32_variant(v) : force_f64_variant(v); }
// Time of one CTA per j-block while k CTAs share its SM, relative to the same with all occ CTAs resident:
// k * cpi(k) / (occ * cpi(occ)), cpi = measured cycles per interaction at that many warps per sub-partition
// (tools/microbench/loop.cu, profiles/r01_microbench_loop.jsonl; I = 8 with the re-scheduled loop: 13.0 alone, 12.0
// in pairs).  A CTA that has the SM to itself runs ~1.85x as fast as one of a resident pair.
double alone_factor(const ForceVariant& v, int k, int occ) {
    if (k >= occ) return 1.0;
    static const double cpi1[] = {0, 22.2, 16.1, 15.2, 14.5, 14.2, 14.0, 13.8};      // I = 1, 128 threads: 1..7 CTAs/SM
    static const double cpi2[] = {0, 15.3, 13.7, 13.4, 13.2};                        // I = 2, 128 threads: 1..4
    if (v.threads == 128 && v.i_per_thread == 1 && occ <= 7) return k * cpi1[k] / (occ * cpi1[occ]);
    if (v.threads == 128 && v.i_per_thread == 2 && occ <= 4) return k * cpi2[k] / (occ * cpi2[occ]);
    return (double)k / occ * (1.0 + 0.085 * (occ - k));                              // I >= 4: 13.0 vs 12.0 at occ = 2
}

// Choose the number of j-splits for one force launch.  Costs are in units of "one layout block of j for one CTA
// with the SM fully occupied", with a fixed prologue/epilogue cost per CTA.
//  * More CTAs than resident slots: whole waves in lock step when the CTAs are long (>= 4 blocks of j), otherwise
//    dynamic scheduling, max(perfectly balanced time, one CTA) + half a CTA of tail (sweeps of S at N = 4096 ... 1M,
//    profiles/r01_small_n_probe.jsonl, r01_mid_n_sweep*.jsonl).
//  * At most one resident wave (few i-tiles: N below ~25 000 per GPU): every CTA starts at once and the launch
//    lasts as long as one CTA on the busiest SM, which hosts k = ceil(CTAs / SMs) of them -- and a CTA that shares
//    its SM with fewer than occ others runs faster (alone_factor).  profiles/r01_mid_n_sweep.jsonl: at N = 6144 the
//    1024-body tiles with 24 splits (144 CTAs, one per SM) take 20.6 us, 48 splits (288 CTAs, two per SM) 24.6,
//    32 splits (192 CTAs: 44 SMs get two) 34.9.
// Bounds: accumulator chains no longer than CHAIN_BODIES j (accuracy), at most 48 splits per launch.
int choose_splits(const ForceVariant& v, int i_tiles, int j_len, int sms, int occ, int forced) {
    if (j_len <= 0) return 0;
    const int CHAIN_BODIES = 65536;
    const int wave_slots = sms * occ;
    int smin = std::max(1, (int)(((long long)j_len * BLK + CHAIN_BODIES - 1) / CHAIN_BODIES));
    smin = std::min(smin, 48);
    const int smax = std::max(smin, std::min(j_len, 48));
    if (forced > 0) return std::max(1, std::min(forced, std::min(j_len, 48)));
    double best = 1e300; int best_s = smin;
    for (int s = smin; s <= smax; s++) {
        const double unit = (double)((j_len + s - 1) / s) + 0.5;      // blocks per CTA + fixed cost
        const long long ctas = (long long)i_tiles * s;
        double cost;
        if (ctas <= wave_slots) {
            cost = (unit - 0.4) * alone_factor(v, (int)((ctas + sms - 1) / sms), occ);     // fixed cost per CTA ~0.1 block
        } else if ((double)j_len / s >= 4.0) {
            // several waves of long CTAs run in lock step: full waves at full occupancy, then the remainder, which
            // shares the SMs more thinly (fits the sweeps at N = 32 768 ... 1M to ~1 %: e.g. N = 131 072, 128 tiles:
            // 37 splits = 16 waves exactly 5 556 us, 32 splits 5 609, 41 splits 5 615)
            const double u_avg = (double)j_len / s, u_max = (double)((j_len + s - 1) / s);
            const long long full = ctas / wave_slots, rem = ctas % wave_slots;
            cost = full * (u_avg + 0.1) + (u_max - u_avg);
            if (rem) cost += (u_max + 0.1) * alone_factor(v, (int)((rem + sms - 1) / sms), occ);
        } else {
            // many short CTAs of uneven length desynchronise: dynamic scheduling, balanced time + half a CTA of tail
            const double balanced = (double)ctas * unit / wave_slots;
            cost = 0.85 * (std::max(balanced, unit) + 0.5 * unit);   // 0.85: this estimate runs 10-17 % above measured times
        }
        cost += 0.02 * s;                                             // integrate reads s more slots
        if (cost < best * (1.0 - 1e-9)) { best = cost; best_s = s; }
    }
    return best_s;
}

int make_plan(int n, int precision, int rank, int world, int sms, int variant, int forced_splits, int overlap, int ctas_per_sm, nbody_plan_t* out, int forced_grid = 0) {
    if (n <= 0) return fail(-1, "n must be positive (got %d)", n);
    if (world < 1 || rank < 0 || rank >= world) return fail(-1, "bad rank/world %d/%d", rank, world);
    if (precision != NBODY_F32 && precision != NBODY_F64) return fail(-1, "precision must be NBODY_F32 or NBODY_F64");
    if (variant < 0 || variant >= variant_count(precision)) return fail(-1, "variant %d out of range [0,%d)", variant, variant_count(precision));
    if (sms <= 0) sms = 148;
    const ForceVariant& v = variant_of(precision, variant);
    nbody_plan_t p{};
    p.n = n; p.world = world; p.rank = rank; p.blk = BLK;
    const int nblocks = (n + BLK - 1) / BLK;
    p.local_blocks = (nblocks + world - 1) / world;
    p.total_blocks = p.local_blocks * world;
    p.i_begin = std::min(n, rank * p.local_blocks * BLK);
    p.i_end = std::min(n, (rank + 1) * p.local_blocks * BLK);
    p.tile_bodies = v.tile_bodies();
    const int ib = p.tile_bodies / BLK;
    p.i_tiles = (p.local_blocks + ib - 1) / ib;
    if (ctas_per_sm <= 0) ctas_per_sm = v.ctas_per_sm_hint;
    if (v.stream) {
        // stream-K: G persistent CTAs share the (i-tile, j-granule) space of each phase evenly; one phase, or the
        // rank's own j-slice first and the other ranks' slices second (the exchange hides under the first)
        // two phases (own j-slice first) hide the exchange of the previous step under 1/world of this one -- worth its second set
        // of cut tiles only while that share is long against the exchange: C3 x 8 (0.5 ms per step) runs 1 % faster with one
        // (profiles/r02c_bench_c3_*_n8.json: 7 620 vs 7 541 G inter/s)
        const double est_us = (double)std::min(n, p.local_blocks * BLK) * n / (precision == NBODY_F32 ? 3.1e6 : 1.08e6);
        p.stream_phases = (world == 1 || !overlap || (overlap < 2 && est_us < 1000.0)) ? 1 : 2;
        const long long gl = (long long)p.local_blocks * GPB, gt = (long long)p.total_blocks * GPB;
        const long long umin = (long long)p.i_tiles * (p.stream_phases == 1 ? gt : std::min(gl, gt - gl));
        long long g = forced_grid > 0 ? (long long)forced_grid : (long long)sms * ctas_per_sm;
        g = std::min(g, std::max(1LL, umin / GPB));          // at least one layout block of j per CTA and phase
        p.stream_grid = (int)std::max(1LL, g);
        p.splits_local = p.splits_remote = p.slots = 0;
        *out = p;
        return 0;
    }
    if (world == 1 || !overlap) {
        p.splits_local = choose_splits(v, p.i_tiles, p.total_blocks, sms, ctas_per_sm, forced_splits);
        p.splits_remote = 0;
    } else {
        p.splits_local = choose_splits(v, p.i_tiles, p.local_blocks, sms, ctas_per_sm, forced_splits);
        p.splits_remote = choose_splits(v, p.i_tiles, p.total_blocks - p.local_blocks, sms, ctas_per_sm, forced_splits);
    }
    p.slots = p.splits_local + p.splits_remote;
    if (p.slots > MAX_SLOTS) return fail(-5, "internal: %d slots exceed MAX_SLOTS", p.slots);
    *out = p;
    return 0;
}

int set_dev(const Rank& r) { CU(cudaSetDevice(r.device)); return 0; }

int free_rank(Rank& r) {
    cudaSetDevice(r.device);
    if (r.st) cudaStreamSynchronize(r.st);
    if (r.st_comm) cudaStreamSynchronize(r.st_comm);
    if (r.comm && g_nccl.so) g_nccl.CommDestroy(r.comm);
    for (int b = 0; b < 2; b++) { if (r.pos[b]) cudaFree(r.pos[b]); if (r.peer_pos_dev[b]) cudaFree(r.peer_pos_dev[b]); }
    for (void* p : r.ipc_opened) cudaIpcCloseMemHandle(p);
    void* ptrs[] = {r.vel, r.part, r.acc, r.staging, r.gather_tmp, r.energy, r.flags, r.peer_flags_dev, r.peer_epoch_dev, r.done_counter, r.tile_counter, r.err_flag, r.prof};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& e : r.evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaEvent_t es[] = {r.ev_local, r.ev_gather, r.ev_t0, r.ev_t1};
    for (cudaEvent_t e : es) if (e) cudaEventDestroy(e);
    if (r.owns_streams) { if (r.st) cudaStreamDestroy(r.st); if (r.st_comm) cudaStreamDestroy(r.st_comm); }
    r = Rank{};
    return 0;
}

bool is_stream(const nbody_ctx* h);
bool fuse_applies(const nbody_ctx* h);

int ensure_part(nbody_ctx* h, Rank& r) {
    // split-grid variants: one slot of partial accelerations per j-split; stream-K: (T + G) segment slots of one
    // tile each per phase (rewritten every pass, L2-resident)
    size_t need = (size_t)std::max(1, h->plan.slots) * h->local_blocks * h->block_bytes();
    if (is_stream(h))
        need = (size_t)h->plan.stream_phases * (h->plan.i_tiles + h->plan.stream_grid) * h->plan.tile_bodies * 3 * h->esize;
    else if (fuse_applies(h))       // fused split-grid mode: a ring of tiles x slots, L2-resident
        need = (size_t)(h->fuse_order == 1 ? h->fuse_ring : h->plan.i_tiles) * std::max(1, h->plan.slots) * h->plan.tile_bodies * 3 * h->esize;
    if (need <= r.part_bytes) return 0;
    OK(set_dev(r));
    CU(cudaStreamSynchronize(r.st));
    if (r.part) CU(cudaFree(r.part));
    r.part = nullptr; r.part_bytes = 0;
    CU(cudaMalloc(&r.part, need));
    r.part_bytes = need;
    return 0;
}

int ensure_tile_counter(nbody_ctx* h, Rank& r) {
    if (r.tile_counter) return 0;
    OK(set_dev(r));
    // [0, local_blocks): per-tile arrival counters (>= i_tiles of every variant); [local_blocks, 2*local_blocks): per-tile
    // "reduced in pass <epoch>" marks of the fused split-grid mode
    CU(cudaMalloc(&r.tile_counter, (size_t)2 * h->local_blocks * sizeof(unsigned int)));
    CU(cudaMemsetAsync(r.tile_counter, 0, (size_t)2 * h->local_blocks * sizeof(unsigned int), r.st));
    return 0;
}

// default force-kernel instantiation: the widest register blocking (re-scheduled loop, 1024-body tiles) as soon as
// tiles x splits can give every SM a CTA -- profiles/r01_mid_n_sweep.jsonl: from 6144 bodies per GPU it beats the
// narrower tiles at every size (by 13 / 6 / 17 / 2 / 6 % at N = 6144 / 8192 / 12288 / 16384 / 20480) -- and the
// one-body-per-thread shape below that (launch-bound sizes, fused step kernel); with a softening other than the
// reference's 1e-9 the FP32 twins that read it from the kernel arguments (15/17) take their place
int default_variant(const nbody_ctx* h) {
    const int n_local = (h->n + h->world - 1) / h->world;
    if (h->precision != NBODY_F32) {
        // FP64: the stream-K kernel (1024-body tiles, one persistent CTA per SM) from 8192 bodies per GPU -- it beats the
        // split grid + integrate kernel at every size from there (profiles/r02_stream_probe_f64.jsonl: 0.90 / 0.95 / 0.98 /
        // 0.99 of its time at N = 8192 / 16384 / 32768 / 65536) and needs one launch per step
        if (h->opt_stream != 0 && n_local >= 8192) return 5;
        return n_local >= 24576 ? 4 : (n_local >= 16384 ? 1 : 2);                               // profiles/r01_f64_mid_sweep.jsonl
    }
    // FP32: the split grid stays the default.  Its re-scheduled loop runs at 12.0 cycles per interaction only while the two
    // CTAs of an SM are of different age (the yield hints hand the issue slots to the older warp, the younger one fills its
    // bubbles, and short CTAs relay); two persistent stream-K CTAs of the same age end up at 12.3-12.4
    // (profiles/r02_stream_fair.jsonl), so variants 19 / 20 are kept as the measured alternative (opt_stream = 1 selects them)
    const bool dflt = h->softening == 1.0e-9;
    if (n_local >= 6144) return h->opt_stream == 1 ? (dflt ? 19 : 20) : (dflt ? 14 : 15);
    return dflt ? 6 : 17;
}

bool is_stream(const nbody_ctx* h) { return variant_of(h->precision, h->variant).stream != 0; }

// Sizes at which the fused multi-step kernel beats CUDA-graph replay of the two launches on B200
// (tools/fused_probe.py, profiles/r01_fused_probe.jsonl: 8 / 12 / 17 % at N = 2048 / 3072 / 4096, none at 1024 or
// from 6144 up).  The persistent kernel wants ~2 units per SM, not whole waves of small CTAs: 8 j-splits.
bool fused_pays(const nbody_ctx* h) {
    return h->world == 1 && h->precision == NBODY_F32 && (h->variant == 6 || h->variant == 17) && h->n >= 1536 && h->n <= 5120;
}

int replan(nbody_ctx* h) {
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }     // captured launches are stale
    nbody_plan_t p;
    int occ = 0;
    for (auto& r : h->ranks) {
        OK(set_dev(r));
        if (h->precision == NBODY_F32) CU(force_f32_setup(h->variant)); else CU(force_f64_setup(h->variant));
        occ = h->precision == NBODY_F32 ? force_f32_occupancy(h->variant) : force_f64_occupancy(h->variant);
    }
    const int splits = (h->opt_splits == 0 && h->opt_fused < 0 && fused_pays(h)) ? 8 : h->opt_splits;
    // fused split-grid mode sweeps the whole rotated j-range in one launch: one set of splits
    const bool fuse = fuse_applies(h);
    OK(make_plan(h->n, h->precision, h->ranks[0].rank, h->world, h->sms, h->variant, splits, fuse ? 0 : h->opt_overlap, occ, &p, h->opt_grid));
    h->plan = p; h->ctas_per_sm = occ;
    if (fuse) {
        // ring of partial-sum slots = two groups of tiles: a group is swept split-major (whole waves of equal CTAs) while its
        // predecessor's slots are still being read back, so a group must cover more tiles than can be in flight at once
        // The ring is sized in BYTES: 28 MiB stays in L2 for the length of a pass (measured with ncu at N = 1M: 10 MB of DRAM
        // writes per launch), 57 MiB does not (217 MB: lines that wait 40 ms for their rewrite get written back).
        const int s = std::max(1, p.slots), in_flight = (h->sms * std::max(1, occ) + s - 1) / s + 1;
        const size_t tile_slots = (size_t)s * p.tile_bodies * 3 * h->esize;
        const int group = std::max(2 * in_flight, (int)(((size_t)14 << 20) / tile_slots));
        h->fuse_ring = 2 * group;
        // CTA order: one group (= split-major over all tiles, every tile its own slots) when the tiles fit the ring anyway;
        // groups of ring/2 tiles otherwise
        h->fuse_order = h->opt_order >= 0 ? h->opt_order : (p.i_tiles <= h->fuse_ring ? 0 : 1);
    }
    for (auto& r : h->ranks) { OK(ensure_part(h, r)); if (is_stream(h) || fuse) OK(ensure_tile_counter(h, r)); }
    return 0;
}

int ensure_staging(nbody_ctx* h, Rank& r, size_t bytes) {
    if (bytes <= r.staging_bytes) return 0;
    OK(set_dev(r));
    if (r.st) CU(cudaStreamSynchronize(r.st));
    if (r.staging) CU(cudaFree(r.staging));
    r.staging = nullptr; r.staging_bytes = 0;
    CU(cudaMalloc(&r.staging, bytes));
    r.staging_bytes = bytes;
    (void)h;
    return 0;
}

int init_rank(nbody_ctx* h, Rank& r) {
    OK(set_dev(r));
    if (h->virtual_ranks && &r != &h->ranks[0] && h->ranks[0].device == r.device) {
        // One stream for all ranks of the device: rank A's step-t kernel spins for rank B's step-(t-1) flag, and on one
        // device nothing guarantees that B's CTAs get SM slots while A's are spinning -- stream order does
        r.st = h->ranks[0].st; r.st_comm = h->ranks[0].st_comm; r.owns_streams = false;
    } else {
        CU(cudaStreamCreateWithFlags(&r.st, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&r.st_comm, cudaStreamNonBlocking));
    }
    CU(cudaEventCreateWithFlags(&r.ev_local, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&r.ev_gather, cudaEventDisableTiming));
    CU(cudaEventCreate(&r.ev_t0));
    CU(cudaEventCreate(&r.ev_t1));
    const size_t full = (size_t)h->total_blocks * h->block_bytes();
    const size_t loc = (size_t)h->local_blocks * h->block_bytes();
    CU(cudaMalloc(&r.pos[0], full));
    CU(cudaMalloc(&r.pos[1], full));
    CU(cudaMalloc(&r.vel, loc));
    CU(cudaMalloc(&r.acc, loc));
    OK(ensure_staging(h, r, (size_t)std::min<long long>(h->n, (long long)h->local_blocks * BLK) * 6 * h->esize));
    CU(cudaMalloc(&r.energy, 2 * sizeof(double)));
    // [0, MAX_WORLD) step flags written by the peers, [MAX_WORLD, 2*MAX_WORLD) epoch flags, [2*MAX_WORLD] "all peers seen at step" (local)
    CU(cudaMalloc(&r.flags, (2 * MAX_WORLD + 8) * sizeof(unsigned long long)));
    CU(cudaMemset(r.flags, 0, (2 * MAX_WORLD + 8) * sizeof(unsigned long long)));
    CU(cudaMalloc(&r.done_counter, sizeof(unsigned int)));
    CU(cudaMemset(r.done_counter, 0, sizeof(unsigned int)));
    CU(cudaMalloc(&r.err_flag, sizeof(int)));
    CU(cudaMemset(r.err_flag, 0, sizeof(int)));
    CU(cudaMemsetAsync(r.vel, 0, loc, r.st));
    return 0;
}

int create_common(int n, int precision, int world, nbody_ctx** out) {
    if (!out) return fail(-1, "out handle is NULL");
    *out = nullptr;
    if (n <= 0) return fail(-1, "n must be positive (got %d)", n);
    if (precision != NBODY_F32 && precision != NBODY_F64) return fail(-1, "precision must be NBODY_F32 (0) or NBODY_F64 (1)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(-2, "no CUDA device available (%s); libnbody_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    nbody_ctx* h = new nbody_ctx();
    h->n = n; h->precision = precision; h->world = world;
    h->esize = precision == NBODY_F32 ? 4 : 8;
    // default force-kernel instantiation: the widest register blocking once there are enough i-bodies
    // per GPU to fill the machine with its 1024-body tiles, narrower tiles for small problems
    h->variant = default_variant(h);
    if (const char* v = getenv("NBODY_VARIANT")) h->variant = atoi(v);
    if (h->variant < 0 || h->variant >= variant_count(precision)) h->variant = 0;
    *out = h;
    return 0;
}

void record_begin(nbody_ctx* h, Rank& r, int kind) {
    if (!h->opt_timing || &r != &h->ranks[0]) return;
    if (r.ev_used == r.evs.size()) {
        if (r.evs.size() >= 4096) return;
        EvPair p{}; p.kind = kind;
        if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
        r.evs.push_back(p);
    }
    r.evs[r.ev_used].kind = kind;
    cudaEventRecord(r.evs[r.ev_used].a, r.st);
}
void record_end(nbody_ctx* h, Rank& r) {
    if (!h->opt_timing || &r != &h->ranks[0]) return;
    if (r.ev_used >= r.evs.size()) return;
    cudaEventRecord(r.evs[r.ev_used].b, r.st);
    r.ev_used++;
}

int launch_force(nbody_ctx* h, Rank& r, const ForceArgs& a) {
    record_begin(h, r, 0);
    cudaError_t e = h->precision == NBODY_F32 ? force_f32_launch(h->variant, a, r.st) : force_f64_launch(h->variant, a, r.st);
    record_end(h, r);
    CU(e);
    h->launches++;
    return 0;
}

// force passes of one rank for the positions in pos[cur]; partial sums land in r.part
int enqueue_forces(nbody_ctx* h, Rank& r) {
    OK(set_dev(r));
    ForceArgs a{};
    a.eps32 = (float)h->softening; a.eps64 = h->softening;
    a.pos = r.pos[h->cur]; a.part = r.part;
    a.total_blocks = h->total_blocks;
    a.i_blk0 = r.rank * h->local_blocks; a.n_iblk = h->local_blocks;
    auto wait_remote = [&]() -> int {
        if (h->gather_pending) CU(cudaStreamWaitEvent(r.st, r.ev_gather, 0));
        if (h->flag_pending) { CU(flag_wait_launch(r.flags, h->world, r.rank, h->flag_pending, r.err_flag, r.st)); h->launches++; }
        return 0;
    };
    if (h->plan.splits_remote == 0) {
        OK(wait_remote());
        a.j_rot0 = 0; a.j_len = h->total_blocks; a.nsplit = h->plan.splits_local; a.slot0 = 0;
        OK(launch_force(h, r, a));
    } else {
        a.j_rot0 = r.rank * h->local_blocks; a.j_len = h->local_blocks; a.nsplit = h->plan.splits_local; a.slot0 = 0;
        OK(launch_force(h, r, a));
        OK(wait_remote());
        a.j_rot0 = ((r.rank + 1) % h->world) * h->local_blocks; a.j_len = h->total_blocks - h->local_blocks;
        a.nsplit = h->plan.splits_remote; a.slot0 = h->plan.splits_local;
        OK(launch_force(h, r, a));
    }
    return 0;
}

int enqueue_integrate(nbody_ctx* h, Rank& r, int slots, double dt_v, double dt_x, bool write_pos, bool write_vel, void* acc_out) {
    OK(set_dev(r));
    IntegrateArgs ia{};
    ia.part = r.part; ia.slots = slots; ia.n_iblk = h->local_blocks; ia.i_blk0 = r.rank * h->local_blocks; ia.n = h->n;
    ia.pos_cur = r.pos[h->cur]; ia.pos_next = write_pos ? r.pos[h->cur ^ 1] : nullptr;
    ia.vel = write_vel ? r.vel : nullptr; ia.acc_out = acc_out;
    ia.dt_v = dt_v; ia.dt_x = dt_x;
    ia.peer_pos_next = nullptr; ia.peer_flags = nullptr; ia.n_peers = 0;
    if (write_pos && h->opt_exchange == 1 && h->world > 1) {
        ia.peer_pos_next = r.peer_pos_dev[h->cur ^ 1]; ia.peer_flags = r.peer_flags_dev; ia.n_peers = r.n_peers;
        ia.done_counter = r.done_counter; ia.flag_value = h->step_counter + 1; ia.flag_index = r.rank;
    }
    record_begin(h, r, 1);
    cudaError_t e = integrate_launch(h->precision, ia, r.st);
    record_end(h, r);
    CU(e);
    h->launches++;
    return 0;
}

// what happens to a tile's accelerations once they are complete (the integrate step, or parts of it)
struct Epilogue { double dt_v, dt_x; bool write_pos, write_vel; void* acc_out; };

// the integrate step (or parts of it) as the record the kernels' tile epilogue takes
TileEpilogue make_epilogue(nbody_ctx* h, Rank& r, const Epilogue& ep) {
    TileEpilogue e{};
    e.pos = r.pos[h->cur];
    e.pos_next = ep.write_pos ? r.pos[h->cur ^ 1] : nullptr;
    e.vel = ep.write_vel ? r.vel : nullptr;
    e.acc_out = ep.acc_out;
    e.dt_v = ep.dt_v; e.dt_x = ep.dt_x;
    e.i_blk0 = r.rank * h->local_blocks; e.n_iblk = h->local_blocks; e.n = h->n; e.i_tiles = h->plan.i_tiles;
    if (ep.write_pos && h->opt_exchange == 1 && h->world > 1) {
        e.peer_pos_next = r.peer_pos_dev[h->cur ^ 1]; e.peer_flags = r.peer_flags_dev; e.n_peers = r.n_peers;
        e.done_counter = r.done_counter; e.flag_value = h->step_counter + 1; e.flag_index = r.rank;
    }
    return e;
}
PeerWait make_peer_wait(nbody_ctx* h, Rank& r) {
    PeerWait w{};
    w.err = r.err_flag;
    if (h->world > 1 && h->flag_pending) { w.flags = r.flags; w.count = h->world; w.skip = r.rank; w.value = h->flag_pending; w.seen = r.flags + 2 * MAX_WORLD; }
    return w;
}

// stream-K force pass with the fused epilogue: one persistent launch (two when the NCCL all-gather has to be
// waited for between the phases -- a kernel cannot wait for a stream event half way)
int enqueue_stream_pass(nbody_ctx* h, Rank& r, const Epilogue& ep) {
    OK(set_dev(r));
    StreamArgs a{};
    a.pos = r.pos[h->cur];
    a.total_blocks = h->total_blocks; a.i_blk0 = r.rank * h->local_blocks; a.n_iblk = h->local_blocks; a.n = h->n;
    a.i_tiles = h->plan.i_tiles; a.grid = h->plan.stream_grid;
    a.nphase = h->plan.stream_phases;
    if (a.nphase == 1) { a.ph_rot0[0] = r.rank * h->local_blocks; a.ph_len[0] = h->total_blocks * GPB; }
    else {
        a.ph_rot0[0] = r.rank * h->local_blocks; a.ph_len[0] = h->local_blocks * GPB;
        a.ph_rot0[1] = ((r.rank + 1) % h->world) * h->local_blocks; a.ph_len[1] = (h->total_blocks - h->local_blocks) * GPB;
    }
    a.eps32 = (float)h->softening; a.eps64 = h->softening;
    a.ws = r.part; a.tile_counter = r.tile_counter; a.store_all = h->opt_twin;
    a.ep = make_epilogue(h, r, ep);
    // remote positions: phase 1 (or the only phase of a sharded pass without overlap) reads the other ranks' slices;
    // push exchange: the CTAs acquire the peers' step flags themselves
    const int remote_from = a.nphase == 2 ? 1 : 0;
    a.wait = make_peer_wait(h, r); a.wait_from = remote_from;
    if (h->opt_profile) {
        if (!r.prof) CU(cudaMalloc(&r.prof, (size_t)65536 * 8 * sizeof(unsigned long long)));
        a.prof = r.prof;
    }
    auto launch = [&](int p0, int p1) -> int {
        a.ph_begin = p0; a.ph_end = p1;
        record_begin(h, r, 0);
        cudaError_t e = h->precision == NBODY_F32 ? force_f32_stream_launch(h->variant, a, r.st) : force_f64_stream_launch(h->variant, a, r.st);
        record_end(h, r);
        CU(e);
        h->launches++;
        return 0;
    };
    // cut tiles are reduced cooperatively by their contributors when the CTAs may wait for each other: one launch for all
    // phases and every CTA resident at once
    const bool one_launch = !(h->world > 1 && h->gather_pending && remote_from > 0);
    a.coop = (h->opt_coop != 0 && !a.store_all && one_launch && a.grid <= h->sms * std::max(1, h->ctas_per_sm)) ? 1 : 0;
    if (h->world > 1 && h->gather_pending) {      // NCCL exchange: stream-ordered wait in front of the first remote phase
        if (remote_from > 0) OK(launch(0, remote_from));
        CU(cudaStreamWaitEvent(r.st, r.ev_gather, 0));
        OK(launch(remote_from, a.nphase));
    } else {
        OK(launch(0, a.nphase));
    }
    if (a.store_all) {
        record_begin(h, r, 1);
        cudaError_t e = h->precision == NBODY_F32 ? force_f32_stream_reduce_launch(h->variant, a, r.st) : force_f64_stream_reduce_launch(h->variant, a, r.st);
        record_end(h, r);
        CU(e);
        h->launches++;
    }
    return 0;
}

// Split-grid force pass in fused mode (FP32): ONE launch over the rank's whole rotated j-range (own slice first), CTAs
// in tile-major order, partial sums in an L2-resident ring, last-arriver reduction + integrate epilogue in the kernel.
// The peers' step flags (push exchange) are acquired by the CTAs whose j-range leaves the rank's own slice.
bool fuse_applies(const nbody_ctx* h) {
    if (h->precision != NBODY_F32 || is_stream(h) || h->opt_fuse == 0) return false;
    if (h->world > 1 && h->opt_exchange != 1) return false;     // the NCCL all-gather is waited for between two launches
    if (h->opt_fuse == 1) return true;
    // auto: where the pass lasts long enough that a tile's last arriver adding its slots alone is noise.  One GPU (measured,
    // profiles/r02_fused_ab.jsonl): +15 / +4 / +2.6 % at N = 8192 / 16384 / 32768 against slot array + integrate kernel,
    // +0.1 % at 131072, 0 at 1M, where it removes 440 MB of DRAM traffic per step.  Sharded (profiles/r02_multi_probe_n2.jsonl):
    // one launch instead of four wins from 16384 bodies per GPU (208 vs 226 us; 733 vs 744; 2823 vs 2842), loses at 8192 (95 vs 79)
    return (h->n + h->world - 1) / h->world >= (h->world > 1 ? 16384 : 65536);
}

int enqueue_fused_pass(nbody_ctx* h, Rank& r, const Epilogue& ep) {
    OK(set_dev(r));
    ForceArgs a{};
    a.eps32 = (float)h->softening; a.eps64 = h->softening;
    a.pos = r.pos[h->cur]; a.part = nullptr;
    a.total_blocks = h->total_blocks;
    a.i_blk0 = r.rank * h->local_blocks; a.n_iblk = h->local_blocks;
    a.j_rot0 = r.rank * h->local_blocks; a.j_len = h->total_blocks; a.nsplit = h->plan.slots; a.slot0 = 0;
    a.fuse = 1; a.order = h->fuse_order; a.ring = h->fuse_order == 1 ? h->fuse_ring : h->plan.i_tiles; a.ws = r.part; a.tile_counter = r.tile_counter; a.tile_done = r.tile_counter + h->local_blocks;
    a.epoch = ++h->fuse_epoch;
    a.local_len = h->world > 1 ? h->local_blocks : h->total_blocks;
    a.wait = make_peer_wait(h, r);
    a.ep = make_epilogue(h, r, ep);
    return launch_force(h, r, a);
}

// force pass + epilogue of one rank for the positions in pos[cur]
int enqueue_pass(nbody_ctx* h, Rank& r, const Epilogue& ep) {
    if (is_stream(h)) return enqueue_stream_pass(h, r, ep);
    if (fuse_applies(h)) return enqueue_fused_pass(h, r, ep);
    OK(enqueue_forces(h, r));
    return enqueue_integrate(h, r, h->plan.slots, ep.dt_v, ep.dt_x, ep.write_pos, ep.write_vel, ep.acc_out);
}

// all-gather the freshly written local slices of pos[cur^1] (or any blocked array) across ranks
int enqueue_allgather(nbody_ctx* h, void* (*buf_of)(Rank&, nbody_ctx*), bool on_comm_stream) {
    if (h->world == 1) return 0;
    const size_t count = (size_t)h->local_blocks * 3 * BLK;
    const ncclDataType_t dt = h->precision == NBODY_F32 ? ncclFloat : ncclDouble;
    if (h->virtual_ranks) {
        // no NCCL between ranks that share a device: plain copies (cold paths only: upload, download, nbody_accel)
        if (on_comm_stream) return fail(-5, "the NCCL position exchange is not available with virtual ranks (use exchange = 1)");
        for (auto& r : h->ranks) { OK(set_dev(r)); CU(cudaStreamSynchronize(r.st)); }
        for (auto& r : h->ranks)
            for (auto& q : h->ranks) {
                if (q.rank == r.rank) continue;
                const size_t off = (size_t)q.rank * count * h->esize;
                CU(cudaMemcpyAsync(static_cast<char*>(buf_of(r, h)) + off, static_cast<char*>(buf_of(q, h)) + off, count * h->esize, cudaMemcpyDeviceToDevice, r.st));
            }
        for (auto& r : h->ranks) { OK(set_dev(r)); CU(cudaStreamSynchronize(r.st)); }
        return 0;
    }
    for (auto& r : h->ranks) {
        OK(set_dev(r));
        if (on_comm_stream) {
            CU(cudaEventRecord(r.ev_local, r.st));
            CU(cudaStreamWaitEvent(r.st_comm, r.ev_local, 0));
        }
    }
    NC(g_nccl.GroupStart());
    for (auto& r : h->ranks) {
        char* base = static_cast<char*>(buf_of(r, h));
        const void* send = base + (size_t)r.rank * count * h->esize;
        ncclResult_t rc = g_nccl.AllGather(send, base, count, dt, r.comm, on_comm_stream ? r.st_comm : r.st);
        if (rc != ncclSuccess) { g_nccl.GroupEnd(); return fail(-3, "ncclAllGather failed: %s", g_nccl.GetErrorString(rc)); }
    }
    NC(g_nccl.GroupEnd());
    if (on_comm_stream)
        for (auto& r : h->ranks) { OK(set_dev(r)); CU(cudaEventRecord(r.ev_gather, r.st_comm)); }
    return 0;
}
void* buf_pos_next(Rank& r, nbody_ctx* h) { return r.pos[h->cur ^ 1]; }
void* buf_pos_cur(Rank& r, nbody_ctx* h) { return r.pos[h->cur]; }
void* buf_gather_tmp(Rank& r, nbody_ctx*) { return r.gather_tmp; }

int ensure_gather_tmp(nbody_ctx* h) {
    for (auto& r : h->ranks) {
        if (r.gather_tmp) continue;
        OK(set_dev(r));
        CU(cudaMalloc(&r.gather_tmp, (size_t)h->total_blocks * h->block_bytes()));
    }
    return 0;
}


// ---- push exchange set-up ---------------------------------------------------------------------------
struct IpcBlob { cudaIpcMemHandle_t pos[2]; cudaIpcMemHandle_t flags; };
static_assert(sizeof(IpcBlob) <= NBODY_IPC_BLOB_BYTES, "blob size");

int install_peers(nbody_ctx* h, Rank& r, const std::vector<void*>& pos0, const std::vector<void*>& pos1,
                  const std::vector<unsigned long long*>& flags) {
    // pos0/pos1/flags: one entry per rank of the world (own entry ignored)
    OK(set_dev(r));
    std::vector<void*> p0, p1; std::vector<unsigned long long*> fl, ep;
    for (int q = 0; q < h->world; q++) {
        if (q == r.rank) continue;
        p0.push_back(pos0[q]); p1.push_back(pos1[q]); fl.push_back(flags[q]); ep.push_back(flags[q] + MAX_WORLD);
    }
    r.n_peers = (int)p0.size();
    const size_t nb = sizeof(void*) * (size_t)std::max(1, r.n_peers);
    if (!r.peer_pos_dev[0]) { CU(cudaMalloc(&r.peer_pos_dev[0], nb)); CU(cudaMalloc(&r.peer_pos_dev[1], nb)); CU(cudaMalloc(&r.peer_flags_dev, nb)); CU(cudaMalloc(&r.peer_epoch_dev, nb)); }
    CU(cudaMemcpy(r.peer_pos_dev[0], p0.data(), sizeof(void*) * r.n_peers, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(r.peer_pos_dev[1], p1.data(), sizeof(void*) * r.n_peers, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(r.peer_flags_dev, fl.data(), sizeof(void*) * r.n_peers, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(r.peer_epoch_dev, ep.data(), sizeof(void*) * r.n_peers, cudaMemcpyHostToDevice));
    r.push_ready = true;
    return 0;
}

int setup_push_single_process(nbody_ctx* h) {
    std::vector<void*> pos0(h->world), pos1(h->world); std::vector<unsigned long long*> flags(h->world);
    for (auto& r : h->ranks) { pos0[r.rank] = r.pos[0]; pos1[r.rank] = r.pos[1]; flags[r.rank] = r.flags; }
    for (auto& r : h->ranks) {
        OK(set_dev(r));
        for (auto& q : h->ranks) {
            if (q.rank == r.rank || q.device == r.device) continue;      // ranks sharing a device see each other's memory directly
            int can = 0; CU(cudaDeviceCanAccessPeer(&can, r.device, q.device));
            if (!can) return fail(-2, "device %d cannot access device %d peer memory: push exchange unavailable", r.device, q.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(q.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
            cudaGetLastError();
        }
    }
    for (auto& r : h->ranks) OK(install_peers(h, r, pos0, pos1, flags));
    return 0;
}

// all ranks of the world meet here (flag handshake on the epoch flags); used where the NCCL path would
// be synchronised by its collective: after uploads and when the exchange mode is switched
int epoch_barrier(nbody_ctx* h) {
    if (h->world == 1) return 0;
    h->epoch++;
    for (auto& r : h->ranks) { OK(set_dev(r)); CU(flag_signal_launch(r.peer_epoch_dev, r.n_peers, r.rank, h->epoch, r.st)); h->launches++; }
    for (auto& r : h->ranks) { OK(set_dev(r)); CU(flag_wait_launch(r.flags + MAX_WORLD, h->world, r.rank, h->epoch, r.err_flag, r.st)); h->launches++; }
    for (auto& r : h->ranks) { OK(set_dev(r)); CU(cudaStreamSynchronize(r.st)); }
    return 0;
}

int check_push_errors(nbody_ctx* h) {
    if (h->opt_exchange != 1 || h->world == 1) return 0;
    for (auto& r : h->ranks) {
        int e = 0;
        OK(set_dev(r));
        CU(cudaMemcpy(&e, r.err_flag, sizeof e, cudaMemcpyDeviceToHost));
        if (e) { cudaMemset(r.err_flag, 0, sizeof(int)); return fail(-5, "peer exchange timed out on rank %d (a peer never published its positions)", r.rank); }
    }
    return 0;
}

// positions of the next step are in place locally: make them so everywhere
int exchange_positions(nbody_ctx* h) {
    if (h->world == 1) return 0;
    h->step_counter++;
    if (h->opt_exchange == 1) { h->flag_pending = h->step_counter; return 0; }   // pushed by the integrate kernel
    OK(enqueue_allgather(h, buf_pos_next, true));
    h->gather_pending = true;
    return 0;
}

int sync_all(nbody_ctx* h) {
    for (auto& r : h->ranks) {
        OK(set_dev(r));
        // push exchange: a rank is only quiescent once every peer's slice of the latest step has landed
        if (h->world > 1 && h->flag_pending && r.st && r.flag_waited != h->flag_pending) {
            CU(flag_wait_launch(r.flags, h->world, r.rank, h->flag_pending, r.err_flag, r.st)); h->launches++; r.flag_waited = h->flag_pending;
        }
        CU(cudaStreamSynchronize(r.st));
        CU(cudaStreamSynchronize(r.st_comm));
    }
    return 0;
}

int check_handle(nbody_ctx* h, bool need_state) {
    if (!h) return fail(-1, "handle is NULL");
    if (need_state && !h->have_state) return fail(-5, "no bodies uploaded yet (call nbody_upload first)");
    return 0;
}

int upload_any(nbody_ctx* h, const void* p) {
    OK(check_handle(h, false));
    if (!p) return fail(-1, "body pointer is NULL");
    OK(sync_all(h));
    h->cur = 0; h->gather_pending = false;
    for (auto& r : h->ranks) {
        OK(set_dev(r));
        if (h->world == 1) {
            CU(cudaMemcpyAsync(r.staging, p, (size_t)h->n * 6 * h->esize, cudaMemcpyHostToDevice, r.st));
            CU(aos_to_blocked_launch(h->precision, r.staging, h->n, 0, h->local_blocks, h->total_blocks, r.pos[0], r.vel, r.st));
        } else {
            // sharded: every rank reads only ITS slice of the caller's array over PCIe (1/world of the bytes),
            // converts its own layout blocks (padding included) and the positions are all-gathered on the device
            const long long i0 = std::min<long long>(h->n, (long long)r.rank * h->local_blocks * BLK);
            const long long i1 = std::min<long long>(h->n, (long long)(r.rank + 1) * h->local_blocks * BLK);
            if (i1 > i0)
                CU(cudaMemcpyAsync(r.staging, static_cast<const char*>(p) + (size_t)i0 * 6 * h->esize, (size_t)(i1 - i0) * 6 * h->esize,
                                   cudaMemcpyHostToDevice, r.st));
            CU(aos_to_blocked_launch(h->precision, r.staging, h->n, r.rank * h->local_blocks, h->local_blocks, h->total_blocks,
                                     r.pos[0], r.vel, r.st, r.rank * h->local_blocks, h->local_blocks, i0));
        }
        h->launches++;
    }
    if (h->world > 1) OK(enqueue_allgather(h, buf_pos_cur, false));
    for (auto& r : h->ranks) {
        OK(set_dev(r));
        // the other buffer must hold valid padding too (only local slices + gathered slices get rewritten)
        CU(cudaMemcpyAsync(r.pos[1], r.pos[0], (size_t)h->total_blocks * h->block_bytes(), cudaMemcpyDeviceToDevice, r.st));
    }
    OK(sync_all(h));
    h->flag_pending = 0;
    if (h->opt_exchange == 1) { OK(epoch_barrier(h)); OK(check_push_errors(h)); }   // nobody pushes into a rank that is still uploading
    h->have_state = true;
    return 0;
}

// this rank's bodies only (layout order = caller's order): no collective, 1/world of the PCIe bytes
int download_local_any(nbody_ctx* h, void* p) {
    OK(check_handle(h, true));
    if (!p) return fail(-1, "body pointer is NULL");
    if (h->ranks.size() != 1) return fail(-5, "nbody_download_local needs a handle that drives one GPU (nbody_create_rank, or ngpus = 1)");
    OK(sync_all(h));
    Rank& r = h->ranks[0];
    OK(set_dev(r));
    const long long i0 = std::min<long long>(h->n, (long long)r.rank * h->local_blocks * BLK);
    const long long i1 = std::min<long long>(h->n, (long long)(r.rank + 1) * h->local_blocks * BLK);
    if (i1 <= i0) return 0;
    const char* pos_local = static_cast<const char*>(r.pos[h->cur]) + (size_t)r.rank * h->local_blocks * h->block_bytes();
    CU(blocked_to_aos_launch(h->precision, pos_local, r.vel, (int)(i1 - i0), r.staging, r.st));
    h->launches++;
    CU(cudaMemcpyAsync(p, r.staging, (size_t)(i1 - i0) * 6 * h->esize, cudaMemcpyDeviceToHost, r.st));
    OK(sync_all(h));
    return 0;
}

int download_any(nbody_ctx* h, void* p) {
    OK(check_handle(h, true));
    if (!p) return fail(-1, "body pointer is NULL");
    OK(sync_all(h));
    const void* velfull = nullptr;
    if (h->world > 1) {
        OK(ensure_gather_tmp(h));
        for (auto& r : h->ranks) {
            OK(set_dev(r));
            char* dst = static_cast<char*>(r.gather_tmp) + (size_t)r.rank * h->local_blocks * h->block_bytes();
            CU(cudaMemcpyAsync(dst, r.vel, (size_t)h->local_blocks * h->block_bytes(), cudaMemcpyDeviceToDevice, r.st));
        }
        OK(enqueue_allgather(h, buf_gather_tmp, false));
    }
    Rank& r = h->ranks[0];
    OK(set_dev(r));
    OK(ensure_staging(h, r, (size_t)h->n * 6 * h->esize));
    velfull = h->world > 1 ? r.gather_tmp : r.vel;
    CU(blocked_to_aos_launch(h->precision, r.pos[h->cur], velfull, h->n, r.staging, r.st));
    h->launches++;
    CU(cudaMemcpyAsync(p, r.staging, (size_t)h->n * 6 * h->esize, cudaMemcpyDeviceToHost, r.st));
    OK(sync_all(h));
    return 0;
}

int accel_any(nbody_ctx* h, void* a3) {
    OK(check_handle(h, true));
    if (!a3) return fail(-1, "output pointer is NULL");
    if (h->world > 1) OK(ensure_gather_tmp(h));
    for (auto& r : h->ranks) {
        void* dst = r.acc;
        if (h->world > 1) dst = static_cast<char*>(r.gather_tmp) + (size_t)r.rank * h->local_blocks * h->block_bytes();
        OK(enqueue_pass(h, r, Epilogue{0.0, 0.0, false, false, dst}));
    }
    if (h->world > 1) OK(enqueue_allgather(h, buf_gather_tmp, false));
    Rank& r = h->ranks[0];
    OK(set_dev(r));
    const void* accfull = h->world > 1 ? r.gather_tmp : r.acc;
    OK(ensure_staging(h, r, (size_t)h->n * 3 * h->esize));
    CU(blocked_to_a3_launch(h->precision, accfull, h->n, r.staging, r.st));
    h->launches++;
    CU(cudaMemcpyAsync(a3, r.staging, (size_t)h->n * 3 * h->esize, cudaMemcpyDeviceToHost, r.st));
    OK(sync_all(h));
    return 0;
}

}  // namespace

int try_small_steps(nbody_ctx* h, double dt, int nsteps, bool write_pos);

// =================================================================================================
extern "C" {

const char* nbody_last_error(void) { return g_err.c_str(); }
const char* nbody_version(void) { return "nbody_b200 0.1 (sm_100a)"; }

int nbody_plan(int n, int precision, int rank, int world, int sms, int variant, nbody_plan_t* out) {
    if (!out) return fail(-1, "out is NULL");
    return make_plan(n, precision, rank, world, sms, variant, 0, 1, 0, out);
}

// Host-side walk of the stream-K decomposition (the same arithmetic stream_run executes on the device,
// stream.cuh): the segments CTA `cta` works on, in order.  Rows of 6 ints: phase, tile, ja, jb (granules of 16 j
// inside the phase's j-range), workspace slot, segments of that tile over the whole pass.  Returns the number of
// rows (negative on error).  Used by the CPU tests of the decomposition; needs no GPU.
int nbody_stream_segments(const nbody_plan_t* p, int cta, int* rows, int cap) {
    if (!p || !rows) return fail(-1, "NULL argument");
    if (p->stream_grid <= 0) return fail(-1, "plan is not a stream-K plan");
    if (cta < 0 || cta >= p->stream_grid) return fail(-1, "cta %d out of range [0,%d)", cta, p->stream_grid);
    int ph_len[STREAM_MAX_PHASES] = {p->total_blocks * GPB, 0};
    if (p->stream_phases == 2) { ph_len[0] = p->local_blocks * GPB; ph_len[1] = (p->total_blocks - p->local_blocks) * GPB; }
    const int T = p->i_tiles, G = p->stream_grid;
    int n = 0;
    for (int ph = 0; ph < p->stream_phases; ph++) {
        const long long L = ph_len[ph], U = (long long)T * L;
        long long u0 = stream_lo(cta, U, G);
        const long long u1 = stream_lo(cta + 1, U, G);
        for (int t = (int)(u0 / L); u0 < u1; t++) {
            const long long tl = (long long)t * L, ue = std::min(u1, tl + L);
            if (n >= cap) return fail(-1, "row buffer too small");
            int* r = rows + 6 * n++;
            r[0] = ph; r[1] = t; r[2] = (int)(u0 - tl); r[3] = (int)(ue - tl); r[4] = ph * (T + G) + t + cta;
            r[5] = stream_tile_segments(t, p->stream_phases, ph_len, T, G);
            u0 = ue;
        }
    }
    return n;
}

// the CTA -> (tile, split) map of the fused split-grid pass (nbody_internal.cuh: fused_cta_of), for the CPU tests
int nbody_fused_cta(int i_tiles, int nsplit, int ring, int order, int bid, int* tile, int* split) {
    if (!tile || !split) return fail(-1, "NULL argument");
    if (i_tiles < 1 || nsplit < 1 || ring < 1 || bid < 0 || bid >= i_tiles * nsplit || (order != 0 && order != 1)) return fail(-1, "bad argument");
    fused_cta_of(bid, i_tiles, nsplit, ring, order, tile, split);
    return 0;
}

int nbody_nccl_unique_id(void* out_id) {
    if (!out_id) return fail(-1, "out_id is NULL");
    static_assert(sizeof(ncclUniqueId) <= NBODY_NCCL_ID_BYTES, "id size");
    OK(nccl_load());
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memset(out_id, 0, NBODY_NCCL_ID_BYTES);
    memcpy(out_id, &id, sizeof id);
    return 0;
}

int nbody_create(int n, int precision, int ngpus, nbody_handle* out) {
    DeviceGuard guard_;
    nbody_ctx* h = nullptr;
    if (ngpus < 1) return fail(-1, "ngpus must be >= 1");
    OK(create_common(n, precision, ngpus, &h));
    int ndev = 0; cudaGetDeviceCount(&ndev);
    const char* vr = getenv("NBODY_VIRTUAL_RANKS");
    h->virtual_ranks = ngpus > 1 && vr && atoi(vr) != 0;
    if (ngpus > ndev && !h->virtual_ranks) { delete h; return fail(-1, "ngpus=%d but only %d CUDA devices visible", ngpus, ndev); }
    if (ngpus > 32) { delete h; return fail(-1, "at most 32 ranks per process"); }
    h->single_process = true;
    h->ranks.resize(ngpus);
    int dev0 = 0;
    if (ngpus == 1 || h->virtual_ranks) if (const char* d = getenv("NBODY_DEVICE")) dev0 = atoi(d);
    // NBODY_VIRTUAL_RANKS=1: all ranks on ONE device (testing aid: the sharded path -- slices, push exchange, flags, in-kernel
    // waits -- on a box with a single GPU); no NCCL between them, so the exchange is the peer-memory push
    for (int g = 0; g < ngpus; g++) { h->ranks[g].device = h->virtual_ranks ? dev0 : dev0 + g; h->ranks[g].rank = g; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->ranks[0].device) != cudaSuccess) { delete h; return fail(-2, "cudaGetDeviceProperties failed"); }
    if (prop.major != 10) { delete h; return fail(-2, "device %s is sm_%d%d; libnbody_b200 is built for sm_100a only", prop.name, prop.major, prop.minor); }
    h->sms = prop.multiProcessorCount;
    nbody_plan_t p;
    int rc = make_plan(n, precision, 0, ngpus, h->sms, h->variant, 0, 1, 0, &p);
    if (rc) { delete h; return rc; }
    h->total_blocks = p.total_blocks; h->local_blocks = p.local_blocks;
    for (auto& r : h->ranks) { rc = init_rank(h, r); if (rc) { nbody_destroy(h); return rc; } }
    if (h->virtual_ranks) {
        rc = setup_push_single_process(h);
        if (rc) { nbody_destroy(h); return rc; }
        h->opt_exchange = 1;
    } else if (ngpus > 1) {
        rc = nccl_load();
        if (rc) { nbody_destroy(h); return rc; }
        std::vector<ncclComm_t> comms(ngpus); std::vector<int> devs(ngpus);
        for (int g = 0; g < ngpus; g++) devs[g] = h->ranks[g].device;
        ncclResult_t nr = g_nccl.CommInitAll(comms.data(), ngpus, devs.data());
        if (nr != ncclSuccess) { nbody_destroy(h); return fail(-3, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(nr)); }
        for (int g = 0; g < ngpus; g++) h->ranks[g].comm = comms[g];
    }
    rc = replan(h);
    if (rc) { nbody_destroy(h); return rc; }
    *out = h;
    return 0;
}

int nbody_create_rank(int n, int precision, int rank, int world, int device, const void* nccl_id, nbody_handle* out) {
    DeviceGuard guard_;
    nbody_ctx* h = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(-1, "bad rank/world %d/%d", rank, world);
    if (world > 1 && !nccl_id) return fail(-1, "nccl_id is required when world > 1");
    OK(create_common(n, precision, world, &h));
    int ndev = 0; cudaGetDeviceCount(&ndev);
    if (device < 0 || device >= ndev) { delete h; return fail(-1, "device %d out of range (have %d)", device, ndev); }
    h->single_process = false;
    h->ranks.resize(1);
    h->ranks[0].device = device; h->ranks[0].rank = rank;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return fail(-2, "cudaGetDeviceProperties failed"); }
    if (prop.major != 10) { delete h; return fail(-2, "device %s is sm_%d%d; libnbody_b200 is built for sm_100a only", prop.name, prop.major, prop.minor); }
    h->sms = prop.multiProcessorCount;
    nbody_plan_t p;
    int rc = make_plan(n, precision, rank, world, h->sms, h->variant, 0, 1, 0, &p);
    if (rc) { delete h; return rc; }
    h->total_blocks = p.total_blocks; h->local_blocks = p.local_blocks;
    rc = init_rank(h, h->ranks[0]);
    if (rc) { nbody_destroy(h); return rc; }
    if (world > 1) {
        rc = nccl_load();
        if (rc) { nbody_destroy(h); return rc; }
        ncclUniqueId id; memcpy(&id, nccl_id, sizeof id);
        cudaSetDevice(device);
        ncclResult_t nr = g_nccl.CommInitRank(&h->ranks[0].comm, world, id, rank);
        if (nr != ncclSuccess) { nbody_destroy(h); return fail(-3, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(nr)); }
    }
    rc = replan(h);
    if (rc) { nbody_destroy(h); return rc; }
    *out = h;
    return 0;
}

int nbody_destroy(nbody_handle h) {
    DeviceGuard guard_;
    if (!h) return 0;
    if (h->graph) { cudaSetDevice(h->ranks.empty() ? 0 : h->ranks[0].device); cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    for (auto it = h->ranks.rbegin(); it != h->ranks.rend(); ++it) free_rank(*it);      // sharers of rank 0's streams first
    delete h;
    return 0;
}

int nbody_upload(nbody_handle h, const Body* p) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    if (h->precision != NBODY_F32) return fail(-1, "handle is FP64: use nbody_upload_d");
    return upload_any(h, p);
}
int nbody_upload_d(nbody_handle h, const BodyD* p) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    if (h->precision != NBODY_F64) return fail(-1, "handle is FP32: use nbody_upload");
    return upload_any(h, p);
}
int nbody_download(nbody_handle h, Body* p) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    if (h->precision != NBODY_F32) return fail(-1, "handle is FP64: use nbody_download_d");
    return download_any(h, p);
}
int nbody_download_local(nbody_handle h, Body* p) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    if (h->precision != NBODY_F32) return fail(-1, "handle is FP64: use nbody_download_local_d");
    return download_local_any(h, p);
}
int nbody_download_local_d(nbody_handle h, BodyD* p) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    if (h->precision != NBODY_F64) return fail(-1, "handle is FP32: use nbody_download_local");
    return download_local_any(h, p);
}
int nbody_download_d(nbody_handle h, BodyD* p) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    if (h->precision != NBODY_F64) return fail(-1, "handle is FP32: use nbody_download");
    return download_any(h, p);
}

int nbody_step_async(nbody_handle h, double dt, int nsteps) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    if (nsteps < 0) return fail(-1, "nsteps must be >= 0");
    Rank& r0 = h->ranks[0];
    OK(set_dev(r0));
    CU(cudaEventRecord(r0.ev_t0, r0.st));
    int s0 = 0;
    {
        const int took = try_small_steps(h, dt, nsteps, true);
        if (took < 0) return took;
        if (took) { CU(cudaEventRecord(r0.ev_t1, r0.st)); return 0; }
    }
    // launch-bound sizes on one GPU: all nsteps in ONE cooperative launch (force units + last-arriver integrate +
    // one grid barrier per step); same instantiation and splits as the two-kernel path => bit-identical state
    if (h->world == 1 && h->precision == NBODY_F32 && !h->opt_timing && nsteps >= 1 && h->plan.splits_remote == 0 &&
        force_f32_fused_supported(h->variant) && (h->opt_fused == 1 || (h->opt_fused < 0 && fused_pays(h)))) {
        OK(ensure_tile_counter(h, r0));
        FusedStepArgs fa{};
        fa.pos[0] = r0.pos[0]; fa.pos[1] = r0.pos[1]; fa.vel = r0.vel; fa.part = r0.part; fa.tile_counter = r0.tile_counter;
        fa.n = h->n; fa.n_iblk = h->local_blocks; fa.i_tiles = h->plan.i_tiles; fa.nsplit = h->plan.splits_local;
        fa.cur = h->cur; fa.nsteps = nsteps; fa.dt_v = (float)dt; fa.dt_x = (float)dt; fa.eps32 = (float)h->softening;
        int grid = 0;
        cudaError_t fe = force_f32_fused_launch(h->variant, fa, h->sms, r0.st, &grid);
        if (fe == cudaSuccess) {
            h->launches += 1; h->fused_launches += 1;
            h->cur ^= (nsteps & 1);
            CU(cudaEventRecord(r0.ev_t1, r0.st));
            return 0;
        }
        if (h->opt_fused == 1) return fail(-(int)fe, "fused step kernel: %s", cudaGetErrorString(fe));
        cudaGetLastError();                        // auto mode: fall back to the two-kernel path of the same library
    }
    // (a captured fused pass replays with the epoch it was captured with: fine while every tile has its own ring position,
    // not with slot reuse, where the reuse wait compares against the epoch)
    const bool want_graph = h->world == 1 && !h->opt_timing && nsteps >= 4 && !(fuse_applies(h) && h->fuse_order == 1) &&
                            (h->opt_graph == 1 || (h->opt_graph < 0 && h->n < 65536));
    if (want_graph) {
        // two steps leave pos[cur] where it started, so one captured pair can be replayed nsteps/2 times
        if (!h->graph || h->graph_dt != dt || h->graph_variant != h->variant || h->graph_slots != h->plan.slots || h->graph_cur != h->cur) {
            if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
            cudaGraph_t g = nullptr;
            const long long launches_before = h->launches;
            CU(cudaStreamBeginCapture(r0.st, cudaStreamCaptureModeThreadLocal));
            int rc = 0;
            for (int k = 0; k < 2 && !rc; k++) {
                rc = enqueue_pass(h, r0, Epilogue{dt, dt, true, true, nullptr});
                h->cur ^= 1;
            }
            cudaError_t ce = cudaStreamEndCapture(r0.st, &g);
            h->launches = launches_before;
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            CU(ce);
            ce = cudaGraphInstantiate(&h->graph, g, 0);
            cudaGraphDestroy(g);
            CU(ce);
            h->graph_dt = dt; h->graph_variant = h->variant; h->graph_slots = h->plan.slots; h->graph_cur = h->cur;
        }
        for (; s0 + 2 <= nsteps; s0 += 2) { CU(cudaGraphLaunch(h->graph, r0.st)); h->launches += is_stream(h) ? 2 : 4; }
    }
    for (int s = s0; s < nsteps; s++) {
        for (auto& r : h->ranks) OK(enqueue_pass(h, r, Epilogue{dt, dt, true, true, nullptr}));
        OK(exchange_positions(h));
        h->cur ^= 1;
    }
    OK(set_dev(r0));
    if (h->world > 1 && h->gather_pending) CU(cudaStreamWaitEvent(r0.st, r0.ev_gather, 0));
    if (h->world > 1 && h->flag_pending && r0.flag_waited != h->flag_pending) {      // the timed region ends when every peer's slice has landed
        CU(flag_wait_launch(r0.flags, h->world, r0.rank, h->flag_pending, r0.err_flag, r0.st)); h->launches++; r0.flag_waited = h->flag_pending;
    }
    CU(cudaEventRecord(r0.ev_t1, r0.st));
    return 0;
}

}  // extern "C" (helper below is internal)

// Small systems on one GPU: all nsteps in ONE cooperative launch, every CTA sweeps all j for its own 28-128 bodies out of a
// shared-memory copy of the whole position array -- no partial sums between CTAs (step_small.cu).  Auto up to SMALL_AUTO_MAX
// bodies and only where the caller asked for no particular path; write_pos = false is the kick alone (bodyForce), so that
// nbody_step == nbody_body_force + nbody_integrate stays true bit for bit.  Returns 1 if the kernel took the call, 0 if the
// tiled paths should, < 0 on error.
int try_small_steps(nbody_ctx* h, double dt, int nsteps, bool write_pos) {
    int ipc = 0, ctas = 0, il = 0;
    constexpr int SMALL_AUTO_MAX = 8192;
    if (!(h->world == 1 && h->precision == NBODY_F32 && !h->opt_timing && nsteps >= 1 && h->opt_small != 0)) return 0;
    if (!(h->opt_small == 1 || (h->n <= SMALL_AUTO_MAX && h->variant == default_variant(h) && h->opt_splits == 0 && h->opt_fused < 0 &&
                                h->opt_fuse < 0 && h->opt_graph < 0 && h->opt_stream < 0))) return 0;
    if (!step_small_plan(h->n, h->sms, &ipc, &ctas, &il)) return 0;
    // auto: one body per lane (N <= 32 * SMs = 4736), or two with the lanes >= 85 % used (N from ~8000): measured
    // (profiles/r02_small_probe.jsonl) 9.8 vs 13.3 us at C1, 4.1 vs 7.2 at 1024, 31.0 vs 33.8 at 8192, but 23.9 vs 21.6 at
    // 6144 (42 bodies on 64 lanes) and 86 vs 60 at 12288
    if (h->opt_small != 1 && !(il == 1 || (il == 2 && ipc * 100 >= 85 * 64))) return 0;
    Rank& r0 = h->ranks[0];
    SmallStepArgs sa{};
    sa.pos[0] = r0.pos[0]; sa.pos[1] = r0.pos[1]; sa.vel = r0.vel; sa.n = h->n; sa.n_iblk = h->local_blocks; sa.ipc = ipc;
    sa.cur = h->cur; sa.nsteps = write_pos ? nsteps : 1; sa.write_pos = write_pos ? 1 : 0;
    sa.dt_v = (float)dt; sa.dt_x = (float)dt; sa.eps32 = (float)h->softening;
    cudaError_t se = step_small_launch(sa, ctas, il, h->softening != 1.0e-9, r0.st);
    if (se == cudaSuccess) {
        h->launches += 1; h->small_launches += 1;
        if (write_pos) h->cur ^= (nsteps & 1);
        return 1;
    }
    if (h->opt_small == 1) return fail(-(int)se, "small-system step kernel: %s", cudaGetErrorString(se));
    cudaGetLastError();                    // auto mode: the tiled paths of the same library take over
    return 0;
}

extern "C" {

int nbody_sync(nbody_handle h) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    OK(sync_all(h));
    OK(check_push_errors(h));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ranks[0].ev_t0, h->ranks[0].ev_t1) == cudaSuccess) h->last_step_ms = ms;
    else cudaGetLastError();
    return 0;
}

int nbody_step(nbody_handle h, double dt, int nsteps) {
    OK(nbody_step_async(h, dt, nsteps));
    return nbody_sync(h);
}

int nbody_body_force(nbody_handle h, double dt) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    {
        OK(set_dev(h->ranks[0]));
        const int took = try_small_steps(h, dt, 1, false);
        if (took < 0) return took;
        if (took) return sync_all(h);
    }
    for (auto& r : h->ranks) OK(enqueue_pass(h, r, Epilogue{dt, 0.0, false, true, nullptr}));
    return sync_all(h);
}

int nbody_integrate(nbody_handle h, double dt) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    for (auto& r : h->ranks) OK(enqueue_integrate(h, r, 0, 0.0, dt, true, true, nullptr));
    OK(exchange_positions(h));
    h->cur ^= 1;
    OK(sync_all(h));
    return check_push_errors(h);
}

int nbody_set_softening(nbody_handle h, double eps) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    if (!(eps > 0.0) || !(eps < 1.0e30)) return fail(-1, "softening must be a positive finite number (it is added to dist^2; the self-pair relies on it)");
    // the self-pair evaluates 0 * eps^(-3/2): that power must stay finite in the working type, or every acceleration is NaN
    if (h->precision == NBODY_F32 && !((float)eps >= 1.0e-25f)) return fail(-1, "softening %g is too small for FP32: eps^(-3/2) overflows (need >= 1e-25)", eps);
    if (h->precision == NBODY_F64 && !(eps >= 1.0e-200)) return fail(-1, "softening %g is too small for FP64: eps^(-3/2) overflows (need >= 1e-200)", eps);
    OK(sync_all(h));
    h->softening = eps;
    h->variant = default_variant(h);
    return replan(h);
}

int nbody_get_softening(nbody_handle h, double* eps) {
    OK(check_handle(h, false));
    if (!eps) return fail(-1, "output pointer is NULL");
    *eps = h->softening;
    return 0;
}

// kick-drift-kick leapfrog from the same two kernels: K(dt/2) D(dt) [K(dt) D(dt)]^(n-1) K(dt/2); the inner
// pairs are the reference's own step (v += dt*F(x); x += dt*v), so only the two half kicks are extra
int nbody_step_kdk(nbody_handle h, double dt, int nsteps) {
    if (nsteps < 0) return fail(-1, "nsteps must be >= 0");
    if (nsteps == 0) return 0;
    OK(nbody_body_force(h, 0.5 * dt));
    OK(nbody_integrate(h, dt));
    OK(nbody_step(h, dt, nsteps - 1));
    return nbody_body_force(h, 0.5 * dt);
}

int nbody_accel(nbody_handle h, float* a3) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    if (h->precision != NBODY_F32) return fail(-1, "handle is FP64: use nbody_accel_d");
    return accel_any(h, a3);
}
int nbody_accel_d(nbody_handle h, double* a3) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    if (h->precision != NBODY_F64) return fail(-1, "handle is FP32: use nbody_accel");
    return accel_any(h, a3);
}

int nbody_energy(nbody_handle h, double* ke, double* pe) {
    DeviceGuard guard_;
    OK(check_handle(h, true));
    if (!ke || !pe) return fail(-1, "output pointer is NULL");
    OK(sync_all(h));
    double k = 0, u = 0;
    for (auto& r : h->ranks) {
        OK(set_dev(r));
        CU(cudaMemsetAsync(r.energy, 0, 2 * sizeof(double), r.st));
        CU(energy_launch(h->precision, r.pos[h->cur], r.vel, h->n, r.rank * h->local_blocks, h->local_blocks, h->total_blocks, h->softening, r.energy, r.st));
        h->launches++;
    }
    for (auto& r : h->ranks) {
        OK(set_dev(r));
        if (!h->single_process && h->world > 1) {
            NC(g_nccl.AllReduce(r.energy, r.energy, 2, ncclDouble, ncclSum, r.comm, r.st));
        }
        double e[2];
        CU(cudaMemcpyAsync(e, r.energy, sizeof e, cudaMemcpyDeviceToHost, r.st));
        CU(cudaStreamSynchronize(r.st));
        k += e[0]; u += e[1];
    }
    *ke = k; *pe = u;
    return 0;
}

int nbody_ipc_export(nbody_handle h, void* blob) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    if (!blob) return fail(-1, "blob is NULL");
    if (h->single_process) return fail(-5, "nbody_ipc_export is for handles made with nbody_create_rank");
    Rank& r = h->ranks[0];
    OK(set_dev(r));
    IpcBlob b; memset(&b, 0, sizeof b);
    CU(cudaIpcGetMemHandle(&b.pos[0], r.pos[0]));
    CU(cudaIpcGetMemHandle(&b.pos[1], r.pos[1]));
    CU(cudaIpcGetMemHandle(&b.flags, r.flags));
    memset(blob, 0, NBODY_IPC_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    return 0;
}

int nbody_ipc_import(nbody_handle h, const void* all_blobs) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    if (!all_blobs) return fail(-1, "all_blobs is NULL");
    if (h->single_process) return fail(-5, "nbody_ipc_import is for handles made with nbody_create_rank");
    if (h->world > MAX_WORLD) return fail(-1, "world too large for the push exchange (at most %d ranks)", MAX_WORLD);
    Rank& r = h->ranks[0];
    OK(set_dev(r));
    std::vector<void*> pos0(h->world, nullptr), pos1(h->world, nullptr); std::vector<unsigned long long*> flags(h->world, nullptr);
    for (int q = 0; q < h->world; q++) {
        if (q == r.rank) continue;
        IpcBlob b; memcpy(&b, static_cast<const char*>(all_blobs) + (size_t)q * NBODY_IPC_BLOB_BYTES, sizeof b);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, b.pos[0], cudaIpcMemLazyEnablePeerAccess)); r.ipc_opened.push_back(p); pos0[q] = p;
        CU(cudaIpcOpenMemHandle(&p, b.pos[1], cudaIpcMemLazyEnablePeerAccess)); r.ipc_opened.push_back(p); pos1[q] = p;
        CU(cudaIpcOpenMemHandle(&p, b.flags, cudaIpcMemLazyEnablePeerAccess)); r.ipc_opened.push_back(p); flags[q] = static_cast<unsigned long long*>(p);
    }
    return install_peers(h, r, pos0, pos1, flags);
}

int nbody_set_option(nbody_handle h, const char* key, long long value) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    if (!key) return fail(-1, "key is NULL");
    OK(sync_all(h));
    const std::string k(key);
    if (k == "variant") {
        if (value < 0 || value >= variant_count(h->precision)) return fail(-1, "variant %lld out of range [0,%d)", value, variant_count(h->precision));
        if (h->precision == NBODY_F32 && h->softening != 1.0e-9 && !variant_of(h->precision, (int)value).eps_rt)
            return fail(-1, "variant %lld carries the reference softening 1e-9 as an immediate; with nbody_set_softening(%g) use a run-time-softening variant (15, 16, 17)", value, h->softening);
        h->variant = (int)value; return replan(h);
    }
    if (k == "splits") { if (value < 0 || value > 48) return fail(-1, "splits must be in [0,48]"); h->opt_splits = (int)value; return replan(h); }
    if (k == "overlap") { h->opt_overlap = value <= 0 ? 0 : (value >= 2 ? 2 : 1); return replan(h); }   // 2: own-slice-first even where the step is short
    if (k == "stream") {                 // -1: stream-K where it is the default (FP64); 1: also for FP32; 0: never
        h->opt_stream = value < 0 ? -1 : (value ? 1 : 0); h->variant = default_variant(h); return replan(h);
    }
    if (k == "grid") { if (value < 0 || value > 65535) return fail(-1, "grid must be in [0,65535]"); h->opt_grid = (int)value; return replan(h); }
    if (k == "stream_twin") { h->opt_twin = value ? 1 : 0; return 0; }
    if (k == "coop") { h->opt_coop = value ? 1 : 0; if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; } return 0; }
    if (k == "profile") { h->opt_profile = value ? 1 : 0; return 0; }
    if (k == "timing") { h->opt_timing = value ? 1 : 0; return 0; }
    if (k == "graph") { h->opt_graph = value < 0 ? -1 : (value ? 1 : 0); return 0; }
    if (k == "small") { h->opt_small = value < 0 ? -1 : (value ? 1 : 0); return 0; }
    if (k == "fused") { h->opt_fused = value < 0 ? -1 : (value ? 1 : 0); return replan(h); }
    if (k == "exchange") {
        if (value != 0 && value != 1) return fail(-1, "exchange must be 0 (NCCL all-gather) or 1 (peer-memory push)");
        if (value == 0 && h->virtual_ranks) return fail(-5, "virtual ranks share one device: there is no NCCL between them, the exchange is the peer-memory push");
        if (value == 1 && h->world > 1) {
            if (h->single_process) { if (!h->ranks[0].push_ready) OK(setup_push_single_process(h)); }
            else if (!h->ranks[0].push_ready) return fail(-5, "exchange=1 with one process per GPU needs nbody_ipc_export / nbody_ipc_import first");
        }
        // finish what is in flight under the old mode, then agree on the switch
        if (h->world > 1 && h->ranks[0].push_ready) { OK(epoch_barrier(h)); OK(check_push_errors(h)); }
        h->opt_exchange = (int)value; h->gather_pending = false; h->flag_pending = 0;
        return replan(h);                  // the fused one-launch pass needs the push exchange when sharded
    }
    if (k == "order") { h->opt_order = value < 0 ? -1 : (value ? 1 : 0); return replan(h); }
    if (k == "fuse") { h->opt_fuse = value < 0 ? -1 : (value ? 1 : 0); return replan(h); }
    return fail(-1, "unknown option '%s'", key);
}

int nbody_get_info(nbody_handle h, const char* key, long long* value) {
    OK(check_handle(h, false));
    if (!key || !value) return fail(-1, "NULL argument");
    const std::string k(key);
    if (k == "n") *value = h->n;
    else if (k == "precision") *value = h->precision;
    else if (k == "world") *value = h->world;
    else if (k == "i_begin") *value = std::min<long long>(h->n, (long long)h->ranks[0].rank * h->local_blocks * BLK);
    else if (k == "i_end") *value = std::min<long long>(h->n, (long long)(h->ranks[0].rank + 1) * h->local_blocks * BLK);
    else if (k == "rank") *value = h->ranks[0].rank;
    else if (k == "sms") *value = h->sms;
    else if (k == "variant") *value = h->variant;
    else if (k == "num_variants") *value = variant_count(h->precision);
    else if (k == "tile_bodies") *value = h->plan.tile_bodies;
    else if (k == "i_tiles") *value = h->plan.i_tiles;
    else if (k == "splits_local") *value = h->plan.splits_local;
    else if (k == "splits_remote") *value = h->plan.splits_remote;
    else if (k == "slots") *value = h->plan.slots;
    else if (k == "virtual_ranks") *value = h->virtual_ranks ? 1 : 0;
    else if (k == "stream") *value = is_stream(h) ? 1 : 0;
    else if (k == "fuse") *value = fuse_applies(h) ? 1 : 0;
    else if (k == "ring") *value = h->fuse_order == 1 ? h->fuse_ring : h->plan.i_tiles;
    else if (k == "order") *value = h->fuse_order;
    else if (k == "grid") *value = h->plan.stream_grid;
    else if (k == "phases") *value = h->plan.stream_phases;
    else if (k == "workspace_bytes") *value = (long long)h->ranks[0].part_bytes;
    else if (k == "total_blocks") *value = h->total_blocks;
    else if (k == "local_blocks") *value = h->local_blocks;
    else if (k == "launches") *value = h->launches;
    else if (k == "fused_launches") *value = h->fused_launches;
    else if (k == "small_launches") *value = h->small_launches;
    else if (k == "ctas_per_sm") *value = h->ctas_per_sm;
    else if (k == "exchange") *value = h->opt_exchange;
    else if (k == "packed") *value = variant_of(h->precision, h->variant).packed;
    else return fail(-1, "unknown info key '%s'", key);
    return 0;
}

int nbody_timing_reset(nbody_handle h) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    OK(sync_all(h));
    h->ranks[0].ev_used = 0;
    h->launches = 0;
    return 0;
}

int nbody_timing_get(nbody_handle h, double* force_ms, double* integrate_ms, long long* launches) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    OK(sync_all(h));
    double f = 0, g = 0;
    Rank& r = h->ranks[0];
    for (size_t i = 0; i < r.ev_used; i++) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, r.evs[i].a, r.evs[i].b));
        (r.evs[i].kind == 0 ? f : g) += ms;
    }
    if (force_ms) *force_ms = f;
    if (integrate_ms) *integrate_ms = g;
    if (launches) *launches = h->launches;
    return 0;
}

// per-CTA timeline of the last stream-K pass of rank 0 (option "profile" = 1): rows of 8 u64 per CTA
// {entry ns, last segment done ns, segments, reductions done by this CTA, ns spent in them, exit ns, SM id, 0}
int nbody_stream_profile(nbody_handle h, unsigned long long* rows, int max_ctas) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    if (!rows) return fail(-1, "rows is NULL");
    Rank& r = h->ranks[0];
    if (!r.prof || !is_stream(h)) return fail(-5, "no stream-K profile recorded (set option profile=1 and run a pass)");
    OK(sync_all(h));
    OK(set_dev(r));
    const int g = std::min(max_ctas, h->plan.stream_grid);
    CU(cudaMemcpy(rows, r.prof, (size_t)g * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return g;
}

int nbody_last_step_ms(nbody_handle h, double* ms) {
    OK(check_handle(h, false));
    if (!ms) return fail(-1, "ms is NULL");
    *ms = h->last_step_ms;
    return 0;
}

int nbody_probe_fp32_peak(nbody_handle h, double* ffma_lane_ops_per_s, double* sm_clock_mhz) {
    DeviceGuard guard_;
    OK(check_handle(h, false));
    Rank& r = h->ranks[0];
    OK(set_dev(r));
    OK(sync_all(h));
    const int grid = h->sms * 4, iters = 4000;
    float* out = nullptr; long long* cyc = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double best = 1e30, clk = 0;
    auto run = [&]() -> int {
        CU(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float)));
        CU(cudaMalloc(&cyc, sizeof(long long)));
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        for (int rep = 0; rep < 5; rep++) {
            // reps 0..3: 4 CTAs/SM for the throughput; rep 4: one CTA per SM so that CTA 0 spans the
            // whole launch and cycles / time is the SM clock under FP32 load
            const int g = rep < 4 ? grid : h->sms;
            CU(cudaEventRecord(e0, r.st));
            CU(ffma_probe_launch(out, cyc, iters, g, r.st));
            CU(cudaEventRecord(e1, r.st));
            CU(cudaStreamSynchronize(r.st));
            float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
            long long c; CU(cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost));
            if (rep > 0 && rep < 4 && ms < best) best = ms;
            if (rep == 4) clk = (double)c / (ms * 1e3);
        }
        return 0;
    };
    const int rc = run();                       // release the scratch objects on the error paths too
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (out) cudaFree(out);
    if (cyc) cudaFree(cyc);
    if (rc) return rc;
    const double lane_ops = (double)grid * 256 * (double)iters * 16 * 8 * 2;   // 2 FMA lanes per FFMA2
    if (ffma_lane_ops_per_s) *ffma_lane_ops_per_s = lane_ops / (best * 1e-3);
    if (sm_clock_mhz) *sm_clock_mhz = clk;
    return 0;
}

int nbody_mailbox_forces(const float* words_in, float* words_out, int n) {
    DeviceGuard guard_;
    if (!words_in || !words_out) return fail(-1, "NULL argument");
    nbody_handle h = nullptr;
    OK(nbody_create(n, NBODY_F32, 1, &h));
    Rank& r = h->ranks[0];
    float* dwords = nullptr;
    auto run = [&]() -> int {
        OK(set_dev(r));
        CU(cudaMalloc(&dwords, (size_t)n * 16));
        CU(cudaMemcpyAsync(dwords, words_in, (size_t)n * 16, cudaMemcpyHostToDevice, r.st));
        CU(mailbox_to_blocked_launch(dwords, n, h->total_blocks, (float*)r.pos[0], r.st));
        h->have_state = true; h->cur = 0;
        OK(enqueue_pass(h, r, Epilogue{0.0, 0.0, false, false, r.acc}));
        CU(blocked_to_mailbox_launch((const float*)r.acc, n, dwords, r.st));
        CU(cudaMemcpyAsync(words_out, dwords, (size_t)n * 16, cudaMemcpyDeviceToHost, r.st));
        CU(cudaStreamSynchronize(r.st));
        return 0;
    };
    const int rc = run();
    const std::string err = g_err;              // nbody_destroy must not clobber the message
    if (dwords) cudaFree(dwords);
    nbody_destroy(h);
    if (rc) g_err = err;
    return rc;
}

// The reference's mailbox handshake as one call (S/top_level.vhd:176-272): `ram` is the shared-RAM image the host
// prepared -- 128-bit words, word 0 = control {bit 0 BEGIN, bits 46:32 NUM_PTS} (:184-185), words 1..N = bodies
// {x,y,z,pad} (:206-208) -- and `results` the image behind the fabric's write port, where words 1..N receive
// {Fx,Fy,Fz,0} (S/compute_store.vhd:220-242; word 0 is never written).  On completion word 0 of `ram` is overwritten the
// way the `complete` state does it (:255-259, din from :146): BEGIN = 0, the elapsed count in bits 63:32 (here: device
// microseconds, at least 1; the RTL counts thousands of fabric clocks), everything else 0.  BEGIN = 0 on entry means the
// fabric is still `waiting`: nothing happens, return value 1.  NUM_PTS is a 15-bit field (ram_depth = 32768 words, word 0
// reserved: at most 32767 bodies, :45,55); images shorter than NUM_PTS + 1 words are rejected.
int nbody_mailbox_run(void* ram, void* results, int depth_words) {
    DeviceGuard guard_;
    if (!ram || !results) return fail(-1, "NULL argument");
    if (depth_words < 1 || depth_words > NBODY_MAILBOX_DEPTH) return fail(-1, "mailbox depth must be in [1, %d] words (S/top_level.vhd:45)", NBODY_MAILBOX_DEPTH);
    uint32_t* w0 = static_cast<uint32_t*>(ram);
    if (!(w0[0] & 1u)) return 1;                                  // BEGIN not raised: still waiting
    if (w0[1] >> 15) return fail(-1, "NUM_PTS field holds %u: the mailbox takes at most %d bodies (15-bit field, S/top_level.vhd:185)", w0[1], NBODY_MAILBOX_DEPTH - 1);
    const int n = (int)(w0[1] & 0x7FFFu);
    if (n + 1 > depth_words) return fail(-1, "NUM_PTS = %d does not fit a %d-word image", n, depth_words);
    double us = 1.0;
    if (n > 0) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { if (e0) cudaEventDestroy(e0); return fail(-2, "cudaEventCreate failed"); }
        cudaEventRecord(e0, 0);
        const int rc = nbody_mailbox_forces(static_cast<const float*>(ram) + 4, static_cast<float*>(results) + 4, n);
        cudaEventRecord(e1, 0);
        float ms = 0.f;
        if (rc == 0 && cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) us = std::max(1.0, (double)ms * 1e3);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (rc) return rc;
    }
    w0[0] = 0u; w0[1] = (uint32_t)std::min(us, 4294967295.0); w0[2] = 0u; w0[3] = 0u;
    return 0;
}

// ---- reference-shaped drop-in entry points -------------------------------------------------------
// One cached handle per precision, shared by all callers of the reference-shaped entry points: serialised by a mutex (the
// reference's loop is single-threaded; concurrent callers queue up) and released at process exit.
static nbody_handle g_dropin[2] = {nullptr, nullptr};
static std::mutex g_dropin_mutex;
static void dropin_atexit() {
    std::lock_guard<std::mutex> lock(g_dropin_mutex);
    for (nbody_handle& h : g_dropin) { if (h) { nbody_destroy(h); h = nullptr; } }
}

static nbody_handle dropin_handle(int n, int precision) {
    nbody_handle& h = g_dropin[precision];
    if (h && h->n != n) { nbody_destroy(h); h = nullptr; }
    if (!h) {
        if (nbody_create(n, precision, 1, &h) != 0) {
            fprintf(stderr, "libnbody_b200: %s\n", nbody_last_error());
            abort();
        }
        // registered after the CUDA runtime is up, so that it runs BEFORE the runtime's own exit handler
        static const bool registered = (atexit(dropin_atexit), true);
        (void)registered;
    }
    return h;
}
static void must(int rc, const char* what) {
    if (rc != 0) { fprintf(stderr, "libnbody_b200: %s failed: %s\n", what, nbody_last_error()); abort(); }
}

static inline uint64_t splitmix64_next(uint64_t* state) {
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void randomizeBodiesSeeded(float* data, long long n, uint64_t seed) {
    uint64_t st = seed;
    for (long long k = 0; k < n; k++) {
        const uint32_t r = (uint32_t)(splitmix64_next(&st) >> 40);
        data[k] = (float)r * (1.0f / 8388608.0f) - 1.0f;
    }
}
void randomizeBodies(float* data, int n) {
    uint64_t seed = 42;
    if (const char* s = getenv("NBODY_SEED")) seed = strtoull(s, nullptr, 10);
    randomizeBodiesSeeded(data, n, seed);
}

void bodyForce(Body* p, float dt, int n) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> lock(g_dropin_mutex);
    nbody_handle h = dropin_handle(n, NBODY_F32);
    must(nbody_upload(h, p), "bodyForce/upload");
    must(nbody_body_force(h, (double)dt), "bodyForce");
    must(nbody_download(h, p), "bodyForce/download");
}
void integrate(Body* p, float dt, int n) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> lock(g_dropin_mutex);
    nbody_handle h = dropin_handle(n, NBODY_F32);
    must(nbody_upload(h, p), "integrate/upload");
    must(nbody_integrate(h, (double)dt), "integrate");
    must(nbody_download(h, p), "integrate/download");
}
void bodyForceD(BodyD* p, double dt, int n) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> lock(g_dropin_mutex);
    nbody_handle h = dropin_handle(n, NBODY_F64);
    must(nbody_upload_d(h, p), "bodyForceD/upload");
    must(nbody_body_force(h, dt), "bodyForceD");
    must(nbody_download_d(h, p), "bodyForceD/download");
}
void integrateD(BodyD* p, double dt, int n) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> lock(g_dropin_mutex);
    nbody_handle h = dropin_handle(n, NBODY_F64);
    must(nbody_upload_d(h, p), "integrateD/upload");
    must(nbody_integrate(h, dt), "integrateD");
    must(nbody_download_d(h, p), "integrateD/download");
}

}  // extern "C"
