/*
 * nbody.h -- C ABI of libnbody_b200.so: the B200-native drop-in for mini-nbody's one hot path,
 *            the O(N^2) all-pairs softened-gravity force evaluation (bodyForce) and the explicit
 *            velocity/position integrate step it feeds.
 *
 * Plain C: pointers, ints, floats.  No CUDA, torch or C++ types cross this boundary.
 *
 * What each entry point replaces in the reference (citations relative to
 * /root/reference/vec_add.srcs/sources_1/new/):
 *
 *   - Body / bodyForce / integrate / randomizeBodies: the host-side C entry points and the
 *     Body{x,y,z,vx,vy,vz} layout named by BASELINE.json north_star.  The host C program is
 *     ABSENT from the reference mount, so there is no file:line to cite; the only in-tree
 *     statement of the arithmetic bodyForce performs is the VHDL force pipeline:
 *        dxy.vhd:94-122, dzsoft.vhd:177-202, dxyz_soft.vhd:149-150   dist^2 = |r_j - r_i|^2 + 1e-9
 *        fxyz.vhd:101-102 (rsqrt), cube.vhd:66-70 (inv^3), fxyz.vhd:120-127 (F += d * inv^3)
 *        top_level.vhd:187-254  sweep: every i against all N j, self-pair included
 *        compute_store.vhd:203-242  result record {Fx,Fy,Fz,0} per body, masked for padding
 *   - nbody_create / nbody_upload / nbody_step / nbody_accel / nbody_download: the mailbox
 *     protocol of top_level.vhd:176-272 (host writes bodies to words 1..N, sets BEGIN, polls,
 *     reads forces) restated as a resident-state API, so bodies stay in HBM between steps.
 *   - nbody_mailbox_forces: the literal 16-byte word image of that mailbox
 *     (top_level.vhd:206-208,238-240 body word {x,y,z,pad}; compute_store.vhd:242 result word).
 *
 * Semantics fixed by the reference / north_star: SOFTENING = 1e-9 (added to dist^2), unit masses,
 * no G, d = r_j - r_i, self-pair included (contributes 0), bodyForce applies v += dt*F, integrate
 * applies x += dt*v.
 *
 * Error behaviour: int-returning functions return 0 on success and a negative code on failure
 * (-1 invalid argument, -2 no CUDA device / CUDA failure, -3 NCCL failure, -4 out of memory,
 * -5 bad state); nbody_last_error() gives the message.  The reference-shaped void functions
 * print the message to stderr and abort() on failure -- there is no CPU fallback.
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z, vx, vy, vz; } Body;    /* 24 bytes, AoS, no padding */
typedef struct { double x, y, z, vx, vy, vz; } BodyD;  /* 48 bytes, FP64 accuracy path */

#define NBODY_SOFTENING 1.0e-9f
#define NBODY_F32 0
#define NBODY_F64 1
#define NBODY_NCCL_ID_BYTES 128

/* ---- reference-shaped entry points (host buffers in, host buffers out) ------------------- */

/* Fill data[0..n) with uniform values in [-1, 1).  n counts floats (6 per body).  Deterministic:
 * seed = $NBODY_SEED or 42; same stream as oracle_randomize (splitmix64, 24-bit mantissas). */
void randomizeBodies(float *data, int n);
void randomizeBodiesSeeded(float *data, long long n, uint64_t seed);

/* For every body i: F_i = sum_j (r_j - r_i) * (|r_j - r_i|^2 + 1e-9)^(-3/2); v_i += dt * F_i.
 * Positions are not modified.  Runs on GPU 0 (or $NBODY_DEVICE). */
void bodyForce(Body *p, float dt, int n);
/* x_i += dt * v_i for every body. */
void integrate(Body *p, float dt, int n);
void bodyForceD(BodyD *p, double dt, int n);
void integrateD(BodyD *p, double dt, int n);

/* ---- resident-state API ------------------------------------------------------------------- */

typedef struct nbody_ctx *nbody_handle;

/* One process driving `ngpus` devices (0..ngpus-1) of this box; i-bodies are sharded across
 * them, every GPU holds all positions, one all-gather of the new positions per step. */
int nbody_create(int n, int precision, int ngpus, nbody_handle *out);

/* One process per GPU (torchrun style): this process is `rank` of `world` and drives CUDA device
 * `device`.  nccl_id is the NBODY_NCCL_ID_BYTES blob from nbody_nccl_unique_id() on rank 0,
 * broadcast by the caller (any transport).  world == 1 accepts nccl_id == NULL. */
int nbody_nccl_unique_id(void *out_id);
int nbody_create_rank(int n, int precision, int rank, int world, int device, const void *nccl_id,
                      nbody_handle *out);

int nbody_destroy(nbody_handle h);

/* Peer-memory ("push") exchange for one-process-per-GPU handles: every rank exports a blob of CUDA IPC
 * handles (its two position buffers and its flag array), the caller all-gathers the blobs (rank-major,
 * NBODY_IPC_BLOB_BYTES each) and every rank imports them; then nbody_set_option(h, "exchange", 1).
 * Single-process multi-GPU handles need neither call (peer access is enabled directly). */
#define NBODY_IPC_BLOB_BYTES 256
int nbody_ipc_export(nbody_handle h, void *blob);
int nbody_ipc_import(nbody_handle h, const void *all_blobs);

/* Host AoS -> device (tile-blocked SoA).  Every rank passes the full n-body array; when the bodies are sharded
 * a rank reads only its own slice of it over PCIe and the positions are all-gathered on the device. */
int nbody_upload(nbody_handle h, const Body *p);
int nbody_upload_d(nbody_handle h, const BodyD *p);
/* Device -> host AoS, full array on every rank (velocities are all-gathered first if sharded). */
int nbody_download(nbody_handle h, Body *p);
int nbody_download_d(nbody_handle h, BodyD *p);
/* This rank's bodies only, p[0 .. i_end - i_begin) = bodies [i_begin, i_end) (nbody_get_info "i_begin"/"i_end",
 * or nbody_plan): no collective, 1/world of the bytes.  For handles that drive one GPU (nbody_create_rank). */
int nbody_download_local(nbody_handle h, Body *p);
int nbody_download_local_d(nbody_handle h, BodyD *p);

/* nsteps x { bodyForce ; integrate } on the resident state.  nbody_step returns after the work
 * has finished; nbody_step_async only enqueues (pair with nbody_sync). */
int nbody_step(nbody_handle h, double dt, int nsteps);
int nbody_step_async(nbody_handle h, double dt, int nsteps);
int nbody_sync(nbody_handle h);

/* The two halves separately (used by the reference-shaped wrappers and by parity tests). */
int nbody_body_force(nbody_handle h, double dt);   /* v += dt * F(x) */
int nbody_integrate(nbody_handle h, double dt);    /* x += dt * v    */

/* SURVEY.md section 8(f) n4 -- steps either side of the reference's fixed choices (explicit Euler, softening
 * 1e-9 hard-wired at S/dzsoft.vhd:177).  nbody_set_softening changes the constant added to dist^2 (forces and
 * nbody_energy); FP32 handles then run the kernel instantiations that read it from their arguments instead of
 * carrying 1e-9 as an immediate.  nbody_step_kdk is a kick-drift-kick leapfrog built from nbody_body_force /
 * nbody_integrate / nbody_step: K(dt/2) D(dt) [K(dt) D(dt)]^(n-1) K(dt/2). */
int nbody_set_softening(nbody_handle h, double eps);
int nbody_get_softening(nbody_handle h, double *eps);
int nbody_step_kdk(nbody_handle h, double dt, int nsteps);

/* Accelerations at the current positions, no state change.  a3 has 3*n floats (doubles for the
 * _d form) laid out {ax,ay,az} per body; full array on every rank. */
int nbody_accel(nbody_handle h, float *a3);
int nbody_accel_d(nbody_handle h, double *a3);

/* Total energy in FP64 (unit masses): ke = 1/2 sum |v|^2, pe = -sum_{i<j} (r_ij^2 + eps)^(-1/2),
 * eps = 1e-9 unless nbody_set_softening changed it. */
int nbody_energy(nbody_handle h, double *ke, double *pe);

/* FPGA mailbox image (n <= 32767 in the reference; no limit here): words_in[n] are 16-byte body
 * words {x,y,z,pad}; words_out[n] receive {Fx,Fy,Fz,0}.  Stateless, runs on GPU 0. */
int nbody_mailbox_forces(const float *words_in, float *words_out, int n);

/* The reference's mailbox handshake as one call (S/top_level.vhd:176-272), images exactly as the fabric sees them:
 *   ram      128-bit words; word 0 = control: bit 0 BEGIN, bits 46:32 NUM_PTS (:184-185); words 1..NUM_PTS = bodies
 *            {x,y,z,pad} (:206-208,238-240)
 *   results  the image behind the fabric's write port: words 1..NUM_PTS receive {Fx,Fy,Fz,0}
 *            (S/compute_store.vhd:220-242); word 0 is never written
 *   depth_words  words in each image, at most NBODY_MAILBOX_DEPTH (ram_depth, :45)
 * Returns 1 and touches nothing while BEGIN = 0 (state `waiting`).  Otherwise computes, then overwrites word 0 of `ram`
 * as the `complete` state does (:255-259): BEGIN = 0, elapsed count in bits 63:32 (device microseconds >= 1 here; the RTL
 * counts thousands of fabric clocks, :140-146), all other bits 0 -- and returns 0.  NUM_PTS > 32767 cannot be expressed
 * in the 15-bit field: a word 0 with higher bits set in 63:47 is rejected (-1), as is NUM_PTS + 1 > depth_words. */
#define NBODY_MAILBOX_DEPTH 32768
int nbody_mailbox_run(void *ram, void *results, int depth_words);

/* Tuning / introspection.  Keys: "variant" (force kernel instantiation), "splits" (j-splits per
 * launch, 0 = planner decides), "overlap" (sharded: 1 = own j-slice first while the exchange is in flight, 0 = one pass,
 * 2 = own-slice-first even where the step is short), "exchange" (0 = NCCL all-gather on a side stream; 1 = the force
 * kernel's integrate epilogue stores every finished tile straight into every peer's next-step buffer over NVLink and
 * publishes a flag, the peers' flags are acquired inside the next force kernel: no collective, one launch per step),
 * "fuse" (FP32 split-grid kernels: in-kernel fixed-order reduction of the j-splits + integrate; -1 auto, 0 off = slot array +
 * integrate kernel, 1 on), "order" (CTA order of the fused pass), "stream" (stream-K kernels: -1 where default, 1 also FP32,
 * 0 never), "grid" / "stream_twin" / "coop" / "profile" (stream-K), "small" (small-system multi-step kernel: -1 auto, 0, 1),
 * "fused" (tiled multi-step kernel for 1536..5120 bodies when "small" is off),
 * "graph" (CUDA-graph replay of step pairs in multi-step calls on one GPU: -1 auto = below 65 536 bodies, 0 off, 1 on),
 * "timing" (1 = record per-kernel CUDA events for nbody_timing_get; default 0 -- the event records cost ~10 us
 * per step, which matters below N ~ 30 000).  INTEGRATION.md section 2 has the table with the defaults. */
int nbody_set_option(nbody_handle h, const char *key, long long value);
int nbody_get_info(nbody_handle h, const char *key, long long *value);

/* Device-side timing of the steps since the last reset (CUDA events on the launching stream):
 * total force-kernel ms, total integrate-kernel ms, number of kernel launches of this library. */
int nbody_timing_reset(nbody_handle h);
int nbody_timing_get(nbody_handle h, double *force_ms, double *integrate_ms, long long *launches);
/* Wall-to-wall device time of one nbody_step call, measured by events on stream 0 of rank 0. */
int nbody_last_step_ms(nbody_handle h, double *ms);

/* FFMA throughput probe on the handle's first device: fills lane-FMA/s (peak FP32 = 2x that). */
int nbody_probe_fp32_peak(nbody_handle h, double *ffma_lane_ops_per_s, double *sm_clock_mhz);

/* Stream-K force pass, option "profile" = 1: per-CTA timeline of the last pass of the handle's first rank, rows of
 * 8 uint64 {entry ns, last segment done ns, segments, reductions done by this CTA, ns spent in them, exit ns, SM id, 0}.
 * Returns the number of CTAs written (<= max_ctas), negative on error.  Development aid. */
int nbody_stream_profile(nbody_handle h, unsigned long long *rows, int max_ctas);

/* Host-only planning (no GPU needed): how n bodies are sharded and how the force pass of one rank
 * is cut into CTAs.  Used by the CPU tests of the N>1 path. */
typedef struct {
    int n;              /* bodies */
    int world, rank;
    int blk;            /* bodies per layout block (128) */
    int total_blocks;   /* blocks in the replicated position array, incl. padding */
    int local_blocks;   /* i-blocks owned by each rank */
    int i_begin, i_end; /* owned body range [i_begin, i_end) clipped to n */
    int tile_bodies;    /* i-bodies per CTA */
    int i_tiles;        /* CTAs along i */
    int splits_local;   /* j-splits of the pass over the rank's own j-slice */
    int splits_remote;  /* j-splits of the pass over the other ranks' slices (0 if world == 1) */
    int slots;          /* partial-acceleration slots summed by integrate */
    int stream_grid;    /* stream-K variants: persistent CTAs of the force pass (0 for the split-grid variants) */
    int stream_phases;  /* stream-K variants: 1, or 2 = own j-slice first, other ranks' slices second */
} nbody_plan_t;
int nbody_plan(int n, int precision, int rank, int world, int sms, int variant, nbody_plan_t *out);
/* Stream-K plans (stream_grid > 0): the segments persistent CTA `cta` works on, in order, as rows of 6 ints
 * {phase, tile, ja, jb, workspace slot, segments of that tile}; ja/jb count granules of 16 j-bodies inside the
 * phase's j-range.  Returns the number of rows, negative on error.  Host-only (tests of the decomposition). */
int nbody_stream_segments(const nbody_plan_t *plan, int cta, int *rows, int cap);
/* Fused split-grid pass: (tile, split) of CTA `bid` of the 1-D grid of i_tiles * nsplit CTAs; order 0 = split-major over all
 * tiles, 1 = split-major inside groups of ring/2 tiles (tile t keeps its partial sums at ring position t % ring).  The map the
 * kernel itself uses; host-only (tests). */
int nbody_fused_cta(int i_tiles, int nsplit, int ring, int order, int bid, int *tile, int *split);

const char *nbody_last_error(void);
const char *nbody_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
