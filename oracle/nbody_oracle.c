/*
 * nbody_oracle.c -- CPU restatement of the reference's all-pairs softened-gravity force path
 *                   plus the explicit integrate step it feeds.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call it.  The product
 * (mini-nbody_b200/csrc, libnbody_b200.so) never links or calls anything in oracle/.
 *
 * PARITY UNPINNED.  The reference mount (/root/reference) holds only a VHDL pipeline; it cannot
 * be compiled or simulated here (no VHDL simulator, vendor floating-point IP absent) and its
 * testbenches check only "not X when valid" -- they carry stimuli but no expected outputs.  The
 * stimuli are hand-derivable and are pinned in tests/test_oracle_kat.py; nothing stronger exists.
 *
 * Citations are relative to /root/reference/vec_add.srcs/ (S = sources_1/new, T = sim_1/new).
 *
 * Per-pair dataflow restated here (binary32, round-to-nearest-even, no contraction beyond the
 * reference's own fused multiply-adds):
 *     dx = x_j - x_i ; dy = y_j - y_i                         S/dxy.vhd:94-98   (target - this)
 *     sxy = dx*dx + dy*dy   (two rounded products, rounded add) S/dxy.vhd:113-122
 *     dz = z_j - z_i ; sz = fma(dz, dz, SOFTENING)             S/dzsoft.vhd:177,186-202
 *     dist2 = sxy + sz                                         S/dxyz_soft.vhd:149-150
 *     inv = rsqrt(dist2)                                       S/fxyz.vhd:101-102
 *     inv3 = inv * (inv * inv)                                 S/cube.vhd:66-70
 *     F{x,y,z} = fma(d{x,y,z}, inv3, F{x,y,z})                 S/fxyz.vhd:120-127
 * j runs over ALL bodies including j == i (S/top_level.vhd:233-249); unit masses, no G.
 *
 * Body{x,y,z,vx,vy,vz}, dt semantics (v += dt*F inside bodyForce, then x += dt*v) and the seeded
 * uniform [-1,1) initial state come from BASELINE.json north_star/configs (host C code is absent
 * from the mount, so there is no file:line to cite for them).
 *
 * Build flavours (oracle/Makefile):
 *   parity: -O2 -ffp-contract=off -fno-fast-math [-fopenmp]  (bit-stable arithmetic; OpenMP only
 *           spreads the i loop, each i keeps its sequential-j order)
 *   speed : -O3 -ffast-math -fopenmp -march=x86-64-v3         (the timed CPU baseline, "port")
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z, vx, vy, vz; } Body;
typedef struct { double x, y, z, vx, vy, vz; } BodyD;

/* SOFTENING = 1.0e-9 rounded to binary32 = 0x3089705F (S/dzsoft.vhd:177) */
#define ORACLE_SOFTENING_F32 1.0e-9f
#define ORACLE_SOFTENING_F64 1.0e-9

/* Softening used by the force loops and the energy: the reference constant unless a test of the run-time
 * softening option (include/nbody.h: nbody_set_softening) changes it.  The
 * entity-level functions (oracle_dzsoft, ...) always use the reference constant. */
static float g_soft32 = ORACLE_SOFTENING_F32;
static double g_soft64 = ORACLE_SOFTENING_F64;
void oracle_set_softening(double eps) { g_soft64 = eps; g_soft32 = (float)eps; }
double oracle_get_softening(void) { return g_soft64; }

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline legs of bench.py ask for the box's cores back */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

uint32_t oracle_softening_bits(void) {
    float s = ORACLE_SOFTENING_F32; uint32_t u; memcpy(&u, &s, 4); return u;
}

/* ---------------------------------------------------------------------------------------------
 * Seeded initial state: all 6n floats i.i.d. uniform in [-1, 1).  splitmix64 stream, top 24 bits
 * of each output -> k * 2^-23 - 1 (exact in binary32), so every build sees identical bits.
 * (BASELINE.json configs: "FP32 seeded random init")
 * ------------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64_next(uint64_t *state) {
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_randomize(float *data, long long count, uint64_t seed) {
    uint64_t st = seed;
    for (long long k = 0; k < count; k++) {
        uint32_t r = (uint32_t)(splitmix64_next(&st) >> 40);          /* 24 bits */
        data[k] = (float)r * (1.0f / 8388608.0f) - 1.0f;                /* r * 2^-23 - 1 */
    }
}

/* ---------------------------------------------------------------------------------------------
 * Stage functions, one per reference entity (used by the KAT tests and by the force loops).
 * ------------------------------------------------------------------------------------------- */
/* S/dxy.vhd:94-122 */
float oracle_dxy(float x_this, float x_target, float y_this, float y_target, float *dx, float *dy) {
    float ddx = x_target - x_this;
    float ddy = y_target - y_this;
    float dx2 = ddx * ddx;
    float dy2 = ddy * ddy;
    if (dx) *dx = ddx;
    if (dy) *dy = ddy;
    return dx2 + dy2;
}
/* S/dzsoft.vhd:186-202 */
float oracle_dzsoft(float z_this, float z_target, float *dz) {
    float ddz = z_target - z_this;
    if (dz) *dz = ddz;
    return fmaf(ddz, ddz, ORACLE_SOFTENING_F32);
}
/* S/dxyz_soft.vhd:87-150 */
float oracle_dxyz_soft(const float this_[3], const float target[3], float d[3]) {
    float sxy = oracle_dxy(this_[0], target[0], this_[1], target[1], &d[0], &d[1]);
    float sz = oracle_dzsoft(this_[2], target[2], &d[2]);
    return sxy + sz;
}
/* rsqrt IP, S/fxyz.vhd:101-102; special values per the comments in T/tb_sqrt.vhd:528-541 */
float oracle_rsqrt(float s) { return 1.0f / sqrtf(s); }
/* S/cube.vhd:66-70 */
float oracle_cube(float inv) { float inv2 = inv * inv; return inv * inv2; }

static inline void pair_f32(float xi, float yi, float zi, float xj, float yj, float zj,
                            float *fx, float *fy, float *fz) {
    float dx = xj - xi, dy = yj - yi, dz = zj - zi;
    float dx2 = dx * dx, dy2 = dy * dy;
    float sxy = dx2 + dy2;
    float sz = fmaf(dz, dz, g_soft32);
    float s = sxy + sz;
    float inv = 1.0f / sqrtf(s);
    float inv2 = inv * inv;
    float inv3 = inv * inv2;
    *fx = fmaf(dx, inv3, *fx);
    *fy = fmaf(dy, inv3, *fy);
    *fz = fmaf(dz, inv3, *fz);
}

static inline void pair_f64(double xi, double yi, double zi, double xj, double yj, double zj,
                            double *fx, double *fy, double *fz) {
    double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    double dx2 = dx * dx, dy2 = dy * dy;
    double sxy = dx2 + dy2;
    double sz = fma(dz, dz, g_soft64);
    double s = sxy + sz;
    double inv = 1.0 / sqrt(s);
    double inv2 = inv * inv;
    double inv3 = inv * inv2;
    *fx = fma(dx, inv3, *fx);
    *fy = fma(dy, inv3, *fy);
    *fz = fma(dz, inv3, *fz);
}

/* ---------------------------------------------------------------------------------------------
 * Accelerations for i in [i0, i1) over all j in [0, n): a3[(i-i0)*3 + {0,1,2}].
 * Sequential-j accumulation (the order a plain C host loop would use).
 * ------------------------------------------------------------------------------------------- */
void oracle_accel_f32(const Body *p, int n, int i0, int i1, float *a3) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = i0; i < i1; i++) {
        float fx = 0.f, fy = 0.f, fz = 0.f;
        const float xi = p[i].x, yi = p[i].y, zi = p[i].z;
        for (int j = 0; j < n; j++) pair_f32(xi, yi, zi, p[j].x, p[j].y, p[j].z, &fx, &fy, &fz);
        a3[(size_t)(i - i0) * 3 + 0] = fx; a3[(size_t)(i - i0) * 3 + 1] = fy; a3[(size_t)(i - i0) * 3 + 2] = fz;
    }
}

/* The reference hardware's own summation order: the FMA output is fed back as its addend through a
 * 16-deep pipeline, giving 16 interleaved partial sums (j mod 16) per (body, dim)
 * (S/fxyz.vhd:120-145, fma_latency = 16 at S/top_level.vhd:40), which a balanced binary tree then
 * adds (S/final_adder.vhd:42-68,88-104).  Shows how much summation order alone moves the result. */
void oracle_accel_f32_fpga_order(const Body *p, int n, int i0, int i1, float *a3) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = i0; i < i1; i++) {
        float px[16], py[16], pz[16];
        for (int k = 0; k < 16; k++) px[k] = py[k] = pz[k] = 0.f;     /* flush-to-zero at burst start, S/fxyz.vhd:129-145 */
        const float xi = p[i].x, yi = p[i].y, zi = p[i].z;
        for (int j = 0; j < n; j++) pair_f32(xi, yi, zi, p[j].x, p[j].y, p[j].z, &px[j & 15], &py[j & 15], &pz[j & 15]);
        for (int w = 8; w >= 1; w >>= 1)
            for (int k = 0; k < w; k++) { px[k] = px[2 * k] + px[2 * k + 1]; py[k] = py[2 * k] + py[2 * k + 1]; pz[k] = pz[2 * k] + pz[2 * k + 1]; }
        a3[(size_t)(i - i0) * 3 + 0] = px[0]; a3[(size_t)(i - i0) * 3 + 1] = py[0]; a3[(size_t)(i - i0) * 3 + 2] = pz[0];
    }
}

/* Kahan-compensated sequential-j sum of the same FP32 pair terms (term = round(d * inv3), then a compensated add):
 * what the order-sensitivity report (tools/order_report.py) puts beside the sequential, FPGA
 * and GPU orders -- how much of the FP32 error is summation order at all.  Meaningful in the parity build only
 * (-ffp-contract=off -fno-fast-math keeps the compensation alive). */
void oracle_accel_f32_kahan(const Body *p, int n, int i0, int i1, float *a3) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = i0; i < i1; i++) {
        volatile float s[3] = {0.f, 0.f, 0.f}, c[3] = {0.f, 0.f, 0.f};
        const float xi = p[i].x, yi = p[i].y, zi = p[i].z;
        for (int j = 0; j < n; j++) {
            float t[3] = {0.f, 0.f, 0.f};
            pair_f32(xi, yi, zi, p[j].x, p[j].y, p[j].z, &t[0], &t[1], &t[2]);      /* fma(d, inv3, 0) = round(d * inv3) */
            for (int d = 0; d < 3; d++) {
                const float y = t[d] - c[d];
                const float u = s[d] + y;
                c[d] = (u - s[d]) - y;
                s[d] = u;
            }
        }
        a3[(size_t)(i - i0) * 3 + 0] = s[0]; a3[(size_t)(i - i0) * 3 + 1] = s[1]; a3[(size_t)(i - i0) * 3 + 2] = s[2];
    }
}

/* FP32 inputs, FP64 arithmetic: the ground truth the FP32 paths are measured against. */
void oracle_accel_f64_from_f32(const Body *p, int n, int i0, int i1, double *a3) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = i0; i < i1; i++) {
        double fx = 0., fy = 0., fz = 0.;
        const double xi = p[i].x, yi = p[i].y, zi = p[i].z;
        for (int j = 0; j < n; j++) pair_f64(xi, yi, zi, (double)p[j].x, (double)p[j].y, (double)p[j].z, &fx, &fy, &fz);
        a3[(size_t)(i - i0) * 3 + 0] = fx; a3[(size_t)(i - i0) * 3 + 1] = fy; a3[(size_t)(i - i0) * 3 + 2] = fz;
    }
}

void oracle_accel_f64(const BodyD *p, int n, int i0, int i1, double *a3) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = i0; i < i1; i++) {
        double fx = 0., fy = 0., fz = 0.;
        const double xi = p[i].x, yi = p[i].y, zi = p[i].z;
        for (int j = 0; j < n; j++) pair_f64(xi, yi, zi, p[j].x, p[j].y, p[j].z, &fx, &fy, &fz);
        a3[(size_t)(i - i0) * 3 + 0] = fx; a3[(size_t)(i - i0) * 3 + 1] = fy; a3[(size_t)(i - i0) * 3 + 2] = fz;
    }
}

/* long-double accumulation of FP64 pair terms: checks the FP64 oracle's own summation error */
void oracle_accel_f80(const BodyD *p, int n, int i0, int i1, double *a3) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = i0; i < i1; i++) {
        long double fx = 0., fy = 0., fz = 0.;
        const long double xi = p[i].x, yi = p[i].y, zi = p[i].z;
        for (int j = 0; j < n; j++) {
            long double dx = p[j].x - xi, dy = p[j].y - yi, dz = p[j].z - zi;
            long double s = dx * dx + dy * dy + dz * dz + (long double)g_soft64;
            long double inv = 1.0L / sqrtl(s);
            long double inv3 = inv * inv * inv;
            fx += dx * inv3; fy += dy * inv3; fz += dz * inv3;
        }
        a3[(size_t)(i - i0) * 3 + 0] = (double)fx; a3[(size_t)(i - i0) * 3 + 1] = (double)fy; a3[(size_t)(i - i0) * 3 + 2] = (double)fz;
    }
}

/* ---------------------------------------------------------------------------------------------
 * The two host-visible steps (north_star: "bodyForce ... and the explicit position/velocity
 * integrate step it feeds"): bodyForce applies v += dt*F for every body from the positions as
 * they stand; integrate applies x += dt*v with the updated velocities.
 * ------------------------------------------------------------------------------------------- */
void oracle_body_force_f32(Body *p, float dt, int n) {
    float *a = (float *)malloc(sizeof(float) * 3 * (size_t)n);
    oracle_accel_f32(p, n, 0, n, a);
    for (int i = 0; i < n; i++) {
        p[i].vx = fmaf(dt, a[3 * (size_t)i + 0], p[i].vx);
        p[i].vy = fmaf(dt, a[3 * (size_t)i + 1], p[i].vy);
        p[i].vz = fmaf(dt, a[3 * (size_t)i + 2], p[i].vz);
    }
    free(a);
}
void oracle_integrate_f32(Body *p, float dt, int n) {
    for (int i = 0; i < n; i++) {
        p[i].x = fmaf(p[i].vx, dt, p[i].x);
        p[i].y = fmaf(p[i].vy, dt, p[i].y);
        p[i].z = fmaf(p[i].vz, dt, p[i].z);
    }
}
void oracle_body_force_f64(BodyD *p, double dt, int n) {
    double *a = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    oracle_accel_f64(p, n, 0, n, a);
    for (int i = 0; i < n; i++) {
        p[i].vx = fma(dt, a[3 * (size_t)i + 0], p[i].vx);
        p[i].vy = fma(dt, a[3 * (size_t)i + 1], p[i].vy);
        p[i].vz = fma(dt, a[3 * (size_t)i + 2], p[i].vz);
    }
    free(a);
}
void oracle_integrate_f64(BodyD *p, double dt, int n) {
    for (int i = 0; i < n; i++) {
        p[i].x = fma(p[i].vx, dt, p[i].x);
        p[i].y = fma(p[i].vy, dt, p[i].y);
        p[i].z = fma(p[i].vz, dt, p[i].z);
    }
}
void oracle_run_f32(Body *p, float dt, int n, int steps) {
    for (int s = 0; s < steps; s++) { oracle_body_force_f32(p, dt, n); oracle_integrate_f32(p, dt, n); }
}
void oracle_run_f64(BodyD *p, double dt, int n, int steps) {
    for (int s = 0; s < steps; s++) { oracle_body_force_f64(p, dt, n); oracle_integrate_f64(p, dt, n); }
}

/* Total energy in FP64 (diagnostic; unit masses): KE = 1/2 sum |v|^2, PE = - sum_{i<j} (r^2+eps)^-1/2 */
void oracle_energy_f64(const BodyD *p, int n, double *ke, double *pe) {
    double k = 0., u = 0.;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : k, u)
    for (int i = 0; i < n; i++) {
        k += 0.5 * (p[i].vx * p[i].vx + p[i].vy * p[i].vy + p[i].vz * p[i].vz);
        double ui = 0.;
        for (int j = i + 1; j < n; j++) {
            double dx = p[j].x - p[i].x, dy = p[j].y - p[i].y, dz = p[j].z - p[i].z;
            ui -= 1.0 / sqrt(dx * dx + dy * dy + dz * dz + g_soft64);
        }
        u += ui;
    }
    *ke = k; *pe = u;
}
void oracle_energy_f32in(const Body *p, int n, double *ke, double *pe) {
    BodyD *d = (BodyD *)malloc(sizeof(BodyD) * (size_t)n);
    for (int i = 0; i < n; i++) { d[i].x = p[i].x; d[i].y = p[i].y; d[i].z = p[i].z; d[i].vx = p[i].vx; d[i].vy = p[i].vy; d[i].vz = p[i].vz; }
    oracle_energy_f64(d, n, ke, pe);
    free(d);
}

/* ---------------------------------------------------------------------------------------------
 * Timed CPU baseline ("port"): same formula, host-friendly layout (SoA gather once per call) so
 * the compiler can vectorise the j loop; meaningful only in the speed build.  Computes the
 * accelerations of i in [i0,i1) over all n j-bodies and applies the velocity update to those
 * bodies, i.e. a bounded sample of one bodyForce call.  Returns interactions evaluated.
 * ------------------------------------------------------------------------------------------- */
double oracle_body_force_f32_fast(Body *p, float dt, int n, int i0, int i1) {
    float *x = (float *)aligned_alloc(64, sizeof(float) * (((size_t)n + 15) & ~(size_t)15));
    float *y = (float *)aligned_alloc(64, sizeof(float) * (((size_t)n + 15) & ~(size_t)15));
    float *z = (float *)aligned_alloc(64, sizeof(float) * (((size_t)n + 15) & ~(size_t)15));
    for (int j = 0; j < n; j++) { x[j] = p[j].x; y[j] = p[j].y; z[j] = p[j].z; }
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = i0; i < i1; i++) {
        float fx = 0.f, fy = 0.f, fz = 0.f;
        const float xi = x[i], yi = y[i], zi = z[i];
#pragma omp simd reduction(+ : fx, fy, fz)
        for (int j = 0; j < n; j++) {
            float dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
            float s = dx * dx + dy * dy + dz * dz + ORACLE_SOFTENING_F32;
            float inv = 1.0f / sqrtf(s);
            float inv3 = inv * inv * inv;
            fx += dx * inv3; fy += dy * inv3; fz += dz * inv3;
        }
        p[i].vx += dt * fx; p[i].vy += dt * fy; p[i].vz += dt * fz;
    }
    free(x); free(y); free(z);
    return (double)(i1 - i0) * (double)n;
}
