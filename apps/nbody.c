/*
 * nbody.c -- plain-C host driver of the hot path, linked against libnbody_b200.so.
 *
 * This is the caller the reference's absent host program would be: it owns the
 * Body array, calls bodyForce() and integrate() once per iteration through the C ABI of
 * include/nbody.h and prints the per-iteration time and the final
 * "<N> Bodies: average <X> Billion Interactions / second" line.  No CUDA in this file.
 *
 *   nbody [nBodies] [nIters] [--resident] [--fp64] [--gpus G] [--check] [--softening EPS] [--kdk]
 *
 *   default          the literal drop-in loop: bodyForce(p, dt, n); integrate(p, dt, n);  (host
 *                    buffers cross PCIe on every call)
 *   --resident       keep the bodies in HBM: nbody_upload once, nbody_step per iteration, nbody_download
 *   --gpus G         shard the i-bodies over G GPUs of this box (resident mode)
 *   --check          print total energy before and after (FP64 diagnostic kernel)
 *   --softening EPS  constant added to dist^2 instead of the reference's 1e-9 (resident mode)
 *   --kdk            kick-drift-kick leapfrog from the same kernels instead of the reference's kick-drift step
 *                    (resident mode)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/nbody.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, nbody_last_error()); return 1; } } while (0)

int main(int argc, char **argv) {
    int nBodies = 30000, nIters = 10, resident = 0, fp64 = 0, gpus = 1, check = 0, npos = 0, kdk = 0;
    double softening = 0.0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--resident")) resident = 1;
        else if (!strcmp(argv[i], "--fp64")) fp64 = 1;
        else if (!strcmp(argv[i], "--check")) check = 1;
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) { gpus = atoi(argv[++i]); resident = 1; }
        else if (!strcmp(argv[i], "--softening") && i + 1 < argc) { softening = atof(argv[++i]); resident = 1; }
        else if (!strcmp(argv[i], "--kdk")) { kdk = 1; resident = 1; }
        else if (npos == 0) { nBodies = atoi(argv[i]); npos++; }
        else if (npos == 1) { nIters = atoi(argv[i]); npos++; }
    }
    if (nBodies <= 0 || nIters <= 0 || gpus <= 0) { fprintf(stderr, "usage: nbody [nBodies] [nIters] [--resident] [--fp64] [--gpus G] [--check] [--softening EPS] [--kdk]\n"); return 2; }

    const float dt = 0.01f;                       /* time step */
    const size_t nfl = 6 * (size_t)nBodies;
    float *buf = (float *)malloc(nfl * sizeof(float));
    Body *p = (Body *)buf;
    randomizeBodies(buf, (int)nfl);               /* seeded: $NBODY_SEED or 42 */
    BodyD *pd = NULL;
    if (fp64) {
        pd = (BodyD *)malloc(sizeof(BodyD) * (size_t)nBodies);
        for (int i = 0; i < nBodies; i++) { pd[i].x = p[i].x; pd[i].y = p[i].y; pd[i].z = p[i].z; pd[i].vx = p[i].vx; pd[i].vy = p[i].vy; pd[i].vz = p[i].vz; }
    }

    nbody_handle h = NULL;
    double e0 = 0, e1 = 0, ke, pe;
    if (resident || check) {
        CHECK(nbody_create(nBodies, fp64 ? NBODY_F64 : NBODY_F32, gpus, &h));
        if (softening > 0.0) CHECK(nbody_set_softening(h, softening));
        if (fp64) CHECK(nbody_upload_d(h, pd)); else CHECK(nbody_upload(h, p));
        if (check) { CHECK(nbody_energy(h, &ke, &pe)); e0 = ke + pe; }
    }

    double totalTime = 0.0;
    for (int iter = 1; iter <= nIters; iter++) {
        const double t0 = now_s();
        if (resident) {
            if (kdk) CHECK(nbody_step_kdk(h, (double)dt, 1)); else CHECK(nbody_step(h, (double)dt, 1));
        } else if (fp64) {
            bodyForceD(pd, (double)dt, nBodies);  /* compute interbody forces, v += dt*F */
            integrateD(pd, (double)dt, nBodies);  /* integrate position */
        } else {
            bodyForce(p, dt, nBodies);
            integrate(p, dt, nBodies);
        }
        const double tElapsed = now_s() - t0;
        if (iter > 1) totalTime += tElapsed;      /* first iteration is warm-up */
        printf("Iteration %d: %.6f seconds\n", iter, tElapsed);
    }
    const double avgTime = nIters > 1 ? totalTime / (double)(nIters - 1) : totalTime;
    if (resident) { if (fp64) CHECK(nbody_download_d(h, pd)); else CHECK(nbody_download(h, p)); }
    if (check) {
        if (!resident) { if (fp64) CHECK(nbody_upload_d(h, pd)); else CHECK(nbody_upload(h, p)); }
        CHECK(nbody_energy(h, &ke, &pe)); e1 = ke + pe;
        printf("Energy: %.9e -> %.9e (relative drift %.3e)\n", e0, e1, (e1 - e0) / (e0 < 0 ? -e0 : e0));
    }
    if (nIters > 1)
        printf("%d Bodies: average %0.3f Billion Interactions / second\n", nBodies, 1e-9 * (double)nBodies * (double)nBodies / avgTime);
    if (h) nbody_destroy(h);
    free(buf); free(pd);
    return 0;
}
