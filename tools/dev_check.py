"""Development check run on the GPU box: parity vs the oracle at small N and a variant sweep."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc

out = []
def log(**kw):
    print(json.dumps(kw)); out.append(kw); sys.stdout.flush()

# parity, N=4096 and ragged N
for n in (4096, 1000, 131072):
    b = orc.randomize(n, 42)
    samp = min(n, 2048)
    ref64 = orc.accel_f64_from_f32(b, 0, samp)
    ref32 = orc.accel_f32(b, 0, samp)
    with nb.NBody(n) as h:
        h.upload(b)
        h.set_option("timing", 1)
        for v in range(h.info("num_variants")):
            h.set_option("variant", v)
            a = h.accel()[:samp]
            e64 = orc.rel_err(a, ref64); e32 = orc.rel_err(a, ref32)
            log(test="parity_f32", n=n, variant=v, max_vs_f64=float(e64.max()), p99_vs_f64=float(np.percentile(e64, 99)), max_vs_f32=float(e32.max()),
                cpu32_vs_f64=float(orc.rel_err(ref32, ref64).max()), splits=h.info("splits_local"))
for n in (4096, 1000, 65536):
    b = orc.widen(orc.randomize(n, 42))
    samp = min(n, 1024)
    ref = orc.accel_f64(b, 0, samp)
    with nb.NBody(n, nb.F64) as h:
        h.upload(b)
        for v in range(h.info("num_variants")):
            h.set_option("variant", v)
            a = h.accel()[:samp]
            log(test="parity_f64", n=n, variant=v, max_err=float(orc.rel_err(a, ref).max()))

# perf sweep
for n, prec in ((131072, nb.F32), (1048576, nb.F32), (65536, nb.F64)):
    b = orc.randomize(n, 42)
    if prec == nb.F64: b = orc.widen(b)
    with nb.NBody(n, prec) as h:
        h.upload(b)
        print(h.probe_fp32_peak())
        for v in range(h.info("num_variants")):
            h.set_option("variant", v)
            steps = 3 if n >= 1000000 else 10
            h.step(0.01, 1)
            h.timing_reset()
            h.step(0.01, steps)
            ms = h.last_step_ms(); t = h.timing()
            log(test="perf", n=n, prec=prec, variant=v, ms_per_step=ms / steps, G_inter_s=n * n * steps / (ms * 1e-3) / 1e9,
                force_ms=t["force_ms"] / steps, integ_ms=t["integrate_ms"] / steps, splits=h.info("splits_local"), tile=h.info("tile_bodies"))
json.dump(out, open("gpurun_out/dev_check.json", "w"), indent=1)
