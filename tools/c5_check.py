"""C5: N = 4 194 304 FP32, 2 steps on one GPU: throughput + sampled parity + momentum conservation."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc
n = 4194304
b = orc.randomize(n, 42)
with nb.NBody(n) as h:
    h.upload(b)
    a = h.accel()
    i0 = 2000000; i1 = i0 + 128
    e = orc.rel_err(a[i0:i1], orc.accel_f64_from_f32(b, i0, i1))
    a64 = a.astype(np.float64)
    mom = float(np.abs(a64.sum(axis=0)).max() / np.abs(a64).sum(axis=0).max())
    h.timing_reset(); h.step(0.01, 2); ms = h.last_step_ms() / 2
    out = {"n": n, "max_rel_err_vs_fp64_128_sample": float(e.max()), "momentum_residual": mom, "ms_per_step": ms,
           "G_inter_s": n * float(n) / (ms * 1e-3) / 1e9, "slots": h.info("slots"), "variant": h.info("variant")}
print(json.dumps(out)); json.dump(out, open("gpurun_out/c5_check.json", "w"))
