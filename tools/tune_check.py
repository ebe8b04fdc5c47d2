"""GPU check of the post-ptxas tuned loop: variant 3 (tuned by tools/sass_tune.py) must be bit-identical to
variant 12 (same arithmetic, unroll 1, untouched by the tuner); then time both."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc
ok = True
for n in (5000, 131072):
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        h.set_option("variant", 3); a3 = h.accel()
        h.set_option("variant", 12); a12 = h.accel()
    same = np.array_equal(a3, a12)
    ref = orc.accel_f64_from_f32(b, 0, 2048)
    print(json.dumps({"n": n, "bit_identical_v3_v12": bool(same), "max_rel_err_v3": float(orc.rel_err(a3[:2048], ref).max()), "max_abs_diff": float(np.abs(a3 - a12).max())}))
    ok &= same
n = 1048576
b = orc.randomize(n, 42)
with nb.NBody(n) as h:
    h.upload(b); h.set_option("timing", 1)
    for v in (3, 12, 7, 3, 12):
        h.set_option("variant", v); h.step(0.01, 1)
        h.timing_reset(); h.step(0.01, 2); t = h.timing()["force_ms"] / 2
        print(json.dumps({"n": n, "variant": v, "force_ms": t, "G_inter_s": n * n / (t * 1e-3) / 1e9, "cyc_per_inter": 148 * 128 * 1.965e9 / (n * n / (t * 1e-3))}))
sys.exit(0 if ok else 1)
