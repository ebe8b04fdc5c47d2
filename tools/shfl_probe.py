"""Broadcast LDS.128 (variant 12: same loop, ptxas schedule, unroll 1; variant 3: unroll 2) vs warp-shuffle broadcast (variant 18)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, mini_nbody_b200 as nb, oracle_lib as orc
n = 262144
b = orc.randomize(n, 42)
with nb.NBody(n) as h:
    h.upload(b); h.set_option("timing", 1)
    acc = {}
    for v in (12, 3, 18, 14):                      # all accelerations first: the timed steps below move the bodies
        h.set_option("variant", v); acc[v] = h.accel()
    for v in (12, 3, 18, 14):
        h.set_option("variant", v); same = bool(np.array_equal(acc[v], acc[12]))
        h.step(0.01, 2); best = 1e9
        for rep in range(3):
            h.timing_reset(); h.step(0.01, 4); best = min(best, h.timing()["force_ms"] / 4)
        print(json.dumps({"variant": v, "cyc_per_interaction": round(148 * 128 * 1.965e9 / (n * n / (best * 1e-3)), 3), "bit_identical_to_v12": same, "splits": h.info("splits_local"), "ctas_per_sm": h.info("ctas_per_sm")}))
