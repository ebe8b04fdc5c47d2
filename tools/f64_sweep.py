"""FP64 force kernel: variant sweep at C3 (N = 65 536)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, mini_nbody_b200 as nb, oracle_lib as orc
n = 65536
b = orc.widen(orc.randomize(n, 42))
ref = orc.accel_f64(b, 1000, 1256)
with nb.NBody(n, nb.F64) as h:
    h.upload(b); h.set_option("timing", 1)
    nv = h.info("num_variants")
    acc = {}
    for v in range(nv):
        h.set_option("variant", v); acc[v] = h.accel()
    for v in range(nv):
        h.set_option("variant", v); h.step(0.01, 2); best = 1e9
        for rep in range(3):
            h.timing_reset(); h.step(0.01, 4); best = min(best, h.timing()["force_ms"] / 4)
        print(json.dumps({"variant": v, "force_ms": round(best, 4), "G_inter_s": round(n * n / (best * 1e-3) / 1e9, 1), "err": float(orc.rel_err(acc[v][1000:1256], ref).max()),
                          "splits": h.info("splits_local"), "ctas_per_sm": h.info("ctas_per_sm"), "tile": h.info("tile_bodies")}), flush=True)
