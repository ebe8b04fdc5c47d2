"""Fused multi-step kernel at C1-like sizes: (variant, j-splits) grid, per-step time (follow-up of fused_probe.py)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
for n in (2048, 3072, 4096, 5120, 6144, 8192):
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        def t(steps=40):
            h.step(0.01, steps); best = 1e9
            for _ in range(3):
                h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
            return round(best * 1e3, 2)
        row = {"n": n, "auto_us": t(), "auto_variant": h.info("variant"), "auto_splits": h.info("splits_local")}
        h.set_option("fused", 1)
        for v in (6, 4):
            for sp in (4, 8, 12, 16, 24, 32):
                try:
                    h.set_option("variant", v); h.set_option("splits", sp)
                    row["v%d_s%d" % (v, h.info("splits_local"))] = t()
                except nb.NBodyError as e:
                    row["err_v%d_s%d" % (v, sp)] = str(e)[:60]
        print(json.dumps(row), flush=True)
