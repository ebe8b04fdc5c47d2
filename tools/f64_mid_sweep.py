"""FP64 force kernel: every variant x forced split counts at several sizes against the automatic plan."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
sizes = [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384, 32768, 49152, 65536]
for n in sizes:
    b = orc.widen(orc.randomize(n, 42))
    with nb.NBody(n, nb.F64) as h:
        h.upload(b)
        steps = max(2, min(20, int(1.5e9 / (n * n))) // 2 * 2)
        def t():
            h.step(0.01, steps); best = 1e9
            for _ in range(3):
                h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
            return round(best * 1e3, 2)
        row = {"n": n, "auto_us": t(), "auto_variant": h.info("variant"), "auto_splits": h.info("splits_local"), "ideal_us_1140": round(n * n / 1140e9 * 1e6, 1)}
        best = (1e9, None)
        for v in range(h.info("num_variants")):
            for sp in (0, 8, 12, 16, 24, 32, 37, 48):
                h.set_option("variant", v); h.set_option("splits", sp)
                us = t(); key = "v%d_s%d%s" % (v, h.info("splits_local"), "p" if sp == 0 else "")
                row[key] = us
                if us < best[0]: best = (us, key)
        row["best"] = best[1]; row["best_us"] = best[0]; row["auto_over_best"] = round(row["auto_us"] / best[0], 3)
        print(json.dumps(row), flush=True)
