"""Launch-bound sizes on one GPU: per-step time of the fused multi-step kernel vs CUDA-graph replay vs plain launches,
for several split counts (the fused kernel is persistent, so its best split count differs from the wave-based plan)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
for n in (1024, 2048, 3072, 4096, 6144, 8192):
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        row = {"n": n, "variant": h.info("variant"), "splits_plan": h.info("splits_local")}
        def t(steps=40):
            h.step(0.01, steps); best = 1e9
            for _ in range(3):
                h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
            return round(best * 1e3, 2)
        h.set_option("fused", 0); h.set_option("graph", 0); row["plain_us"] = t()
        h.set_option("graph", 1); row["graph_us"] = t()
        h.set_option("fused", -1); row["auto_us"] = t(); row["auto_fused"] = h.info("fused_launches") > 0
        h.set_option("fused", 1)
        for sp in (0, 4, 8, 16, 32):
            try:
                h.set_option("splits", sp); row["fused_us_s%d" % (sp or h.info("splits_local"))] = t()
            except nb.NBodyError as e:
                row["err_s%d" % sp] = str(e)[:60]
        if n <= 8192:
            for v in (4, 6):
                h.set_option("variant", v); h.set_option("splits", 0); row["fused_us_v%d" % v] = t()
        row["ideal_us_at_3100G"] = round(n * n / 3100e9 * 1e6, 2)
        print(json.dumps(row), flush=True)
