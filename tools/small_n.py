"""Where the time goes at launch-bound sizes: per-kernel event times vs whole-step time."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
for n in (1024, 4096, 16384, 32768):
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        for v, sp in ((h.info("variant"), 0),):
            h.set_option("variant", v); h.set_option("splits", sp)
            h.step(0.01, 3)
            h.set_option("timing", 1); h.timing_reset(); h.step(0.01, 20); t = h.timing(); ms_t = h.last_step_ms() / 20
            h.set_option("timing", 0); h.set_option("graph", 0); h.step(0.01, 20); ms = h.last_step_ms() / 20
            h.set_option("graph", 1); h.step(0.01, 20); h.step(0.01, 20); ms_g = h.last_step_ms() / 20
            print(json.dumps({"n": n, "variant": v, "splits": h.info("splits_local"), "step_us": round(ms * 1e3, 2), "step_us_graph": round(ms_g * 1e3, 2), "step_us_with_events": round(ms_t * 1e3, 2),
                              "force_us": round(t["force_ms"] / 20 * 1e3, 2), "integrate_us": round(t["integrate_ms"] / 20 * 1e3, 2),
                              "G_inter_s": round(n * n / (ms * 1e-3) / 1e9, 1)}))
