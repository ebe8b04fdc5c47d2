"""A/B on one GPU: split-grid kernel with slot array + integrate kernel, the same in fused mode (last-arriver reduction and
integrate in the force kernel), and the stream-K kernel."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
for n in [int(x) for x in sys.argv[1:]] or [8192, 16384, 32768, 131072, 1048576]:
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        steps = max(2, min(40, int(6e10 / (float(n) * n))) // 2 * 2)
        def t():
            h.step(0.01, steps); best = 1e9
            for _ in range(3):
                h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
            return round(best * 1e3, 2)
        row = {"n": n, "steps": steps, "ideal_us_3100": round(float(n) * n / 3100e9 * 1e6, 1)}
        h.set_option("stream", 0); h.set_option("fuse", 0); row["split_grid_us"] = t(); row["splits"] = h.info("splits_local")
        h.set_option("fuse", 1); row["split_grid_fused_us"] = t(); row["ring"] = h.info("ring")
        h.set_option("order", 0); row["split_grid_fused_split_major_us"] = t(); h.set_option("order", 1)
        h.set_option("stream", 1)
        if not h.info("stream"):
            h.set_option("variant", 19)
        row["stream_us"] = t()
        row["fused_over_split_grid"] = round(row["split_grid_fused_us"] / row["split_grid_us"], 4)
        print(json.dumps(row), flush=True)
