"""N scan on one GPU: per-step device time and G inter/s for the default variant choice (and all-variant best)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc
res = []
for n in (1024, 4096, 16384, 32768, 65536, 131072, 262144, 524288):
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        steps = 20 if n <= 65536 else 5
        row = {"n": n, "default_variant": h.info("variant")}
        best = (1e30, -1)
        for v in ([h.info("variant")] + [x for x in (1, 3, 4, 6, 13, 14) if x != h.info("variant")]):
            h.set_option("variant", v)
            h.step(0.01, 2)
            t = min(_t for _t in [(h.step(0.01, steps), h.last_step_ms() / steps)[1] for _ in range(3)])
            row["v%d_us" % v] = round(t * 1e3, 2)
            if t < best[0]: best = (t, v)
        row["best_variant"] = best[1]; row["best_G_inter_s"] = round(n * n / (best[0] * 1e-3) / 1e9, 1)
        row["default_G_inter_s"] = round(n * n / (row["v%d_us" % row["default_variant"]] * 1e-6) / 1e9, 1)
        print(json.dumps(row)); res.append(row)
json.dump(res, open("gpurun_out/scan_n.json", "w"), indent=1)
