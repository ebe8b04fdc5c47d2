"""Item 8 of VERDICT r01: the library's automatic choice at 8 192 ... 24 576 bodies on one GPU against the N^2 / 3100 G line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
for n in [int(x) for x in sys.argv[1:]] or [8192, 10240, 12288, 16384, 20480, 24576, 32768, 65536]:
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        steps = max(4, min(40, int(4e9 / (float(n) * n))) // 2 * 2)
        h.step(0.01, steps); best = 1e9
        for _ in range(4):
            h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
        ideal = n * float(n) / 3100e9 * 1e6
        print(json.dumps({"n": n, "auto_us": round(best * 1e3, 2), "ideal_us_3100": round(ideal, 1), "auto_over_ideal": round(best * 1e3 / ideal, 3),
                          "path": "small" if h.info("small_launches") else ("fused" if h.info("fuse") else "split grid + integrate (graph replay)"),
                          "variant": h.info("variant"), "splits": h.info("splits_local")}), flush=True)
