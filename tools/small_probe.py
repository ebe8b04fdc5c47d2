"""Small systems on one GPU: per-step time of the whole-array-in-shared-memory multi-step kernel (step_small.cu) against the
tiled paths (fused tiled kernel / CUDA-graph replay of force + integrate) it replaces by default up to 8192 bodies."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
for n in [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 3072, 4096, 4736, 6144, 8192, 12288, 16384]:
    b = orc.randomize(n, 42)
    row = {"n": n, "ideal_us_3100": round(n * n / 3100e9 * 1e6, 2)}
    for small in (1, 0):
        with nb.NBody(n) as h:
            h.set_option("small", small); h.upload(b)
            steps = 200 if n <= 4096 else 100
            h.step(0.01, steps); best = 1e9
            for _ in range(4):
                h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
            row["small_us" if small else "tiled_us"] = round(best * 1e3, 2)
            if small:
                row["small_took_it"] = h.info("small_launches") > 0
    row["G_inter_s_small"] = round(n * n / (row["small_us"] * 1e-6) / 1e9, 1)
    print(json.dumps(row), flush=True)
