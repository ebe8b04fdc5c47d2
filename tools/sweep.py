"""Variant sweep on the GPU box: python tools/sweep.py N steps [variants...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
n = int(sys.argv[1]); steps = int(sys.argv[2]); prec = int(os.environ.get("PREC", "0"))
b = orc.randomize(n, 42)
if prec: b = orc.widen(b)
res = []
with nb.NBody(n, prec) as h:
    h.upload(b)
    h.set_option("timing", 1)
    print(json.dumps(h.probe_fp32_peak()))
    vs = [int(v) for v in sys.argv[3:]] or range(h.info("num_variants"))
    for v in vs:
        h.set_option("variant", v)
        h.step(0.01, 1)
        best = 1e30
        for rep in range(2):
            h.timing_reset(); h.step(0.01, steps)
            best = min(best, h.timing()["force_ms"] / steps)
        r = {"n": n, "variant": v, "force_ms": best, "G_inter_s": n * n / (best * 1e-3) / 1e9, "splits": h.info("splits_local"),
             "tile": h.info("tile_bodies"), "occ": h.info("ctas_per_sm"), "cyc_per_inter": 148 * 128 * 1.965e9 / (n * n / (best * 1e-3))}
        print(json.dumps(r)); res.append(r)
json.dump(res, open("gpurun_out/sweep_%d_p%d.json" % (n, prec), "w"), indent=1)
