"""One process driving G GPUs (nbody_create): per-step time of the sharded FP32 pass, fused (one launch per rank and step)
against unfused (force A, flag wait, force B, integrate), push exchange.  usage: multi_probe.py G [N ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
G = int(sys.argv[1])
for n in [int(x) for x in sys.argv[2:]] or [32768, 65536, 131072, 262144]:
    b = orc.randomize(n, 42)
    row = {"n": n, "gpus": G, "n_local": n // G}
    for prec, tag in ((0, "f32"), (1, "f64")):
        if prec and n > 131072:
            continue
        bb = orc.widen(b) if prec else b
        for fuse in ((0, 1) if prec == 0 else (-1,)):
            with nb.NBody(n, prec, ngpus=G) as h:
                h.set_option("exchange", 1)
                if prec == 0:
                    h.set_option("fuse", fuse)
                h.upload(bb)
                steps = max(4, min(40, int(2e10 * G / (float(n) * n))))
                h.step(0.01, steps); best = 1e9
                for _ in range(3):
                    h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
                key = "%s_%s_us" % (tag, {0: "unfused", 1: "fused", -1: "default"}[fuse])
                row[key] = round(best * 1e3, 2)
                row[key.replace("_us", "_variant")] = h.info("variant")
    row["ideal_f32_us"] = round(float(n) * n / G / 3100e9 * 1e6, 1)
    print(json.dumps(row), flush=True)
