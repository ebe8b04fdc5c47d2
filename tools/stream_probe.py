"""Stream-K force pass against the (i-tile, j-split) grid + integrate kernel it replaces: per-step time on one GPU
over N, FP32 and FP64, and the persistent-grid size.  Prints one JSON line per size.
usage: stream_probe.py [f32|f64] [N ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
args = sys.argv[1:]
prec = 1 if (args and args[0] == "f64") else 0
if args and args[0] in ("f32", "f64"):
    args = args[1:]
sizes = [int(x) for x in args] or ([6144, 8192, 12288, 16384, 24576, 32768, 65536, 131072] if prec == 0 else [4096, 8192, 16384, 32768, 65536])
rate = 3100e9 if prec == 0 else 1074e9
for n in sizes:
    b = orc.randomize(n, 42)
    if prec:
        b = orc.widen(b)
    with nb.NBody(n, prec) as h:
        h.upload(b)
        steps = max(4, min(40, int((4e9 if prec == 0 else 1.5e9) / (float(n) * n))) // 2 * 2)
        def t():
            h.step(0.01, steps); best = 1e9
            for _ in range(3):
                h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
            return round(best * 1e3, 2)
        row = {"n": n, "prec": "f64" if prec else "f32", "ideal_us": round(float(n) * n / rate * 1e6, 1)}
        h.set_option("stream", 0)
        row["split_grid_us"] = t(); row["split_grid_variant"] = h.info("variant"); row["splits"] = h.info("splits_local")
        h.set_option("stream", 1)
        if not h.info("stream"):
            h.set_option("variant", 19 if prec == 0 else 5)
        row["stream_us"] = t(); row["stream_variant"] = h.info("variant"); row["grid"] = h.info("grid")
        sms, occ = h.info("sms"), h.info("ctas_per_sm")
        for g in sorted({sms, sms * occ // 2, sms * occ} - {h.info("grid")}):
            h.set_option("grid", g); row["stream_g%d_us" % h.info("grid")] = t()
        h.set_option("grid", 0)
        for v in ((21, 23) if prec == 0 else (6, 7)):
            h.set_option("variant", v); row["stream_v%d_g%d_us" % (v, h.info("grid"))] = t()
        row["stream_over_ideal"] = round(row["stream_us"] / row["ideal_us"], 3)
        row["stream_over_split_grid"] = round(row["stream_us"] / row["split_grid_us"], 3)
        print(json.dumps(row), flush=True)
