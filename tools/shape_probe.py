"""Per-rank kernel shapes of the 8-GPU runs, timed on ONE GPU with virtual ranks (NBODY_VIRTUAL_RANKS=1: the ranks' launches run
one after the other on one stream, so step time / G is the time of one rank's pass without any exchange latency):
C2 x 8 (16 384 i-bodies against 131 072 j) over forced j-split counts, C3 x 8 (FP64, 8 192 against 65 536) over stream-K shapes."""
import json, os, sys
os.environ["NBODY_VIRTUAL_RANKS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
G = 8
def timed(h, steps):
    h.step(0.01, steps); best = 1e9
    for _ in range(3):
        h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
    return round(best * 1e3 / G, 2)
which = sys.argv[1:] or ["c2", "c3"]
if "c2" in which:
    n = 131072; b = orc.randomize(n, 42)
    row = {"shape": "C2 x 8: 16384 i x 131072 j FP32 per rank", "ideal_us_3100": round(n * n / G / 3100e9 * 1e6, 1)}
    with nb.NBody(n, ngpus=G) as h:
        h.upload(b)
        for fuse in (1, 0):
            h.set_option("fuse", fuse)
            for sp in (0, 9, 18, 19, 27, 37, 48):
                h.set_option("splits", sp)
                row["fuse%d_splits%d(%d)_us" % (fuse, sp, h.info("splits_local") + h.info("splits_remote"))] = timed(h, 6)
    print(json.dumps(row), flush=True)
if "c3" in which:
    n = 65536; b = orc.widen(orc.randomize(n, 42))
    row = {"shape": "C3 x 8: 8192 i x 65536 j FP64 per rank", "ideal_us_1084": round(n * n / G / 1084e9 * 1e6, 1)}
    with nb.NBody(n, nb.F64, ngpus=G) as h:
        h.upload(b)
        for v in (5, 6, 7, 4, 1, 2):
            for ov in (1, 0):
                h.set_option("variant", v); h.set_option("overlap", ov)
                row["v%d_overlap%d_us" % (v, ov)] = timed(h, 10)
        h.set_option("variant", 5); h.set_option("overlap", 1)
        for g in (74, 111, 148):
            h.set_option("grid", g); row["v5_grid%d_us" % g] = timed(h, 10)
    print(json.dumps(row), flush=True)
