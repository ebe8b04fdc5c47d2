"""Sharded FP32 pass at N = 1M on G GPUs of one process: unfused (two force launches + integrate), fused tile-major + ring,
fused split-major (slots of all tiles).  usage: multi_order_probe.py G [N]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
G = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
b = orc.randomize(n, 42)
row = {"n": n, "gpus": G, "ideal_ms_at_3104": round(float(n) * n / G / 3104e9 * 1e3, 3)}
for name, opts in (("unfused", {"fuse": 0}), ("fused_tile_major", {"fuse": 1, "order": 1}), ("fused_split_major", {"fuse": 1, "order": 0})):
    with nb.NBody(n, ngpus=G) as h:
        h.set_option("exchange", 1)
        for k, v in opts.items():
            h.set_option(k, v)
        h.upload(b)
        h.step(0.01, 2); one = []; many = 1e9
        for _ in range(3):
            h.step(0.01, 1); one.append(h.last_step_ms())
        for _ in range(2):
            h.step(0.01, 4); many = min(many, h.last_step_ms() / 4)
        row[name + "_ms_single_step"] = round(min(one), 3); row[name + "_ms_back_to_back"] = round(many, 3)
        row[name + "_ws_mb"] = round(h.info("workspace_bytes") / 1e6, 1)
print(json.dumps(row), flush=True)
