import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, mini_nbody_b200 as nb, oracle_lib as orc
va, vb = int(sys.argv[1]), int(sys.argv[2])
for n in [int(x) for x in os.environ.get("DIAG_N", "1024,4096,65536").split(",")]:
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        h.set_option("variant", va); a = h.accel()
        h.set_option("variant", vb); c = h.accel()
        a2 = h.accel()
    d = np.argwhere(a != c)
    rel = np.abs(a - c) / (np.abs(a) + 1e-30)
    idx = np.unique(d[:, 0]) if d.size else np.array([], dtype=int)
    print(json.dumps({"n": n, "mismatch_elems": int(len(d)), "bodies": int(len(idx)), "repeatable": bool(np.array_equal(c, a2)), "max_rel": float(rel.max()),
                      "first": idx[:12].tolist(), "mod128_hist_nonzero": int((np.bincount(idx % 128, minlength=128) > 0).sum()) if len(idx) else 0,
                      "tile_hist": np.bincount(idx // 1024).tolist()[:16] if len(idx) else []}))
