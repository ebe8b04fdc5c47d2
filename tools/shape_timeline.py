"""Per-CTA timeline of ONE rank's stream-K pass in the C3 x 8 shape (8192 i x 65536 j, FP64), on one GPU with virtual ranks."""
import json, os, sys
os.environ["NBODY_VIRTUAL_RANKS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc
G, n = 8, 65536
b = orc.widen(orc.randomize(n, 42))
for ov in (1, 0):
    with nb.NBody(n, nb.F64, ngpus=G) as h:
        h.set_option("overlap", ov); h.upload(b); h.step(0.01, 3)
        h.set_option("profile", 1); h.step(0.01, 1)
        r = h.stream_profile().astype(np.float64)
        t0 = r[:, 0].min()
        entry, seg_done, nseg, nred, tred, exit_ = (r[:, 0] - t0) / 1e3, (r[:, 1] - t0) / 1e3, r[:, 2], r[:, 3], r[:, 4] / 1e3, (r[:, 5] - t0) / 1e3
        q = lambda a: [round(float(x), 1) for x in (a.min(), np.median(a), a.max())]
        print(json.dumps({"phases": 2 if ov else 1, "ctas": len(r), "entry_us": q(entry), "last_segment_done_us": q(seg_done), "exit_us": q(exit_),
                          "segments_per_cta": q(nseg), "reductions": int(nred.sum()), "reduction_us_each": q(tred[nred > 0] / nred[nred > 0]),
                          "step_ms_all_ranks": round(h.last_step_ms(), 4)}), flush=True)
