"""Per-CTA timeline of the stream-K force pass (option profile=1, nbody_stream_profile): where the time of a pass
goes that is not the inner loop -- launch skew, segment work, last-arriver reductions, tail."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc
prec = 0
args = sys.argv[1:]
if args and args[0] in ("f32", "f64"):
    prec = 1 if args[0] == "f64" else 0; args = args[1:]
for n in [int(x) for x in args] or [8192, 16384, 131072, 1048576]:
    b = orc.randomize(n, 42)
    if prec: b = orc.widen(b)
    with nb.NBody(n, prec) as h:
        h.set_option("stream", 1)
        if not h.info("stream"):
            h.set_option("variant", 19 if prec == 0 else 5)
        h.set_option("graph", 0)
        h.upload(b)
        h.step(0.01, 2)
        h.set_option("profile", 1)
        h.step(0.01, 1); ms = h.last_step_ms()
        r = h.stream_profile().astype(np.float64)
        t0 = r[:, 0].min()
        entry, seg_done, nseg, nred, tred, exit_, smid = (r[:, 0] - t0) / 1e3, (r[:, 1] - t0) / 1e3, r[:, 2], r[:, 3], r[:, 4] / 1e3, (r[:, 5] - t0) / 1e3, r[:, 6]
        per_sm = np.bincount(smid.astype(int))
        print(json.dumps({"n": n, "prec": prec, "grid": len(r), "i_tiles": h.info("i_tiles"), "step_ms_event": round(ms, 4),
                          "span_us": round(exit_.max(), 2), "entry_us_max": round(entry.max(), 2),
                          "seg_done_us_min_med_max": [round(float(x), 2) for x in (seg_done.min(), np.median(seg_done), seg_done.max())],
                          "exit_us_min_med_max": [round(float(x), 2) for x in (exit_.min(), np.median(exit_), exit_.max())],
                          "segments_per_cta_max": int(nseg.max()), "reductions_total": int(nred.sum()), "reduction_us_mean_max": [round(float(tred[nred > 0].mean()) if (nred > 0).any() else 0, 2), round(float(tred.max()), 2)],
                          "ctas_per_sm_min_max": [int(per_sm[per_sm > 0].min()), int(per_sm.max())], "sms_used": int((per_sm > 0).sum())}), flush=True)
