"""A/B of the stream-K kernel's experiment switches (option "tune") against the split-grid path."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb
import oracle_lib as orc
for n in [int(x) for x in sys.argv[1:]] or [131072, 1048576]:
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h:
        h.upload(b)
        steps = max(2, min(20, int(6e10 / (float(n) * n))) // 2 * 2)
        def t():
            h.step(0.01, steps); best = 1e9
            for _ in range(3):
                h.step(0.01, steps); best = min(best, h.last_step_ms() / steps)
            return round(best * 1e3, 2)
        row = {"n": n, "steps": steps}
        h.set_option("stream", 0); row["split_grid_us"] = t()
        h.set_option("stream", 1)
        for tune in (0, 1, 2, 3):
            h.set_option("tune", tune); row["stream_tune%d_us" % tune] = t()
        h.set_option("tune", 0); h.set_option("grid", 148); row["stream_g148_us"] = t()
        print(json.dumps(row), flush=True)
