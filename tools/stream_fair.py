"""Fairness of the two CTAs that share an SM in the persistent stream-K kernel, per library build
(NBODY_B200_LIB): pass time and the spread of the CTAs' finish times (equal work => equal finish if fair)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, mini_nbody_b200 as nb, oracle_lib as orc
n = int(os.environ.get("FAIR_N", "131072"))
b = orc.randomize(n, 42)
with nb.NBody(n) as h:
    h.upload(b); h.set_option("variant", 19); h.set_option("graph", 0)
    for kv in os.environ.get("FAIR_OPTS", "").split(","):
        if kv:
            h.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    h.step(0.01, 2); best = 1e9
    for rep in range(3):
        h.step(0.01, 2); best = min(best, h.last_step_ms() / 2)
    h.set_option("profile", 1); h.step(0.01, 1)
    r = h.stream_profile().astype(np.float64)
    t0 = r[:, 0].min(); done = (r[:, 1] - t0) / 1e3; sm = r[:, 6].astype(int)
    first = np.array([done[sm == s].min() for s in np.unique(sm)]); last = np.array([done[sm == s].max() for s in np.unique(sm)])
    h.set_option("profile", 0); h.set_option("stream", 0); h.step(0.01, 2); sg = 1e9
    for rep in range(3):
        h.step(0.01, 2); sg = min(sg, h.last_step_ms() / 2)
    print(json.dumps({"lib": os.path.basename(os.environ.get("NBODY_B200_LIB", "default")), "opts": os.environ.get("FAIR_OPTS", ""), "n": n,
                      "stream_us": round(best * 1e3, 1), "split_grid_us": round(sg * 1e3, 1),
                      "cyc_per_inter": round(148 * 128 * 1.965e9 / (n * float(n) / (best * 1e-3)), 3),
                      "first_cta_of_sm_done_us_med": round(float(np.median(first)), 1), "second_cta_of_sm_done_us_med": round(float(np.median(last)), 1),
                      "first_over_second": round(float(np.median(first / last)), 3)}))
''' % (ROOT, ROOT)
for lib in sys.argv[1:]:
    env = dict(os.environ, NBODY_B200_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", code], env=env, timeout=300)
