import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, mini_nbody_b200 as nb, oracle_lib as orc
n = 262144
b = orc.randomize(n, 42)
with nb.NBody(n) as h:
    h.upload(b); h.set_option("timing", 1)
    vs = [int(x) for x in os.environ.get("AB_VARIANTS", "3").split(",")]
    h.set_option("variant", 12); a12 = h.accel()
    out = {"lib": os.path.basename(os.environ.get("NBODY_B200_LIB", "default"))}
    for v in vs:
        h.set_option("variant", v); out["v%%d_bit_identical" %% v] = bool(np.array_equal(h.accel(), a12))
    for v in vs + [12]:
        h.set_option("variant", v); h.step(0.01, 2)
        best = 1e9
        for rep in range(3):
            h.timing_reset(); h.step(0.01, 4); best = min(best, h.timing()["force_ms"] / 4)
        out["v%%d_cyc" %% v] = round(148 * 128 * 1.965e9 / (n * n / (best * 1e-3)), 3)
print(json.dumps(out))
''' % (ROOT, ROOT)
for lib in sys.argv[1:]:
    env = dict(os.environ, NBODY_B200_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", code], env=env, timeout=300)
