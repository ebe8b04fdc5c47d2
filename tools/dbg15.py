import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, mini_nbody_b200 as nb, oracle_lib as orc
eps = 1e-2
for n in (3000, 12000, 30000):
    b = orc.randomize(n, 77)
    i1 = min(n, 4096)
    with orc.softening(eps):
        ref = orc.accel_f64_from_f32(b, 0, i1)
    ref0 = orc.accel_f64_from_f32(b, 0, i1)
    with nb.NBody(n) as h:
        h.upload(b)
        a_default = h.accel()
        print(n, "default variant", h.info("variant"), "err vs eps=1e-9 oracle %.3e" % orc.rel_err(a_default[:i1], ref0).max())
        h.set_softening(eps)
        a = h.accel()
        e = orc.rel_err(a[:i1], ref)
        print(n, "variant", h.info("variant"), "fuse", h.info("fuse"), "splits", h.info("splits_local"), "max err %.3e" % e.max(), "n bad", int((e > 1e-5).sum()),
              "err vs eps=1e-9 oracle %.3e" % orc.rel_err(a[:i1], ref0).max())
        a2 = h.accel()
        print("   second call identical:", np.array_equal(a, a2), " max err %.3e" % orc.rel_err(a2[:i1], ref).max())
