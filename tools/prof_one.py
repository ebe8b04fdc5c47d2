"""Run a few steps of one configuration (profiling / clock-sampling target)."""
import argparse, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=131072)
ap.add_argument("--prec", type=int, default=0)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--splits", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--clocks", default="")
a = ap.parse_args()

b = orc.randomize(a.n, 42)
if a.prec == 1:
    b = orc.widen(b)
with nb.NBody(a.n, a.prec) as h:
    h.upload(b)
    h.set_option("timing", 1)
    h.set_option("variant", a.variant)
    if a.splits:
        h.set_option("splits", a.splits)
    h.step(0.01, 1)
    mon = None
    if a.clocks:
        mon = subprocess.Popen(["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                "--format=csv", "-lms", "100"], stdout=open(a.clocks, "w"))
        time.sleep(0.5)
    h.timing_reset()
    h.step(0.01, a.steps)
    ms = h.last_step_ms(); t = h.timing()
    if mon:
        mon.terminate(); mon.wait()
    print(json.dumps({"n": a.n, "prec": a.prec, "variant": a.variant, "ms_per_step": ms / a.steps,
                      "G_inter_s": a.n * a.n * a.steps / (ms * 1e-3) / 1e9, "force_ms": t["force_ms"] / a.steps,
                      "splits": h.info("splits_local"), "tile": h.info("tile_bodies")}))
