"""Generates tests/golden/c1_small.npz from the oracle's parity build (gcc -O2 -ffp-contract=off).
The reference itself (VHDL + absent vendor IP) cannot be executed, so these are regression pins of
the oracle, not reference outputs.  Run: python tests/golden/make_golden.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as orc
n, seed, steps = 512, 42, 10
b = orc.randomize(n, seed)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_small.npz"),
                    n=n, seed=seed, steps=steps, bodies=b.view(np.float32), accel_f32=orc.accel_f32(b),
                    accel_f64=orc.accel_f64_from_f32(b), after_1_step=orc.run(b, 0.01, 1).view(np.float32),
                    after_steps=orc.run(b, 0.01, steps).view(np.float32))
print("wrote c1_small.npz")
