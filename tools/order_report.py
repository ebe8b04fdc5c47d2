"""how much summation order alone moves the force.  Per-body relative error against the FP64 oracle of
the same FP32 inputs for: the sequential-j CPU loop, the reference hardware's own order (16 interleaved partials + adder
tree, S/fxyz.vhd:120-145, S/final_adder.vhd:88-104), a Kahan-compensated sequential sum, and the shipped GPU kernels
(three-level: 256-add register chains -> per-stage shared-memory sums -> per-split slots added in order)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/order_report.json"
rows = []
for n, samp in ((4096, 4096), (32768, 4096), (131072, 2048), (1048576, 512)):
    b = orc.randomize(n, 42)
    i0 = (n - samp) // 2; i1 = i0 + samp
    ref = orc.accel_f64_from_f32(b, i0, i1)
    cpu = {k: orc.accel_f32(b, i0, i1, order=k) for k in ("sequential", "fpga", "kahan")}
    with nb.NBody(n) as h:
        h.upload(b); gpu = h.accel()[i0:i1]
        info = {k: h.info(k) for k in ("variant", "splits_local", "fuse", "stream", "tile_bodies")}
    def st(a):
        e = orc.rel_err(a, ref); return {"max": float(e.max()), "p99": float(np.percentile(e, 99)), "median": float(np.median(e))}
    row = {"n": n, "sample": [i0, i1], "cpu_fp32_sequential_j": st(cpu["sequential"]), "cpu_fp32_fpga_order_16_interleaved_tree": st(cpu["fpga"]),
           "cpu_fp32_kahan_sequential_j": st(cpu["kahan"]), "gpu_fp32_three_level": st(gpu), "gpu_plan": info}
    print(json.dumps(row), flush=True); rows.append(row)
json.dump(rows, open(out, "w"), indent=1)
