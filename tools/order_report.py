"""how much summation order alone moves the force (all against the FP64 oracle)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mini_nbody_b200 as nb
import oracle_lib as orc
rows = []
for n, samp in ((4096, 4096), (32768, 4096), (131072, 2048), (1048576, 512)):
    b = orc.randomize(n, 42)
    i0 = (n - samp) // 2; i1 = i0 + samp
    ref = orc.accel_f64_from_f32(b, i0, i1)
    seq = orc.accel_f32(b, i0, i1); fpga = orc.accel_f32(b, i0, i1, order="fpga")
    with nb.NBody(n) as h:
        h.upload(b); gpu = h.accel()[i0:i1]; slots = h.info("slots")
    def st(a):
        e = orc.rel_err(a, ref); return {"max": float(e.max()), "p99": float(np.percentile(e, 99)), "median": float(np.median(e))}
    row = {"n": n, "sample": samp, "cpu_fp32_sequential_j": st(seq), "cpu_fp32_fpga_order_16_interleaved_tree": st(fpga), "gpu_fp32": st(gpu), "gpu_slots": slots}
    print(json.dumps(row)); rows.append(row)
json.dump(rows, open("gpurun_out/order_report.json", "w"), indent=1)
