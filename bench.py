#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native mini-nbody hot path.

Metric (BASELINE.json): billion interactions / s of  bodyForce + integrate  at N = 1 048 576 FP32,
strong scaling over 1/2/4/8 B200 (i-bodies sharded, positions replicated, one all-gather per step).
A "step" is one time step (force over all N^2 pairs incl. self-pairs, then integrate).

    python bench.py --gpus N --steps K --warmup W            # this framework (libnbody_b200.so)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One JSON line on stdout (rank 0).  Timing: CUDA events on the library's launching stream
(nbody_last_step_ms), each timed step bracketed by barrier + device synchronize, L2 flushed
between timed steps, MAX over ranks.  `e2e` goes through the reference-shaped C-ABI calls with
pinned HOST buffers (bodyForce / integrate at N=1; nbody_upload / nbody_step / nbody_download when
sharded), copies inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_DEFAULT = 1048576
DT = 0.01
SEED = 42
FLOP_PER_INTERACTION = 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--bodies", dest="n", type=int, default=N_DEFAULT, help="bodies (use --bodies under torchrun: its parser claims --n)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--exchange", default="push", choices=["nccl", "push"], help="per-step position exchange when sharded")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-energy", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--fuse", type=int, default=-1, help="split-grid FP32 kernels: 1/0 force the in-kernel reduction + integrate on/off")
    ap.add_argument("--overlap", type=int, default=-1, help="sharded: 0 = one pass over all j instead of own-slice-first (two stream-K phases / two launches)")
    ap.add_argument("--stream", type=int, default=-1, help="1/0: allow / forbid the stream-K kernels as the default variant")
    ap.add_argument("--cpu-seconds", type=float, default=None, help="CPU time budget of the cpu_baseline / reference leg")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle-reason sampler running during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v == "Active":
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            # samples under load = those drawing more than the idle-ish 40% of the maximum power seen
            thr = 0.6 * max(pw)
            load = [s for s, p in zip(sm, pw) if p >= thr] or sm
            out = {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def workload_string(n, precision):
    """config.workload: the same string from both arms (the driver compares them)"""
    return ("N=%d %s all-pairs bodyForce+integrate per step, dt=0.01, softening=1e-9, seeded uniform [-1,1) init (BASELINE.json configs[3])"
            % (n, precision.upper()))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_leg(n, seconds, steps=None, warmup=0):
    """Times the oracle's speed build (the reference algorithm, plain C + OpenMP, all host threads)
    on a bounded i-sample of the N-body workload.  Returns (G interactions/s, info dict)."""
    import numpy as np
    import oracle_lib as orc
    lib = orc.load("speed")
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: ask for the box's cores explicitly
    lib.oracle_set_num_threads(host_cores())
    threads = lib.oracle_num_threads()
    b = orc.randomize(n, SEED)
    # calibrate the i-sample so that one sample step takes ~seconds/steps
    probe = min(n, 64 * threads)
    t0 = time.perf_counter(); lib.oracle_body_force_f32_fast(orc._p(b), DT, n, 0, probe); t1 = time.perf_counter()
    rate = probe * n / (t1 - t0)
    nsteps = steps if steps else 3
    per_step = seconds / max(1, nsteps + warmup)
    m = int(max(threads, min(n, rate * per_step / n)))
    m -= m % max(1, threads)
    m = max(m, threads)
    for _ in range(warmup):
        lib.oracle_body_force_f32_fast(orc._p(b), DT, n, 0, m)
    times = []
    for s in range(nsteps):
        t0 = time.perf_counter()
        lib.oracle_body_force_f32_fast(orc._p(b), DT, n, 0, m)
        orc.load("speed").oracle_integrate_f32(orc._p(b[:m]), DT, m)
        times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    val = m * n / t / 1e9
    info = {"value": val, "unit": "G interactions/s", "cores": threads, "kind": "port",
            "sample": "bodyForce+integrate of %d i-bodies against all %d j-bodies per step (%.3g interactions), %d steps, gcc -O3 -ffast-math -fopenmp -march=x86-64-v3; full step extrapolates as N/%d"
                      % (m, n, m * n, nsteps, m),
            "ms_per_sample_step": t * 1e3, "ms_per_full_step_extrapolated": t * 1e3 * n / m, "sample_i_bodies": m,
            "nproc": os.cpu_count(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")}
    return val, info


def run_reference(a, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, info = cpu_reference_leg(a.n, seconds=a.cpu_seconds if a.cpu_seconds else max(20.0, 6.0 * (a.steps + a.warmup)), steps=a.steps, warmup=a.warmup)
    line = {
        "impl": "reference", "metric": "billion_interactions_per_s", "value": val, "unit": "G interactions/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": info["ms_per_sample_step"], "ms_per_full_step_extrapolated": info["ms_per_full_step_extrapolated"],
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(a.n, "f32"),
                   "note": "reference algorithm (oracle port of the VHDL pipeline + host integrate) on the box's host cores; the reference's own implementation is FPGA RTL and cannot run here"},
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": "G interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    a = parse()
    # stdout carries exactly one JSON line: anything libraries print on fd 1 (NCCL version banner ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if a.impl == "reference":
        return run_reference(a, emit)

    import numpy as np
    import torch
    import mini_nbody_b200 as nb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (a.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    if local_rank >= torch.cuda.device_count():          # launcher restricted CUDA_VISIBLE_DEVICES to one device per rank
        local_rank = local_rank % max(1, torch.cuda.device_count())
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    prec = nb.F32 if a.precision == "f32" else nb.F64
    n = a.n
    # host state in pinned memory (every rank holds the same seeded array)
    dtype = nb.body_dtype if prec == nb.F32 else nb.bodyd_dtype
    pinned = torch.empty(n * dtype.itemsize, dtype=torch.uint8).pin_memory()
    host = pinned.numpy().view(dtype)
    host[:] = nb.randomizeBodies(n, SEED, dtype)
    init = host.copy()

    nccl_id = None
    if world > 1:
        ids = [nb.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    h = nb.NBody(n, prec, rank=rank, world=world, device=local_rank, nccl_id=nccl_id)
    if a.stream >= 0:
        h.set_option("stream", a.stream)
    if a.variant >= 0:
        h.set_option("variant", a.variant)
    exchange = "nccl" if world > 1 else "none"
    if world > 1 and a.exchange == "push":
        # peer-memory exchange needs CUDA IPC + NVLink peer access on every rank; if any rank cannot set it up,
        # all ranks stay on the NCCL all-gather (still a GPU path of this library, not a fallback to anything else)
        ok = 1
        try:
            blobs = [None] * world
            dist.all_gather_object(blobs, h.ipc_export())
            h.ipc_import(blobs)
        except Exception as e:
            sys.stderr.write("rank %d: push exchange unavailable (%s)\n" % (rank, e))
            ok = 0
        t = torch.tensor([ok], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 1:
            h.set_option("exchange", 1)
            exchange = "push"
    if a.fuse >= 0:
        h.set_option("fuse", a.fuse)
    if a.overlap >= 0:
        h.set_option("overlap", a.overlap)
    h.set_option("timing", 1)      # per-kernel CUDA events on the launching stream (roofline)
    h.upload(host)

    peaks, peaks_src = measured_peaks()
    probe = h.probe_fp32_peak()
    flush = None if a.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    energy0 = None
    if not a.no_energy:
        ke, pe = h.energy(); energy0 = ke + pe

    # ---- parity (outside the timed region; the oracle is the checker, never the thing measured) --------
    # accelerations of the uploaded state on sampled bodies against the FP64 oracle (rank 0), and, when sharded,
    # against a single-GPU handle on rank 0's device for ALL bodies; fails loudly above the north_star tolerance
    parity = None
    if not a.no_parity:
        import oracle_lib as orc
        acc = h.accel()                                    # collective when sharded: every rank calls it
        if rank == 0:
            tol = 1e-5 if prec == nb.F32 else 1e-12
            ns, errs = min(64, n), []
            for i0 in sorted({0, max(0, n // 3 - ns // 2), max(0, (2 * n) // 3 - ns // 2), n - ns}):
                ref = orc.accel_f64_from_f32(init, i0, i0 + ns) if prec == nb.F32 else orc.accel_f64(init, i0, i0 + ns)
                errs.append(orc.rel_err(acc[i0:i0 + ns], ref))
            errs = np.concatenate(errs)
            parity = {"max_rel_err": float(errs.max()), "n_sample": int(len(errs)), "tolerance": tol, "vs": "FP64 oracle, same inputs",
                      "vs_single_gpu_max": None}
            if world > 1:
                with nb.NBody(n, prec, rank=0, world=1, device=local_rank) as h1:
                    h1.upload(init)
                    parity["vs_single_gpu_max"] = float(orc.rel_err(acc, h1.accel()).max())
            if not parity["max_rel_err"] <= tol or (parity["vs_single_gpu_max"] is not None and not parity["vs_single_gpu_max"] <= 2 * tol):
                raise SystemExit("bench.py: PARITY FAILURE %s" % json.dumps(parity))
        barrier()

    # ---- warm-up ---------------------------------------------------------------------------------
    for _ in range(a.warmup):
        h.step(DT, 1)
    barrier()

    # ---- timed region: K steps, device-timed, L2 flushed in between, max over ranks ---------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    h.timing_reset()
    step_ms, own_ms = [], []
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        if flush is not None:
            flush.zero_()
        barrier()
        h.step(DT, 1)
        barrier()
        own_ms.append(h.last_step_ms())
        step_ms.append(allmax(own_ms[-1]))
    wall1 = time.perf_counter()
    tim = h.timing()
    clocks = sampler.stop() if sampler else None
    # every rank's own device time per step and the time of its force launches (diagnosis of the max over ranks)
    by_rank = None
    if dist is not None:
        t = torch.tensor([sum(own_ms) / a.steps, tim["force_ms"] / a.steps], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        by_rank = {"step_ms": [round(float(x[0]), 3) for x in allt], "force_ms": [round(float(x[1]), 3) for x in allt]}
    total_ms = sum(step_ms)
    ms_per_step = total_ms / a.steps
    value = n * float(n) * a.steps / (total_ms * 1e-3) / 1e9

    # ---- back-to-back variant: one call, all-gather overlapped with the next step's local pass ----
    barrier()
    h.step(DT, a.steps)
    b2b_ms = allmax(h.last_step_ms())
    value_b2b = n * float(n) * a.steps / (b2b_ms * 1e-3) / 1e9

    energy1 = None
    if not a.no_energy:
        ke, pe = h.energy(); energy1 = ke + pe

    # ---- e2e: host buffers through the reference-facing calls --------------------------------------
    host[:] = init
    nbytes = n * dtype.itemsize
    e2e_steps = a.steps
    if world == 1:
        nb.bodyForce(host, DT); nb.integrate(host, DT)          # warm the drop-in handle
        host[:] = init
        barrier(); t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nb.bodyForce(host, DT)
            nb.integrate(host, DT)
        barrier(); t1 = time.perf_counter()
        h2d, d2h = 2 * nbytes, 2 * nbytes
        e2e_each = None
        e2e_api = "bodyForce(Body*,dt,n) + integrate(Body*,dt,n) on a pinned host array"
    else:
        i0, i1 = h.info("i_begin"), h.info("i_end")
        mine = host[i0:i1]                                     # this rank's slice of the pinned host array
        h.upload(host); h.step(DT, 1); h.download_local(mine)  # warm the path
        host[:] = init
        e2e_each = []
        barrier(); t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            # every rank reads its slice of the inputs over PCIe (positions all-gathered on the device), steps,
            # and reads its slice of the result back: the whole job moves the array once each way per step
            h.upload(host); h.step(DT, 1); h.download_local(mine)
            e2e_each.append(round((time.perf_counter() - ts) * 1e3, 3))
        barrier(); t1 = time.perf_counter()
        h2d, d2h = nbytes, nbytes
        e2e_api = "nbody_upload + nbody_step + nbody_download_local on a pinned host array: every rank moves its own slice (1/%d of the bytes) each way" % world
    e2e_s = allmax(t1 - t0)
    e2e_val = n * float(n) * e2e_steps / e2e_s / 1e9

    # ---- roofline of the dominant kernel (force pass of rank 0) ------------------------------------
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    sms = h.info("sms")
    peak_tflops = sms * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    n_local = h.info("local_blocks") * 128
    force_ms_step = tim["force_ms"] / a.steps
    flops_per_step_rank = FLOP_PER_INTERACTION * float(min(n_local, n)) * n
    achieved = flops_per_step_rank / (force_ms_step * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp) and world == 1 and n == N_DEFAULT and prec == nb.F32:
        try:
            traffic = json.load(open(tp)).get("force_f32_dram_bytes_per_launch")
        except Exception:
            traffic = None
    fused = bool(h.info("fuse")) or bool(h.info("stream"))
    launches_per_step = tim["launches"] / float(a.steps)
    force_launches_per_step = 1 if (world == 1 or fused) else 2
    roofline = {
        "kernel": "%s (%s)" % ("force_stream_f%s_kernel" % ("32" if prec == nb.F32 else "64") if h.info("stream") else "force_f%s_kernel" % ("32" if prec == nb.F32 else "64"),
                               "FFMA2+MUFU.RSQ" if prec == nb.F32 else "DFMA+MUFU.RSQ64H"),
        "bound": "fp32" if prec == nb.F32 else "fp64", "achieved": achieved,
        "peak": peak_tflops if prec == nb.F32 else peak_tflops / 2, "unit": "TFLOP/s",
        "frac": achieved / (peak_tflops if prec == nb.F32 else peak_tflops / 2), "traffic": traffic,
        "peak_source": "SMs(%d) x 128 FP32 lanes x 2 x sm_max_mhz(%.0f, MEASURED_PEAKS.json %s); MEASURED_PEAKS holds no FP32 figure, so the "
                       "in-run FFMA2 probe is given beside it" % (sms, sm_max_mhz, peaks_src),
        "peak_measured_ffma2_tflops": probe["ffma_lane_ops_per_s"] * 2 / 1e12,
        "frac_of_measured_ffma2": achieved / (probe["ffma_lane_ops_per_s"] * 2 / 1e12) if prec == nb.F32 else None,
        "sm_clock_mhz_probe": probe["sm_clock_mhz"],
        "algorithmic": "20 flop x N_local x N interactions per step; %d force launch(es) per step, avg %.3f ms per step on rank 0; the integrate step "
                       "(48 B/body: read + write pos, vel) %s" % (force_launches_per_step, force_ms_step,
                       "is this kernel's epilogue: accelerations and partial sums never reach HBM" if fused else "runs as integrate_kernel behind it"),
        "algorithmic_bytes_per_launch": float(n) * 12 + float(min(n_local, n)) * 48 if fused else None,
        "cycles_per_interaction_per_lane": sms * 128 * (clocks["sm_mhz"] if clocks and clocks["sm_mhz"] else sm_max_mhz) * 1e6
                                           / (float(min(n_local, n)) * n / (force_ms_step * 1e-3)),
    }
    slots = h.info("slots")
    es = 4 if prec == nb.F32 else 8
    if fused and world == 1:
        # no integrate kernel in the step: the HBM-bound kernel left on the path is the standalone drift (integrate(): x += dt*v),
        # timed here on its own -- 36 B/body algorithmic (read pos, vel; write pos)
        h.timing_reset()
        reps = 20
        for _ in range(reps):
            if flush is not None:
                flush.zero_()
                torch.cuda.synchronize()               # the flush runs on torch's stream, the kernel on the library's
            h.integrate(0.0)
        t2 = h.timing()
        drift_ms = t2["integrate_ms"] / reps
        drift_bytes = float(min(n_local, n)) * 9 * es
        roofline_integrate = {
            "kernel": "integrate_kernel (drift only; in a step the integrate is fused into the force kernel's epilogue)", "bound": "hbm",
            "achieved": drift_bytes / (drift_ms * 1e-3) / 1e9 if drift_ms > 0 else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": (drift_bytes / (drift_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if drift_ms > 0 else None,
            "algorithmic": "per body: read pos + vel, write pos (%d B); L2 flushed before each launch" % (9 * es),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (%s)" % peaks_src, "ms": drift_ms, "in_step": "fused",
        }
    elif fused:
        roofline_integrate = {"kernel": "none in the step (integrate fused into the force kernel's epilogue)", "in_step": "fused"}
    else:
        integ_ms_step = tim["integrate_ms"] / a.steps
        integ_bytes = float(min(n_local, n)) * (12 * es + 3 * es * slots)
        roofline_integrate = {
            "kernel": "integrate_kernel", "bound": "hbm", "achieved": integ_bytes / (integ_ms_step * 1e-3) / 1e9 if integ_ms_step > 0 else None,
            "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": (integ_bytes / (integ_ms_step * 1e-3) / 1e9) / peaks["hbm_gbs"] if integ_ms_step > 0 else None,
            "algorithmic": "per body: read pos+vel, write pos+vel (%d B) + %d partial slots x %d B (the slots are NOT algorithmic: 60 B/body is)" % (12 * es, slots, 3 * es),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (%s)" % peaks_src, "ms": integ_ms_step,
        }

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------------
    cpu = None
    if world == 1 and not a.no_cpu_baseline and prec == nb.F32:
        try:
            _, cpu = cpu_reference_leg(n, seconds=a.cpu_seconds if a.cpu_seconds else 12.0)
        except Exception as e:  # the oracle is a checker, not the product: report but do not fail
            cpu = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "billion_interactions_per_s", "value": value, "unit": "G interactions/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": a.precision, "data": "synthetic",
            "config": {
                "workload": workload_string(n, a.precision),
                "parallelism": ("i-sharded x%d, positions replicated, %s" % (world, "the force kernel's integrate epilogue pushes every finished tile into every peer's next-step buffer over NVLink (peer memory + flag); the peers' flags are acquired inside the next force kernel, own j-slice first" if exchange == "push" else "NCCL all-gather per step overlapped with the local-j force pass")) if world > 1 else "single GPU",
                "exchange": exchange,
                "l2": "flushed between timed steps (256 MiB memset)" if flush is not None else "not flushed",
                "force_variant": h.info("variant"), "tile_bodies": h.info("tile_bodies"), "splits_local": h.info("splits_local"),
                "splits_remote": h.info("splits_remote"), "ctas_per_sm": h.info("ctas_per_sm"),
                "reduction": ("stream-K: last-arriver fixed-order reduction of cut tiles + integrate in the force kernel" if h.info("stream") else
                              "last-arriver fixed-order reduction of the j-split slots (L2-resident ring of %d tiles, %s) + integrate in the force kernel"
                              % (h.info("ring"), "split-major inside groups of ring/2 tiles" if h.info("order") else "split-major over all tiles")) if fused else "slot array in HBM + integrate_kernel",
                "launches_per_step": launches_per_step, "stream_grid": h.info("grid"), "stream_phases": h.info("phases"),
            },
            "tflops_20flop": value * FLOP_PER_INTERACTION / 1e3,
            "frac_fp32_peak": value * FLOP_PER_INTERACTION / 1e3 / (peak_tflops * world) if prec == nb.F32 else None,
            "value_back_to_back": value_b2b, "ms_per_step_back_to_back": b2b_ms / a.steps,
            "wall_s_timed_region": wall1 - wall0, "by_rank": by_rank,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "G interactions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": e2e_api,
                    "ms_per_step": e2e_s * 1e3 / e2e_steps, "ms_each_rank0": e2e_each},
            "gpu_launches": tim["launches"],
            "roofline": roofline, "roofline_integrate": roofline_integrate, "parity": parity,
            "cpu_baseline": cpu,
            "energy": {"e0": energy0, "e1": energy1, "steps": 2 * a.steps + a.warmup,
                       "rel_drift": (energy1 - energy0) / abs(energy0) if energy0 else None,
                       "note": "softening 1e-9 makes close encounters unresolved at dt=0.01: reported, never gated on"} if energy0 is not None else None,
        }
        emit(line)
    h.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
