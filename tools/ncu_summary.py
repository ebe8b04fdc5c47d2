"""Compact summary of an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py rep [kernel-idx]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = [
    "Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
]
for ri, vals in enumerate(rows[2:]):
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print("## launch %d" % ri)
    for k in WANT:
        if k in d:
            print("| %s | %s | %s |" % (k, d[k], u[k]))
    st = [(float(v.replace(",", "")), h) for h, v in d.items() if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and v not in ("", "n/a")]
    if not st:
        st = [(float(v.replace(",", "")), h) for h, v in d.items() if "warp_issue_stalled" in h and "per_warp_active" in h and v not in ("", "n/a")]
    for v, h in sorted(st, reverse=True)[:10]:
        print("| %s | %.4f | |" % (h, v))
