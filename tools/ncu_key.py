"""Key metrics of every launch in an ncu report:  python tools/ncu_key.py report.ncu-rep [units-per-launch]"""
import csv, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh, un = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum"]
stalls = [w for w in hh if "issue_stalled" in w and "per_issue_active" in w and "not_issued" not in w]
for r in rr[2:]:
    print("##", r[hh.index("Kernel Name")][:120])
    for w in want:
        if w in hh:
            v = r[hh.index(w)]
            extra = ""
            if units and w.endswith(".sum") and "time" not in w and "bytes" not in w:
                try:
                    extra = "   (%.1f per unit)" % (float(v.replace(",", "")) / units)
                except ValueError:
                    pass
            print("  %-82s %14s %s%s" % (w, v, un[hh.index(w)], extra))
    st = sorted(((float(r[hh.index(w)] or 0), w) for w in stalls), reverse=True)[:9]
    for v, w in st:
        print("  stall %-40s %.3f" % (w.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
