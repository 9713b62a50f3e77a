"""Key metrics of every kernel launch in an .ncu-rep (raw page), one line per metric.  usage: ncu_summary.py rep [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum",
        "smsp__inst_executed_op_branch.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
stalls = [c for c in h if "average_warps_issue_stalled" in c and "per_issue_active" in c and "not_issued" not in c]
ki = h.index("Kernel Name")
for r in rows[2:]:
    if len(r) != len(h) or (pat and not pat.search(r[ki])):
        continue
    print("==", r[ki][:70])
    for w in want:
        if w in h:
            print("  %-70s %s" % (w, r[h.index(w)]))
    st = sorted(((float(r[h.index(c)]), c.split("issue_stalled_")[1].split("_per_issue")[0]) for c in stalls if r[h.index(c)]), reverse=True)
    print("  stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:9]))
