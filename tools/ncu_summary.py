"""Headline counters of every kernel in an .ncu-rep (raw page): python tools/ncu_summary.py file.ncu-rep [out.csv]"""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum",
        "lts__t_sectors_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "sm__inst_executed_pipe_xu.sum"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name"))
    for k in want:
        if k in d:
            print("   %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[h]) for h in hdr if "issue_stalled_" in h and h.endswith("per_issue_active.ratio") and d[h]}
    print("   stalls per issue:", ", ".join("%s %.2f" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
