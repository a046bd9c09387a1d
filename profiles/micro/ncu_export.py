# Writes the counters bench.py and DESIGN.md quote from an .ncu-rep (ncu --set full, one kernel) as JSON {metric: [value, unit]}:
#   python profiles/micro/ncu_export.py REPORT.ncu-rep OUT.json
import csv, json, subprocess, sys
WANT = ("gpu__time_duration.sum", "launch__", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_", "l1tex__data_pipe_lsu_wavefronts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ldgsts_cache_access.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__average_warps_issue_stalled", "smsp__warps_issue_stalled")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
d = {"kernel": [r[hdr.index("Kernel Name")], ""]}
for i, h in enumerate(hdr):
    if any(w in h for w in WANT) and ".peak_sustained" not in h and ".per_second" not in h and not h.endswith((".min", ".max")) or h == "sm__cycles_elapsed.max":
        d[h] = [r[i], units[i]]
json.dump(d, open(sys.argv[2], "w"), indent=1, sort_keys=True)
print(len(d), "metrics ->", sys.argv[2])
