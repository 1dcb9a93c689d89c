"""Developer tool: the metrics the profiles/ summaries quote, one column per kernel launch, from an .ncu-rep (--set full).
usage: python tools/ncu_table.py REPORT.ncu-rep [KERNEL_REGEX] > table.md"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
kern = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
launches = [r for r in rows[2:] if len(r) == len(hdr) and (kern is None or kern.search(r[ik]))]
names = []
for r in launches:
    n = r[ik].split("(")[0].replace("void ", "").replace("lab::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    names.append(n)
print("| metric | " + " | ".join(names) + " |")
print("|---|" + "---:|" * len(names))
for m in WANT:
    if m not in hdr:
        continue
    i = hdr.index(m)
    vals = []
    for r in launches:
        v = r[i].replace(",", "")
        try:
            f = float(v)
            vals.append(f"{f:.0f}" if abs(f) >= 1000 or f == int(f) else f"{f:.2f}")
        except ValueError:
            vals.append(v)
    print(f"| {m} [{units[i]}] | " + " | ".join(vals) + " |")
