"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py <rep> [out.txt]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
out = []
for w in want:
    if w in hdr:
        i = hdr.index(w)
        out.append(f"{w} [{units[i]}]: " + " | ".join(r[i] for r in data))
out.append("-- warp stall reasons (cycles per issued instruction), first launch --")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") or (h.startswith("smsp__average_warp") and h.endswith(".ratio")):
        out.append(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '')}: {data[0][i]}")
text = "\n".join(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
try:
    print(text)
except BrokenPipeError:
    pass
