"""Summarise an ncu report (--set full) into text: per-kernel key metrics + stall breakdown."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("kernel:", r[idx["Kernel Name"]])
    for k in keys:
        if k in idx:
            print(f"  {k:70s} {r[idx[k]]} {units[idx[k]]}")
    st = [(float(r[idx[h]].replace(",", "") or 0), h) for h in hdr
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[idx[h]] not in ("", "n/a")]
    print("  warp stall reasons (warps stalled per issue-active cycle):")
    for v, h in sorted(st, reverse=True)[:7]:
        print(f"    {v:7.3f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
