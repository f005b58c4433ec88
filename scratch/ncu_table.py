"""ncu -i report --page raw --csv -> markdown table of the metrics the roofline discussion uses.
usage: python scratch/ncu_table.py report.ncu-rep [title]"""
import csv, subprocess, sys, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
M = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
     "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
kern = []
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]); name = name.replace("void ", "").replace("xeq::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    kern.append((name[:44], r))
print("| metric | " + " | ".join(k for k, _ in kern) + " |")
print("|---|" + "---|" * len(kern))
for m in M:
    if m not in ix: continue
    vals = []
    for _, r in kern:
        v = r[ix[m]]
        try: v = f"{float(v.replace(',', '')):.4g}"
        except ValueError: pass
        vals.append(v)
    print(f"| {m} [{units[ix[m]]}] | " + " | ".join(vals) + " |")
