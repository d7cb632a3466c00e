"""Key counters of every kernel in an .ncu-rep (ncu --set full) as a text table.  usage: ncu_summary.py rep [title]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers / thread"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
        ("smsp__inst_executed.sum", "warp instructions"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue")]
idx = {h: i for i, h in enumerate(hdr)}
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else rep))
print("# ncu -i %s --page raw --csv" % rep.split("/")[-1])
names = [r[idx["Kernel Name"]][:60] for r in rows[2:]]
print("%-42s" % "metric" + "".join(" | %-40s" % n[:40] for n in names))
for key, label in want:
    if key not in idx: continue
    print("%-42s" % label + "".join(" | %-40s" % ("%s %s" % (r[idx[key]], units[idx[key]])) for r in rows[2:]))
