"""Print selected metrics of every kernel in an ncu report (raw page).
    python tools/ncu_metrics.py X.ncu-rep [regex ...]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or [r"gpu__time_duration.sum$", r"dram__bytes_(read|write).sum$", r"launch__registers_per_thread$",
                        r"launch__occupancy_limit", r"sm__warps_active.avg.pct_of_peak_sustained_active", r"smsp__issue_active.avg.pct",
                        r"smsp__inst_executed.sum$", r"issue_stalled.*per_issue_active", r"sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|lsu|xu|fp64|uniform|adu|cbu)\.avg.pct_of_peak_sustained_active",
                        r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$", r"smsp__thread_inst_executed.sum$",
                        r"sm__throughput.avg.pct", r"dram__throughput.avg.pct", r"lts__t_sector_hit_rate.pct", r"sm__cycles_elapsed.max$",
                        r"l1tex__throughput.avg.pct", r"lts__throughput.avg.pct", r"sm__inst_executed_pipe_.*pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("==", name[:100])
    for h, u, v in zip(hdr, units, r):
        if any(re.search(p, h) for p in pats) and v not in ("", "0"):
            print(f"  {h:90s} {v:>16s} {u}")
