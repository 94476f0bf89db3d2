"""one line per kernel (its LAST captured launch) of an ncu report: python scripts/ncu_table.py rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("lts__t_sector_hit_rate.pct", "l2_hit"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_act"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act"),
        ("launch__registers_per_thread", "regs"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier")]
last = {}
for r in rows[2:]:
    last[r[hdr.index("Kernel Name")]] = r
print(" | ".join(c[1] for c in cols))
for name, r in last.items():
    cells = []
    for k, _ in cols:
        i = hdr.index(k)
        v = r[i]
        if k == "Kernel Name":
            v = v.split("(")[0][:48]
        elif units[i] not in ("", "register/thread"):
            v = f"{v} {units[i]}"
        cells.append(v)
    print(" | ".join(cells))
