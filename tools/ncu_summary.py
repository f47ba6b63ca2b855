#!/usr/bin/env python
"""Summarise .ncu-rep files (raw page) into a markdown table: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue % (active)"), ("smsp__inst_executed.sum", "warp instr"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("lts__t_bytes.sum", "L2 bytes")]
print("| report | kernel | " + " | ".join(k[1] for k in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        name = name.replace("void ", "").split("(")[0][-70:]
        vals = []
        for k, _ in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                    vals.append(("%.4g" % v) + " " + units[i])
                except ValueError:
                    vals.append(r[i])
            else:
                vals.append("-")
        print("| %s | `%s` | %s |" % (rep.split("/")[-1], name, " | ".join(vals)))
