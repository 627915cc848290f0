"""Per-launch table from an `ncu -i X.ncu-rep --page raw --csv` export (selected metrics).
    python tools/summarize_ncu_raw.py raw.csv > profiles/NAME_summary.txt
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("launch__cluster_size", "cluster"),
        ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"), ("launch__registers_per_thread", "regs"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"), ("lts__t_sector_hit_rate.pct", "l2hit_%")]
idx = [(hdr.index(k), lab) for k, lab in want if k in hdr]
print(f"# {sys.argv[1]}: {len(data)} launches (ncu --set full, --clock-control none; cold caches, serialised)")
print(" ".join(lab.rjust(10) if i else lab.ljust(18) for i, (_, lab) in enumerate(idx)))
print(" ".join((units[j] or "-").rjust(10) if i else "".ljust(18) for i, (j, _) in enumerate(idx)))
tot = 0.0
for r in data:
    name = r[idx[0][0]].replace("void ", "").replace("lsa::", "").split("<")[0].split("(")[0]
    cells = [name[:18].ljust(18)]
    for j, lab in idx[1:]:
        v = r[j].replace(" ", "")
        try:
            f = float(v)
            v = f"{f:.1f}" if lab != "regs" and lab != "cluster" else f"{int(f)}"
            if lab == "us":
                tot += f
        except ValueError:
            pass
        cells.append(v[:10].rjust(10))
    print(" ".join(cells))
print(f"# total {tot:.1f} us")
