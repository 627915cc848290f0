"""Per-launch DRAM traffic of one triangular-solve sweep from
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file X.csv ...
Prints the table and, with --json NAME, the entry for profiles/traffic.json.
    python tools/summarize_dram.py X.csv [--skip N] [--json cfg2]
"""
import csv
import gzip
import json
import sys

path = sys.argv[1]
skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
op = gzip.open if path.endswith(".gz") else open
rows = [r for r in csv.reader(op(path, "rt")) if len(r) > 10 and r[0].isdigit()]
launches = {}
for r in rows:
    d = launches.setdefault(int(r[0]), {"kernel": r[4]})
    val = float(r[-1].replace(",", ""))
    unit = r[-2]
    if r[-3] == "gpu__time_duration.sum":
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    else:
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r[-3]] = val
ids = sorted(launches)[skip:]
print(f"# {path}: launches {ids[0]}..{ids[-1]} (cold-cache, serialised; compare shares)")
print("launch kernel                    dram MB        us     GB/s")
tr = tw = tt = 0.0
for i in ids:
    d = launches[i]
    name = d["kernel"].replace("void ", "").replace("lsa::", "").split("<")[0].split("(")[0]
    rd, wr, us = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0), d.get("gpu__time_duration.sum", 0.0)
    tr, tw, tt = tr + rd, tw + wr, tt + us
    print(f"{i:6d} {name[:22]:22s} {(rd + wr) / 1e6:10.2f} {us:9.2f} {(rd + wr) / 1e3 / max(us, 1e-9):8.0f}")
print(f"# total: read {tr / 1e9:.3f} GB, write {tw / 1e9:.3f} GB, {tt:.1f} us over {len(ids)} launches")
if "--json" in sys.argv:
    print(json.dumps({sys.argv[sys.argv.index("--json") + 1]: {"dram_read_bytes": tr, "dram_write_bytes": tw, "launches": len(ids)}}, indent=1))
