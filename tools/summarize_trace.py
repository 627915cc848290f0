"""Per-kernel and per-level totals of an LSA_TRACE=1 factor / sweep trace (stderr of tools/trace_solve.py)."""
import collections
import re
import sys

tot = collections.defaultdict(float)
lvl = collections.defaultdict(lambda: collections.defaultdict(float))
for line in open(sys.argv[1]):
    m = re.match(r"TRACE level\s+(\d+)\s+(\w+)\s+j0\s+(\d+)\s+grid\s+(\d+) x\s+(\d+)\s+([\d.]+) us", line)
    if m:
        tot[m[2]] += float(m[6])
        lvl[int(m[1])][m[2]] += float(m[6])
print({k: round(v / 1000, 1) for k, v in tot.items()}, round(sum(tot.values()) / 1000, 1))
for L in sorted(lvl):
    print(L, round(sum(lvl[L].values()) / 1000, 2), {k: round(v / 1000, 2) for k, v in lvl[L].items()})
