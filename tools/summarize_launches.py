"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares, not absolutes)."""
import collections
import csv
import re
import sys


def main(path: str) -> None:
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r["Kernel Name"])).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {total / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':40s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in tot.most_common():
        print(f"{k:40s} {cnt[k]:8d} {v / 1e6:10.3f} {100 * v / total:6.1f}% {v / cnt[k] / 1e3:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
