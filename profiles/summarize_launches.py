"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / share."""
import collections
import csv
import re
import sys


def main(path, top=25):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit == "ns" else (v * 1000.0 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:58]
        agg[(name, row.get("Grid Size", ""), row.get("Block Size", ""))].append(v)
    total = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    print("# %s: %d launches, %.1f us total (cold-cache, serialised: compare shares)" % (path, n, total))
    print("%-58s %-14s %5s %9s %10s %6s" % ("kernel", "grid", "n", "avg_us", "sum_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
        print("%-58s %-14s %5d %9.2f %10.1f %5.1f%%" % (k[0], k[1], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / total))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
