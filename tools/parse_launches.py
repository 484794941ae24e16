"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)
            if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        k = re.sub(r"\(.*", "", k)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%d launches, %.1f us total (cold-cache, serialised: compare shares)" % (len(rows), tot / 1e3))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-60s n=%4d total=%9.1f us avg=%8.1f us share=%5.1f%%" % (k[:60], n, t / 1e3, t / 1e3 / n, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
