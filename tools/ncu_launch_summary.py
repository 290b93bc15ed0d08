"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel
(count, mean/min microseconds, share of the total)."""
import collections
import csv
import sys


def main(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(row["Metric Unit"], 1.0)
        agg.setdefault(row["Kernel Name"].split("(")[0][:64], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':64s} {'n':>5s} {'mean_us':>9s} {'min_us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:64s} {len(v):5d} {sum(v)/len(v):9.2f} {min(v):9.2f} {100*sum(v)/tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
