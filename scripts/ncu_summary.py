"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit.startswith("n") else v * 1000 if unit.startswith("m") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("gtav::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us (cold-cache, serialised)")
    print(f"{'us':>10} {'share':>6} {'n':>5} {'avg us':>8}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d} {v[1] / v[0]:8.2f}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
