"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (per kernel: count, total, average, share)."""
import collections
import csv
import sys


def main(path, window=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        rows.append((row["Kernel Name"], v, row.get("Grid Size"), row.get("Block Size")))
    if window:
        rows = rows[: int(window)]
    agg = collections.OrderedDict()
    for name, v, g, b in rows:
        a = agg.setdefault(name.split("(")[0][:64], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':66s} {'n':>5s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:66s} {n:5d} {t:10.1f} {t / n:8.2f} {t / tot:6.3f}")


if __name__ == "__main__":
    main(*sys.argv[1:])
