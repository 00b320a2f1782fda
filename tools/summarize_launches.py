"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (second half = the warmed-up pass)."""
import collections
import csv
import re
import sys


def main(path, second_half=True):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if l.startswith('"')))
    if second_half:
        rows = rows[len(rows) // 2:]
    agg = collections.OrderedDict()
    for x in rows:
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1000 if x["Metric Unit"] == "ns" else (v * 1000 if x["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0, x["Grid Size"], x["Block Size"]])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"{'us':>10} {'share':>6} {'n':>4}  kernel  (last grid, block)")
    for name, a in agg.items():
        print(f"{a[1]:10.1f} {100 * a[1] / total:5.1f}% {a[0]:4d}  {name[:80]}  {a[2]} {a[3]}")
    print(f"{total:10.1f} total")


if __name__ == "__main__":
    main(sys.argv[1], "--all" not in sys.argv)
