"""Per-kernel (and per-grid) totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_summary.py gpurun_out/launches.csv [--by-grid] [--filter substr]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    by_grid = "--by-grid" in sys.argv
    flt = sys.argv[sys.argv.index("--filter") + 1] if "--filter" in sys.argv else None
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void vidil::<unnamed>::", "").replace("vidil::<unnamed>::", "")
        if flt and flt not in name:
            continue
        key = (name, row["Grid Size"]) if by_grid else (name, "")
        agg[key][0] += 1
        agg[key][1] += v
        total += v
    print(f"total {total / 1e3:.2f} ms over {sum(c for c, _ in agg.values())} launches")
    for (name, grid), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{t / 1e3:9.2f} ms {100 * t / total:5.1f}% {c:6d} x {t / c:9.1f} us  {name[:70]} {grid}")


if __name__ == "__main__":
    main()
