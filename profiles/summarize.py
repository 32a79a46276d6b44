"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: summarize.py <csv> [--seq N]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        seq.append((name, v, row["Grid Size"]))
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0
    for n, v, _ in seq:
        agg[n][0] += 1
        agg[n][1] += v
        tot += v
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {n} | {t:.0f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    print(f"\ntotal {tot / 1e3:.2f} ms")
    if "--seq" in sys.argv:
        n = int(sys.argv[sys.argv.index("--seq") + 1])
        for i, s in enumerate(seq[:n]):
            print(i, s)
