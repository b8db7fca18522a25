"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python profiles/summarize_launches.py gpurun_out/launches_r1.csv > profiles/launches_r1_summary.md
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        rows.append((name, ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        agg[name][0] += 1
        agg[name][1] += ns
    print(f"# ncu launch list summary: {path}\n")
    print(f"{len(rows)} launches, {tot/1e6:.1f} ms total (cold-cache, serialised: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n} | {ns/1e6:.2f} | {100*ns/tot:.1f}% | {ns/n/1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
