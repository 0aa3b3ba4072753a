#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): tools/launch_shares.py launches.csv [top_n]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        k = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} ms of kernel time (cold-cache, serialised under ncu)\n")
    print("| kernel | launches | total ms | share | mean ms |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"| `{k[:100]}` | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f} % | {a[1] / a[0]:.3f} |")


if __name__ == "__main__":
    main()
