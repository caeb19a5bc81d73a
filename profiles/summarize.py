#!/usr/bin/env python
"""Fold an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
usage: python profiles/summarize.py gpurun_out/launches_X.csv [bench_X.json] > profiles/rNN_X_launches.md"""
import collections
import csv
import json
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[i_val].replace(",", ""))
        except ValueError:
            continue
        a = acc.setdefault(r[i_name].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in acc.values())
    print("# ncu launch list: %s" % sys.argv[1])
    print()
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches; shares are what counts).")
    print("%d launches captured, %.3f ms of kernel time in total." % (sum(a[0] for a in acc.values()), tot / 1e6))
    print()
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for n, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f%% |" % (n, a[0], a[1] / 1e6, 100 * a[1] / tot))
    if len(sys.argv) > 2:
        d = json.load(open(sys.argv[2]))
        print()
        print("## bench.py line of the same build (not under ncu)")
        print()
        print("```json")
        print(json.dumps(d, indent=1))
        print("```")


if __name__ == "__main__":
    main()
