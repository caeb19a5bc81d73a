#!/usr/bin/env python
"""per-source-line stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name K ...`"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
out = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 7 and r[0] not in ("", "Line No"):
        try:
            s = int(r[4])
            inst = int(r[7])
        except ValueError:
            continue
        out.append((s, inst, cur_file, r[0], r[1].strip()))
tot = sum(o[0] for o in out)
print("total samples", tot)
for s, inst, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%6d %5.1f%% %10d  %s:%s  %s" % (s, 100.0 * s / max(tot, 1), inst, f, ln, src[:110]))
