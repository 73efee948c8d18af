#!/usr/bin/env python
"""Per-instruction view of an ncu report's source page: SASS with stall samples and executed counts.
usage: python profiles/sass_hot.py rep.ncu-rep [kernel-index] [--all]"""
import csv, io, subprocess, sys
path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
idx = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
blk = blocks[1 + idx]
lines = blk.split("\n")
print("kernel:", lines[0][:150])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
S, X = ci["# Samples"], ci["Instructions Executed"]
tot_s = sum(int(r[S]) for r in rows[1:] if len(r) > S and r[S].isdigit())
tot_x = sum(int(r[X]) for r in rows[1:] if len(r) > X and r[X].isdigit())
print("total samples %d, warp-instructions %d, SASS lines %d" % (tot_s, tot_x, len(rows) - 1))
show_all = "--all" in sys.argv
for n, r in enumerate(rows[1:]):
    if len(r) <= X or not r[S].isdigit():
        continue
    s, x = int(r[S]), int(r[X])
    if show_all or s >= 0.01 * tot_s:
        print("%4d %6.2f%% x%-9d %s" % (n, 100.0 * s / max(tot_s, 1), x, r[ci["Source"]].strip()[:110]))
