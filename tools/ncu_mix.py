#!/usr/bin/env python3
"""Dynamic instruction mix of one kernel from an ncu report captured with --import-source on.
usage: ncu_mix.py report.ncu-rep kernel-regex"""
import collections
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot, smp, n, first = collections.Counter(), collections.Counter(), 0, None
for r in rows:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    op = r[iS].strip().split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.split(".")[0]
    e = int(r[iE])
    if first is None:
        first = e
    tot[o] += e
    smp[o] += int(r[iSt]) if r[iSt].isdigit() else 0
    n += e
print("warps", first, "instr/warp %.1f" % (n / first))
for o, c in tot.most_common(30):
    print("%-8s %8.1f per warp  %5.1f%%  stall samples %d" % (o, c / first, 100 * c / n, smp[o]))
