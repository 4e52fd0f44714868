#!/usr/bin/env python
"""Hot-code footprint of a kernel from an `ncu --set full --import-source on` report: how many static SASS
instructions cover 90 / 99 % of the executed warp instructions, and the execution profile along the address space
(the I-cache levels are ~6 KB L0 / 32 KB L1.5: a hot loop beyond that shows up as the `no_instruction` stall).
usage: footprint.py <report.ncu-rep> [bucket=64]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ie, ism, ith = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
data = [(int(r[ie]), int(r[ism]), float(r[ith] or 0)) for r in rows[2:] if len(r) > ie and r[ia].startswith("0x")]
tot = sum(d[0] for d in data); smp = sum(d[1] for d in data)
print("static instructions %d (%.1f KB), executed %.3e" % (len(data), len(data)*16/1024, tot))
s = sorted((d[0] for d in data), reverse=True)
for frac in (0.90, 0.99):
    acc = 0
    for k, e in enumerate(s):
        acc += e
        if acc >= frac*tot:
            print("%2.0f %% of executed instructions come from %d static instructions = %.1f KB" % (100*frac, k + 1, (k + 1)*16/1024))
            break
for i in range(0, len(data), B):
    c = data[i:i + B]
    e = sum(d[0] for d in c)
    if e/tot < 0.002:
        continue
    th = sum(d[2]*d[0] for d in c)/max(e, 1)
    print("%5d-%5d  exec %5.2f%%  samples %5.2f%%  threads %4.1f" % (i, i + B, 100*e/tot, 100*sum(d[1] for d in c)/smp, th))
