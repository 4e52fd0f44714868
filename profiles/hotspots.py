#!/usr/bin/env python
"""Joins an ncu SASS source page with nvdisasm line info and prints stall samples per source line.

usage: hotspots.py <report.ncu-rep> <object-or-so with the kernel> <kernel mangled-name substring> [top]
Needs the kernel compiled with -lineinfo.  Runs where there is no GPU (reads the report only).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, obj, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    lines = {}
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "-g", cb], capture_output=True, text=True).stdout.splitlines()
        in_fn, cur = False, ("?", 0)
        for ln in txt:
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                in_fn = kname in m.group(1)
                continue
            if not in_fn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
            if m:
                lines[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    base = None
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[hdr_i + 1:]:
        if len(r) <= ci["Thread Instructions Executed"]:
            continue
        addr = int(r[ci["Address"]], 16)
        base = addr if base is None else base
        key = lines.get(addr - base, ("?", 0))
        v = [int(float(r[ci["# Samples"]] or 0)), int(float(r[ci["Instructions Executed"]] or 0)), int(float(r[ci["Thread Instructions Executed"]] or 0))]
        for k in range(3):
            agg[key][k] += v[k]; tot[k] += v[k]
    print("total samples %d, warp instructions %d, avg active threads %.2f" % (tot[0], tot[1], tot[2]/max(tot[1], 1)))
    byfile = defaultdict(lambda: [0, 0, 0])
    for (f, l), v in agg.items():
        for k in range(3):
            byfile[f][k] += v[k]
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  %-22s samples %5.1f%%  instr %5.1f%%  active threads %.1f" % (f, 100*v[0]/tot[0], 100*v[1]/tot[1], v[2]/max(v[1], 1)))
    print("top lines by samples:")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-22s:%-4d samples %5.2f%%  instr %5.2f%%  active threads %5.1f" % (f, l, 100*v[0]/tot[0], 100*v[1]/tot[1], v[2]/max(v[1], 1)))


if __name__ == "__main__":
    main()
