#!/usr/bin/env python
"""Hottest SASS instructions (by stall samples) of a kernel in an .ncu-rep, with their source lines."""
import csv, io, re, subprocess, sys
rep, kre, cubin, mang = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ci = {n: i for i, n in enumerate(hdr)}
sass = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    sass.append(r)
dis = subprocess.run(["nvdisasm", "-g", cubin], stdout=subprocess.PIPE, text=True).stdout
inside = False; lines = []; cur = ("?", 0, "")
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+(\S+?),", ln)
    if m: inside = m.group(1) == ".text." + mang; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): lines.append(cur)
S = "Warp Stall Sampling (All Samples)"
order = sorted(range(len(sass)), key=lambda i: -int(sass[i][ci[S]] or 0))[:top]
for i in sorted(order):
    r = sass[i]
    l = lines[i] if i < len(lines) else ("?", 0, "")
    print("%5d %6s %9s  %-22s %-40s | %s" % (i, r[ci[S]], r[ci["Instructions Executed"]], "%s:%d" % l[:2], l[2][:40], r[ci["Source"]][:60]))
