#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS listing by CUDA source line.

ncu's CSV source page is per SASS instruction; this joins it (by instruction order) with
`nvdisasm -g` line info of the same kernel, so stall samples and executed instructions can be read per
line of the .cuh files.

  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled kernel name> [top N]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


SORT = 1


def main():
    global SORT
    if "--by-inst" in sys.argv:
        sys.argv.remove("--by-inst")
        SORT = 0
    dump = None                       # --sass FILE:LINE[,LINE...] | all : list the SASS of those lines with their counts
    if "--sass" in sys.argv:
        i = sys.argv.index("--sass")
        dump = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    rep, kre, cubin, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the report may hold several launches of the kernel: keep the first table only
    tables, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            tables.append(cur)
        elif cur is not None:
            cur.append(r)
    t = tables[0]
    hdr = t[0]
    ci = {n: i for i, n in enumerate(hdr)}
    sass = t[1:]
    dis = subprocess.run(["nvdisasm", "-g", cubin], stdout=subprocess.PIPE, text=True).stdout
    line_of = []
    curline = ("?", 0)
    inl = ""
    inside = False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+(\S+?),", ln)
        if m:
            inside = m.group(1) == ".text." + mangled
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            curline = (m.group(1).split("/")[-1], int(m.group(2)))
            inl = m.group(3)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            line_of.append(curline)
    if len(line_of) != len(sass):
        print("warning: %d SASS rows in the report vs %d in nvdisasm" % (len(sass), len(line_of)))
    if dump:
        want = None if dump == "all" else set(int(x) for x in dump.split(","))
        for i, r in enumerate(sass):
            key = line_of[i] if i < len(line_of) else ("?", 0)
            if want is None or key[1] in want:
                print("%5d %-24s %10s %6s  %s" % (i, "%s:%d" % key, r[ci["Instructions Executed"]],
                                                  r[ci["Warp Stall Sampling (All Samples)"]], r[ci["Source"]][:90]))
        return
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for i, r in enumerate(sass):
        key = line_of[i] if i < len(line_of) else ("?", 0)
        vals = [int(float(r[ci[c]] or 0)) for c in ("Instructions Executed", "Warp Stall Sampling (All Samples)",
                                                      "Warp Stall Sampling (Not-issued Samples)")]
        for k in range(3):
            agg[key][k] += vals[k]
            tot[k] += vals[k]
    print("total: inst %d, samples %d, not-issued %d" % tuple(tot))
    print("%-28s %12s %6s %10s %6s" % ("file:line", "warp inst", "%", "samples", "%"))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
        print("%-28s %12d %6.2f %10d %6.2f" % ("%s:%d" % key, v[0], 100.0 * v[0] / max(tot[0], 1), v[1],
                                              100.0 * v[1] / max(tot[1], 1)))


if __name__ == "__main__":
    main()
