#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean, share."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    d[r[ki][:90]].append(v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0))
ours = {k: v for k, v in d.items() if any(t in k for t in ("terse_", "prolix_", "publish"))}
tot = sum(sum(v) for v in ours.values())
print("kernels of libtrpx_b200.so (share = of their sum; cold-cache, serialised under ncu)")
for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
    print("%-92s n=%3d mean=%10.1f us share=%.3f" % (k, len(v), sum(v) / len(v), sum(v) / tot))
other = sum(sum(v) for k, v in d.items() if k not in ours)
print("other kernels (torch: synthetic input generation, checks): n=%d, total %.1f us" % (sum(len(v) for k, v in d.items() if k not in ours), other))
