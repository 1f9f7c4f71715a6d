#!/usr/bin/env python
"""Developer tool: build variants of libtrpx_b200.so with extra -D flags into trpx_b200/_variants/ (git-ignored)
for A/B runs on the GPU box:  python tools/ab_build.py name1:-DFOO=1,-DBAR=2 name2: ...
Then on the box:  bash tools/ab_run.sh name1 name2 ...  (copies each over the in-tree library and runs bench.py)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trpx_b200 import build as B
os.makedirs(os.path.join(B.HERE, "_variants"), exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out = os.path.join(B.HERE, "_variants", "lib_%s.so" % name)
    cmd = [B.nvcc(), "-ccbin", "/usr/bin/g++"] + B.NVCC_FLAGS + [f for f in flags.split(",") if f] + B.SOURCES + ["-o", out]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    o, _ = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + o)
