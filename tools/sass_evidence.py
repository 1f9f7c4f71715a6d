#!/usr/bin/env python
"""SASS listings and mnemonic counts of the benchmark's kernels -> profiles/<round>_sass_*.txt (cuobjdump, no GPU needed)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "trpx_b200", "libtrpx_b200.so")
want = ["terse_encode_kernelItLi192", "prolix_walk_kernelILi256", "prolix_resolve_kernelILi256", "prolix_unpack_seg_kernelItLb0",
        "prolix_frame_spec_kernelILi256", "prolix_frame_chain_kernel", "prolix_segments_kernelILi1024"]
allsass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
bodies, cur = {}, None
for l in allsass.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1)
        bodies[cur] = []
    elif cur is not None:
        bodies[cur].append(l)
funcs = sorted(f for f in bodies if any(w in f for w in want))
KEYS = ["UBLKCP", "SYNCS", "REDUX", "SHFL", "VOTE", "BAR", "ATOMS", "ATOMG", "LDS", "STS", "LDG", "STG", "LDC", "SHF", "LOP3", "IMAD", "NANOSLEEP", "BPT"]
ev = ["SASS evidence, round 2 (cuobjdump -sass of trpx_b200/libtrpx_b200.so, sm_100a, final round-2 binaries; full listings of the",
      "three main kernels: %s_sass_u16_kernels.txt).  UBLKCP = TMA bulk copy, SYNCS = mbarrier; BPT (trap) must be 0." % R, ""]
full = ["cuobjdump -sass trpx_b200/libtrpx_b200.so (sm_100a), the three main u16 kernels of the benchmark", ""]
for f in funcs:
    lines = [l for l in bodies[f] if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
    cnt = collections.Counter()
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            cnt[m.group(1).split(".")[0]] += 1
    ev.append(f)
    ev.append("  instructions: %d" % len(lines))
    for k in KEYS:
        n = sum(v for kk, v in cnt.items() if kk == k or kk.startswith(k))
        if n or k == "BPT":
            ev.append("  %-10s %d" % (k, n))
    ev.append("")
    if any(w in f for w in want[:1] + want[1:2] + want[3:4]):
        full.append("Function : " + f)
        full += lines
        full.append("")
open(os.path.join(ROOT, "profiles", R + "_sass_evidence.txt"), "w").write("\n".join(ev) + "\n")
open(os.path.join(ROOT, "profiles", R + "_sass_u16_kernels.txt"), "w").write("\n".join(full) + "\n")
print("\n".join(ev[:60]))
