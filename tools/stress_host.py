#!/usr/bin/env python
"""Developer tool (GPU box): randomized stress of the host-pointer pipelines (batch sizes, lane counts, frame counts,
pixel types, partial decodes, two contexts at once) against the CPU oracle.  Usage: stress_host.py [seconds] [seed]"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc, trpx_b200

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t_end = time.time() + budget
trials = 0
while time.time() < t_end:
    os.environ["TRPX_BATCH_MB"] = str(int(rng.choice([0, 0, 1, 2])))
    os.environ["TRPX_ENC_LANES"] = str(int(rng.integers(1, 9)))
    os.environ["TRPX_DEC_LANES"] = str(int(rng.integers(1, 9)))
    a, b = trpx_b200.Codec(0), trpx_b200.Codec(0)
    dt = int(rng.choice([orc.U8, orc.U16, orc.U16, orc.I16, orc.U32, orc.I32]))
    npdt = orc.NP_OF[dt]
    n = int(rng.choice([12 * 700 + 8, 256 * 256, 5001, 12 * 4096]))
    F = int(rng.integers(1, 90))
    st = np.stack([orc.kat_fill(dt, n, 7000 + trials * 100 + f) >> int(rng.integers(0, 8 * np.dtype(npdt).itemsize - 2)) for f in range(F)])
    want, per, pbw = orc.encode_stack(st)
    res = {}

    def work(c, tag):
        p, fb, pb = c.encode(st)
        ok = pb == pbw and np.array_equal(fb, per) and np.array_equal(p, want)
        f0 = int(rng.integers(0, F)); nf = int(rng.integers(1, F - f0 + 1))
        d, _ = c.decode(p, n, F, np.dtype(npdt).kind == "i", npdt, frame_bytes=fb, first_frame=f0, n_frames=nf)
        res[tag] = ok and np.array_equal(d, st[f0:f0 + nf])

    th = [threading.Thread(target=work, args=(a, "a")), threading.Thread(target=work, args=(b, "b"))]
    [t.start() for t in th]; [t.join() for t in th]
    a.close(); b.close()
    if not (res.get("a") and res.get("b")):
        print("FAILED trial", trials, dict(dt=dt, n=n, F=F, env={k: os.environ[k] for k in ("TRPX_BATCH_MB", "TRPX_ENC_LANES", "TRPX_DEC_LANES")}), res)
        sys.exit(1)
    trials += 1
print("stress ok: %d trials" % trials)
