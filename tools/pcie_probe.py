#!/usr/bin/env python
"""Developer tool (GPU box): raw pinned-memory PCIe bandwidth, one direction at a time and both at once --
the bound of bench.py's e2e number."""
import time
import torch
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


run(True, True, 1)
print("H2D only   %.1f GB/s" % run(True, False))
print("D2H only   %.1f GB/s" % run(False, True))
print("both       %.1f GB/s per direction" % run(True, True))

# the traffic pattern of bench.py's streamed e2e: one host thread keeps `depth` pixel-sized H2D copies queued on its
# own streams; a second thread runs [small H2D -> big D2H] pairs on its streams, `depth2` pairs in flight
import threading
big, small = 128 << 20, 27 << 20
for depth, depth2 in ((4, 6), (2, 6), (4, 2), (1, 1)):
    sa = [torch.cuda.Stream() for _ in range(depth)]
    sb = [torch.cuda.Stream() for _ in range(depth2)]
    nrep = 40

    def side_a():
        for i in range(nrep):
            st = sa[i % depth]
            st.synchronize()
            with torch.cuda.stream(st):
                d1[:big].copy_(h1[(i % 8) * big:(i % 8 + 1) * big], non_blocking=True)
        for st in sa:
            st.synchronize()

    def side_b():
        for i in range(nrep):
            st = sb[i % depth2]
            st.synchronize()
            with torch.cuda.stream(st):
                d2[:small].copy_(h1[:small], non_blocking=True)
                h2[(i % 8) * big:(i % 8 + 1) * big].copy_(d2[big:2 * big], non_blocking=True)
        for st in sb:
            st.synchronize()

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=side_a), threading.Thread(target=side_b)]
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("pattern depth %d/%d: %.1f ms for %d x (128 MB H2D | 27 MB H2D + 128 MB D2H): %.1f GB/s H2D, %.1f GB/s D2H"
          % (depth, depth2, 1e3 * dt, nrep, nrep * (big + small) / dt / 1e9, nrep * big / dt / 1e9))
