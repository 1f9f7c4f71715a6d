#!/usr/bin/env python
"""Developer tool (GPU box): the encoder reading its pixels straight from pinned host memory (TMA bulk loads over PCIe)
and the decoder's unpack kernel writing its pixels straight to pinned host memory (TMA bulk stores over PCIe) -- no
copy engine, no staging in HBM -- against the copy-engine pipelines of trpx_encode_host / trpx_decode_host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, trpx_b200
import bench

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = bench.N_VALUES
dev = torch.device("cuda", 0)
px = bench.synth_stack(torch, F, 1000, dev)
h_px = torch.empty((F, N), dtype=torch.int16, pin_memory=True); h_px.copy_(px)
h_back = torch.empty((F, N), dtype=torch.int16, pin_memory=True)
enc, dec = trpx_b200.Codec(0), trpx_b200.Codec(0)
cap = trpx_b200.max_compressed_bytes(N, np.uint16, 12, F)
payload = torch.empty(cap, dtype=torch.uint8, device=dev)
ends = torch.zeros(F, dtype=torch.int64, device=dev)
small = torch.zeros(4, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
raw = F * N * 2


def enc_from_host(stream):
    enc.encode_device(h_px.data_ptr(), np.uint16, N, F, payload.data_ptr(), cap, ends.data_ptr(), small.data_ptr(),
                      small.data_ptr() + 4, stream.cuda_stream)


def dec_to_host(stream, nbytes):
    dec.decode_device(payload.data_ptr(), nbytes, False, N, F, ends.data_ptr(), h_back.data_ptr(), np.uint16,
                      small.data_ptr() + 8, stream.cuda_stream)


def timeit(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


enc_from_host(s1); torch.cuda.synchronize()
cb = int(ends[F - 1]); assert int(small[1]) == 0
t = timeit(lambda: enc_from_host(s1))
print("encode, pixels read from pinned host memory by the kernel: %.1f ms, %.1f GB/s of pixels" % (1e3 * t, raw / t / 1e9))
dec_to_host(s2, cb); torch.cuda.synchronize()
assert int(small[2]) == 0 and torch.equal(h_back, h_px), "zero-copy decode differs"
t = timeit(lambda: dec_to_host(s2, cb))
print("decode, pixels written to pinned host memory by the kernel: %.1f ms, %.1f GB/s of pixels" % (1e3 * t, raw / t / 1e9))
# both directions at once (two contexts, two streams): the encoder re-encodes into a second payload buffer
payload2 = torch.empty(cap, dtype=torch.uint8, device=dev); ends2 = torch.zeros(F, dtype=torch.int64, device=dev)


def both():
    enc.encode_device(h_px.data_ptr(), np.uint16, N, F, payload2.data_ptr(), cap, ends2.data_ptr(), small.data_ptr(),
                      small.data_ptr() + 4, s1.cuda_stream)
    dec_to_host(s2, cb)


t = timeit(both)
print("both at once: %.1f ms for %d frames each way -> %.1f GB/s per direction" % (1e3 * t, F, raw / t / 1e9))
L = trpx_b200.lib()

# the question that decides whether zero-copy helps the duplex e2e: how fast is the kernel-driven direction while the
# copy engine drives the other one?
big = 128 << 20
h_src = torch.empty(8 * big, dtype=torch.uint8, pin_memory=True)
d_dst = torch.empty(big, dtype=torch.uint8, device=dev)
d_src = torch.empty(big, dtype=torch.uint8, device=dev)
h_dst = torch.empty(8 * big, dtype=torch.uint8, pin_memory=True)
s3 = torch.cuda.Stream()
nrep = int(raw / big) + 1


def ce_h2d():
    with torch.cuda.stream(s3):
        for i in range(nrep):
            d_dst.copy_(h_src[(i % 8) * big:(i % 8 + 1) * big], non_blocking=True)


def ce_d2h():
    with torch.cuda.stream(s3):
        for i in range(nrep):
            h_dst[(i % 8) * big:(i % 8 + 1) * big].copy_(d_src, non_blocking=True)


t = timeit(lambda: (ce_h2d(), dec_to_host(s2, cb)))
print("copy-engine H2D (%.2f GB) || kernel-written D2H (%.2f GB): %.1f ms -> %.1f / %.1f GB/s"
      % (nrep * big / 1e9, raw / 1e9, 1e3 * t, nrep * big / t / 1e9, raw / t / 1e9))
t = timeit(lambda: (ce_d2h(), enc_from_host(s1)))
print("copy-engine D2H (%.2f GB) || kernel-read H2D (%.2f GB): %.1f ms -> %.1f / %.1f GB/s"
      % (nrep * big / 1e9, raw / 1e9, 1e3 * t, nrep * big / t / 1e9, raw / t / 1e9))
t = timeit(lambda: (ce_h2d(), ce_d2h()))
print("copy-engine H2D || copy-engine D2H on ONE stream (serial): %.1f ms" % (1e3 * t))
