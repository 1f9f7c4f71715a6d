#!/usr/bin/env python
"""Secondary workloads (BASELINE.json configs[2..4]) on one GPU, device-resident: informational numbers for DESIGN.md
(the headline metric lives in bench.py).  Each line: encode / decode ms, GB/s of algorithmic bytes, round-trip check."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, trpx_b200

dev = torch.device("cuda", 0)
codec = trpx_b200.Codec(0)
codec.set_profiling(True)
g = torch.Generator(device=dev); g.manual_seed(7)


def run(name, px, np_dtype, signed=False, reps=5):
    F, N = px.shape
    so = np.dtype(np_dtype).itemsize
    cap = trpx_b200.max_compressed_bytes(N, np_dtype, 12, F)
    payload = torch.empty(cap, dtype=torch.uint8, device=dev)
    ends = torch.zeros(F, dtype=torch.int64, device=dev)
    small = torch.zeros(4, dtype=torch.int32, device=dev)
    back = torch.empty_like(px)
    st = torch.cuda.current_stream().cuda_stream
    enc = lambda: codec.encode_device(px.data_ptr(), np_dtype, N, F, payload.data_ptr(), cap, ends.data_ptr(), small.data_ptr(), small.data_ptr() + 4, st)
    enc(); torch.cuda.synchronize()
    cb = int(ends[F - 1])
    dec = lambda: codec.decode_device(payload.data_ptr(), cb, signed, N, F, ends.data_ptr(), back.data_ptr(), np_dtype, small.data_ptr() + 8, st, lane=1)
    dec(); torch.cuda.synchronize()
    ok = bool(torch.equal(back, px)) and int(small[1]) == 0 and int(small[2]) == 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    te = td = 0.0
    for _ in range(reps):
        ev[0].record(); enc(); ev[1].record(); dec(); ev[2].record(); torch.cuda.synchronize()
        te += ev[0].elapsed_time(ev[1]); td += ev[1].elapsed_time(ev[2])
    te /= reps; td /= reps
    kt = {k: round(v, 3) for k, v in codec.last_kernel_times(0) + codec.last_kernel_times(1)}
    alg = F * N * so + cb
    print("%-46s F=%-5d ratio %.3f  encode %7.3f ms %7.1f GB/s | decode %7.3f ms %7.1f GB/s | roundtrip %s" %
          (name, F, cb / (F * N * so), te, alg / te / 1e6, td, alg / td / 1e6, "ok" if ok else "FAILED"), flush=True)
    print("    ", kt, flush=True)


def poisson(F, N, lam, dtype):
    return torch.poisson(torch.full((F, N), lam, device=dev), generator=g).to(torch.int32).to(dtype)


which = sys.argv[1:] or ["c3", "c4u8", "c4u16", "c5i16", "c5i32"]
if "c3" in which:      # Eiger2-16M-class u32
    N = 4148 * 4362
    px = poisson(8, N, 0.5, torch.int32)
    hot = torch.randint(0, N, (8, 2000), device=dev, generator=g)
    px.scatter_(1, hot, torch.randint(1000, 1000000, (8, 2000), device=dev, generator=g, dtype=torch.int32))
    run("C3 4148x4362 u32, Poisson(0.5)+2000 peaks", px, np.uint32)
    del px
if "c4u8" in which:    # cryo-EM counting movie, mostly 0/1
    run("C4 5760x4092 u8, Poisson(0.02)", poisson(40, 5760 * 4092, 0.02, torch.uint8), np.uint8)
if "c4u16" in which:
    run("C4 5760x4092 u16, Poisson(0.02)", poisson(40, 5760 * 4092, 0.02, torch.int16), np.uint16)
if "c5i16" in which:   # dark-subtracted signed frames
    px = (poisson(4000, 512 * 512, 3.0, torch.int16) - 3 + torch.round(2 * torch.randn((4000, 512 * 512), device=dev, generator=g)).to(torch.int16))
    run("C5 512x512 i16 dark-subtracted", px, np.int16, signed=True)
    del px
if "c5i32" in which:
    px = (poisson(16, 4148 * 4362, 3.0, torch.int32) - 3)
    run("C5 4148x4362 i32 dark-subtracted", px, np.int32, signed=True)
