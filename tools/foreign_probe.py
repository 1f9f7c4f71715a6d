"""Decode of a foreign multi-frame payload (frame sizes unknown) against the known-sizes decode, device-resident.
Tuning / evidence helper, not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import trpx_b200, bench

F = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cfg = bench.CONFIGS["c2"]
dev = torch.device("cuda", 0)
codec = trpx_b200.Codec(0)
codec.set_profiling(True)
px = bench.synth_stack(torch, cfg, 0, F, dev)
N = px.shape[1]
cap = trpx_b200.max_compressed_bytes(N, np.uint16, 12, F)
payload = torch.empty(cap, dtype=torch.uint8, device=dev)
ends = torch.zeros(F, dtype=torch.int64, device=dev)
ends2 = torch.zeros(F, dtype=torch.int64, device=dev)
small = torch.zeros(4, dtype=torch.int32, device=dev)
back = torch.empty_like(px)
st = torch.cuda.current_stream().cuda_stream
codec.encode_device(px.data_ptr(), np.uint16, N, F, payload.data_ptr(), cap, ends.data_ptr(), small.data_ptr(), small.data_ptr() + 4, st)
torch.cuda.synchronize()
cb = int(ends[F - 1])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for name, fe, fe_out in (("known sizes", ends.data_ptr(), None), ("sizes unknown", None, ends2.data_ptr())):
    ts = []
    for k in range(6):
        back.zero_()
        e0.record()
        codec.decode_device(payload.data_ptr(), cb, False, N, F, fe, back.data_ptr(), np.uint16, small.data_ptr() + 8, st, lane=1, d_frame_ends_out=fe_out)
        e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            ts.append(e0.elapsed_time(e1))
    ok = bool(torch.equal(back, px)) and int(small[2]) == 0 and (fe_out is None or bool(torch.equal(ends, ends2)))
    res[name] = min(ts)
    print("F=%d %-14s decode %.3f ms  %s  %s" % (F, name, min(ts), "ok" if ok else "FAILED", {k: round(v, 3) for k, v in codec.last_kernel_times(1)}), flush=True)
print("ratio unknown / known: %.2f" % (res["sizes unknown"] / res["known sizes"]))
