"""Single-call latency (device-resident, CUDA events): F frames per call, for P1 geometries given by TRPX_SEG_BYTES /
TRPX_WARM_BYTES in the environment.  Tuning helper, not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import trpx_b200, bench

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = bench.CONFIGS["c2"]
dev = torch.device("cuda", 0)
codec = trpx_b200.Codec(0)
codec.set_profiling(True)
px = bench.synth_stack(torch, cfg, 0, F, dev)
N = px.shape[1]
cap = trpx_b200.max_compressed_bytes(N, np.uint16, 12, F)
payload = torch.empty(cap, dtype=torch.uint8, device=dev)
ends = torch.zeros(F, dtype=torch.int64, device=dev)
small = torch.zeros(4, dtype=torch.int32, device=dev)
back = torch.empty_like(px)
st = torch.cuda.current_stream().cuda_stream
codec.encode_device(px.data_ptr(), np.uint16, N, F, payload.data_ptr(), cap, ends.data_ptr(), small.data_ptr(), small.data_ptr() + 4, st)
torch.cuda.synchronize()
cb = int(ends[F - 1])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for k in range(25):
    e0.record()
    codec.decode_device(payload.data_ptr(), cb, False, N, F, ends.data_ptr(), back.data_ptr(), np.uint16, small.data_ptr() + 8, st, lane=1)
    e1.record()
    torch.cuda.synchronize()
    if k >= 5:
        ts.append(e0.elapsed_time(e1))
ts.sort()
ok = bool(torch.equal(back, px)) and int(small[2]) == 0
print("F=%d seg=%s warm=%s decode median %.1f us  %s  kernels %s" % (F, os.environ.get("TRPX_SEG_BYTES", "auto"), os.environ.get("TRPX_WARM_BYTES", "auto"),
      1e3 * ts[len(ts) // 2], "ok" if ok else "FAILED", {k: round(1e3 * v, 1) for k, v in codec.last_kernel_times(1)}))
