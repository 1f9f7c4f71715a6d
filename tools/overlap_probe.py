"""Does the decoder gain from running the P1 walk of one half of a stack while the other half is unpacked?  Two lanes /
two streams of ONE context, device-resident.  Experiment helper, not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import trpx_b200, bench

F = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = bench.CONFIGS["c2"]
dev = torch.device("cuda", 0)
codec = trpx_b200.Codec(0)
px = bench.synth_stack(torch, cfg, 0, F, dev)
N = px.shape[1]
cap = trpx_b200.max_compressed_bytes(N, np.uint16, 12, F)
payload = torch.empty(cap, dtype=torch.uint8, device=dev)
ends = torch.zeros(F, dtype=torch.int64, device=dev)
small = torch.zeros(16, dtype=torch.int32, device=dev)
back = torch.empty_like(px)
st0 = torch.cuda.current_stream()
codec.encode_device(px.data_ptr(), np.uint16, N, F, payload.data_ptr(), cap, ends.data_ptr(), small.data_ptr(), small.data_ptr() + 4, st0.cuda_stream)
torch.cuda.synchronize()
h_ends = ends.cpu().numpy()
cb = int(h_ends[-1])
# chunk tables: payload slices must start 16-byte aligned for the device flavour -> copy each chunk's slab to its own buffer
cuts = [F * k // K for k in range(K + 1)]
slabs, cends = [], []
for k in range(K):
    b0 = int(h_ends[cuts[k] - 1]) if cuts[k] else 0
    b1 = int(h_ends[cuts[k + 1] - 1])
    s = torch.zeros(b1 - b0 + 64, dtype=torch.uint8, device=dev)
    s[:b1 - b0] = payload[b0:b1]
    slabs.append((s, b1 - b0))
    cends.append((ends[cuts[k]:cuts[k + 1]] - b0).contiguous())
streams = [torch.cuda.Stream() for _ in range(2)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def run(overlap):
    back.zero_()
    torch.cuda.synchronize()
    e0.record(st0)
    for s in streams:
        s.wait_stream(st0)
    for k in range(K):
        s = streams[k % 2] if overlap else streams[0]
        lane = 1 + (k % 2) if overlap else 1
        nf = cuts[k + 1] - cuts[k]
        codec.decode_device(slabs[k][0].data_ptr(), slabs[k][1], False, N, nf, cends[k].data_ptr(), back.data_ptr() + cuts[k] * N * 2,
                            np.uint16, small.data_ptr() + 8 + 4 * k, s.cuda_stream, lane=lane)
    for s in streams:
        st0.wait_stream(s)
    e1.record(st0)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)

for ov in (False, True):
    ts = [run(ov) for _ in range(6)][2:]
    print("F=%d chunks=%d %s: %.3f ms (ok=%s)" % (F, K, "two streams" if ov else "one stream ", min(ts), bool(torch.equal(back, px))), flush=True)
