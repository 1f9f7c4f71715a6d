"""Debug helper (not part of the product): device-resident encode of F synthetic frames, status word and a
byte comparison of the first frames with the CPU oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import trpx_b200, orc
import bench

def run(F, check=4):
    dev = torch.device("cuda", 0)
    codec = trpx_b200.Codec(0)
    px = bench.synth_stack(torch, F, 1000, dev)
    N = px.shape[1]
    cap = trpx_b200.max_compressed_bytes(N, np.uint16, 12, F)
    payload = torch.zeros(cap, dtype=torch.uint8, device=dev)
    ends = torch.zeros(F, dtype=torch.int64, device=dev)
    small = torch.zeros(4, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    codec.encode_device(px.data_ptr(), np.uint16, N, F, payload.data_ptr(), cap, ends.data_ptr(), small.data_ptr(), small.data_ptr() + 4, st)
    torch.cuda.synchronize()
    status = int(small[1]) & 0xffffffff
    e = ends.cpu().numpy()
    ok = True
    if status == 0:
        host = px[:check].cpu().numpy().view(np.uint16)
        want, per, pb = orc.encode_stack(host)
        got = payload[:int(e[check - 1])].cpu().numpy()
        ok = np.array_equal(got, want) and np.array_equal(e[:check], np.cumsum(per).astype(np.int64))
    if os.environ.get("TRPX_STATS"):
        capb = cap & ~15
        dbg = payload[capb - 256:capb - 96].cpu().numpy().view(np.uint64)
        names = ["wait_full", "wait_ticket", "forced_drain", "early_drain", "final_drain", "n_forced", "n_early", "rounds", "total", "pack", "n_tail_fetch", "drain_wait_resolved", "L:load", "L:ticket+issue", "L:K1-K3", "L:post+alloc", "L:pack", "L:merge", "L:publish", "L:early"]
        tot = float(dbg[8]) or 1.0
        print("  stats (cycles summed over worker warps; %% of total): " + ", ".join("%s=%.3g (%.1f%%)" % (n, float(v), 100 * float(v) / tot) for n, v in zip(names, dbg) if n != "-"))
    print("F=%d status=0x%x (code %d warp %d round %d) bytes=%d first-frames-ok=%s pb=%d" % (F, status, status & 15, (status >> 4) & 31, status >> 12, int(e[-1]), ok, int(small[0])), flush=True)
    codec.close()

if __name__ == "__main__":
    for F in [int(x) for x in sys.argv[1:]] or [1, 8, 64, 1000]:
        run(F, min(F, 4))
