"""N > 1 host logic on CPU: world_size-2 gloo processes shard a stack by frame, each encodes its slab
(with the CPU oracle standing in for the device codec -- this test checks the bookkeeping, not the
kernels), and the merged stack must equal the single-process stack byte for byte."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import orc
from trpx_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stack = np.stack([orc.kat_fill(orc.U16, 5000, 100 + f) for f in range(7)])   # 7 frames over 2 ranks: 3 + 4
    enc = lambda st: orc.encode_stack(st)
    payload, fb, pb = shard.encode_sharded(enc, stack, dist)
    whole = orc.encode_stack(stack)
    ok = np.array_equal(payload, whole[0]) and np.array_equal(fb, whole[1]) and pb == whole[2]

    def dec(slab, sizes):
        out, off = [], 0
        for sz in sizes:
            out.append(orc.decode_frame(slab[off:off + int(sz)], 5000, False, np.uint16)[0])
            off += int(sz)
        return np.stack(out)

    back = shard.decode_sharded(dec, payload, fb, 7, dist)
    ok = ok and np.array_equal(back, stack)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_by_frame():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_frame_ranges_and_slabs():
    assert [shard.frame_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    fb = np.array([5, 7, 1, 9], np.uint64)
    assert shard.payload_slab(fb, 1, 3) == (5, 13)
    p, f, b = shard.merge_encoded([(np.arange(3, dtype=np.uint8), [3], 4), (np.arange(2, dtype=np.uint8), [2], 9)])
    assert p.tolist() == [0, 1, 2, 0, 1] and f.tolist() == [3, 2] and b == 9
