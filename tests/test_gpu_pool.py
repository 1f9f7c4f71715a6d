"""Multi-GPU frame sharding through the C ABI (trpx_pool_*): sharded == single-GPU == oracle, byte for byte
(test 4).  Frames are independent in the reference (Terse.hpp:290-302, :505, :547); the only shared scalars
are prolix_bits (max, :516) and memory_size (sum, :459)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return torch.cuda.device_count()


def _stack(dt, frames, n, seed):
    return np.stack([orc.kat_fill(dt, n, seed + f) for f in range(frames)])


@pytest.mark.parametrize("dt,frames,n", [(orc.U16, 37, 512 * 64), (orc.I32, 9, 12 * 4096 + 8), (orc.U8, 5, 16 * 7000)])
def test_pool_equals_single_device_and_oracle(dt, frames, n):
    import trpx_b200
    st = _stack(dt, frames, n, 77)
    want, per, pb = orc.encode_stack(st)
    with_pool = trpx_b200.Pool()                                      # every visible device (1 on a single-GPU box)
    p, fb, qb = with_pool.encode(st)
    assert qb == pb and np.array_equal(fb, per) and p.size == want.size and np.array_equal(p, want)
    single = trpx_b200.Codec(0)
    p1, fb1, qb1 = single.encode(st)
    assert qb1 == qb and np.array_equal(fb1, fb) and np.array_equal(p1, p)
    back, _ = with_pool.decode(p, n, frames, dt >= orc.I8, st.dtype, frame_bytes=fb)
    assert np.array_equal(back, st)
    # a sub-range of the frames, and a foreign stack (sizes unknown: the first device recovers them)
    sub, _ = with_pool.decode(p, n, frames, dt >= orc.I8, st.dtype, frame_bytes=fb, first_frame=2, n_frames=frames - 3)
    assert np.array_equal(sub, st[2:frames - 1])
    again, rec = with_pool.decode(p, n, frames, dt >= orc.I8, st.dtype)
    assert np.array_equal(again, st) and np.array_equal(rec, per)
    single.close()
    with_pool.close()


def test_pool_shards_over_two_devices():
    """The sharded path proper: needs at least two GPUs (the driver's multi-GPU tier; skipped on a one-GPU box)."""
    if _devices() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    import trpx_b200
    st = np.stack([orc.synth_frame(orc.U16, 512, 512, 2.0, 200, 4000 + f) for f in range(23)])
    want, per, pb = orc.encode_stack(st)
    pool = trpx_b200.Pool([0, 1])
    assert pool.size == 2
    p, fb, qb = pool.encode(st)
    assert qb == pb and np.array_equal(fb, per) and np.array_equal(p, want)      # sharded == oracle
    single = trpx_b200.Codec(1)                                                  # the other device, on its own
    p1, _, _ = single.encode(st)
    assert np.array_equal(p1, p)                                                 # == single GPU
    back, _ = pool.decode(p, st.shape[1], st.shape[0], False, np.uint16, frame_bytes=fb)
    assert np.array_equal(back, st)
    rec_back, rec = pool.decode(p, st.shape[1], st.shape[0], False, np.uint16)
    assert np.array_equal(rec_back, st) and np.array_equal(rec, per)
    single.close()
    pool.close()


def test_pinned_ranges():
    import ctypes as C
    import trpx_b200
    L = trpx_b200.lib()
    a = np.zeros(1 << 20, np.uint8)
    assert L.trpx_host_pin(a.ctypes.data, a.size) == trpx_b200.OK
    assert L.trpx_host_pin(a.ctypes.data, a.size) == 7            # TRPX_ALREADY: not an error, and not ours to unpin
    assert L.trpx_host_unpin(a.ctypes.data) == trpx_b200.OK
    p = L.trpx_host_alloc(4096)
    assert p
    C.memset(p, 0, 4096)
    L.trpx_host_free(p)
