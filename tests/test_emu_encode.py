"""Kernel-logic tests WITHOUT a GPU: trpx_b200/csrc/terse_encode.cuh compiled for the host with the
test-only SIMT emulator, compared bit-for-bit with the oracle.  (The real parity tests run the CUDA
build through the C ABI: tests/test_gpu_*.py.)"""
import numpy as np
import pytest

import emu_lib
import golden_util as G
import orc


def check(stack, block=12, **kw):
    p, ends, pb, st, fast = emu_lib.encode(stack, block, **kw)
    q, per, qb = orc.encode_stack(stack, block)
    assert st == 0
    assert pb == qb
    assert np.array_equal(ends, np.cumsum(per))
    assert p.size == q.size and np.array_equal(p, q)
    return fast


@pytest.mark.parametrize("c", G.load("kat_small"), ids=lambda c: c["name"])
def test_small_kats(c):
    a = G.small_input(c)
    check(a[None, :], c["block"])


@pytest.mark.parametrize("dt", list(range(8)), ids=lambda d: str(np.dtype(orc.NP_OF[d])))
def test_fast_kernel_all_types(dt):
    n = 12 * 700 + 8 if orc.NP_OF[dt]().itemsize >= 2 else 12 * 1500 + 16   # multiple of 16 bytes, partial last block
    n -= n % 16
    n += 8 if (n * orc.NP_OF[dt]().itemsize) % 16 == 0 else 0
    while (n * orc.NP_OF[dt]().itemsize) % 16:
        n += 1
    st = np.stack([orc.kat_fill(dt, n, 40 + f) for f in range(3)])
    assert check(st) is True
    assert check(st, incl_stride=7) is True              # exercises the aggregate walk of the look-back


def test_fast_kernel_synthetic_frames_multi_tile():
    st = np.stack([orc.synth_frame(orc.U16, 128, 96, 2.0, 12, 1000 + f) for f in range(4)])
    assert check(st) is True
    assert check(st, incl_stride=5) is True


def test_sparse_and_zero_frames():
    z = np.zeros((3, 12 * 4096), np.uint8)
    assert check(z) is True                              # 1 bit per block, many threads per word
    z[1, 5000] = 1
    z[2, ::977] = 3
    check(z)
    check(z.astype(np.uint16))
    check(z.astype(np.uint32), incl_stride=3)


def test_tiny_frames_many():
    st = np.stack([orc.kat_fill(orc.U16, 8, 7 + f) for f in range(40)])     # 16-byte frames
    assert check(st) is True
    st = np.stack([orc.kat_fill(orc.U16, 24, 7 + f) for f in range(9)])
    check(st, incl_stride=4)
    check(np.zeros((50, 8), np.uint16), incl_stride=6)   # one 1-bit block per frame


def test_signed_extremes_outside_reference_domain():
    a = np.array([-32768, 32767, -1, 0, 5, -5, 100, -100, 1, 2, 3, 4] * 4, np.int16)
    check(a[None, :])                                    # width 17 (App. C4: correct stream, not reference's)
    b = np.array([-128, 127, 0, 1] * 12, np.int8)
    check(b[None, :])
    c = np.array([-2**31, 2**31 - 1, 0, -1] * 6, np.int32)
    check(c[None, :])
    d = np.array([-2**63, 2**63 - 1, 0, -1] * 6, np.int64)
    check(d[None, :])
    e = np.array([2**64 - 1, 0, 1, 2**63] * 6, np.uint64)
    check(e[None, :])


@pytest.mark.parametrize("dt", [orc.U8, orc.U16, orc.I16, orc.U32, orc.I64])
def test_generic_kernel_blocks_and_alignment(dt):
    rng = np.random.default_rng(5 + dt)
    for block, n, frames, mis in [(12, 1001, 3, 0), (7, 500, 2, 0), (1, 77, 2, 0), (40, 999, 3, 0),
                                  (12, 1024, 2, 2 if orc.NP_OF[dt]().itemsize <= 2 else 8), (5, 3, 4, 0),
                                  (300, 5000, 2, 0)]:
        st = np.stack([orc.kat_fill(dt, n, int(rng.integers(1, 1 << 30))) for _ in range(frames)])
        fast = check(st, block, misalign=mis, incl_stride=int(rng.integers(0, 4)))
        assert fast is False


@pytest.mark.parametrize("ctas,sms", [(3, 1), (7, 3)])
def test_concurrent_ctas(ctas, sms):
    """Several CTAs resident at once (their fibers interleaved): tickets, the look-back between tiles and the
    boundary-word hand-off between tiles of different CTAs run concurrently, as on the device."""
    st = np.stack([orc.synth_frame(orc.U16, 256, 64, 2.0, 4, 1000 + f) for f in range(40)])
    assert check(st, ctas=ctas, sms=sms) is True
    rng = np.random.default_rng(9)
    n = 3072 * 67 + 100
    big = (rng.random((3, n)) < 0.05).astype(np.uint32) * rng.integers(1, 5000, size=(3, n), dtype=np.uint32)
    assert check(big, ctas=ctas, sms=sms) is True
    assert check(np.stack([orc.kat_fill(orc.I16, 12 * 3000 + 8, 5 + f) for f in range(6)]), ctas=ctas, sms=sms) is True


def test_capacity_error():
    st = np.stack([orc.kat_fill(orc.U16, 12 * 800, 3)])
    p, ends, pb, status, fast = emu_lib.encode(st, cap=1024)
    assert status == 2


def test_groups_of_more_than_64_tiles():
    """Frames longer than one look-back group (64 tiles) and a group that ends mid-frame."""
    rng = np.random.default_rng(9)
    n = 3072 * 67 + 100                                  # u32 tile = 3072 values -> 68 tiles, 2 groups per frame
    st = (rng.random((3, n)) < 0.05).astype(np.uint32) * rng.integers(1, 5000, size=(3, n), dtype=np.uint32)
    assert check(st) is True
    assert check(st, incl_stride=3) is True


def test_staging_ring_wraps_and_blocks():
    """A staging ring barely larger than two worst-case tiles: allocations wrap and wait for stores."""
    st = np.stack([orc.kat_fill(orc.U16, 6144 * 9 + 40, 11 + f) for f in range(2)])   # 10 tiles per frame
    for ring in (8192, 16384):                           # the smallest legal ring (a power of two >= 2 worst-case tiles) wraps often
        assert check(st, incl_stride=(ring << 16)) is True
    z = np.zeros((2, 6144 * 20), np.uint16)                                          # tiny tiles: depth-limited
    z[1, ::4099] = 9
    assert check(z, incl_stride=8192 << 16) is True
