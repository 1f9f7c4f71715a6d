"""Kernel-logic tests WITHOUT a GPU: trpx_b200/csrc/prolix_decode.cuh compiled for the host with the
test-only SIMT emulator.  Streams come from the oracle (== reference bytes); tiny segments and short
warm-ups force the speculative walkers to disagree so the verify / re-walk loop is exercised."""
import numpy as np
import pytest

import emu_lib
import golden_util as G
import orc


def roundtrip(stack, block=12, out_dtype=None, known_ends=True, **kw):
    stack = np.ascontiguousarray(stack)
    F, N = stack.shape
    p, per, pb = orc.encode_stack(stack, block)
    ends = np.cumsum(per).astype(np.uint64)
    od = stack.dtype if out_dtype is None else np.dtype(out_dtype)
    got, st, staged, fe = emu_lib.decode(p, N, F, stack.dtype.kind == "i", od, block,
                                         ends if known_ends else None, **kw)
    assert st == 0
    if not known_ends:
        assert np.array_equal(fe, ends)
    want = np.stack([orc.decode_frame(p[int(ends[f] - per[f]):int(ends[f])], N, stack.dtype.kind == "i", od, block)[0]
                     for f in range(F)])
    assert np.array_equal(got, want)
    if od == stack.dtype:
        assert np.array_equal(got, stack)
    return staged


@pytest.mark.parametrize("c", G.load("kat_small"), ids=lambda c: c["name"])
def test_small_kats(c):
    a = G.small_input(c)
    roundtrip(a[None, :], c["block"])


@pytest.mark.parametrize("dt", list(range(8)), ids=lambda d: str(np.dtype(orc.NP_OF[d])))
@pytest.mark.parametrize("seg,warm", [(64, 32), (256, 512), (4096, 4096)])
def test_all_types_staged(dt, seg, warm):
    isz = orc.NP_OF[dt]().itemsize
    n = 12 * (1400 if seg > 64 or isz <= 2 else 350) + 8     # (64-byte segments of wide pixels: thousands of emulated walkers)
    while (n * isz) % 16:
        n += 1
    st = np.stack([orc.kat_fill(dt, n, 90 + f) for f in range(3)])
    assert roundtrip(st, seg_bytes=seg, warm_bytes=warm) is True


def test_synthetic_frames_tiny_segments():
    st = np.stack([orc.synth_frame(orc.U16, 128, 96, 2.0, 12, 1000 + f) for f in range(3)])
    for seg, warm in [(16, 0), (48, 16), (512, 128), (100000, 64)]:
        assert roundtrip(st, seg_bytes=seg, warm_bytes=warm) is True


def test_sparse_and_zero_frames():
    z = np.zeros((3, 12 * 4096), np.uint8)
    roundtrip(z, seg_bytes=32, warm_bytes=16)            # zero-run skipping across segments and tiles
    z[1, 5000] = 1
    z[2, ::977] = 3
    roundtrip(z, seg_bytes=32, warm_bytes=16)
    roundtrip(z.astype(np.uint16), seg_bytes=64, warm_bytes=64)
    roundtrip(z.astype(np.uint32), seg_bytes=4096, warm_bytes=64)


def test_tiny_frames_many():
    roundtrip(np.stack([orc.kat_fill(orc.U16, 8, 7 + f) for f in range(40)]))
    roundtrip(np.zeros((50, 8), np.uint16))
    roundtrip(np.stack([orc.kat_fill(orc.I16, 24, 7 + f) for f in range(9)]), seg_bytes=16, warm_bytes=8)


def test_unknown_frame_sizes_are_recovered():
    st = np.stack([orc.kat_fill(orc.U16, 1000, 7 + f) for f in range(5)])
    roundtrip(st, known_ends=False)
    roundtrip(st[:1], known_ends=False)


@pytest.mark.parametrize("seg,warm", [(64, 64), (256, 128), (1024, 512)])
def test_frame_recovery_follows_the_parallel_walk(seg, warm):
    """Frame sizes unknown (a foreign .trpx, Terse.hpp:562-585): the payload is walked in parallel as one run of blocks and
    one warp follows the frames along the checkpoints.  Ragged last blocks (N % 12 != 0), frames shorter than the distance a
    walker needs to re-synchronise, sparse frames (runs of one-bit headers), wide and signed blocks, many frames."""
    kw = dict(known_ends=False, seg_bytes=seg, warm_bytes=warm)
    emu_lib.set_spec(8, 4, 6, 64)                            # (small candidate windows: the default geometry has its own test below)
    few = seg == 64                                          # (64-byte segments: thousands of emulated walkers per frame)
    roundtrip(np.stack([orc.synth_frame(orc.U16, 64, 52, 2.0, 3, 100 + f) for f in range(12 if few else 40)]), **kw)   # 3328 = 277 * 12 + 4
    roundtrip(np.stack([orc.kat_fill(orc.U16, 30, 5 + f) for f in range(40 if few else 120)]), **kw)         # tiny frames
    z = np.zeros((9, 12 * 500 + 7), np.uint8)
    z[3, 100] = 1
    z[5, ::97] = 3
    z[8, -1] = 200
    roundtrip(z, **kw)                                                                                       # sparse, ragged
    roundtrip(np.stack([orc.kat_fill(orc.I32, 12 * 77 + 5, 11 + f) for f in range(17)]), **kw)              # wide signed blocks
    roundtrip(np.stack([orc.kat_fill(orc.U32, 1999, 3 + f) for f in range(6)]), block=7, **kw)              # another block size
    roundtrip(np.stack([orc.kat_fill(orc.U64, 24, 9 + f) for f in range(30)]), **kw)
    alt = np.stack([orc.synth_frame(orc.U16, 64, 52, 2.0, 3, 300 + f) if f % 3 != 1 else np.zeros(3328, np.uint16) for f in range(14)])
    roundtrip(alt, **kw)                                     # compressed sizes 50:1 apart: the segment guess overshoots and restarts
    emu_lib.set_spec()


@pytest.mark.parametrize("spec", [(64, 32, 80, 256), (8, 2, 3, 256), (3, 0, 0, 1), (16, 40, 0, 24)])
def test_frame_chain_is_followed_speculatively(spec):
    """The chain over frames as table look-ups: F (the header that ends the next frame) evaluated for windows of candidate
    headers in parallel, then followed by one thread; window misses end a batch early, frames the candidates cannot decide
    are left to the serial chain.  Tiny windows force misses and re-anchoring, (3, 0, 0, 1) one exact candidate per frame; a
    candidate whose T runs out of steps (24, 1) is walked out by the following thread from where it stopped."""
    emu_lib.set_spec(*spec)
    try:
        kw = dict(known_ends=False, seg_bytes=256, warm_bytes=128)
        emu_lib.spec_followed()
        roundtrip(np.stack([orc.synth_frame(orc.U16, 128, 104, 2.0, 3, 700 + f) for f in range(12)]), **kw)   # 1110 blocks a frame
        assert emu_lib.spec_followed() == 12                 # every frame end came from the table, none from the serial chain
        roundtrip(np.stack([orc.synth_frame(orc.U16, 64, 52, 2.0, 3, 500 + f) for f in range(37)]), **kw)
        alt = np.stack([orc.synth_frame(orc.U16, 64, 52, 2.0, 3, 300 + f) if f % 3 != 1 else np.zeros(3328, np.uint16) for f in range(14)])
        roundtrip(alt, **kw)                                 # empty frames: runs of one-bit headers on both sides of a boundary
        roundtrip(np.stack([orc.kat_fill(orc.I32, 12 * 77 + 5, 11 + f) for f in range(17)]), **kw)
        roundtrip(np.stack([orc.kat_fill(orc.U16, 12 * 64 + 1, 5 + f) for f in range(21)]), **kw)    # 65 blocks: just above the pass's minimum
    finally:
        emu_lib.set_spec()


def test_frame_recovery_flags_a_stream_that_ends_early():
    st = np.stack([orc.kat_fill(orc.U16, 3000, 7 + f) for f in range(6)])
    p, per, pb = orc.encode_stack(st)
    got, status, staged, fe = emu_lib.decode(p[:int(per[:4].sum()) + 10], 3000, 6, False, np.uint16, 12, None, seg_bytes=256, warm_bytes=128)
    assert status == 4                                       # TRPX_ERR_MALFORMED, no crash, no hang


@pytest.mark.parametrize("src,dst", [(np.uint16, np.uint8), (np.uint16, np.uint64), (np.uint16, np.int32),
                                     (np.int16, np.int8), (np.int16, np.int64), (np.uint32, np.uint16),
                                     (np.int32, np.int16), (np.uint8, np.uint32), (np.int64, np.int32)])
def test_output_conversion_clamps_like_get_range(src, dst):
    rng = np.random.default_rng(3)
    info = np.iinfo(src)
    a = rng.integers(max(info.min, -70000), min(info.max, 70000), size=(2, 12 * 300), endpoint=True).astype(src)
    a[:, ::50] = info.max
    if info.min < 0:
        a[:, 7::50] = info.min + 1
    roundtrip(a, out_dtype=dst)


def test_signed_extremes_roundtrip():
    a = np.array([-32768, 32767, -1, 0, 5, -5, 100, -100, 1, 2, 3, 4] * 4, np.int16)
    roundtrip(a[None, :])
    d = np.array([-2**63, 2**63 - 1, 0, -1] * 6, np.int64)
    roundtrip(d[None, :])
    e = np.array([2**64 - 1, 0, 1, 2**63] * 6, np.uint64)
    roundtrip(e[None, :])


@pytest.mark.parametrize("dt", [orc.U8, orc.U16, orc.I16, orc.U32, orc.I64])
def test_generic_blocks_and_alignment(dt):
    rng = np.random.default_rng(5 + dt)
    for block, n, frames, mis in [(12, 1001, 3, 0), (7, 500, 2, 0), (1, 77, 2, 0), (40, 999, 3, 0),
                                  (12, 1024, 2, 8), (5, 3, 4, 0), (300, 5000, 2, 0)]:
        st = np.stack([orc.kat_fill(dt, n, int(rng.integers(1, 1 << 30))) for _ in range(frames)])
        staged = roundtrip(st, block, seg_bytes=int(rng.integers(16, 300)), warm_bytes=int(rng.integers(0, 100)),
                           misalign_out=mis)
        if mis or block != 12 or (n * st.dtype.itemsize) % 16:
            assert staged is False


def test_truncated_payload_is_flagged():
    st = np.stack([orc.kat_fill(orc.U16, 12 * 500, 3)])
    p, per, pb = orc.encode_stack(st)
    cut = p[: p.size // 2]
    got, status, _, _ = emu_lib.decode(cut, st.shape[1], 1, False, np.uint16, 12, np.array([cut.size], np.uint64))
    assert status == 4


@pytest.mark.parametrize("sub_shift", [5, 6, 7, 8])
def test_checkpoint_spacing_is_a_runtime_choice(sub_shift):
    """4 / 8 / 16 / 32-byte sub-segments (sparse streams use the small ones: a thread then still owns ~6 blocks)."""
    rng = np.random.default_rng(17)
    sparse = (rng.random((3, 12 * 4096 + 4)) < 0.03).astype(np.uint16) * rng.integers(1, 9, size=(3, 12 * 4096 + 4), dtype=np.uint16)
    n = sparse.shape[1]
    while (n * 2) % 16:
        n += 1
    sparse = np.ascontiguousarray(np.pad(sparse, ((0, 0), (0, n - sparse.shape[1]))))
    assert roundtrip(sparse, seg_bytes=256, warm_bytes=128, sub_shift=sub_shift) is True
    dense = np.stack([orc.synth_frame(orc.U16, 128, 96, 2.0, 12, 500 + f) for f in range(2)])
    assert roundtrip(dense, seg_bytes=512, warm_bytes=256, sub_shift=sub_shift) is True
    wide = np.stack([orc.kat_fill(orc.U32, 12 * 400, 3 + f) for f in range(2)])
    assert roundtrip(wide, seg_bytes=1024, warm_bytes=512, sub_shift=sub_shift) is True


@pytest.mark.parametrize("fdt", [np.float32, np.float64])
def test_floating_point_outputs(fdt):
    """Decoder-only output types (Terse.hpp:379-383): through a 64-bit integer and a double, never clamped; staged
    (block 12) and generic paths, signed and unsigned streams, values past the mantissa of the output type."""
    rng = np.random.default_rng(8)
    u16 = np.stack([orc.kat_fill(orc.U16, 12 * 200 + 8, 60 + f) for f in range(3)])
    assert roundtrip(u16, out_dtype=fdt, seg_bytes=256, warm_bytes=128)
    i32 = rng.integers(-2 ** 29, 2 ** 29, (2, 12 * 64)).astype(np.int32)
    i32[0, :48] = 0
    roundtrip(i32, out_dtype=fdt, seg_bytes=256, warm_bytes=128)
    u64 = rng.integers(0, 2 ** 62, (2, 12 * 40), dtype=np.uint64)
    roundtrip(u64, out_dtype=fdt, seg_bytes=512, warm_bytes=256)
    roundtrip(u16[:, :1001], block=7, out_dtype=fdt, seg_bytes=128, warm_bytes=64)
