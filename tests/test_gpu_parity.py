"""Parity tests proper: the CUDA path, called through the C ABI (libtrpx_b200.so, include/trpx_b200.h),
against the CPU oracle (== reference bytes, tests/test_oracle.py) and the committed golden vectors.
Bit-exact: TERSE payload bytes, per-frame sizes and prolix_bits must equal the oracle's; PROLIX must
return the original pixels (and the oracle's values for converting decodes).

Needs a GPU; nothing here reads /root/reference."""
import ctypes as C

import numpy as np
import pytest

import golden_util as G
import orc
import trpx_b200

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    c = trpx_b200.Codec(0)
    yield c
    c.close()


def enc_check(codec, stack, block=12):
    stack = np.ascontiguousarray(stack)
    p, fb, pb = codec.encode(stack, block)
    q, per, qb = orc.encode_stack(stack, block)
    assert pb == qb
    assert np.array_equal(fb, per)
    assert p.size == q.size and np.array_equal(p, q)
    return p, fb, pb


def roundtrip(codec, stack, block=12, out_dtype=None, known_sizes=True):
    stack = np.ascontiguousarray(stack)
    F, N = stack.shape
    p, fb, pb = enc_check(codec, stack, block)
    od = stack.dtype if out_dtype is None else np.dtype(out_dtype)
    sgn = stack.dtype.kind == "i"
    got, fb2 = codec.decode(p, N, F, sgn, od, block, fb if known_sizes else None)
    assert np.array_equal(fb2, fb)
    ends = np.cumsum(fb).astype(np.int64)
    if od == stack.dtype:
        assert np.array_equal(got, stack)
    else:
        want = np.stack([orc.decode_frame(p[int(ends[f] - fb[f]):int(ends[f])], N, sgn, od, block)[0] for f in range(F)])
        assert np.array_equal(got, want)


# ---------------------------------------------------------------- golden vectors (reference-produced)
@pytest.mark.parametrize("c", G.load("kat_small"), ids=lambda c: c["name"])
def test_small_kats(codec, c):
    a = G.small_input(c)
    p, fb, pb = codec.encode(a[None, :], c["block"])
    assert pb == c["prolix_bits"] and p.size == c["memory_size"]
    if "payload_hex" in c:
        assert p.tobytes().hex() == c["payload_hex"]
    else:
        assert hex(orc.fnv(p)) == c["fnv1a64"]
    d, _ = codec.decode(p, a.size, 1, c["dtype"][0] == "i", a.dtype, c["block"])
    assert np.array_equal(d[0], a)


@pytest.mark.parametrize("c", G.load("kat_large"),
                         ids=lambda c: "%s-%s-%d-%d" % (c["gen"], c["dtype"], c["seed"], c["n"]))
def test_large_kats(codec, c):
    a = G.large_input(c)
    p, fb, pb = codec.encode(a[None, :])
    assert (pb, p.size, hex(orc.fnv(p)), p[:8].tobytes().hex()) == \
           (c["prolix_bits"], c["memory_size"], c["fnv1a64"], c["first8"])
    d, _ = codec.decode(p, a.size, 1, c["dtype"][0] == "i", a.dtype)
    assert np.array_equal(d[0], a)


def test_stack_kat_is_concatenation(codec):
    c = G.load("kat_files")[3]
    st = np.stack([orc.kat_fill(orc.U16, c["n"], s) for s in c["seeds"]])
    p, fb, pb = codec.encode(st)
    assert hex(orc.fnv(p)) == c["payload_fnv1a64"] and int(fb.sum()) == p.size
    assert orc.header(pb, False, 12, p.size, c["n"], [], c["frames"]).decode() == c["header"]
    d, fb2 = codec.decode(p, c["n"], 2, False, np.uint16)          # frame sizes recovered from the stream
    assert np.array_equal(d, st) and np.array_equal(fb2, fb)


# ---------------------------------------------------------------- differential vs the oracle
@pytest.mark.parametrize("dt", list(range(8)), ids=lambda d: str(np.dtype(orc.NP_OF[d])))
def test_all_types_fast_path(codec, dt):
    isz = orc.NP_OF[dt]().itemsize
    n = 12 * 5000 + 8
    while (n * isz) % 16:
        n += 1
    st = np.stack([orc.kat_fill(dt, n, 40 + f) for f in range(5)])
    roundtrip(codec, st)


def test_c2_shape_frames(codec):
    """512x512 u16 Poisson + Bragg peaks (BASELINE configs[0]/[1] shape), 24 frames."""
    st = np.stack([orc.synth_frame(orc.U16, 512, 512, 2.0, 200, 1000 + f) for f in range(24)])
    roundtrip(codec, st)
    roundtrip(codec, st[:3], known_sizes=False)


def test_foreign_stack_frame_sizes_are_recovered_from_the_stream(codec):
    """A multi-frame payload as it comes out of a .trpx file: only the total size is known (Terse.hpp:459)."""
    import time
    st = np.stack([orc.synth_frame(orc.U16, 512, 512, 2.0, 100, 5000 + f) for f in range(48)])
    st[7] = 0                                                      # an empty frame: one long run of 1-bit blocks
    p, fb, pb = codec.encode(st)
    t0 = time.time()
    d, fb2 = codec.decode(p, st.shape[1], st.shape[0], False, np.uint16)
    dt = time.time() - t0
    assert np.array_equal(fb2, fb) and np.array_equal(d, st)
    assert dt < 5.0, "frame-boundary recovery took %.1f s" % dt
    d, _ = codec.decode(p, st.shape[1], st.shape[0], False, np.uint16, first_frame=40, n_frames=3)
    assert np.array_equal(d, st[40:43])


def test_foreign_stack_many_frames_mixed_content(codec):
    """Several batches of the speculative frame chain (64 frames each), with what makes its windows miss or its candidates
    give up: empty frames, frames of constant width (T and G never re-synchronise: the serial chain takes over), dense and
    sparse frames side by side, a ragged last block.  The recovered sizes must be the encoder's, frame by frame."""
    rng = np.random.default_rng(5)
    n = 128 * 130 + 5
    frames = []
    for f in range(330):
        kind = f % 11
        if kind == 3:
            a = np.zeros(n, np.uint16)
        elif kind == 7:
            a = rng.integers(0, 4, n).astype(np.uint16)
            a[::12] = 3                                            # every block 2 bits wide: nothing but "same width" headers
        elif kind == 9:
            a = (rng.random(n) < 0.002).astype(np.uint16) * 9     # sparse
        else:
            a = rng.poisson(2.0 + (f % 5), n).astype(np.uint16)
        frames.append(a)
    st = np.stack(frames)
    p, fb, pb = codec.encode(st)
    d, fb2 = codec.decode(p, n, st.shape[0], False, np.uint16)
    assert np.array_equal(fb2, fb)
    assert np.array_equal(d, st)
    with pytest.raises(trpx_b200.TrpxError) as e:                  # the stream ends inside frame 200: flagged, no hang
        codec.decode(p[:int(fb[:200].sum()) + 7], n, st.shape[0], False, np.uint16)
    assert e.value.status == trpx_b200.ERR_MALFORMED


def test_signed_dark_subtracted_frames(codec):
    for dt in (orc.I16, orc.I32):
        st = np.stack([orc.synth_frame(dt, 512, 512, 3.0, 0, 77 + f) for f in range(6)])
        roundtrip(codec, st)


def test_sparse_counting_frames(codec):
    rng = np.random.default_rng(11)
    for dt in (np.uint8, np.uint16):
        st = (rng.random((3, 1536 * 1024)) < 0.02).astype(dt)
        st[1] = 0
        roundtrip(codec, st)
    roundtrip(codec, np.zeros((7, 512 * 512), np.uint16))


def test_tiny_and_ragged_frames(codec):
    for n in (1, 3, 8, 11, 12, 13, 24, 100, 1000, 4097):
        st = np.stack([orc.kat_fill(orc.U16, n, 7 + f) for f in range(9)])
        roundtrip(codec, st)
        roundtrip(codec, st, known_sizes=False)
    roundtrip(codec, np.zeros((50, 8), np.uint16))
    roundtrip(codec, np.zeros((1, 96), np.uint16))                 # 'ff 00': the extra byte rule


@pytest.mark.parametrize("dt", [orc.U8, orc.U16, orc.I16, orc.U32, orc.I64])
def test_generic_blocks(codec, dt):
    rng = np.random.default_rng(5 + dt)
    for block, n, frames in [(12, 1001, 3), (7, 500, 2), (1, 77, 2), (40, 999, 3), (5, 3, 4), (300, 5000, 2),
                             (13, 100003, 2)]:
        st = np.stack([orc.kat_fill(dt, n, int(rng.integers(1, 1 << 30))) for _ in range(frames)])
        roundtrip(codec, st, block)


def test_signed_extremes(codec):
    a = np.array([-32768, 32767, -1, 0, 5, -5, 100, -100, 1, 2, 3, 4] * 4, np.int16)
    roundtrip(codec, a[None, :])
    roundtrip(codec, np.array([-2**63, 2**63 - 1, 0, -1] * 6, np.int64)[None, :])
    roundtrip(codec, np.array([2**64 - 1, 0, 1, 2**63] * 6, np.uint64)[None, :])
    roundtrip(codec, np.array([2**32 - 1] * 12, np.uint32)[None, :])   # s == W (reference decode fails, App. C5)


@pytest.mark.parametrize("src,dst", [(np.uint16, np.uint8), (np.uint16, np.uint64), (np.uint16, np.int32),
                                     (np.int16, np.int8), (np.int16, np.int64), (np.uint32, np.uint16),
                                     (np.int32, np.int16), (np.uint8, np.uint32), (np.int64, np.int32)])
def test_output_conversion_clamps_like_get_range(codec, src, dst):
    rng = np.random.default_rng(3)
    info = np.iinfo(src)
    a = rng.integers(max(info.min, -70000), min(info.max, 70000), size=(2, 12 * 3000), endpoint=True).astype(src)
    a[:, ::50] = info.max
    if info.min < 0:
        a[:, 7::50] = info.min + 1
    roundtrip(codec, a, out_dtype=dst)


def test_partial_decode_of_a_stack(codec):
    st = np.stack([orc.kat_fill(orc.U16, 5000, 7 + f) for f in range(10)])
    p, fb, pb = codec.encode(st)
    d, _ = codec.decode(p, 5000, 10, False, np.uint16, frame_bytes=fb, first_frame=3, n_frames=4)
    assert np.array_equal(d, st[3:7])
    d, _ = codec.decode(p, 5000, 10, False, np.uint16, first_frame=9, n_frames=1)
    assert np.array_equal(d, st[9:10])


# ---------------------------------------------------------------- error behaviour
def test_errors(codec):
    st = np.stack([orc.kat_fill(orc.U16, 12 * 800, 3)])
    with pytest.raises(trpx_b200.TrpxError) as e:
        codec.encode(st, capacity=1024)
    assert e.value.status == trpx_b200.ERR_CAPACITY
    p, fb, pb = codec.encode(st)
    with pytest.raises(trpx_b200.TrpxError) as e:
        codec.decode(p[: p.size // 2], st.shape[1], 1, False, np.uint16)
    assert e.value.status == trpx_b200.ERR_MALFORMED
    with pytest.raises(trpx_b200.TrpxError) as e:                  # signed stream into unsigned output
        codec.decode(p, st.shape[1], 1, True, np.uint16)
    assert e.value.status == trpx_b200.ERR_BAD_ARG
    # still usable afterwards
    roundtrip(codec, st)


def test_garbage_payloads_never_crash(codec):
    """The reference reads out of bounds on corrupt input (no checks at all); here a corrupt stream may decode to
    garbage or be flagged TRPX_ERR_MALFORMED, but it must neither fault nor hang, and the context stays usable."""
    rng = np.random.default_rng(99)
    n = 512 * 512
    good = np.stack([orc.synth_frame(orc.U16, 512, 512, 2.0, 50, 300 + f) for f in range(4)])
    p, fb, pb = codec.encode(good)
    cases = [rng.integers(0, 256, p.size, dtype=np.uint8),                 # pure noise
             np.zeros(p.size, np.uint8), np.full(p.size, 0xFF, np.uint8)]  # all-explicit-zero headers / one endless run
    flipped = p.copy()
    flipped[rng.integers(0, p.size, 200)] ^= 0x5A                          # a damaged real stream
    cases.append(flipped)
    for bad in cases:
        for sizes in (fb, None):
            try:
                d, _ = codec.decode(bad, n, 4, False, np.uint16, frame_bytes=sizes)
                assert d.shape == (4, n)
            except trpx_b200.TrpxError as e:
                assert e.status == trpx_b200.ERR_MALFORMED
    d, _ = codec.decode(p, n, 4, False, np.uint16, frame_bytes=fb)
    assert np.array_equal(d, good)


# ---------------------------------------------------------------- device-pointer flavour + big shapes
def test_device_flavour_large_stack_properties(codec):
    """512x512 u16 x 2000 frames resident in HBM: round trip, frame-size bookkeeping, sample frames
    byte-identical to the oracle, and the stack payload == concatenation of independently encoded
    halves (frames are independent: the property multi-GPU sharding relies on)."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    F, N = 2000, 512 * 512
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    px = torch.poisson(torch.full((F, N), 2.0, device=dev), generator=g).to(torch.int32)
    hot = torch.randint(0, N, (F, 64), device=dev, generator=g)
    px.scatter_(1, hot, torch.randint(20, 3000, (F, 64), device=dev, generator=g, dtype=torch.int32))
    px = px.to(torch.uint16)
    cap = trpx_b200.max_compressed_bytes(N, np.uint16, 12, F)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    ends = torch.zeros(F, dtype=torch.int64, device=dev)
    small = torch.zeros(4, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    codec.encode_device(px.data_ptr(), np.uint16, N, F, out.data_ptr(), cap, ends.data_ptr(), small.data_ptr(),
                        small.data_ptr() + 4, s)
    torch.cuda.synchronize()
    assert int(small[1]) == 0
    e = ends.cpu().numpy()
    assert np.all(np.diff(e) > 0)
    total = int(e[-1])
    # sample frames against the oracle
    host = out[:total].cpu().numpy()
    for f in (0, 1, 999, 1999):
        q, qb = orc.encode_frame(px[f].cpu().numpy())
        lo = int(e[f - 1]) if f else 0
        assert np.array_equal(host[lo:int(e[f])], q)
        assert qb <= int(small[0])
    # halves encode independently to the same bytes
    out2 = torch.empty(cap, dtype=torch.uint8, device=dev)
    ends2 = torch.zeros(F, dtype=torch.int64, device=dev)
    h = F // 2
    codec.encode_device(px[h:].data_ptr(), np.uint16, N, F - h, out2.data_ptr(), cap, ends2.data_ptr(),
                        small.data_ptr() + 8, small.data_ptr() + 12, s)
    torch.cuda.synchronize()
    n2 = int(ends2[F - h - 1])
    assert n2 == total - int(e[h - 1])
    assert torch.equal(out2[:n2], out[int(e[h - 1]):total])
    # decode on the device, known frame ends
    back = torch.empty((F, N), dtype=torch.uint16, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    codec.decode_device(out.data_ptr(), total, False, N, F, ends.data_ptr(), back.data_ptr(), np.uint16,
                        st.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(st[0]) == 0
    assert torch.equal(back.view(torch.int16), px.view(torch.int16))


@pytest.mark.parametrize("shape,dt,lam", [((4148, 4362), np.uint32, 0.5), ((5760, 4092), np.uint8, 0.02),
                                          ((5760, 4092), np.uint16, 0.02)])
def test_big_frames_roundtrip_and_oracle(codec, shape, dt, lam):
    """Eiger2-16M-class u32 and cryo-EM movie frames (BASELINE configs[2], [3]): 2 frames each, oracle
    bytes on the first, round trip on both."""
    rng = np.random.default_rng(5)
    n = shape[0] * shape[1]
    st = rng.poisson(lam, size=(2, n)).astype(dt)
    if dt == np.uint32:
        st[0, rng.integers(0, n, 2000)] = rng.integers(1000, 1000000, 2000)
        st[1, ::100003] = 0xFFFFFFFF                               # detector-gap pixels: s == 32
    p, fb, pb = codec.encode(st)
    q, qb = orc.encode_frame(st[0])
    assert np.array_equal(p[: q.size], q) and int(fb[0]) == q.size
    d, fb2 = codec.decode(p, n, 2, False, dt, frame_bytes=fb)
    assert np.array_equal(d, st)
    d1, _ = codec.decode(p[: int(fb[0])], n, 1, False, dt)
    assert np.array_equal(d1[0], st[0])


# ---------------------------------------------------------------- host pipelines: many batches, lane reuse, two contexts
def test_host_flavour_many_batches_and_lane_reuse(monkeypatch):
    """One frame per batch (TRPX_BATCH_MB=0) and few lanes: every lane is reused many times, payload copies of
    uneven sizes land back to back in `out`, every batch has its own status word and frame-end slice."""
    monkeypatch.setenv("TRPX_BATCH_MB", "0")
    monkeypatch.setenv("TRPX_ENC_LANES", "2")
    monkeypatch.setenv("TRPX_DEC_LANES", "3")
    c = trpx_b200.Codec(0)
    try:
        rng = np.random.default_rng(5)
        st = np.stack([orc.kat_fill(orc.U16, 12 * 1000 + 8, 300 + f) >> int(rng.integers(0, 12)) for f in range(37)])
        st[11] = 0
        roundtrip(c, st)
        p, fb, pb = c.encode(st)
        d, _ = c.decode(p, st.shape[1], st.shape[0], False, np.uint16, frame_bytes=fb, first_frame=5, n_frames=29)
        assert np.array_equal(d, st[5:34])
        bad = p.copy()
        bad[int(fb[:20].sum()) + 3:int(fb[:21].sum())] = 0xFF          # frame 20 (its own batch) no longer adds up
        with pytest.raises(trpx_b200.TrpxError) as e:
            c.decode(bad, st.shape[1], st.shape[0], False, np.uint16, frame_bytes=fb)
        assert e.value.status == trpx_b200.ERR_MALFORMED
        sst = np.stack([orc.synth_frame(orc.I16, 64, 96, 3.0, 0, 9 + f) for f in range(9)])
        roundtrip(c, sst)
    finally:
        c.close()


def test_two_contexts_stream_chunks_concurrently(codec):
    """bench.py's streamed e2e in miniature: one host thread encodes chunk k+1 on its context while a second thread
    decodes chunk k on another context; the concatenated payload equals the oracle's."""
    import queue
    import threading
    st = np.stack([orc.synth_frame(orc.U16, 256, 256, 2.0, 50, 700 + f) for f in range(24)])
    other = trpx_b200.Codec(0)
    q, parts, back = queue.Queue(), [], {}

    def enc_side():
        for k in range(0, 24, 4):
            p, fb, pb = codec.encode(st[k:k + 4])
            parts.append(p)
            q.put((k, p, fb))
        q.put(None)

    def dec_side():
        while (it := q.get()) is not None:
            k, p, fb = it
            back[k] = other.decode(p, st.shape[1], 4, False, np.uint16, frame_bytes=fb)[0]

    th = [threading.Thread(target=enc_side), threading.Thread(target=dec_side)]
    [t.start() for t in th]
    [t.join() for t in th]
    other.close()
    want, _, _ = orc.encode_stack(st)
    assert np.array_equal(np.concatenate(parts), want)
    assert np.array_equal(np.concatenate([back[k] for k in range(0, 24, 4)]), st)


@pytest.mark.parametrize("fdt", [np.float32, np.float64])
def test_floating_point_outputs(codec, fdt):
    """TRPX_F32 / TRPX_F64 (decoder outputs only, Terse.hpp:379-383) on the staged and the generic path; the encoder
    refuses them."""
    rng = np.random.default_rng(8)
    st = np.stack([orc.synth_frame(orc.U16, 256, 256, 2.0, 50, 40 + f) for f in range(5)])
    roundtrip(codec, st, out_dtype=fdt)
    roundtrip(codec, np.stack([orc.synth_frame(orc.I32, 128, 96, 3.0, 0, 9 + f) for f in range(3)]), out_dtype=fdt)
    roundtrip(codec, rng.integers(0, 2 ** 62, (2, 12 * 400), dtype=np.uint64), out_dtype=fdt)
    roundtrip(codec, rng.integers(-2 ** 40, 2 ** 40, (2, 12 * 400)).astype(np.int64), out_dtype=fdt)
    roundtrip(codec, st[:2, :5001], block=7, out_dtype=fdt)
    L = trpx_b200.lib()
    a = np.zeros(24, fdt)
    out = np.zeros(4096, np.uint8)
    tot, pb = C.c_size_t(0), C.c_uint(0)
    rc = L.trpx_encode_host(codec._h, a.ctypes.data, trpx_b200.dtype_code(fdt), 24, 1, 12, out.ctypes.data, out.size, None,
                            C.byref(tot), C.byref(pb))
    assert rc == trpx_b200.ERR_BAD_ARG
    assert L.trpx_max_compressed_bytes(24, trpx_b200.dtype_code(fdt), 12, 1) == 0


def test_encode_progress_is_a_consistent_prefix(monkeypatch):
    """trpx_ctx_encode_progress from a second thread while trpx_encode_host runs (one frame per batch): the reported
    prefix only grows, its payload bytes are the oracle's prefix bytes, and it ends at the whole stack."""
    import threading
    monkeypatch.setenv("TRPX_BATCH_MB", "0")
    c = trpx_b200.Codec(0)
    try:
        st = np.stack([orc.synth_frame(orc.U16, 128, 128, 2.0, 20, 900 + f) for f in range(60)])
        want, per, _ = orc.encode_stack(st)
        cum = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
        L = trpx_b200.lib()
        F, N = st.shape
        cap = L.trpx_max_compressed_bytes(N, trpx_b200.U16, 12, F)
        out = np.zeros(cap, np.uint8)
        fb = np.zeros(F, np.uint64)
        tot, pb = C.c_size_t(0), C.c_uint(0)
        seq0 = c.encode_progress()[0]
        seen, stop = [], threading.Event()

        def watcher():
            while not stop.is_set():
                q, f, b = c.encode_progress()
                if q != seq0:
                    seen.append((f, b, bytes(out[:b]) == bytes(want[:b]), np.array_equal(fb[:f], per[:f])))

        th = threading.Thread(target=watcher)
        th.start()
        rc = L.trpx_encode_host(c._h, st.ctypes.data, trpx_b200.U16, N, F, 12, out.ctypes.data, cap, fb.ctypes.data,
                                C.byref(tot), C.byref(pb))
        stop.set()
        th.join()
        assert rc == 0 and tot.value == want.size
        assert c.encode_progress()[1:] == (F, want.size)
        assert seen and all(ok1 and ok2 for _, _, ok1, ok2 in seen)
        assert all(a[0] <= b[0] and a[1] <= b[1] for a, b in zip(seen, seen[1:]))
        assert all(b == cum[f] for f, b, _, _ in seen)
    finally:
        c.close()


def test_unsigned_stream_into_signed_types_keeps_its_values(codec):
    """Deliberate difference to the reference (trpx_b200.h): sign extension follows the STREAM's signedness, so unsigned
    data decoded into a signed type -- wider or as wide -- is value preserving; into a narrower signed type it clamps.
    (The reference's get_range sign-extends whenever the OUTPUT type is signed, Bit_pointer.hpp:784-789.)"""
    a = np.array([5, 7, 4, 6, 5, 7, 4, 6, 5, 7, 4, 6] * 50 + [40000, 65535, 3, 0] * 6, np.uint16)   # top bit of the block width set
    p, fb, pb = codec.encode(a[None, :])
    for out in (np.int32, np.int64, np.int16, np.int8):
        got, _ = codec.decode(p, a.size, 1, False, out, frame_bytes=fb)
        want, used = orc.decode_frame(p, a.size, False, out)
        assert np.array_equal(got[0], want)
        if np.dtype(out).itemsize > 2:
            assert np.array_equal(got[0], a.astype(out))                 # nothing turns negative
    got16, _ = codec.decode(p, a.size, 1, False, np.int16, frame_bytes=fb)
    assert got16[0][600] == np.int16(-25536) and got16[0][0] == 5       # same-width: the documented wrap (Terse.hpp:113-115)
