"""Pins the CPU oracle (oracle/terse_oracle.c): against the committed golden vectors that the
reference produced (tests/golden/) and, where oracle/_ref is built, against the
live reference on fuzzed inputs inside the parity domain."""
import numpy as np
import pytest

import golden_util as G
import orc


@pytest.mark.parametrize("c", G.load("kat_small"), ids=lambda c: c["name"])
def test_small_kats(c):
    a = G.small_input(c)
    p, pb = orc.encode_frame(a, c["block"])
    assert pb == c["prolix_bits"] and p.size == c["memory_size"]
    if "payload_hex" in c:
        assert p.tobytes().hex() == c["payload_hex"]
    else:
        assert hex(orc.fnv(p)) == c["fnv1a64"]
    signed = c["dtype"][0] == "i"
    d, used = orc.decode_frame(p, a.size, signed, a.dtype, c["block"])
    assert used == p.size and np.array_equal(d, a)


@pytest.mark.parametrize("c", [c for c in G.load("kat_large") if c["n"] <= 400000],
                         ids=lambda c: "%s-%s-%d-%d" % (c["gen"], c["dtype"], c["seed"], c["n"]))
def test_large_kats(c):
    a = G.large_input(c)
    if "pixel_fnv1a64" in c:
        assert hex(orc.fnv(a.view(np.uint8))) == c["pixel_fnv1a64"]   # generator is reproducible
    p, pb = orc.encode_frame(a)
    assert (pb, p.size, hex(orc.fnv(p)), p[:8].tobytes().hex()) == \
           (c["prolix_bits"], c["memory_size"], c["fnv1a64"], c["first8"])
    d, used = orc.decode_frame(p, a.size, c["dtype"][0] == "i", a.dtype)
    assert used == p.size and np.array_equal(d, a)
    w, used = orc.frame_widths(p, a.size)
    assert used == p.size and int(w.max()) == pb


def test_stack_is_concatenation_and_header():
    files = G.load("kat_files")
    c = files[3]                                      # App. B stack KAT
    st = np.stack([orc.kat_fill(orc.U16, c["n"], s) for s in c["seeds"]])
    p, per, pb = orc.encode_stack(st)
    hdr = orc.header(pb, False, 12, p.size, c["n"], [], c["frames"])
    assert hdr.decode() == c["header"] and hex(orc.fnv(p)) == c["payload_fnv1a64"]
    assert int(per.sum()) == p.size
    for c in files[:3]:
        img = bytes.fromhex(c["file_hex"])
        h = img.index(b"/>") + 2
        seeds = c.get("seeds", [c.get("seed")])
        dt = G.CODE[c["dtype"]]
        st = np.stack([orc.kat_fill(dt, c["n"], s) for s in seeds])
        p, per, pb = orc.encode_stack(st)
        assert p.tobytes() == img[h:]
        assert orc.header(pb, dt >= 4, 12, p.size, c["n"], c["dims"], c["frames"]) == img[:h]


def test_conversions_clamp_and_widen():
    a = np.array([300, 1, 2, 0, 65535, 2, 3, 0, 1, 2, 3, 1], np.uint16)
    p, _ = orc.encode_frame(a)
    assert np.array_equal(orc.decode_frame(p, 12, False, np.uint8)[0], np.minimum(a, 255))
    assert np.array_equal(orc.decode_frame(p, 12, False, np.uint64)[0], a.astype(np.uint64))
    b = np.array([-300, 1, 2, 0, 1, 2, 300, 0, 1, 2, 3, 1], np.int16)
    p, _ = orc.encode_frame(b)
    assert np.array_equal(orc.decode_frame(p, 12, True, np.int8)[0], np.clip(b, -128, 127))
    assert np.array_equal(orc.decode_frame(p, 12, True, np.int64)[0], b.astype(np.int64))


def test_truncated_stream_is_rejected():
    a = orc.kat_fill(orc.U16, 1000, 3)
    p, _ = orc.encode_frame(a)
    assert orc.decode_frame(p[: p.size // 2], 1000, False, np.uint16)[1] == 0


needs_ref = pytest.mark.skipif(orc.ref() is None, reason="oracle/_ref not built (no reference mount)")


@needs_ref
@pytest.mark.parametrize("dt", list(range(8)), ids=lambda d: str(np.dtype(orc.NP_OF[d])))
def test_fuzz_against_live_reference(dt):
    """Differential fuzz inside the parity domain: unsigned all values; signed |v| < 2^(W-2)
    (App. C4); N >= block for incompressible data (App. C7)."""
    rng = np.random.default_rng(1234 + dt)
    W = 8 * np.dtype(orc.NP_OF[dt]).itemsize
    for it in range(150):
        block = 12 if it % 3 else int(rng.integers(1, 41))
        n = int(rng.integers(block, 400))
        top = 31 if dt == orc.I64 else (W - 1 if dt >= 4 else W + 1)   # i64: DESIGN.md C11
        bits = rng.integers(0, top, size=(n + block - 1) // block)
        bits = np.repeat(bits, block)[:n]
        hi = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
        mag = np.where(bits == 0, 0, hi >> (np.uint64(64) - np.maximum(bits, 1).astype(np.uint64)))
        if dt >= 4:
            sign = rng.integers(0, 2, size=n).astype(bool)
            a = np.where(sign, -mag.astype(np.int64), mag.astype(np.int64)).astype(orc.NP_OF[dt])
        else:
            a = mag.astype(orc.NP_OF[dt])
        p, pb = orc.encode_frame(a, block)
        q, qb = orc.ref_encode_frame(a, block)
        assert pb == qb and np.array_equal(p, q), (it, block, n)
        d, used = orc.decode_frame(p, n, dt >= 4, a.dtype, block)
        assert used == p.size and np.array_equal(d, a)
        # reference DECODE is only sound for widths < min(W, 32) (App. C5; for 64-bit outputs the
        # byte refill `buffer << n` is an int shift, Bit_pointer.hpp:775) -- encode has no such limit
        if W <= 16 or int(pb) < min(W, 32):
            r = orc.ref_decode_frame(p, n, dt >= 4, pb, a.dtype, block)
            assert np.array_equal(r, a)


@needs_ref
def test_reference_decodes_oracle_streams_with_clamp():
    b = np.array([-300, 1, 2, 0, 1, 2, 300, 0, 1, 2, 3, 1] * 3, np.int16)
    p, pb = orc.encode_frame(b)
    assert np.array_equal(orc.ref_decode_frame(p, b.size, True, pb, np.int8),
                          orc.decode_frame(p, b.size, True, np.int8)[0])
    a = np.array([300, 1, 2, 0, 65535, 2, 3, 0, 1, 2, 3, 1], np.uint16)
    p, pb = orc.encode_frame(a)
    assert np.array_equal(orc.ref_decode_frame(p, 12, False, pb, np.uint8),
                          orc.decode_frame(p, 12, False, np.uint8)[0])


@pytest.mark.parametrize("fdt", [np.float32, np.float64])
def test_floating_point_outputs_follow_the_reference(fdt):
    """Terse.hpp:379-383: a floating-point iterator receives double(int64/uint64(field)).  Checked against the live
    reference where it was built (oracle/_ref), inside its parity domain (values below 2^31: its wide reads are
    undefined beyond)."""
    if orc.ref() is None:
        pytest.skip("oracle/_ref not built (no reference mount)")
    rng = np.random.default_rng(3)
    for dt, lo, hi in ((np.uint16, 0, 60000), (np.int16, -3000, 3000), (np.uint32, 0, 2 ** 30), (np.int32, -2 ** 27, 2 ** 27)):
        a = rng.integers(lo, hi, 12 * 40 + 5).astype(dt)
        a[24:60] = 0
        p, pb = orc.ref_encode_frame(a)
        want = orc.ref_decode_frame(p, a.size, a.dtype.kind == "i", pb, fdt)
        got, _ = orc.decode_frame(p, a.size, a.dtype.kind == "i", fdt)
        assert np.array_equal(got, want) and np.array_equal(want, a.astype(np.float64).astype(fdt))
