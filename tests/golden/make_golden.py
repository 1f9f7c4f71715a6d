#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref/libtrpx_ref.so,
built from /root/reference/include by oracle/Makefile) in the build container.  The GPU box has no
/root/reference, so the vectors are committed; re-run this script only where the reference mount
exists:   python tests/golden/make_golden.py
Every vector records what the REFERENCE produced (payload hex or FNV-1a-64 + size + prolix_bits);
the oracle (oracle/terse_oracle.c) and the CUDA path are both checked against them."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

assert orc.ref() is not None, "reference shim not built (needs /root/reference)"

NAME = {orc.U8: "u8", orc.U16: "u16", orc.U32: "u32", orc.U64: "u64",
        orc.I8: "i8", orc.I16: "i16", orc.I32: "i32", orc.I64: "i64"}


def small():
    cases = [
        ("u8 {3,4,2} block 3 (Terse.hpp:54-57)", orc.U8, [3, 4, 2], 3),
        ("i8 {-3,4,2} block 3 (Terse.hpp:54-57)", orc.I8, [-3, 4, 2], 3),
        ("u16 12x0", orc.U16, [0] * 12, 12),
        ("u16 24x0", orc.U16, [0] * 24, 12),
        ("u16 96x0 (8 header bits exactly)", orc.U16, [0] * 96, 12),
        ("u16 12x1", orc.U16, [1] * 12, 12),
        ("u16 12x1 then 12x0", orc.U16, [1] * 12 + [0] * 12, 12),
        ("u16 12x127", orc.U16, [127] * 12, 12),
        ("u16 12x255", orc.U16, [255] * 12, 12),
        ("u16 12x511", orc.U16, [511] * 12, 12),
        ("u16 12x1023", orc.U16, [1023] * 12, 12),
        ("u16 12x65535", orc.U16, [65535] * 12, 12),
        ("u16 1..14 (partial 2nd block)", orc.U16, list(range(1, 15)), 12),
        ("u32 12x0xFFFFFFFF", orc.U32, [0xFFFFFFFF] * 12, 12),
        ("u64 12x~0", orc.U64, [0xFFFFFFFFFFFFFFFF] * 12, 12),
        ("i16 12x-1", orc.I16, [-1] * 12, 12),
        ("i16 12x-4", orc.I16, [-4] * 12, 12),
        ("i32 mixed", orc.I32, [100000, -100000, 0, 1, -1, 2, -2, 3, -3, 4, -4, 5], 12),
        ("u8 37 values block 5", orc.U8, [(i * 37 + 11) % 200 for i in range(37)], 5),
        ("u16 100 values block 7", orc.U16, [(i * i * 131) % 5000 for i in range(100)], 7),
        ("i16 50 values block 16", orc.I16, [((i * 977) % 600) - 300 for i in range(50)], 16),
        ("u32 30 values block 1", orc.U32, [(i * 2654435761) % (1 << 20) for i in range(30)], 1),
    ]
    out = []
    for name, dt, vals, block in cases:
        a = np.array(vals, dtype=orc.NP_OF[dt])
        p, pb = orc.ref_encode_frame(a, block)
        out.append({"name": name, "dtype": NAME[dt], "block": block, "values": [int(v) for v in vals],
                    "prolix_bits": pb, "memory_size": int(p.size), "payload_hex": p.tobytes().hex()})
    a = np.arange(-500, 500, dtype=np.int32)          # Terse.hpp:127-154 doc example
    p, pb = orc.ref_encode_frame(a)
    out.append({"name": "int iota(-500..499) (Terse.hpp:127-154)", "dtype": "i32", "block": 12,
                "iota": [-500, 1000], "prolix_bits": pb, "memory_size": int(p.size),
                "fnv1a64": hex(orc.fnv(p))})
    return out


def large():
    out = []
    for dt, S, n in [(orc.U8, 1, 262144), (orc.U16, 1, 262144), (orc.U16, 2, 262144),
                     (orc.U32, 1, 262144), (orc.U64, 1, 65536), (orc.I8, 1, 262144),
                     (orc.I16, 1, 262144), (orc.I32, 1, 262144), (orc.I64, 1, 65536),
                     (orc.U16, 3, 1000), (orc.U16, 4, 262147), (orc.U8, 5, 100003),
                     (orc.U32, 7, 18093576), (orc.U8, 9, 23569920)]:
        a = orc.kat_fill(dt, n, S)
        if dt == orc.I64:
            # reference defect: Terse.hpp:554 calls ::abs(int) on the int64 OR, so
            # widths are only right while OR|v| < 2^31 -- keep the i64 vector inside that domain
            a = (a >> 33).astype(np.int64)
        p, pb = orc.ref_encode_frame(a)
        out.append({"gen": "kat_fill", "dtype": NAME[dt], "seed": S, "n": n, "prolix_bits": pb,
                    "memory_size": int(p.size), "fnv1a64": hex(orc.fnv(p)),
                    "first8": p[:8].tobytes().hex()})
    # BASELINE.json config C1/C2 shape: 512x512 u16 Poisson(2)+200 Bragg peaks, seeds 1000..1003
    for seed in (1000, 1001, 1002, 1003):
        a = orc.synth_frame(orc.U16, 512, 512, 2.0, 200, seed)
        p, pb = orc.ref_encode_frame(a)
        out.append({"gen": "synth", "dtype": "u16", "seed": seed, "width": 512, "height": 512,
                    "lambda": 2.0, "peaks": 200, "n": 262144, "prolix_bits": pb,
                    "memory_size": int(p.size), "fnv1a64": hex(orc.fnv(p)),
                    "first8": p[:8].tobytes().hex(), "pixel_fnv1a64": hex(orc.fnv(a.view(np.uint8)))})
    # C4-like sparse u8, C5-like signed dark-subtracted i16/i32 (small shapes)
    for dt, w, h, lam, seed in [(orc.U8, 640, 480, 0.02, 2000), (orc.U16, 640, 480, 0.02, 2001),
                                (orc.I16, 512, 512, 3.0, 3000), (orc.I32, 512, 512, 3.0, 3001),
                                (orc.U32, 512, 512, 0.5, 4000)]:
        a = orc.synth_frame(dt, w, h, lam, 50 if dt == orc.U32 else 0, seed, 20.0, 1e6)
        p, pb = orc.ref_encode_frame(a)
        out.append({"gen": "synth", "dtype": NAME[dt], "seed": seed, "width": w, "height": h,
                    "lambda": lam, "peaks": 50 if dt == orc.U32 else 0, "amp_hi": 1e6, "n": w * h,
                    "prolix_bits": pb, "memory_size": int(p.size), "fnv1a64": hex(orc.fnv(p)),
                    "first8": p[:8].tobytes().hex(), "pixel_fnv1a64": hex(orc.fnv(a.view(np.uint8)))})
    return out


def files():
    """Whole-file images written by the reference's own Terse::write (header parity)."""
    out = []
    a = orc.kat_fill(orc.U16, 1000, 3)
    buf = np.zeros(1 << 16, np.uint8)
    k = orc.ref().ref_write_file_image(orc._ptr(a), orc.U16, a.size, 12, None, 0, orc._ptr(buf), buf.size)
    out.append({"name": "single frame, no dims", "dtype": "u16", "seed": 3, "n": 1000, "frames": 1,
                "dims": [], "file_hex": buf[:k].tobytes().hex()})
    d = np.array([40, 25], np.uint64)
    k = orc.ref().ref_write_file_image(orc._ptr(a), orc.U16, a.size, 12, orc._ptr(d), 2, orc._ptr(buf), buf.size)
    out.append({"name": "single frame, dims 40 25", "dtype": "u16", "seed": 3, "n": 1000, "frames": 1,
                "dims": [40, 25], "file_hex": buf[:k].tobytes().hex()})
    st = np.concatenate([orc.kat_fill(orc.I16, 600, 11), orc.kat_fill(orc.I16, 600, 12),
                         orc.kat_fill(orc.I16, 600, 13)])
    k = orc.ref().ref_write_stack_image(orc._ptr(st), orc.I16, 600, 3, orc._ptr(buf), buf.size)
    out.append({"name": "3-frame push_back stack", "dtype": "i16", "seeds": [11, 12, 13], "n": 600,
                "frames": 3, "dims": [], "file_hex": buf[:k].tobytes().hex()})
    # stack KAT of SURVEY App. B: {u16 S=1, u16 S=2}, N=262144
    st = np.concatenate([orc.kat_fill(orc.U16, 262144, 1), orc.kat_fill(orc.U16, 262144, 2)])
    big = np.zeros(1 << 21, np.uint8)
    k = orc.ref().ref_write_stack_image(orc._ptr(st), orc.U16, 262144, 2, orc._ptr(big), big.size)
    img = big[:k].tobytes()
    h = img.index(b"/>") + 2
    out.append({"name": "App. B stack KAT", "dtype": "u16", "seeds": [1, 2], "n": 262144, "frames": 2,
                "dims": [], "header": img[:h].decode(), "payload_fnv1a64": hex(orc.fnv(np.frombuffer(img[h:], np.uint8)))})
    return out


if __name__ == "__main__":
    for name, fn in (("kat_small", small), ("kat_large", large), ("kat_files", files)):
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
