"""Tiny uncompressed greyscale TIFF writer / reader for the CLI tests (classic little-endian TIFF, one strip per
image, SampleFormat tag).  Test infrastructure only."""
import struct

import numpy as np

_FMT = {"u": 1, "i": 2, "f": 3}


def write_tiff(path, frames):
    """frames: (F, H, W) array of an integer or float dtype."""
    frames = np.ascontiguousarray(frames)
    F, H, W = frames.shape
    out = bytearray(b"II" + struct.pack("<HI", 42, 0))
    link = 4
    for f in range(F):
        if len(out) & 1:
            out += b"\0"
        data_off = len(out)
        out += frames[f].astype(frames.dtype.newbyteorder("<")).tobytes()
        if len(out) & 1:
            out += b"\0"
        ifd = len(out)
        out[link:link + 4] = struct.pack("<I", ifd)
        tags = [(256, 4, W), (257, 4, H), (258, 3, frames.dtype.itemsize * 8), (259, 3, 1), (262, 3, 1), (273, 4, data_off),
                (277, 3, 1), (278, 4, H), (279, 4, H * W * frames.dtype.itemsize), (339, 3, _FMT[frames.dtype.kind])]
        out += struct.pack("<H", len(tags))
        for tag, typ, val in tags:
            out += struct.pack("<HHI", tag, typ, 1) + (struct.pack("<I", val) if typ == 4 else struct.pack("<HH", val, 0))
        link = len(out)
        out += struct.pack("<I", 0)
    with open(path, "wb") as fh:
        fh.write(out)


def read_tiff(path):
    """-> list of 2-D arrays (one per IFD).  Little- or big-endian, strips, integer / float samples."""
    b = open(path, "rb").read()
    e = "<" if b[:2] == b"II" else ">"
    assert struct.unpack(e + "H", b[2:4])[0] == 42
    ifd = struct.unpack(e + "I", b[4:8])[0]
    imgs = []
    while ifd:
        n = struct.unpack(e + "H", b[ifd:ifd + 2])[0]
        t = {}
        for i in range(n):
            tag, typ, cnt = struct.unpack(e + "HHI", b[ifd + 2 + 12 * i: ifd + 10 + 12 * i])
            sz = {1: 1, 3: 2, 4: 4}[typ]
            raw = b[ifd + 10 + 12 * i: ifd + 14 + 12 * i]
            off = struct.unpack(e + "I", raw)[0] if sz * cnt > 4 else None
            src = b[off: off + sz * cnt] if off is not None else raw[: sz * cnt]
            t[tag] = list(struct.unpack(e + {1: "B", 3: "H", 4: "I"}[typ] * cnt, src))
        W, H, bits = t[256][0], t[257][0], t[258][0]
        kind = {1: "u", 2: "i", 3: "f"}[t.get(339, [1])[0]]
        dt = np.dtype("%s%s%d" % (e, kind, bits // 8))
        counts = t.get(279) or [W * H * bits // 8]           # (the reference's writer omits StripByteCounts)
        data = b"".join(b[o: o + c] for o, c in zip(t[273], counts))
        imgs.append(np.frombuffer(data, dt, W * H).reshape(H, W).astype(dt.newbyteorder("=")))
        ifd = struct.unpack(e + "I", b[ifd + 2 + 12 * n: ifd + 6 + 12 * n])[0]
    return imgs
