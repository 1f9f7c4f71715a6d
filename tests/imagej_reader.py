"""TEST INFRASTRUCTURE: a restatement of the reference's ImageJ plugin reader (ImageJ/TRPX_Reader.java) -- the one
independent consumer of .trpx files the reference ships.  A file this repo writes must open in it.
Plain Python loops: small images only.

  header scan     TRPX_Reader.java:41-91   first line holding "<Terse ", attributes by regex, data start = bytes before
                                           the element + the element itself (text before it counts one byte per line break)
  admissibility   :93-98                   unsigned, prolix_bits <= 16
  frame loop      :112-131                 per block: 1 / 3 / 2 / 6 header bits, `block` values of `significant_bits` bits,
                                           zero fill for width 0; the width restarts at 0 per frame; every frame ends on
                                           a byte boundary, (1 + (bit >> 3)) << 3
  ToShort         :142-150                 s bits from a 24-bit little-endian window (so s <= 16 + 1: 16-bit data only)
"""
import math
import re

import numpy as np


class NotAdmissible(Exception):
    pass


def read(raw):
    """bytes of a .trpx file -> (frames, dim1, dim0) uint16 array, as the plugin would show it."""
    data_start = 0
    attrs = None
    pos = 0
    while pos <= len(raw):
        nl = raw.find(b"\n", pos)
        line = raw[pos:] if nl < 0 else raw[pos:nl]
        i = line.find(b"<Terse ")
        if i < 0:
            data_start += len(line) + 1                       # :47-48
            if nl < 0:
                break
            pos = nl + 1
            continue
        end = line.find(b"/>", i) + 2                         # :50
        data_start += end
        attrs = line[i:end].decode("latin-1")
        break
    if attrs is None:
        raise ValueError("no <Terse .../> element")

    def num(name, default=None):
        m = re.search(name + r'="(\d+)"', attrs)
        if m is None:
            if default is None:
                raise ValueError("attribute " + name + " missing")
            return default
        return int(m.group(1))

    prolix_bits, signed, block = num("prolix_bits"), num("signed"), num("block")
    size, n_values, n_frames = num("memory_size"), num("number_of_values"), num("number_of_frames", 1)
    m = re.search(r'dimensions="(\d+)(?:\s+(\d+))?(?:\s+(\d+))?"', attrs)
    if m:
        dim0 = int(m.group(1))
        dim1 = int(m.group(2)) if m.group(2) else 0
    else:
        dim0 = dim1 = int(math.sqrt(n_values))                # :84-86
    if signed != 0 or prolix_bits > 16:
        raise NotAdmissible("images must be unsigned 16 bit")  # :93-98
    buf = raw[:data_start + size] + b"\0\0\0"
    bit = data_start * 8

    def to_short(s):                                          # :142-150
        nonlocal bit
        idx = bit >> 3
        w = buf[idx] | (buf[idx + 1] << 8) | (buf[idx + 2] << 16)
        v = (w >> (bit & 7)) & ((1 << s) - 1)
        bit += s
        return v & 0xFFFF

    out = np.zeros((n_frames, n_values), np.uint16)
    for f in range(n_frames):
        sig = 0
        for lo in range(0, n_values, block):
            if to_short(1) == 0:
                sig = to_short(3)
                if sig == 7:
                    sig += to_short(2)
                    if sig == 10:
                        sig += to_short(6)
            hi = min(n_values, lo + block)
            if sig:
                for j in range(lo, hi):
                    out[f, j] = to_short(sig)
        bit = (1 + (bit >> 3)) << 3                           # :131
    return out.reshape(n_frames, dim1, dim0) if dim0 * dim1 == n_values else out
