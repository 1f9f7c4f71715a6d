"""Generates tests/golden/kat_cli.json with the REFERENCE CLIs (oracle/_ref/terse_ref, prolix_ref: the reference's
src/terse.cpp and src/prolix.cpp compiled unchanged by oracle/Makefile).  Run in the build container only:
  python tests/golden/make_golden_cli.py
For each case a TIFF is written (tests/tiff_util.py), `terse_ref` turns it into a .trpx, and the header text, payload
size and FNV-1a-64 are recorded; `prolix_ref` turns it back and the FNV of the pixels of frame 0 is recorded."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402
import tiff_util  # noqa: E402

REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref")

CASES = [
    dict(name="C1: one 512x512 u16 diffraction frame", dtype="u16", width=512, height=512, frames=1, lam=2.0, peaks=200, seed=1000),
    dict(name="3-frame 640x480 u16 stack", dtype="u16", width=640, height=480, frames=3, lam=2.0, peaks=50, seed=2000),
    dict(name="2-frame 256x256 i16 dark-subtracted", dtype="i16", width=256, height=256, frames=2, lam=3.0, peaks=0, seed=3000),
    dict(name="one 300x200 u32 frame", dtype="u32", width=300, height=200, frames=1, lam=0.5, peaks=30, seed=4000, amp_hi=1000000.0),
    dict(name="one 128x128 u8 frame", dtype="u8", width=128, height=128, frames=1, lam=0.02, peaks=0, seed=5000),
]
CODE = {"u8": orc.U8, "u16": orc.U16, "u32": orc.U32, "i16": orc.I16}


def frames_of(c):
    return np.stack([orc.synth_frame(CODE[c["dtype"]], c["width"], c["height"], c["lam"], c["peaks"], c["seed"] + f, 20.0,
                                     c.get("amp_hi", 3000.0)).reshape(c["height"], c["width"]) for f in range(c["frames"])])


def main():
    out = []
    for c in CASES:
        st = frames_of(c)
        with tempfile.TemporaryDirectory() as d:
            tif = os.path.join(d, "x.tif")
            tiff_util.write_tiff(tif, st)
            subprocess.run([os.path.join(REF, "terse_ref"), tif], check=True, stdout=subprocess.DEVNULL)
            img = open(os.path.join(d, "x.trpx"), "rb").read()
            h = img.index(b"/>") + 2
            payload = np.frombuffer(img[h:], np.uint8)
            # the reference decoder mis-addresses frames >= 2 of a stack and may crash there
            r = subprocess.run([os.path.join(REF, "prolix_ref"), os.path.join(d, "x.trpx")], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)
            ref = dict(ref_prolix_rc=r.returncode)
            if r.returncode == 0 and os.path.exists(tif):
                back = tiff_util.read_tiff(tif)
                ref.update(ref_prolix_dtype=str(back[0].dtype), ref_prolix_frames=len(back),
                           ref_prolix_frames_ok=[bool(np.array_equal(back[f].astype(np.int64), st[f].astype(np.int64)))
                                                 for f in range(min(len(back), c["frames"]))])
        e = dict(c)
        e.update(header=img[:h].decode(), memory_size=int(payload.size), payload_fnv1a64=hex(orc.fnv(payload)), **ref)
        out.append(e)
        print(e["name"], e["header"], e["payload_fnv1a64"], ref)
    with open(os.path.join(HERE, "kat_cli.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
