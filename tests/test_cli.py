"""The `terse` / `prolix` command-line tools on top of the GPU codec (cxx/terse.cpp, cxx/prolix.cpp), against golden
vectors made by the REFERENCE CLIs (tests/golden/kat_cli.json, generator tests/golden/make_golden_cli.py): same .trpx
bytes (header text, payload size and FNV-1a-64), same file handling (source deleted, -verbose report), and a prolix
that returns the original pixels -- including the cases the reference's own prolix gets wrong (3-frame stacks crash it,
17..32-bit data is mangled: SURVEY App. C1, C6).  BASELINE configs[0] is the first case."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as G
import orc
import tiff_util
from trpx_b200 import build as B

CODE = {"u8": orc.U8, "u16": orc.U16, "u32": orc.U32, "i16": orc.I16}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def frames_of(c):
    return np.stack([orc.synth_frame(CODE[c["dtype"]], c["width"], c["height"], c["lam"], c["peaks"], c["seed"] + f, 20.0,
                                     c.get("amp_hi", 3000.0)).reshape(c["height"], c["width"]) for f in range(c["frames"])])


@pytest.fixture(scope="module")
def tools():
    B.build()
    outs = B.build_host()
    return {os.path.basename(o): o for o in outs}


def test_tiff_io_roundtrip_on_the_host(tools, tmp_path):
    for dt in (np.uint8, np.uint16, np.int16, np.uint32, np.int32, np.float32):
        st = (np.arange(3 * 7 * 5).reshape(3, 7, 5) * 37 % 251).astype(dt)
        a, b = tmp_path / "a.tif", tmp_path / "b.tif"
        tiff_util.write_tiff(a, st)
        r = subprocess.run([tools["terse_selftest"], "--tiff", str(a), str(b)], stdout=subprocess.PIPE, text=True, timeout=60)
        assert r.returncode == 0 and r.stdout.count("image 5x7") == 3, r.stdout
        back = tiff_util.read_tiff(b)
        assert len(back) == 3 and all(np.array_equal(back[f], st[f]) and back[f].dtype == st.dtype for f in range(3))
    be = tmp_path / "be.tif"                                  # big-endian input
    be.write_bytes(b"MM\x00\x2a\x00\x00\x00\x0c" + b"\x01\x02\x03\x04" +
                   b"\x00\x06" + b"".join(t.to_bytes(2, "big") + ty.to_bytes(2, "big") + (1).to_bytes(4, "big") +
                                          (v.to_bytes(4, "big") if ty == 4 else v.to_bytes(2, "big") + b"\0\0")
                                          for t, ty, v in [(256, 4, 2), (257, 4, 1), (258, 3, 16), (259, 3, 1), (273, 4, 8), (277, 3, 1)]) +
                   b"\x00\x00\x00\x00")
    r = subprocess.run([tools["terse_selftest"], "--tiff", str(be), str(tmp_path / "le.tif")], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0
    assert tiff_util.read_tiff(tmp_path / "le.tif")[0].tolist() == [[0x0102, 0x0304]]


def test_help_needs_no_gpu(tools):
    for t in ("terse", "prolix"):
        r = subprocess.run([tools[t], "-help"], stdout=subprocess.PIPE, text=True, timeout=30)
        assert r.returncode == 0 and "-verbose" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("c", G.load("kat_cli"), ids=lambda c: c["name"])
def test_cli_matches_the_reference_cli(tools, tmp_path, c):
    st = frames_of(c)
    tif, trpx = tmp_path / "x.tif", tmp_path / "x.trpx"
    tiff_util.write_tiff(tif, st)
    other = tmp_path / "notes.txt"
    other.write_text("not a tiff")
    r = subprocess.run([tools["terse"], "-verbose", str(tif), str(other)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert not tif.exists() and trpx.exists() and other.exists()             # src/terse.cpp:81-82
    assert "Deleting original TIFF file" in r.stdout and "Terse compressed: 1 files" in r.stdout and "Compression rate" in r.stdout
    img = trpx.read_bytes()
    h = img.index(b"/>") + 2
    assert img[:h].decode() == c["header"]
    payload = np.frombuffer(img[h:], np.uint8)
    assert payload.size == c["memory_size"] and hex(orc.fnv(payload)) == c["payload_fnv1a64"]
    r = subprocess.run([tools["prolix"], "-verbose", str(trpx)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert tif.exists() and not trpx.exists() and "Prolix expanded : 1 files" in r.stdout   # src/prolix.cpp:104-110
    back = tiff_util.read_tiff(tif)
    assert len(back) == c["frames"]
    want_dtype = {"u8": np.uint16, "u16": np.uint16, "i16": np.int16, "u32": np.uint32}[c["dtype"]]   # src/prolix.cpp:69-97
    for f in range(c["frames"]):
        assert back[f].dtype == want_dtype and np.array_equal(back[f].astype(np.int64), st[f].astype(np.int64))
