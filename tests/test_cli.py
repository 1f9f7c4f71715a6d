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
import imagej_reader
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


def test_imagej_reader_restatement_reads_what_the_reference_writes():
    """The restated plugin reader (tests/imagej_reader.py, TRPX_Reader.java:41-150) against the reference's own bytes: the
    oracle's payload behind the oracle's header, 1 and 3 frames, a ragged last block, text in front of the element."""
    for frames, w, h in ((1, 24, 19), (3, 16, 13)):
        st = np.stack([orc.synth_frame(orc.U16, w, h, 2.0, 2, 40 + f, 20.0, 900.0) for f in range(frames)])
        payload, per, pb = orc.encode_stack(st)
        head = orc.header(pb, False, 12, payload.size, w * h, [w, h], frames)
        for prefix in (b"", b"some text\nmore text\n"):
            got = imagej_reader.read(prefix + head + payload.tobytes())
            assert got.shape == (frames, h, w) and np.array_equal(got.reshape(frames, -1), st)
    signed = orc.header(12, True, 12, 10, 12, [4, 3], 1) + bytes(10)
    with pytest.raises(imagej_reader.NotAdmissible):
        imagej_reader.read(signed)


@pytest.mark.gpu
def test_files_written_by_the_gpu_cli_open_in_the_imagej_reader(tools, tmp_path):
    """the ImageJ plugin (TRPX_Reader.java) is the reference's second, independent reader; a .trpx
    written by cxx/terse on the GPU codec must decode in it to the TIFF's pixels -- header attributes, data start, frame
    alignment and all.  16-bit unsigned only, as the plugin demands."""
    for frames, w, h, peaks in ((1, 64, 48, 3), (5, 40, 37, 2)):
        st = np.stack([orc.synth_frame(orc.U16, w, h, 2.0, peaks, 700 + f, 20.0, 3000.0).reshape(h, w) for f in range(frames)])
        tif, trpx = tmp_path / ("ij%d.tif" % frames), tmp_path / ("ij%d.trpx" % frames)
        tiff_util.write_tiff(tif, st)
        r = subprocess.run([tools["terse"], str(tif)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode == 0 and trpx.exists(), r.stdout
        got = imagej_reader.read(trpx.read_bytes())
        assert got.shape == (frames, h, w) and np.array_equal(got, st)


@pytest.mark.gpu
def test_several_files_per_command_are_prefetched_in_order(tools, tmp_path):
    """The next file is read by a helper thread while the current one is on the GPU (cli_common.hpp); results, messages and
    what gets deleted must be exactly what the one-by-one loop of src/terse.cpp:44 / src/prolix.cpp:42 gives: a corrupt file
    in the middle is reported and kept, its neighbours are converted."""
    stacks = {}
    names = []
    for k, (frames, w, h) in enumerate(((2, 48, 40), (1, 31, 29), (3, 64, 20))):
        st = np.stack([orc.synth_frame(orc.U16, w, h, 2.0, 2, 900 + 10 * k + f, 20.0, 2000.0).reshape(h, w) for f in range(frames)])
        p = tmp_path / ("s%d.tif" % k)
        tiff_util.write_tiff(p, st)
        stacks[p.stem] = st
        names.append(p)
    bad = tmp_path / "s1b.tif"
    bad.write_bytes(b"II*\x00\x08\x00\x00\x00" + b"\xff" * 40)              # a TIFF header in front of garbage
    order = [names[0], names[1], bad, names[2]]
    r = subprocess.run([tools["terse"], "-verbose"] + [str(p) for p in order], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "Terse compressed: 3 files" in r.stdout and "Error processing" in r.stdout and bad.exists()
    deleted = [l for l in r.stdout.splitlines() if l.startswith("Deleting original TIFF file")]
    assert [os.path.basename(l.split('"')[1]) for l in deleted] == ["s0.tif", "s1.tif", "s2.tif"]
    trpx = [p.with_suffix(".trpx") for p in names]
    assert all(t.exists() for t in trpx) and not any(p.exists() for p in names)
    r = subprocess.run([tools["prolix"], "-verbose"] + [str(t) for t in trpx], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "Prolix expanded : 3 files" in r.stdout, r.stdout
    for p in names:
        back = tiff_util.read_tiff(p)
        assert len(back) == len(stacks[p.stem]) and all(np.array_equal(back[f], stacks[p.stem][f]) for f in range(len(back)))
