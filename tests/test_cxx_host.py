"""The host-side drop-in class include/trpx/Terse.hpp (same surface as the reference's jpa::Terse,
Terse.hpp:228-474), compiled with g++ -std=c++20 against the C-ABI library.

CPU: the .trpx container code is host-only -- files the REFERENCE wrote (tests/golden/kat_files.json) are read
and written back byte for byte.  GPU: cxx/terse_selftest --gpu repeats the reference's own test scenario
(test/terse_tests.cpp:15-33) and more through the kernels."""
import os
import subprocess

import pytest

import golden_util as G
from trpx_b200 import build as B


@pytest.fixture(scope="module")
def selftest():
    B.build()
    outs = B.build_host()
    exe = [o for o in outs if o.endswith("terse_selftest")][0]
    return exe


@pytest.mark.parametrize("c", G.load("kat_files")[:3], ids=lambda c: c["name"])
def test_reference_files_roundtrip_through_the_container_code(selftest, tmp_path, c):
    img = bytes.fromhex(c["file_hex"])
    src, dst = tmp_path / "in.trpx", tmp_path / "out.trpx"
    src.write_bytes(b"leading junk that the scanner must skip " + img)
    r = subprocess.run([selftest, "--container", str(src), str(dst)], stdout=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0, r.stdout
    assert dst.read_bytes() == img
    assert "frames=%d" % c["frames"] in r.stdout and "number_of_values=%d" % c["n"] in r.stdout


def test_several_objects_in_one_stream(selftest, tmp_path):
    """Files may hold several <Terse/> objects back to back with junk in between; the reader leaves the stream right
    after each payload (reference Terse.hpp:275-279, :485-498; XML_element.hpp:216-224, :428-452)."""
    cases = G.load("kat_files")
    a, b = bytes.fromhex(cases[0]["file_hex"]), bytes.fromhex(cases[2]["file_hex"])          # different dtypes / frame counts
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(b"junk <Ter before\n" + a + b"\n\n<!-- between -->" + b + b"trailing bytes")
    r = subprocess.run([selftest, "--objects", str(src), str(dst)], stdout=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0, r.stdout
    assert dst.read_bytes() == a + b                                                     # both re-written byte for byte
    assert "object 0:" in r.stdout and "object 1:" in r.stdout and "stopped after 2 object(s)" in r.stdout
    assert "frames=%d" % cases[0]["frames"] in r.stdout.splitlines()[0]
    assert "frames=%d" % cases[2]["frames"] in r.stdout.splitlines()[1]


def test_header_without_number_of_frames_and_malformed_headers(selftest, tmp_path):
    """`number_of_frames` is documented as optional by the reference (Terse.hpp:46) but its reader requires it (:497, App. C8):
    here an absent attribute means one frame.  Malformed attributes make the constructor throw; nothing crashes."""
    img = bytes.fromhex(G.load("kat_files")[0]["file_hex"])
    head, payload = img[:img.index(b"/>") + 2], img[img.index(b"/>") + 2:]
    import re
    no_nf = re.sub(rb' number_of_frames="\d+"', b"", head)
    assert no_nf != head
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(no_nf + payload)
    r = subprocess.run([selftest, "--objects", str(src), str(dst)], stdout=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0 and "frames=1" in r.stdout, r.stdout
    assert dst.read_bytes() == img                     # the writer always states the frame count
    for bad in (head.replace(b'memory_size="', b'memory_size="x'), head.replace(b'block="12"', b'block="0"'),
                head.replace(b'prolix_bits="', b'prolix_bits="-'), re.sub(rb' number_of_values="\d+"', b"", head),
                head[:-2],                               # element never closed
                head + payload[:-1]):                    # truncated payload
        src.write_bytes(bad if bad.endswith(payload[:-1]) else bad + payload)
        r = subprocess.run([selftest, "--objects", str(src), str(dst)], stdout=subprocess.PIPE, text=True, timeout=60)
        assert r.returncode == 3 and "stopped after 0 object(s)" in r.stdout, (bad[:80], r.returncode, r.stdout)


def test_tiff_reader_rejects_malformed_files(selftest, tmp_path):
    """The TIFF reader sits in front of a CLI that deletes its inputs: a tag without a value, a cyclic directory chain or
    an image larger than the file must raise, not crash or loop."""
    import struct
    def tiff(entries, next_ifd=0, data=b"\x01\x02\x03\x04"):
        body = b"II*\x00" + struct.pack("<I", 8 + len(data)) + data
        ifd = struct.pack("<H", len(entries)) + b"".join(struct.pack("<HHII", *e) for e in entries) + struct.pack("<I", next_ifd)
        return body + ifd
    ok = [(256, 4, 1, 2), (257, 4, 1, 2), (258, 3, 1, 8), (259, 3, 1, 1), (273, 4, 1, 8), (277, 3, 1, 1), (279, 4, 1, 4)]
    src, dst = tmp_path / "in.tif", tmp_path / "out.tif"
    src.write_bytes(tiff(ok))
    r = subprocess.run([selftest, "--tiff", str(src), str(dst)], stdout=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0 and "image 2x2 bits=8" in r.stdout, r.stdout
    bad_files = [tiff([(256, 4, 0, 2)] + ok[1:]),                                   # a tag with count 0
                 tiff(ok, next_ifd=12),                                              # the next directory is this one: a cycle
                 tiff([(256, 4, 1, 0x7fffffff), (257, 4, 1, 0x7fffffff)] + ok[2:]),  # width * height overflows / exceeds the file
                 tiff(ok[:4] + [(273, 4, 1, 0xfffffff0)] + ok[5:])]                  # strip offset near 2^32
    for f in bad_files:
        src.write_bytes(f)
        r = subprocess.run([selftest, "--tiff", str(src), str(dst)], stdout=subprocess.PIPE, text=True, timeout=60)
        assert r.returncode == 2 and "exception: TIFF" in r.stdout, r.stdout


def test_no_gpu_means_exception_not_fallback(selftest):
    import torch
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU behaviour")
    r = subprocess.run([selftest, "--gpu"], stdout=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 2 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_selftest_on_gpu(selftest):
    r = subprocess.run([selftest, "--gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "selftest ok" in r.stdout, r.stdout
