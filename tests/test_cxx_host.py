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
