"""Builds tests/emu/libtrpx_emu.so: the kernel sources compiled for the HOST with the test-only SIMT
emulator (tests/emu/emu.hpp).  Test infrastructure only -- see trpx_b200/csrc/simt.cuh."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "emu_lib.cpp")
OUT = os.path.join(ROOT, "tests", "emu", "libtrpx_emu.so")
CSRC = os.path.join(ROOT, "trpx_b200", "csrc")
DEPS = [SRC, os.path.join(ROOT, "tests", "emu", "emu.hpp")] + sorted(
    os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))


def build(force=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DTRPX_EMU", "-x", "c++",
           "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-unused-variable",
           "-I", os.path.join(ROOT, "tests", "emu"), "-I", os.path.join(ROOT, "trpx_b200", "csrc"),
           SRC, "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
