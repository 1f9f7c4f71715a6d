"""Builds trpx_b200/libtrpx_b200.so (the C-ABI library of include/trpx_b200.h) IN-TREE with nvcc for
sm_100a.  nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtrpx_b200.so")
SOURCES = [os.path.join(CSRC, "trpx_api.cu")]
DEPS = SOURCES + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + [
    os.path.join(os.path.dirname(HERE), "include", "trpx_b200.h")]
STAMP = OUT + ".srchash"        # content hash of DEPS + flags the .so was built from (git-ignored, travels with the .so)
INFO = os.path.join(HERE, "build_info.json")   # what the last build() call did: "compiled" or "reused" (and why)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared", "-cudart", "static"]


def nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def src_hash():
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in DEPS:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    return h.hexdigest()


def up_to_date():
    """The .so is reused only when it was built from exactly these sources and flags (content hash, not mtime)."""
    try:
        return os.path.exists(OUT) and open(STAMP).read().strip() == src_hash()
    except OSError:
        return False


def _record(mode, why):
    import json
    import time
    try:
        json.dump({"build_mode": mode, "why": why, "src_hash": src_hash(), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                   "nvcc_flags": NVCC_FLAGS}, open(INFO, "w"), indent=1)
    except OSError:
        pass


def build(force=False, verbose=False):
    force = force or os.environ.get("TRPX_FORCE_BUILD") == "1"
    if not force and up_to_date():
        _record("reused", "libtrpx_b200.so matches the content hash of its sources and flags")
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + SOURCES + ["-o", OUT]
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libtrpx_b200.so")
    open(STAMP, "w").write(src_hash() + "\n")
    _record("compiled", "forced" if force else "sources or flags changed (or no previous build)")
    return OUT


ROOT = os.path.dirname(HERE)
HOST_TARGETS = {"terse_selftest": [os.path.join(ROOT, "cxx", "terse_selftest.cpp")],
                "terse": [os.path.join(ROOT, "cxx", "terse.cpp")],
                "prolix": [os.path.join(ROOT, "cxx", "prolix.cpp")],
                "terse_bench": [os.path.join(ROOT, "cxx", "terse_bench.cpp")]}


def build_host(force=False):
    """Host-side C++ on top of the C ABI (include/trpx/Terse.hpp): g++ -std=c++20, linked against the in-tree .so."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    outs = []
    for name, srcs in HOST_TARGETS.items():
        srcs = [s for s in srcs if os.path.exists(s)]
        if not srcs:
            continue
        out = os.path.join(ROOT, "cxx", name)
        deps = srcs + [os.path.join(ROOT, "include", "trpx", "Terse.hpp"), os.path.join(ROOT, "include", "trpx", "Grey_tiff_io.hpp"), os.path.join(ROOT, "cxx", "cli_common.hpp"),
                       os.path.join(ROOT, "include", "trpx_b200.h"), OUT]
        if force or not os.path.exists(out) or any(os.path.getmtime(out) < os.path.getmtime(d) for d in deps if os.path.exists(d)):
            cmd = [cxx, "-std=c++20", "-O2", "-DNDEBUG", "-Wall", "-I", os.path.join(ROOT, "include")] + srcs + [
                "-L", HERE, "-ltrpx_b200", "-Wl,-rpath," + HERE, "-o", out]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout)
                raise RuntimeError("g++ failed building " + name)
        outs.append(out)
    return outs


def build_variant(name, defines, verbose=False):
    """A/B variant of the library with extra -D flags -> trpx_b200/_variants/<name>.so (git-ignored; select it with
    TRPX_LIB=...).  Tuning only: the product is always the plain build."""
    vdir = os.path.join(HERE, "_variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, name + ".so")
    cmd = [nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + SOURCES + ["-o", out]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building variant " + name)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":      # python -m trpx_b200.build --variant NAME DEF1 DEF2=3 ...
        print(build_variant(sys.argv[2], sys.argv[3:], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
