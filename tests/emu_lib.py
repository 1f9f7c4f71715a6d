"""ctypes bindings for tests/emu/libtrpx_emu.so (kernel sources compiled for the host with the
test-only SIMT emulator).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

import emu_build
import orc

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(emu_build.build())
        L.emu_encode.restype = C.c_int
        L.emu_encode.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_uint, C.c_void_p,
                                 C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint,
                                 C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _aligned(nbytes, align=16, offset=0):
    raw = np.zeros(nbytes + align + 64, np.uint8)
    start = (-raw.ctypes.data) % align + offset
    return raw[start:start + nbytes]


def encode(stack, block=12, incl_stride=0, misalign=0, cap=None):
    """stack: (F, N) array.  Returns (payload, frame_ends, prolix_bits, status, used_fast)."""
    stack = np.ascontiguousarray(stack)
    F, N = stack.shape
    dt = orc.code_of(stack.dtype)
    buf = _aligned(stack.nbytes, 16, misalign)
    buf[:] = stack.view(np.uint8).ravel()
    if cap is None:
        cap = (orc.orc().orc_max_frame_bytes(N, dt, block) * F + 15) // 16 * 16
    out = _aligned(cap + 16)
    out[:] = 0xEE                                        # the encoder must not rely on a zeroed output
    ends = np.zeros(F, np.uint64)
    pb = np.zeros(1, np.uint32)
    st = np.zeros(1, np.uint32)
    fast = C.c_int(0)
    rc = lib().emu_encode(buf.ctypes.data, dt, N, F, block, out.ctypes.data, cap, ends.ctypes.data,
                          pb.ctypes.data, st.ctypes.data, incl_stride, C.byref(fast))
    assert rc == 0
    total = int(ends[-1]) if st[0] == 0 else 0
    return out[:total].copy(), ends, int(pb[0]), int(st[0]), bool(fast.value)
