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


def _set_grid(ctas, sms):
    """ctas: CTAs of a grid that are resident (interleaved) at once; sms: the emulated device's SM count (grid size)."""
    import os
    os.environ["EMU_CTAS"] = str(ctas)
    os.environ["EMU_SMS"] = str(sms)


def encode(stack, block=12, incl_stride=0, misalign=0, cap=None, ctas=1, sms=1):
    """stack: (F, N) array.  Returns (payload, frame_ends, prolix_bits, status, used_fast)."""
    _set_grid(ctas, sms)
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


def decode(payload, n, frames, is_signed, out_dtype, block=12, frame_ends=None, seg_bytes=64,
           warm_bytes=64, misalign_out=0, sub_shift=0):
    """Returns (values (frames, n), status, staged, recovered frame ends)."""
    L = lib()
    if not hasattr(L, "_dec_ready"):
        L.emu_decode.restype = C.c_int
        L.emu_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_size_t, C.c_size_t,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint,
                                 C.c_uint, C.POINTER(C.c_int), C.c_uint]
        L._dec_ready = True
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    buf = _aligned(payload.size + 16)
    buf[:payload.size] = payload
    buf[payload.size:] = 0xEE                            # bytes past the payload must not matter
    npdt = np.dtype(orc.NP_OF[out_dtype] if isinstance(out_dtype, int) else out_dtype)
    outb = _aligned(frames * n * npdt.itemsize, 16, misalign_out)
    outb[:] = 0xCD
    st = np.zeros(1, np.uint32)
    staged = C.c_int(0)
    fe_out = np.zeros(frames, np.uint64)
    fe = None if frame_ends is None else np.ascontiguousarray(frame_ends, dtype=np.uint64)
    rc = L.emu_decode(buf.ctypes.data, payload.size, int(is_signed), block, n, frames,
                      None if fe is None else fe.ctypes.data, fe_out.ctypes.data, outb.ctypes.data,
                      orc.FLOAT_CODE[npdt] if npdt in orc.FLOAT_CODE else orc.code_of(npdt), st.ctypes.data, seg_bytes, warm_bytes, C.byref(staged), sub_shift)
    assert rc == 0
    return outb.view(npdt).reshape(frames, n).copy(), int(st[0]), bool(staged.value), fe_out


def set_spec(b=64, r0=32, rs=80, max_steps=256):
    """Batch geometry of the speculative frame chain (frames per batch, window radius r0 + rs * sqrt(k), T steps per candidate)."""
    lib().emu_set_spec(b, r0, rs, max_steps)


def spec_followed(reset=True):
    """Frames whose end the speculative pass produced since the last reset."""
    L = lib()
    L.emu_spec_followed.restype = C.c_ulonglong
    return int(L.emu_spec_followed(1 if reset else 0))
