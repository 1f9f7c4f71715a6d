"""ctypes bindings for the TEST-ONLY checkers under oracle/ (plain-C oracle and, when built,
the reference shim oracle/_ref/libtrpx_ref.so).  Only tests/, bench.py's cpu_baseline leg and
__graft_entry__.smoke() import this module; the product never does."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

U8, U16, U32, U64, I8, I16, I32, I64 = range(8)
NP_OF = {U8: np.uint8, U16: np.uint16, U32: np.uint32, U64: np.uint64,
         I8: np.int8, I16: np.int16, I32: np.int32, I64: np.int64}
CODE_OF = {np.dtype(v): k for k, v in NP_OF.items()}
F32, F64 = 8, 9                                                    # decoder OUTPUT types of the CUDA library only
FLOAT_CODE = {np.dtype(np.float32): F32, np.dtype(np.float64): F64}


def code_of(dt):
    return CODE_OF[np.dtype(dt)]


def build():
    """Compile liboracle.so (always) and _ref (only if /root/reference is mounted)."""
    subprocess.run(["make", "-s", "-C", ODIR], check=True, stdout=subprocess.DEVNULL)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


_orc = None
_ref = None


def orc():
    global _orc
    if _orc is None:
        p = os.path.join(ODIR, "liboracle.so")
        if not os.path.exists(p):
            build()
        L = C.CDLL(p)
        L.orc_max_frame_bytes.restype = C.c_size_t
        L.orc_max_frame_bytes.argtypes = [C.c_size_t, C.c_int, C.c_uint]
        L.orc_encode_frame.restype = C.c_size_t
        L.orc_encode_frame.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_uint, C.c_void_p,
                                       C.POINTER(C.c_uint)]
        L.orc_encode_stack.restype = C.c_size_t
        L.orc_encode_stack.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_uint,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_uint)]
        L.orc_decode_frame.restype = C.c_size_t
        L.orc_decode_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_size_t,
                                       C.c_void_p, C.c_int]
        L.orc_frame_widths.restype = C.c_size_t
        L.orc_frame_widths.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_size_t, C.c_void_p]
        L.orc_fnv64_frames.restype = None
        L.orc_fnv64_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_header.restype = C.c_size_t
        L.orc_header.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.c_int, C.c_uint, C.c_size_t,
                                 C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_kat_fill.restype = None
        L.orc_kat_fill.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_uint64]
        L.orc_synth_frame.restype = None
        L.orc_synth_frame.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_double,
                                      C.c_uint, C.c_double, C.c_double, C.c_uint64]
        _orc = L
    return _orc


def ref():
    """The reference shim, or None if it has not been built (no /root/reference at build time)."""
    global _ref
    if _ref is None:
        L = _load(os.path.join(ODIR, "_ref", "libtrpx_ref.so"))
        if L is None:
            return None
        L.ref_encode_frame.restype = C.c_size_t
        L.ref_encode_frame.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_uint, C.c_void_p,
                                       C.c_size_t, C.POINTER(C.c_uint)]
        L.ref_write_file_image.restype = C.c_size_t
        L.ref_write_file_image.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_uint, C.c_void_p,
                                           C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_write_stack_image.restype = C.c_size_t
        L.ref_write_stack_image.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t,
                                            C.c_void_p, C.c_size_t]
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_size_t]
        L.ref_prolix.restype = C.c_int
        L.ref_prolix.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_close.restype = None
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_frame_digests.restype = C.c_int
        L.ref_frame_digests.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_uint, C.c_void_p, C.c_void_p]
        L.ref_bench_encode.restype = C.c_double
        L.ref_bench_encode.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_uint,
                                       C.POINTER(C.c_size_t)]
        L.ref_bench_decode.restype = C.c_double
        L.ref_bench_decode.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_uint,
                                       C.c_void_p]
        _ref = L
    return _ref


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def encode_frame(a, block=12):
    """Oracle encode of one frame -> (payload bytes as np.uint8, prolix_bits)."""
    a = np.ascontiguousarray(a).ravel()
    dt = code_of(a.dtype)
    cap = orc().orc_max_frame_bytes(a.size, dt, block)
    out = np.zeros(cap, np.uint8)
    pb = C.c_uint(0)
    n = orc().orc_encode_frame(_ptr(a), dt, a.size, block, _ptr(out), C.byref(pb))
    return out[:n].copy(), pb.value


def encode_stack(a, block=12):
    """a: (F, N) array -> (payload, per_frame_bytes, prolix_bits)."""
    a = np.ascontiguousarray(a)
    F, N = a.shape
    dt = code_of(a.dtype)
    cap = orc().orc_max_frame_bytes(N, dt, block) * F
    out = np.zeros(cap, np.uint8)
    per = np.zeros(F, np.uint64)
    pb = C.c_uint(0)
    n = orc().orc_encode_stack(_ptr(a), dt, N, F, block, _ptr(out), _ptr(per), C.byref(pb))
    return out[:n].copy(), per, pb.value


def decode_frame(payload, n, is_signed, out_dtype, block=12):
    """Oracle decode of one frame -> (values, bytes consumed).  Floating-point outputs follow Terse.hpp:379-383:
    every value through a 64-bit integer (signed or unsigned as the stream is) and a double."""
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    if not isinstance(out_dtype, int) and np.dtype(out_dtype) in FLOAT_CODE:
        v, used = decode_frame(payload, n, is_signed, I64 if is_signed else U64, block)
        return v.astype(np.float64).astype(out_dtype), used
    out = np.zeros(n, NP_OF[out_dtype] if isinstance(out_dtype, int) else out_dtype)
    used = orc().orc_decode_frame(_ptr(payload), payload.size, int(is_signed), block, n, _ptr(out),
                                  code_of(out.dtype))
    return out, used


def frame_widths(payload, n, block=12):
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    w = np.zeros((n + block - 1) // block, np.uint8)
    used = orc().orc_frame_widths(_ptr(payload), payload.size, block, n, _ptr(w))
    return w, used


def header(prolix_bits, is_signed, block, memory_size, n, dims, frames):
    buf = C.create_string_buffer(512)
    d = np.asarray(dims if dims else [], dtype=np.uint64)
    k = orc().orc_header(buf, 512, prolix_bits, int(is_signed), block, memory_size, n,
                         _ptr(d) if d.size else None, d.size, frames)
    return buf.raw[:k]


def fnv(payload):
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    return orc().orc_fnv1a64(_ptr(payload), payload.size)


def kat_fill(dtype, n, seed):
    a = np.zeros(n, NP_OF[dtype])
    orc().orc_kat_fill(_ptr(a), dtype, n, seed)
    return a


def synth_frame(dtype, width, height, lam, n_peaks, seed, amp_lo=20.0, amp_hi=3000.0):
    a = np.zeros(width * height, NP_OF[dtype])
    orc().orc_synth_frame(_ptr(a), dtype, width, height, lam, n_peaks, amp_lo, amp_hi, seed)
    return a


def ref_encode_frame(a, block=12):
    a = np.ascontiguousarray(a).ravel()
    dt = code_of(a.dtype)
    cap = orc().orc_max_frame_bytes(a.size, dt, block) + 64
    out = np.zeros(cap, np.uint8)
    pb = C.c_uint(0)
    n = ref().ref_encode_frame(_ptr(a), dt, a.size, block, _ptr(out), cap, C.byref(pb))
    return out[:n].copy(), pb.value


def ref_decode_frame(payload, n, is_signed, prolix_bits, out_dtype, block=12):
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    h = ref().ref_open(_ptr(payload), payload.size, int(is_signed), block, prolix_bits, n)
    assert h
    out = np.zeros(n, NP_OF[out_dtype] if isinstance(out_dtype, int) else out_dtype)
    ref().ref_prolix(h, _ptr(out), FLOAT_CODE[out.dtype] if out.dtype in FLOAT_CODE else code_of(out.dtype))
    ref().ref_close(h)
    return out
