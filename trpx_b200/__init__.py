"""trpx_b200 -- B200-native TERSE/PROLIX codec.

This package is only a thin ctypes loader for the C-ABI library libtrpx_b200.so
(include/trpx_b200.h), used by the tests and bench.py; the product is the shared library plus the host
C++ class include/trpx/Terse.hpp.  There is no CPU implementation here: without the built CUDA
library, or without a GPU, every call raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build

U8, U16, U32, U64, I8, I16, I32, I64, F32, F64 = range(10)      # F32 / F64: decoder outputs only
_NP = {U8: np.uint8, U16: np.uint16, U32: np.uint32, U64: np.uint64,
       I8: np.int8, I16: np.int16, I32: np.int32, I64: np.int64, F32: np.float32, F64: np.float64}
_CODE = {np.dtype(v): k for k, v in _NP.items()}

OK, ERR_BAD_ARG, ERR_CAPACITY, ERR_CUDA, ERR_MALFORMED, ERR_NO_DEVICE, ERR_NOMEM = range(7)

EXPORTS = ["trpx_abi_version", "trpx_strerror", "trpx_dtype_size", "trpx_dtype_is_signed",
           "trpx_max_compressed_bytes", "trpx_ctx_create", "trpx_ctx_destroy", "trpx_ctx_device",
           "trpx_last_error", "trpx_ctx_lanes", "trpx_ctx_launch_count", "trpx_ctx_scratch_bytes",
           "trpx_encode_host", "trpx_encode_device", "trpx_decode_host", "trpx_decode_device",
           "trpx_ctx_set_profiling", "trpx_ctx_last_kernel_times", "trpx_ctx_encode_progress",
           "trpx_pool_create", "trpx_pool_destroy", "trpx_pool_size", "trpx_pool_device", "trpx_pool_last_error",
           "trpx_pool_encode_host", "trpx_pool_decode_host", "trpx_host_pin", "trpx_host_unpin", "trpx_host_alloc",
           "trpx_host_free"]


class TrpxError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("trpx_b200: %s (status %d)" % (msg, status))
        self.status = status


def dtype_code(dt):
    return _CODE[np.dtype(dt)]


def np_dtype(code):
    return np.dtype(_NP[code])


_lib = None


def lib(build_if_missing=True):
    """Load libtrpx_b200.so (building it in-tree with nvcc if absent).  Raises if that fails --
    there is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TRPX_LIB") or _build.OUT      # TRPX_LIB: an A/B variant built by tools/build_variant.py
    if not os.path.exists(path):
        if not build_if_missing:
            raise TrpxError(ERR_NO_DEVICE, "libtrpx_b200.so has not been built")
        _build.build()
    L = C.CDLL(path)
    vp, sz, u, i = C.c_void_p, C.c_size_t, C.c_uint, C.c_int
    L.trpx_abi_version.restype = i
    L.trpx_strerror.restype = C.c_char_p
    L.trpx_strerror.argtypes = [i]
    L.trpx_dtype_size.restype = sz
    L.trpx_dtype_size.argtypes = [i]
    L.trpx_dtype_is_signed.restype = i
    L.trpx_dtype_is_signed.argtypes = [i]
    L.trpx_max_compressed_bytes.restype = sz
    L.trpx_max_compressed_bytes.argtypes = [sz, i, u, sz]
    L.trpx_ctx_create.restype = i
    L.trpx_ctx_create.argtypes = [i, C.POINTER(vp)]
    L.trpx_ctx_destroy.restype = None
    L.trpx_ctx_destroy.argtypes = [vp]
    L.trpx_ctx_device.restype = i
    L.trpx_ctx_device.argtypes = [vp]
    L.trpx_last_error.restype = C.c_char_p
    L.trpx_last_error.argtypes = [vp]
    L.trpx_ctx_lanes.restype = i
    L.trpx_ctx_lanes.argtypes = [vp]
    L.trpx_ctx_launch_count.restype = C.c_uint64
    L.trpx_ctx_launch_count.argtypes = [vp]
    L.trpx_ctx_scratch_bytes.restype = sz
    L.trpx_ctx_scratch_bytes.argtypes = [vp]
    L.trpx_ctx_set_profiling.restype = i
    L.trpx_ctx_set_profiling.argtypes = [vp, i]
    L.trpx_ctx_last_kernel_times.restype = i
    L.trpx_ctx_last_kernel_times.argtypes = [vp, i, C.POINTER(C.c_char_p), C.POINTER(C.c_float), i]
    L.trpx_ctx_encode_progress.restype = i
    L.trpx_ctx_encode_progress.argtypes = [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
    L.trpx_encode_host.restype = i
    L.trpx_encode_host.argtypes = [vp, vp, i, sz, sz, u, vp, sz, vp, C.POINTER(sz), C.POINTER(u)]
    L.trpx_encode_device.restype = i
    L.trpx_encode_device.argtypes = [vp, i, vp, i, sz, sz, u, vp, sz, vp, vp, vp, vp]
    L.trpx_decode_host.restype = i
    L.trpx_decode_host.argtypes = [vp, vp, sz, i, u, sz, sz, sz, sz, vp, vp, vp, i]
    L.trpx_decode_device.restype = i
    L.trpx_decode_device.argtypes = [vp, i, vp, sz, i, u, sz, sz, vp, vp, vp, i, vp, vp]
    L.trpx_pool_create.restype = i
    L.trpx_pool_create.argtypes = [C.POINTER(i), i, C.POINTER(vp)]
    L.trpx_pool_destroy.restype = None
    L.trpx_pool_destroy.argtypes = [vp]
    L.trpx_pool_size.restype = i
    L.trpx_pool_size.argtypes = [vp]
    L.trpx_pool_device.restype = i
    L.trpx_pool_device.argtypes = [vp, i]
    L.trpx_pool_last_error.restype = C.c_char_p
    L.trpx_pool_last_error.argtypes = [vp]
    L.trpx_pool_encode_host.restype = i
    L.trpx_pool_encode_host.argtypes = [vp, vp, i, sz, sz, u, vp, sz, vp, C.POINTER(sz), C.POINTER(u)]
    L.trpx_pool_decode_host.restype = i
    L.trpx_pool_decode_host.argtypes = [vp, vp, sz, i, u, sz, sz, sz, sz, vp, vp, vp, i]
    L.trpx_host_pin.restype = i
    L.trpx_host_pin.argtypes = [vp, sz]
    L.trpx_host_unpin.restype = i
    L.trpx_host_unpin.argtypes = [vp]
    L.trpx_host_alloc.restype = vp
    L.trpx_host_alloc.argtypes = [sz]
    L.trpx_host_free.restype = None
    L.trpx_host_free.argtypes = [vp]
    _lib = L
    return L


def max_compressed_bytes(n_values, dtype, block=12, n_frames=1):
    return lib().trpx_max_compressed_bytes(n_values, dtype_code(dtype), block, n_frames)


class Codec:
    """One context (= one GPU).  Host-pointer calls take numpy arrays; device-pointer calls take raw
    addresses (e.g. torch.Tensor.data_ptr()) and a CUDA stream handle."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().trpx_ctx_create(device, C.byref(self._h))
        if rc != OK:
            raise TrpxError(rc, lib().trpx_strerror(rc).decode())

    def close(self):
        if self._h:
            lib().trpx_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            detail = lib().trpx_last_error(self._h).decode()
            raise TrpxError(rc, lib().trpx_strerror(rc).decode() + (": " + detail if detail else ""))

    @property
    def launches(self):
        return int(lib().trpx_ctx_launch_count(self._h))

    def encode_progress(self):
        """(calls started, frames done, payload bytes done) of the running / last host-pointer encode; any thread."""
        q, f, b = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        self._check(lib().trpx_ctx_encode_progress(self._h, C.byref(q), C.byref(f), C.byref(b)))
        return q.value, f.value, b.value

    def set_profiling(self, on=True):
        self._check(lib().trpx_ctx_set_profiling(self._h, int(on)))

    def last_kernel_times(self, lane=0):
        """[(kernel name, device ms), ...] of the last *_device call on `lane` (stream must have drained)."""
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = lib().trpx_ctx_last_kernel_times(self._h, lane, names, ms, 32)
        return [(names[k].decode(), float(ms[k])) for k in range(n)]

    # ---- host-pointer flavour
    def encode(self, stack, block=12, capacity=None):
        """stack: (F, N) integer array -> (payload uint8[], frame_bytes uint64[F], prolix_bits)."""
        stack = np.ascontiguousarray(stack)
        if stack.ndim == 1:
            stack = stack[None, :]
        F, N = stack.shape
        dt = dtype_code(stack.dtype)
        cap = capacity if capacity is not None else lib().trpx_max_compressed_bytes(N, dt, block, F)
        out = np.empty(cap, np.uint8)
        fb = np.zeros(F, np.uint64)
        total = C.c_size_t(0)
        pb = C.c_uint(0)
        self._check(lib().trpx_encode_host(self._h, stack.ctypes.data, dt, N, F, block, out.ctypes.data, cap,
                                           fb.ctypes.data, C.byref(total), C.byref(pb)))
        return out[:total.value].copy(), fb, pb.value

    def decode(self, payload, n_values, total_frames, is_signed, out_dtype, block=12, frame_bytes=None,
               first_frame=0, n_frames=None):
        """-> (values (n_frames, n_values) of out_dtype, frame_bytes uint64[total_frames])."""
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        n_frames = total_frames - first_frame if n_frames is None else n_frames
        out = np.empty((n_frames, n_values), np.dtype(out_dtype))
        fb_out = np.zeros(total_frames, np.uint64)
        fb = None if frame_bytes is None else np.ascontiguousarray(frame_bytes, dtype=np.uint64)
        self._check(lib().trpx_decode_host(self._h, payload.ctypes.data, payload.size, int(bool(is_signed)), block,
                                           n_values, total_frames, first_frame, n_frames,
                                           None if fb is None else fb.ctypes.data, fb_out.ctypes.data,
                                           out.ctypes.data, dtype_code(out.dtype)))
        return out, fb_out

    # ---- device-pointer flavour (asynchronous on `stream`)
    def encode_device(self, d_pixels, dtype, n_values, n_frames, d_out, out_capacity, d_frame_ends, d_prolix_bits,
                      d_status, stream=0, block=12, lane=0):
        self._check(lib().trpx_encode_device(self._h, lane, d_pixels, dtype_code(dtype), n_values, n_frames, block,
                                             d_out, out_capacity, d_frame_ends, d_prolix_bits, d_status, stream))

    def decode_device(self, d_payload, payload_bytes, is_signed, n_values, n_frames, d_frame_ends, d_out, out_dtype,
                      d_status, stream=0, block=12, lane=0, d_frame_ends_out=None):
        self._check(lib().trpx_decode_device(self._h, lane, d_payload, payload_bytes, int(bool(is_signed)), block,
                                             n_values, n_frames, d_frame_ends, d_frame_ends_out, d_out,
                                             dtype_code(out_dtype), d_status, stream))


class Pool:
    """Several GPUs of one box: one stack sharded by frame over them (trpx_pool_*), results concatenated on the host.
    devices=None takes every visible device."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        if devices is None:
            rc = lib().trpx_pool_create(None, 0, C.byref(self._h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = lib().trpx_pool_create(arr, len(devices), C.byref(self._h))
        if rc != OK:
            raise TrpxError(rc, lib().trpx_strerror(rc).decode())

    def close(self):
        if self._h:
            lib().trpx_pool_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self):
        return int(lib().trpx_pool_size(self._h))

    def _check(self, rc):
        if rc != OK:
            detail = lib().trpx_pool_last_error(self._h).decode()
            raise TrpxError(rc, lib().trpx_strerror(rc).decode() + (": " + detail if detail else ""))

    def encode(self, stack, block=12):
        stack = np.ascontiguousarray(stack)
        F, N = stack.shape
        dt = dtype_code(stack.dtype)
        cap = lib().trpx_max_compressed_bytes(N, dt, block, F)
        out = np.empty(cap, np.uint8)
        fb = np.zeros(F, np.uint64)
        total = C.c_size_t(0)
        pb = C.c_uint(0)
        self._check(lib().trpx_pool_encode_host(self._h, stack.ctypes.data, dt, N, F, block, out.ctypes.data, cap,
                                                fb.ctypes.data, C.byref(total), C.byref(pb)))
        return out[:total.value].copy(), fb, pb.value

    def decode(self, payload, n_values, total_frames, is_signed, out_dtype, block=12, frame_bytes=None, first_frame=0,
               n_frames=None):
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        n_frames = total_frames - first_frame if n_frames is None else n_frames
        out = np.empty((n_frames, n_values), np.dtype(out_dtype))
        fb_out = np.zeros(total_frames, np.uint64)
        fb = None if frame_bytes is None else np.ascontiguousarray(frame_bytes, dtype=np.uint64)
        self._check(lib().trpx_pool_decode_host(self._h, payload.ctypes.data, payload.size, int(bool(is_signed)), block,
                                                n_values, total_frames, first_frame, n_frames,
                                                None if fb is None else fb.ctypes.data, fb_out.ctypes.data,
                                                out.ctypes.data, dtype_code(out.dtype)))
        return out, fb_out
