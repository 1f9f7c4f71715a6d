"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/trpx_b200.h declares; size helpers agree with the oracle; and without a GPU the compute entry
points fail loudly (there is no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orc
import trpx_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "trpx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trpx_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    L = trpx_b200.lib()
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), n
    assert sorted(names) == sorted(trpx_b200.EXPORTS)
    assert L.trpx_abi_version() == 2


def test_size_helpers_match_the_oracle():
    L = trpx_b200.lib()
    for code in range(8):
        assert L.trpx_dtype_size(code) == orc.orc().orc_dtype_size(code)
        assert L.trpx_dtype_is_signed(code) == int(code >= 4)
        for n, block, frames in [(1, 12, 1), (262144, 12, 3), (1000, 7, 2), (18093576, 12, 1)]:
            cap = L.trpx_max_compressed_bytes(n, code, block, frames)
            assert cap % 16 == 0 and cap >= frames * orc.orc().orc_max_frame_bytes(n, code, block)
    # TRPX_F32 / TRPX_F64 are output types of the decoder only: they have a size, but nothing can be encoded from them
    assert L.trpx_dtype_size(8) == 4 and L.trpx_dtype_size(9) == 8 and L.trpx_dtype_is_signed(8) == 1
    assert L.trpx_max_compressed_bytes(10, 8, 12, 1) == 0 and L.trpx_max_compressed_bytes(10, 9, 12, 1) == 0
    assert L.trpx_dtype_size(10) == 0 and L.trpx_dtype_size(-1) == 0 and L.trpx_max_compressed_bytes(10, 10, 12, 1) == 0


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_device_means_no_codec():
    L = trpx_b200.lib()
    h = C.c_void_p()
    assert L.trpx_ctx_create(0, C.byref(h)) == trpx_b200.ERR_NO_DEVICE
    with pytest.raises(trpx_b200.TrpxError):
        trpx_b200.Codec(0)
    pool = C.c_void_p()
    assert L.trpx_pool_create(None, 0, C.byref(pool)) == trpx_b200.ERR_NO_DEVICE     # the multi-GPU entry points too
    with pytest.raises(trpx_b200.TrpxError):
        trpx_b200.Pool()
    # null context: every compute entry point refuses
    a = np.zeros(24, np.uint16)
    out = np.zeros(256, np.uint8)
    assert L.trpx_encode_host(None, a.ctypes.data, 1, 24, 1, 12, out.ctypes.data, 256, None, None, None) == trpx_b200.ERR_BAD_ARG
