// emu_lib.cpp -- TEST INFRASTRUCTURE ONLY: the kernel sources compiled for the host (-DTRPX_EMU) and
// driven through the same launch sequences (codec_launch.cuh) the product uses, with host memory
// standing in for device memory.  Built by tests/emu_build.py into tests/emu/libtrpx_emu.so.
#include <stdlib.h>
#include <string.h>

#include "codec_launch.cuh"

using namespace trpx;

extern "C" {

int emu_encode(const void* px, int dtype, size_t n, size_t frames, unsigned block, uint8_t* out,
               size_t cap, uint64_t* frame_ends, uint32_t* prolix_bits, uint32_t* status,
               unsigned dbg_incl_stride, int* used_fast)
{
    EncPlan pl = enc_plan(dtype, px, n, frames, block);
    if (!pl.ok) return 1;
    if (used_fast) *used_fast = pl.fast ? 1 : 0;
    const size_t guard = 4096;                               // nothing may be written past the plan's scratch size
    unsigned char* scratch = (unsigned char*)malloc(pl.scratch_bytes + guard);
    memset(scratch, 0x5A, pl.scratch_bytes + guard);
    Launcher L{nullptr, getenv("EMU_SMS") ? (u32)atoi(getenv("EMU_SMS")) : 1u, nullptr, cudaSuccess};
    encode_async(L, dtype, px, n, frames, block, out, cap, (u64*)frame_ends, prolix_bits, status, scratch, pl,
                 3, dbg_incl_stride);
    int rc = 0;
    for (size_t i = 0; i < guard; ++i)
        if (scratch[pl.scratch_bytes + i] != 0x5A) rc = 3;
    free(scratch);
    return rc;
}

}

extern "C" int emu_decode(const uint8_t* payload, size_t payload_bytes, int is_signed, unsigned block,
                          size_t n, size_t frames, const uint64_t* frame_ends, uint64_t* frame_ends_out,
                          void* out, int out_dtype, uint32_t* status, unsigned seg_bytes,
                          unsigned warm_bytes, int* used_staged, unsigned sub_shift)
{
    DecPlan pl = dec_plan(out_dtype, payload_bytes, n, frames, block, out, seg_bytes, warm_bytes, sub_shift);
    if (!pl.ok) return 1;
    if (used_staged) *used_staged = pl.staged ? 1 : 0;
    const size_t need = (dec_scratch_need(pl, out_dtype, payload_bytes, block, frames, frame_ends == nullptr) + 255) / 256 * 256;
    const size_t guard = 4096;                               // nothing may be written past what dec_scratch_need() asked for
    unsigned char* scratch = (unsigned char*)aligned_alloc(256, need + guard);
    memset(scratch, 0x5A, need + guard);
    Launcher L{nullptr, 1, nullptr, cudaSuccess};
    decode_async(L, payload, payload_bytes, is_signed != 0, block, n, frames, (const u64*)frame_ends,
                 (u64*)frame_ends_out, out, out_dtype, status, scratch, pl, 1);
    int rc = 0;
    for (size_t i = 0; i < guard; ++i)
        if (scratch[need + i] != 0x5A) rc = 3;
    free(scratch);
    return rc;
}

// batch geometry of the speculative frame chain (tests shrink it to exercise window misses)
extern "C" void emu_set_spec(unsigned b, unsigned r0, unsigned rs, unsigned max_steps)
{
    trpx::g_spec_params[0] = b; trpx::g_spec_params[1] = r0; trpx::g_spec_params[2] = rs; trpx::g_spec_params[3] = max_steps;
}
extern "C" unsigned long long emu_spec_followed(int reset)
{
    const unsigned long long n = trpx::emu_spec_followed();
    if (reset) trpx::emu_spec_followed() = 0;
    return n;
}
