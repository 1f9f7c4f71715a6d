// codec_launch.cuh -- kernel launch sequences for TERSE (encode) and PROLIX (decode).
//
// Everything here is asynchronous on one stream and touches only device memory, so the same code
// runs under the test-only SIMT emulator (-DTRPX_EMU, host pointers stand in for device pointers).
// Contexts, streams, pinned staging and the C ABI live in trpx_api.cu.
#pragma once

#include "simt.cuh"
#include "terse_encode.cuh"
#include "prolix_decode.cuh"

namespace trpx {

#ifdef TRPX_EMU
using ::emu::cudaMemsetAsync;
#endif

enum { DT_U8 = 0, DT_U16, DT_U32, DT_U64, DT_I8, DT_I16, DT_I32, DT_I64 };

inline size_t dtype_size(int dt) { return dt < 0 || dt > 7 ? 0 : (size_t)1 << (dt & 3); }
inline bool dtype_signed(int dt) { return dt >= DT_I8; }

constexpr int ENC_NT = 256;     // threads per encoder CTA (8 warps, 48 bytes of pixels per thread)
constexpr int GEN_NT = 256;     // threads per generic-encoder CTA (one block per thread)

struct Launcher {
    cudaStream_t stream;
    u32 sm_count;               // SMs of the device (148 on B200)
    u64* launches;              // optional counter of kernels launched
    cudaError_t err;
    void count() { if (launches) ++*launches; }
};

// ---------------------------------------------------------------------------------- encode
struct EncPlan {
    bool fast;                  // block == 12 and 16-byte aligned frames: TMA-staged kernel
    u32 tile_blocks;
    u64 nblocks, tiles_per_frame, n_tiles;
    size_t smem;                // dynamic shared memory per CTA
    size_t scratch_bytes;       // ticket + descriptors + tails
    bool ok;
};

template <typename T>
inline EncPlan enc_plan_t(const void* d_pixels, u64 n_values, u64 n_frames, u32 block)
{
    EncPlan pl;
    pl.ok = true;
    pl.nblocks = div_up(n_values, block);
    pl.fast = block == 12 && ((uintptr_t)d_pixels & 15) == 0 && ((n_values * sizeof(T)) & 15) == 0;
    if (pl.fast) {
        pl.tile_blocks = EncGeom<T, ENC_NT>::TILE_BLOCKS;
        pl.smem = EncGeom<T, ENC_NT>::SMEM_BYTES;
    } else {
        const u64 maxbits = 12 + (u64)block * (Pix<T>::W + (Pix<T>::SGN ? 1 : 0));
        const u64 cap_bits = (u64)GenGeom<T, GEN_NT>::STG_WORDS_MAX * 32;
        u64 tb = cap_bits / maxbits;
        if (tb > (u64)GEN_NT) tb = GEN_NT;
        if (tb == 0) { pl.ok = false; tb = 1; }            // block too large for one CTA's staging
        pl.tile_blocks = (u32)tb;
        pl.smem = SM_HEADER + (size_t)((tb * maxbits + 31) / 32 + 4) * 4;
    }
    pl.tiles_per_frame = div_up(pl.nblocks, pl.tile_blocks);
    pl.n_tiles = pl.tiles_per_frame * n_frames;
    if (pl.n_tiles >= (1ull << 31)) pl.ok = false;
    pl.scratch_bytes = 64 + (size_t)pl.n_tiles * 16;
    return pl;
}

inline EncPlan enc_plan(int dtype, const void* d_pixels, u64 n_values, u64 n_frames, u32 block)
{
    switch (dtype) {
    case DT_U8: return enc_plan_t<uint8_t>(d_pixels, n_values, n_frames, block);
    case DT_U16: return enc_plan_t<uint16_t>(d_pixels, n_values, n_frames, block);
    case DT_U32: return enc_plan_t<uint32_t>(d_pixels, n_values, n_frames, block);
    case DT_U64: return enc_plan_t<uint64_t>(d_pixels, n_values, n_frames, block);
    case DT_I8: return enc_plan_t<int8_t>(d_pixels, n_values, n_frames, block);
    case DT_I16: return enc_plan_t<int16_t>(d_pixels, n_values, n_frames, block);
    case DT_I32: return enc_plan_t<int32_t>(d_pixels, n_values, n_frames, block);
    default: return enc_plan_t<int64_t>(d_pixels, n_values, n_frames, block);
    }
}

template <typename T>
inline void encode_launch_t(Launcher& L, const EncPlan& pl, EncParams p, u32 ctas_per_sm)
{
    u64 grid = (u64)L.sm_count * ctas_per_sm;
    if (grid > pl.n_tiles) grid = pl.n_tiles;
    if (grid == 0) return;
    if (pl.fast)
        L.err = launch(terse_encode_kernel<T, ENC_NT>, (u32)grid, ENC_NT, pl.smem, L.stream, p);
    else
        L.err = launch(terse_encode_generic_kernel<T, GEN_NT>, (u32)grid, GEN_NT, pl.smem, L.stream, p,
                       pl.tile_blocks);
    L.count();
}

// scratch: pl.scratch_bytes of device memory (any content).  d_prolix_bits / d_status: 1 x u32 each.
inline void encode_async(Launcher& L, int dtype, const void* d_pixels, u64 n_values, u64 n_frames, u32 block,
                         void* d_out, u64 out_capacity, u64* d_frame_ends, u32* d_prolix_bits, u32* d_status,
                         void* scratch, const EncPlan& pl, u32 ctas_per_sm, u32 dbg_incl_stride = 0)
{
    L.err = cudaSuccess;
    EncParams p;
    p.pixels = d_pixels;
    p.n_values = n_values;
    p.n_frames = n_frames;
    p.block = block;
    p.nblocks = pl.nblocks;
    p.tiles_per_frame = pl.tiles_per_frame;
    p.n_tiles = pl.n_tiles;
    p.out_words = (u32*)d_out;
    p.out_capacity = out_capacity;
    p.frame_ends = d_frame_ends;
    p.prolix_bits = d_prolix_bits;
    p.status = d_status;
    p.ticket = (u32*)scratch;
    p.desc = (u64*)((unsigned char*)scratch + 64);
    p.tails = p.desc + pl.n_tiles;
    p.dbg_incl_stride = dbg_incl_stride;
    cudaMemsetAsync(scratch, 0, pl.scratch_bytes, L.stream);
    cudaMemsetAsync(d_prolix_bits, 0, sizeof(u32), L.stream);
    cudaMemsetAsync(d_status, 0, sizeof(u32), L.stream);
    switch (dtype) {
    case DT_U8: encode_launch_t<uint8_t>(L, pl, p, ctas_per_sm); break;
    case DT_U16: encode_launch_t<uint16_t>(L, pl, p, ctas_per_sm); break;
    case DT_U32: encode_launch_t<uint32_t>(L, pl, p, ctas_per_sm); break;
    case DT_U64: encode_launch_t<uint64_t>(L, pl, p, ctas_per_sm); break;
    case DT_I8: encode_launch_t<int8_t>(L, pl, p, ctas_per_sm); break;
    case DT_I16: encode_launch_t<int16_t>(L, pl, p, ctas_per_sm); break;
    case DT_I32: encode_launch_t<int32_t>(L, pl, p, ctas_per_sm); break;
    default: encode_launch_t<int64_t>(L, pl, p, ctas_per_sm); break;
    }
}

}  // namespace trpx
