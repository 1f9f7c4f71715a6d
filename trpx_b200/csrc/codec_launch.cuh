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

// pixel types 0..7; DT_F32 / DT_F64 only as OUTPUT types of the decoder (Terse.hpp:379-383)
enum { DT_U8 = 0, DT_U16, DT_U32, DT_U64, DT_I8, DT_I16, DT_I32, DT_I64, DT_F32, DT_F64 };

inline size_t dtype_size(int dt) { return dt < 0 || dt > DT_F64 ? 0 : dt == DT_F32 ? 4 : dt == DT_F64 ? 8 : (size_t)1 << (dt & 3); }
inline bool dtype_signed(int dt) { return dt >= DT_I8; }                 // (floating point takes signed streams)
inline bool dtype_is_pixel(int dt) { return dt >= 0 && dt <= DT_I64; }  // what the encoder accepts

#ifndef ENC_NT_U16
#define ENC_NT_U16 192
#endif
#ifndef ENC_NT_U8
#define ENC_NT_U8 256
#endif
#ifndef ENC_NT_U32
#define ENC_NT_U32 192
#endif
constexpr int ENC_NT = 256;     // worker threads per encoder CTA (8 warps; a thread owns 48 or 96 bytes of pixels)
template <typename T> struct EncNT { static constexpr int V = sizeof(T) == 2 ? ENC_NT_U16 : sizeof(T) == 1 ? ENC_NT_U8 : sizeof(T) == 4 ? ENC_NT_U32 : ENC_NT; };
constexpr int GEN_NT = 256;     // threads per generic-encoder CTA (one block per thread)

struct Launcher {
    cudaStream_t stream;
    u32 sm_count;               // SMs of the device (148 on B200)
    u64* launches;              // optional counter of kernels launched
    cudaError_t err;
    // optional profiling hook: called with a static name after every kernel (and once after the
    // memsets that open a call), so the owner can drop an event on the stream between kernels
    void (*mark_fn)(void* user, const char* name) = nullptr;
    void* mark_user = nullptr;
    void mark(const char* name) { if (mark_fn) mark_fn(mark_user, name); }
    void count(const char* name)
    {
#ifdef TRPX_EMU_TRACE
        fprintf(stderr, "emu: kernel %s done\n", name);
#endif
        if (launches) ++*launches;
        mark(name);
    }
};

// ---------------------------------------------------------------------------------- encode
struct EncPlan {
    bool fast;                  // block == 12 and 16-byte aligned frames: TMA-staged kernel
    u32 tile_blocks;
    u64 nblocks, tiles_per_frame, n_tiles, groups_per_frame, n_groups;
    u32 threads;                // threads per CTA
    size_t smem;                // dynamic shared memory per CTA
    size_t scratch_bytes;       // ticket + tile descriptors + tails + group descriptors
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
        pl.tile_blocks = EncGeom<T, EncNT<T>::V>::TILE_BLOCKS;
        pl.smem = EncGeom<T, EncNT<T>::V>::SMEM_BYTES;
        pl.threads = EncGeom<T, EncNT<T>::V>::THREADS;
    } else {
        pl.threads = GEN_NT;
        const u64 maxbits = 12 + (u64)block * (Pix<T>::W + (Pix<T>::SGN ? 1 : 0));
        const u64 cap_bits = (u64)GenGeom<T, GEN_NT>::STG_WORDS_MAX * 32;
        u64 tb = cap_bits / maxbits;
        if (tb > (u64)GEN_NT) tb = GEN_NT;
        if (tb == 0) { pl.ok = false; tb = 1; }            // block too large for one CTA's staging
        pl.tile_blocks = (u32)tb;
        pl.smem = SM_HEADER + (size_t)((tb * maxbits + 31) / 32 + 12) * 4;
    }
    pl.tiles_per_frame = div_up(pl.nblocks, pl.tile_blocks);
    pl.n_tiles = pl.tiles_per_frame * n_frames;
    if (pl.n_tiles >= (1ull << 31)) pl.ok = false;
    pl.groups_per_frame = div_up(pl.tiles_per_frame, GROUP);
    pl.n_groups = pl.groups_per_frame * n_frames;
    pl.scratch_bytes = 64 + (size_t)pl.n_tiles * 16 + (size_t)pl.n_groups * 8;
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
        L.err = launch(terse_encode_kernel<T, EncNT<T>::V>, (u32)grid, pl.threads, pl.smem, L.stream, p);
    else
        L.err = launch(terse_encode_generic_kernel<T, GEN_NT>, (u32)grid, GEN_NT, pl.smem, L.stream, p,
                       pl.tile_blocks);
    L.count(pl.fast ? "terse_encode" : "terse_encode_generic");
}

template <typename T>
inline const void* enc_kernel_t(bool fast)
{
    return fast ? (const void*)terse_encode_kernel<T, EncNT<T>::V> : (const void*)terse_encode_generic_kernel<T, GEN_NT>;
}
inline const void* enc_kernel(int dtype, bool fast)
{
    switch (dtype) {
    case DT_U8: return enc_kernel_t<uint8_t>(fast);
    case DT_U16: return enc_kernel_t<uint16_t>(fast);
    case DT_U32: return enc_kernel_t<uint32_t>(fast);
    case DT_U64: return enc_kernel_t<uint64_t>(fast);
    case DT_I8: return enc_kernel_t<int8_t>(fast);
    case DT_I16: return enc_kernel_t<int16_t>(fast);
    case DT_I32: return enc_kernel_t<int32_t>(fast);
    default: return enc_kernel_t<int64_t>(fast);
    }
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
    p.groups_per_frame = pl.groups_per_frame;
    p.ticket = (u32*)scratch;
    p.tdesc = (u64*)((unsigned char*)scratch + 64);
    p.tails = p.tdesc + pl.n_tiles;
    p.gdesc = p.tails + pl.n_tiles;
    p.dbg_incl_stride = dbg_incl_stride & 0xffffu;
    p.dbg_ring_words = dbg_incl_stride >> 16;
    cudaMemsetAsync(scratch, 0, pl.scratch_bytes, L.stream);
    cudaMemsetAsync(d_prolix_bits, 0, sizeof(u32), L.stream);
    cudaMemsetAsync(d_status, 0, sizeof(u32), L.stream);
    L.mark("memset");
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

// ---------------------------------------------------------------------------------- decode
#ifndef TRPX_WALK_NT
#define TRPX_WALK_NT 256
#endif
constexpr int WALK_NT = TRPX_WALK_NT;      // threads per CTA of the P1 walkers (one stream segment per thread)
constexpr int RESOLVE_NT = 256;   // threads per CTA of the cooperative verify / scan kernel
constexpr int SEGTAB_NT = 1024;

struct DecPlan {
    bool staged;                  // block == 12 and 16-byte aligned frames: fused re-walk + unpack kernel (TMA store)
    u64 nblocks, tiles_per_frame, n_tiles, max_segs;
    u32 last_cnt, seg_bytes, warm_bytes, subs_per_seg, sub_shift;
    size_t smem_unpack, off_ckpt, off_hdr_tab, off_segd;
    // scratch layout (byte offsets)
    size_t off_frame_ends, off_seg_base, off_seg_frame, off_seg_entry, off_seg_exit, off_seg_count,
        off_seg_b0, off_changed, off_zero_begin, off_widths, off_anchors, off_spec, scratch_bytes;
    bool ok;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// sub_shift: log2 of the checkpoint spacing in bits (8 = 32 bytes ... 5 = 4 bytes; 0 = 32 bytes)
inline DecPlan dec_plan(int out_dtype, u64 payload_bytes, u64 n_values, u64 n_frames, u32 block,
                        const void* d_out, u32 seg_bytes, u32 warm_bytes, u32 sub_shift = 0, bool force_ckpt = false)
{
    DecPlan pl;
    const size_t so = dtype_size(out_dtype);
    pl.ok = so != 0 && block != 0 && n_values != 0 && n_frames != 0;
    pl.nblocks = div_up(n_values, block ? block : 1);
    pl.last_cnt = (u32)(n_values - (pl.nblocks - 1) * block);
    pl.tiles_per_frame = div_up(pl.nblocks, DEC_TB);
    pl.n_tiles = pl.tiles_per_frame * n_frames;
    if (pl.n_tiles >= (1ull << 31)) pl.ok = false;
    pl.staged = force_ckpt || (block == 12 && ((uintptr_t)d_out & 15) == 0 && ((n_values * so) & 15) == 0);
    pl.sub_shift = sub_shift < SUB_SHIFT_MIN || sub_shift > SUB_SHIFT_MAX ? SUB_SHIFT_MAX : sub_shift;
    const u32 sub_bytes = 1u << (pl.sub_shift - 3);
    pl.seg_bytes = (seg_bytes < SUB_BYTES ? SUB_BYTES : seg_bytes + SUB_BYTES - 1) / SUB_BYTES * SUB_BYTES;
    pl.subs_per_seg = pl.seg_bytes / sub_bytes;
    pl.warm_bytes = warm_bytes;
    // the walkers keep lane-relative bit positions in 32 bits
    if ((u64)pl.seg_bytes + warm_bytes >= (1ull << 19) || (u64)block * 73 + 12 >= (1ull << 23)) pl.ok = false;
    pl.max_segs = payload_bytes / pl.seg_bytes + n_frames + 1;
    pl.smem_unpack = 128 + DEC_TB + 16 + (size_t)DEC_TB * 4;
    size_t o = 0;
    pl.off_frame_ends = o; o = align_up(o + n_frames * 8, 256);
    pl.off_seg_base = o;   o = align_up(o + (n_frames + 1) * 8, 256);
    pl.off_seg_frame = o;  o = align_up(o + pl.max_segs * 4, 256);
    pl.off_seg_entry = o;  o = align_up(o + pl.max_segs * 8, 256);
    pl.off_seg_exit = o;   o = align_up(o + pl.max_segs * 8, 256);
    pl.off_seg_count = o;  o = align_up(o + pl.max_segs * 4, 256);
    pl.off_seg_b0 = o;     o = align_up(o + pl.max_segs * 8, 256);
    pl.off_zero_begin = o;                                   // everything from here is zeroed per call
    pl.off_changed = o;    o = align_up(o + 16, 256);
    pl.off_hdr_tab = o;    o = align_up(o + HDR_TAB_BYTES, 256);
    pl.off_segd = o;       o = align_up(o + (pl.staged ? pl.max_segs * 32 : 0), 256);
    if (pl.staged) {                                         // fast path: checkpoints instead of widths + anchors
        pl.off_ckpt = o;   o = align_up(o + pl.max_segs * pl.subs_per_seg * 8, 256);
        pl.off_anchors = pl.off_widths = o;
    } else {
        pl.off_ckpt = o;
        pl.off_anchors = o; o = align_up(o + pl.n_tiles * 8, 256);
        pl.off_widths = o;  o = align_up(o + n_frames * pl.nblocks + 16, 256);
    }
    pl.off_spec = o;                                         // G plans: the speculative frame chain's state and table
    if (force_ckpt) o = align_up(o + spec_scratch_bytes(SPEC_B, SPEC_R0, SPEC_RS), 256);
    pl.scratch_bytes = o;
    return pl;
}

// Batch geometry of the speculative frame chain (frames per batch, window radius R0 + RS * k); the emulator tests
// shrink it so that window misses and re-anchoring are exercised.  Never larger than the defaults (scratch size).
inline u32 g_spec_params[4] = {SPEC_B, SPEC_R0, SPEC_RS, SPEC_MAX_STEPS};   // .. and T steps per candidate
constexpr u64 SPEC_MIN_FRAMES = 4, SPEC_MIN_BLOCKS = 64;

// CTAs of the cooperative resolve kernel: a warp puts its segments right one after another, so a small call wants
// few segments per warp (four), but no more CTAs than that takes -- the grid barriers cost by the CTA; one CTA scans
// one frame at the end.
inline u32 resolve_grid(const DecPlan& pl, u64 n_frames, u32 coop_grid)
{
    u64 g = div_up(pl.max_segs, (u64)(RESOLVE_NT / 32) * 4);
    if (g < n_frames) g = n_frames;
    if (g > coop_grid) g = coop_grid;
    return g ? (u32)g : 1u;
}

// Frame sizes unknown: the recovery pass (find_frames_async) first uses the scratch for the tables of the G plan --
// the payload as one frame, checkpoints forced -- and the recovered frame ends sit behind BOTH layouts.
inline DecPlan dec_plan_chain(const DecPlan& pl, int out_dtype, u64 payload_bytes, u32 block)
{
    return dec_plan(out_dtype, payload_bytes, block, 1, block, nullptr, pl.seg_bytes, pl.warm_bytes, pl.sub_shift, true);
}
inline size_t dec_chain_ends_offset(const DecPlan& pl, const DecPlan& gpl)
{
    return align_up(pl.scratch_bytes > gpl.scratch_bytes ? pl.scratch_bytes : gpl.scratch_bytes, 256);
}
// bytes of scratch a decode call needs
inline size_t dec_scratch_need(const DecPlan& pl, int out_dtype, u64 payload_bytes, u32 block, u64 n_frames, bool frames_unknown)
{
    if (!frames_unknown || n_frames <= 1) return pl.scratch_bytes + 256;
    const DecPlan gpl = dec_plan_chain(pl, out_dtype, payload_bytes, block);
    return dec_chain_ends_offset(pl, gpl) + align_up(n_frames * 8, 256);
}

template <typename O, bool SGN>
inline void unpack_launch_t(Launcher& L, const DecPlan& pl, const DecParams& p)
{
    if (pl.staged) {
        u64 grid = pl.max_segs * div_up(pl.subs_per_seg, UNP_NT);      // slices; persistent CTAs take them round-robin
        if (grid > (u64)L.sm_count * UNP_CTAS) grid = (u64)L.sm_count * UNP_CTAS;
        L.err = launch(prolix_unpack_seg_kernel<O, SGN>, (u32)grid, (u32)UNP_NT, (size_t)UNP_SMEM_BYTES, L.stream, p);
        L.count("prolix_unpack_seg");
    } else {
        L.err = launch(prolix_unpack_kernel<O, SGN, false>, (u32)pl.n_tiles, DEC_NT, pl.smem_unpack, L.stream, p);
        L.count("prolix_unpack");
    }
}
template <bool SGN>
inline void unpack_launch(Launcher& L, int out_dtype, const DecPlan& pl, const DecParams& p)
{
    switch (out_dtype) {
    case DT_U8: unpack_launch_t<uint8_t, SGN>(L, pl, p); break;
    case DT_U16: unpack_launch_t<uint16_t, SGN>(L, pl, p); break;
    case DT_U32: unpack_launch_t<uint32_t, SGN>(L, pl, p); break;
    case DT_U64: unpack_launch_t<uint64_t, SGN>(L, pl, p); break;
    case DT_I8: unpack_launch_t<int8_t, SGN>(L, pl, p); break;
    case DT_I16: unpack_launch_t<int16_t, SGN>(L, pl, p); break;
    case DT_I32: unpack_launch_t<int32_t, SGN>(L, pl, p); break;
    case DT_F32: unpack_launch_t<float, SGN>(L, pl, p); break;
    case DT_F64: unpack_launch_t<double, SGN>(L, pl, p); break;
    default: unpack_launch_t<int64_t, SGN>(L, pl, p); break;
    }
}

inline void fill_walk_params(DecParams& p, unsigned char* sc, const DecPlan& pl)
{
    p.seg_bytes = pl.seg_bytes;
    p.warm_bytes = pl.warm_bytes;
    p.max_segs = pl.max_segs;
    p.seg_base = (u64*)(sc + pl.off_seg_base);
    p.seg_frame = (u32*)(sc + pl.off_seg_frame);
    p.seg_entry = (u64*)(sc + pl.off_seg_entry);
    p.seg_exit = (u64*)(sc + pl.off_seg_exit);
    p.seg_count = (u32*)(sc + pl.off_seg_count);
    p.seg_b0 = (u64*)(sc + pl.off_seg_b0);
    p.changed = (u32*)(sc + pl.off_changed);
    p.widths = sc + pl.off_widths;
    p.anchors = (u64*)(sc + pl.off_anchors);
    p.ckpt = pl.staged ? (u64*)(sc + pl.off_ckpt) : nullptr;
    p.subs_per_seg = pl.subs_per_seg;
    p.sub_shift = pl.sub_shift;
    p.hdr_tab = (unsigned short*)(sc + pl.off_hdr_tab);
    p.segd = pl.staged ? (u64*)(sc + pl.off_segd) : nullptr;
}

// Frame boundaries of a payload whose frame sizes are unknown -> d_ends_out[n_frames] (Terse.hpp:562-585).  The whole
// payload is walked in parallel as ONE run of blocks (speculative walkers + resolve, chain_mode), then one warp follows
// the frames along the recorded checkpoints (prolix_frame_chain_kernel).  `scratch` holds the tables of
// dec_plan(.., n_frames = 1, .., force_ckpt = true) -- a prefix of what the call's own plan needs -- and d_ends_out
// must not overlap them.  Blocks too large for the chain kernel's window fall back to the one-warp walker.
inline void find_frames_async(Launcher& L, const void* d_payload, u64 payload_bytes, u32 block, u64 nblocks, u32 last_cnt,
                              u64 n_frames, u64* d_ends_out, u32* d_status, void* scratch, const DecPlan& gpl, u32 coop_grid)
{
    unsigned char* sc = (unsigned char*)scratch;
    DecParams p{};
    p.payload = (const u32*)d_payload;
    p.payload_bytes = payload_bytes;
    p.block = block;
    p.nblocks = nblocks;
    p.last_cnt = last_cnt;
    p.status = d_status;
    if (12 + (u64)block * 73 + 64 > (u64)FW_WORDS * 32 || !gpl.staged) {
        p.n_frames = n_frames;
        L.err = launch(prolix_find_frames_kernel, 1u, 32u, 0, L.stream, p, d_ends_out);
        L.count("prolix_find_frames");
        return;
    }
    // G: the payload as one frame of (arbitrarily many) full blocks
    fill_walk_params(p, sc, gpl);
    p.chain_mode = 1;
    p.n_frames = 1;
    p.n_values = 0;
    p.nblocks = ~0ull >> 2;
    u64* g_end = (u64*)(sc + gpl.off_frame_ends);
    cudaMemsetAsync(sc + gpl.off_zero_begin, 0, 256, L.stream);
    L.err = launch(prolix_single_frame_kernel, 1u, 32u, 0, L.stream, g_end, payload_bytes);
    if (L.err != cudaSuccess) return;
    p.frame_ends = g_end;
    L.err = launch(prolix_segments_kernel<SEGTAB_NT>, 1u, (u32)SEGTAB_NT, 0, L.stream, p);
    L.count("prolix_segments");
    if (L.err != cudaSuccess) return;
    const u32 walk_grid = (u32)div_up(gpl.max_segs, WALK_NT);
    const size_t walk_smem = HDR_TAB_BYTES + (size_t)(WALK_NT / 32) * WALK_BUF_WORDS * 4;
    L.err = launch(prolix_walk_kernel<WALK_NT>, walk_grid, (u32)WALK_NT, walk_smem, L.stream, p);
    L.count("prolix_walk");
    if (L.err != cudaSuccess) return;
    L.err = launch_coop(prolix_resolve_kernel<RESOLVE_NT>, resolve_grid(gpl, 1, coop_grid), (u32)RESOLVE_NT, 0, L.stream, p);
    L.count("prolix_resolve");
    if (L.err != cudaSuccess) return;
    // T: follow the frames -- speculatively in parallel, then serially whatever that pass left
    p.n_frames = n_frames;
    p.nblocks = nblocks;
    u64* spec = nullptr;
    if (n_frames >= SPEC_MIN_FRAMES && nblocks >= SPEC_MIN_BLOCKS) {
        spec = (u64*)(sc + gpl.off_spec);
        const u32 B = g_spec_params[0] < SPEC_B ? g_spec_params[0] : SPEC_B, R0 = g_spec_params[1] < SPEC_R0 ? g_spec_params[1] : SPEC_R0,
                  RS = g_spec_params[2] < SPEC_RS ? g_spec_params[2] : SPEC_RS;
        const u32 grid = coop_grid < 2 * (u32)L.sm_count ? coop_grid : 2 * (u32)L.sm_count;   // one candidate per thread
        L.err = launch_coop(prolix_frame_spec_kernel<SPEC_NT>, grid ? grid : 1u, (u32)SPEC_NT, 0, L.stream, p, d_ends_out, spec, B ? B : 1u, R0, RS, g_spec_params[3] ? g_spec_params[3] : 1u);
        L.count("prolix_frame_spec");
        if (L.err != cudaSuccess) return;
    }
    L.err = launch(prolix_frame_chain_kernel, 1u, 32u, 0, L.stream, p, d_ends_out, (const u64*)spec);
    L.count("prolix_frame_chain");
}

// d_frame_ends == nullptr: recover the frame boundaries from the stream first (and report them in
// d_frame_ends_out when given).  coop_grid: CTAs of the cooperative resolve kernel (all resident).
inline void decode_async(Launcher& L, const void* d_payload, u64 payload_bytes, bool is_signed, u32 block,
                         u64 n_values, u64 n_frames, const u64* d_frame_ends, u64* d_frame_ends_out,
                         void* d_out, int out_dtype, u32* d_status, void* scratch, const DecPlan& pl,
                         u32 coop_grid)
{
    L.err = cudaSuccess;
    unsigned char* sc = (unsigned char*)scratch;
    DecParams p;
    p.payload = (const u32*)d_payload;
    p.payload_bytes = payload_bytes;
    p.block = block;
    p.n_values = n_values;
    p.n_frames = n_frames;
    p.nblocks = pl.nblocks;
    p.last_cnt = pl.last_cnt;
    p.is_signed = is_signed ? 1 : 0;
    p.seg_bytes = pl.seg_bytes;
    p.warm_bytes = pl.warm_bytes;
    p.max_segs = pl.max_segs;
    p.seg_base = (u64*)(sc + pl.off_seg_base);
    p.seg_frame = (u32*)(sc + pl.off_seg_frame);
    p.seg_entry = (u64*)(sc + pl.off_seg_entry);
    p.seg_exit = (u64*)(sc + pl.off_seg_exit);
    p.seg_count = (u32*)(sc + pl.off_seg_count);
    p.seg_b0 = (u64*)(sc + pl.off_seg_b0);
    p.changed = (u32*)(sc + pl.off_changed);
    p.widths = sc + pl.off_widths;
    p.anchors = (u64*)(sc + pl.off_anchors);
    p.ckpt = pl.staged ? (u64*)(sc + pl.off_ckpt) : nullptr;
    p.subs_per_seg = pl.subs_per_seg;
    p.sub_shift = pl.sub_shift;
    p.hdr_tab = (unsigned short*)(sc + pl.off_hdr_tab);
    p.segd = pl.staged ? (u64*)(sc + pl.off_segd) : nullptr;
    p.tile_blocks = DEC_TB;
    p.tiles_per_frame = pl.tiles_per_frame;
    p.out = d_out;
    p.status = d_status;
    cudaMemsetAsync(d_status, 0, sizeof(u32), L.stream);
    // the fix-up flags always; widths (pre-zeroed: empty blocks are never written) and anchors on the generic path
    cudaMemsetAsync(sc + pl.off_zero_begin, 0, pl.staged ? 256 : pl.scratch_bytes - pl.off_zero_begin, L.stream);
    L.mark("memset");
    if (d_frame_ends == nullptr) {
        u64* fe = d_frame_ends_out ? d_frame_ends_out : (u64*)(sc + pl.off_frame_ends);
        if (n_frames == 1) {
            // a single frame ends where the payload ends; nothing to search for
            L.err = launch(prolix_single_frame_kernel, 1u, 32u, 0, L.stream, fe, payload_bytes);
            L.count("prolix_find_frames");
        } else {
            // (the G tables come first in the scratch, the call's own tables are written over them afterwards; the
            // recovered ends sit behind both: dec_scratch_need())
            const DecPlan gpl = dec_plan_chain(pl, out_dtype, payload_bytes, block);
            if (!d_frame_ends_out) fe = (u64*)(sc + dec_chain_ends_offset(pl, gpl));
            find_frames_async(L, d_payload, payload_bytes, block, pl.nblocks, pl.last_cnt, n_frames, fe, d_status, scratch, gpl, coop_grid);
            cudaMemsetAsync(sc + pl.off_zero_begin, 0, pl.staged ? 256 : pl.scratch_bytes - pl.off_zero_begin, L.stream);
        }
        if (L.err != cudaSuccess) return;
        d_frame_ends = fe;
    }
    p.frame_ends = d_frame_ends;
    L.err = launch(prolix_segments_kernel<SEGTAB_NT>, n_frames >= 64 ? 16u : 1u, (u32)SEGTAB_NT, 0, L.stream, p);
    L.count("prolix_segments");
    if (L.err != cudaSuccess) return;
    const u32 walk_grid = (u32)div_up(pl.max_segs, WALK_NT);
    const size_t walk_smem = HDR_TAB_BYTES + (size_t)(WALK_NT / 32) * WALK_BUF_WORDS * 4;
    L.err = launch(prolix_walk_kernel<WALK_NT>, walk_grid, (u32)WALK_NT, walk_smem, L.stream, p);
    L.count("prolix_walk");
    if (L.err != cudaSuccess) return;
    L.err = launch_coop(prolix_resolve_kernel<RESOLVE_NT>, resolve_grid(pl, n_frames, coop_grid), (u32)RESOLVE_NT, 0, L.stream, p);
    L.count("prolix_resolve");
    if (L.err != cudaSuccess) return;
    if (!pl.staged) {
        L.err = launch(prolix_emit_kernel<WALK_NT>, walk_grid, (u32)WALK_NT, walk_smem, L.stream, p);
        L.count("prolix_emit");
        if (L.err != cudaSuccess) return;
    }
    if (is_signed) unpack_launch<true>(L, out_dtype, pl, p);
    else unpack_launch<false>(L, out_dtype, pl, p);
}

}  // namespace trpx
