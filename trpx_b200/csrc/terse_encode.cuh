// terse_encode.cuh -- TERSE encoder for sm_100a: ONE fused pass over the pixels.
//
// Replaces Terse::f_compress (reference include/Terse.hpp:500-549) and what it calls:
//   K1 block width     <- OR-reduce + f_highest_set_bit       Terse.hpp:508-515, :551-560
//   K2 header/length   <- header emit                         Terse.hpp:517-535
//   K3 offsets         <- the serial Bit_pointer walk         Terse.hpp:504, :519-540, size rule :547
//   K4 bit packing     <- Bit_range::append_range             Bit_pointer.hpp:700-730
//
// Shape of the computation:
//   * The stack is ONE bit stream; a frame end rounds the position up to "1 + floor(bits/8)" bytes
//     (Terse.hpp:547).  A tile is a run of blocks inside one frame; its effect on the stream position
//     is P -> P + bits, or align(P + bits) when it ends its frame.  Those maps compose into
//     (has_end, a, c): P -> has_end ? align(P + a) + c : P + a, which is what the single-pass
//     decoupled look-back carries from tile to tile -- no second pass over the pixels, no
//     per-frame kernel.
//   * A persistent CTA takes tiles by ticket, prefetches the next tile with a TMA bulk copy
//     (cp.async.bulk + mbarrier) while it works on the current one, keeps each thread's 48 bytes
//     (4 / 2 / 1 blocks of 12 u8 / u16 / u32 values) in registers, packs the tile's bits into a
//     shared-memory staging area in tile-relative coordinates (so packing does not wait for the
//     look-back), then shifts the staging words by the tile's global bit offset while storing them
//     coalesced.  The word two tiles share is written by the later tile, which receives the earlier
//     tile's bits through a 64-bit hand-off word -- no global atomics, no pre-zeroed output.
#pragma once

#include "simt.cuh"

#ifndef ENC_UNIT_U8
#define ENC_UNIT_U8 96
#endif
#ifndef ENC_UNIT_U32
#define ENC_UNIT_U32 96
#endif
#ifndef ENC_UNIT_U16
#define ENC_UNIT_U16 96
#endif

namespace trpx {

// ------------------------------------------------------------------ look-back descriptors
// One 64-bit word per tile: [63:62] status, [61] "tile ends its frame", [60:0] value.
// AGG: value = bits of the tile; INCL: value = stream position (bits) after the tile.
constexpr u64 ST_INVALID = 0, ST_AGG = 1, ST_INCL = 2;
constexpr int ST_SHIFT = 62;
constexpr u64 ENDS_BIT = 1ull << 61;
constexpr u64 VAL_MASK = ENDS_BIT - 1;
constexpr u64 TAIL_VALID = 1ull << 32;

// frame end: the frame occupies 1 + floor(bits/8) bytes (Terse.hpp:547); the next frame starts there
TRPX_HD u64 align_frame(u64 bits) { return ((bits >> 3) + 1) << 3; }

struct FrameFn { u32 h; u64 a, c; };   // P -> h ? align_frame(P + a) + c : P + a
TRPX_HD FrameFn fn_identity() { FrameFn f; f.h = 0; f.a = 0; f.c = 0; return f; }
// g o f, where f is one tile (applied first): f(P) = ends ? align_frame(P + bits) : P + bits
TRPX_HD FrameFn fn_after_tile(FrameFn g, bool ends, u64 bits)
{
    FrameFn r;
    if (!ends) { r.h = g.h; r.a = g.a + bits; r.c = g.c; return r; }
    r.h = 1;
    r.a = bits;
    // f(P) is a multiple of 8, so align_frame(f(P) + a) = f(P) + align_frame(a)
    r.c = g.h ? align_frame(g.a) + g.c : g.a;
    return r;
}
TRPX_HD u64 fn_apply(FrameFn g, u64 P) { return g.h ? align_frame(P + g.a) + g.c : P + g.a; }

struct EncParams {
    const void* pixels;     // n_frames x n_values, frame-major
    u64 n_values;           // values per frame
    u64 n_frames;
    u32 block;              // values per block (12 on the fast path)
    u64 nblocks;            // blocks per frame
    u64 tiles_per_frame;
    u64 n_tiles;
    u64 groups_per_frame;   // ceil(tiles_per_frame / GROUP)
    u32* out_words;         // payload, 4-byte aligned
    u64 out_capacity;       // bytes
    u64* frame_ends;        // [n_frames] end byte offset of each frame
    u32* prolix_bits;       // [1], zeroed
    u32* status;            // [1], zeroed
    u64* tdesc;             // [n_tiles] zeroed: level-1 descriptors (bits of a tile)
    u64* gdesc;             // [n_groups] zeroed: level-2 descriptors (groups of <= 64 tiles of one frame)
    u64* tails;             // [n_tiles] zeroed: boundary-word hand-off
    u32* ticket;            // [1] zeroed
    u32 dbg_incl_stride;    // tests only: publish INCL for every k-th group only (0 = always)
    u32 dbg_ring_words;     // tests only: capacity of the staging ring in words (0 = EncGeom::RING_WORDS)
};

// ------------------------------------------------------------------ pixel-type traits
template <typename T>
struct Pix {
    static constexpr int SZ = (int)sizeof(T);
    static constexpr int W = 8 * SZ;
    static constexpr bool SGN = T(-1) < T(0);
    static constexpr int UNIT_BYTES = SZ == 8 ? 96 : SZ == 2 ? ENC_UNIT_U16 : SZ == 1 ? ENC_UNIT_U8 : ENC_UNIT_U32;    // bytes one thread owns (3 or 6 LDS.128)
    static constexpr int UW = UNIT_BYTES / 4;               // 32-bit words per unit
    static constexpr int VPU = UNIT_BYTES / SZ;             // values per unit: 48 / 24 / 12 / 12
    static constexpr int BPU = VPU / 12;                    // blocks per unit:  4 /  2 /  1 /  1
    static constexpr int BW = 12 * SZ / 4;                  // words per block:  3 /  6 / 12 / 24
    static constexpr int MAXBITS = 12 + 12 * (W + (SGN ? 1 : 0));   // worst block: header + data
};

TRPX_HD u32 abs32(u32 x) { u32 sx = (u32)((int)x >> 31); return (x ^ sx) - sx; }
TRPX_HD u64 abs64(u64 x) { u64 sx = (u64)((i64)x >> 63); return (x ^ sx) - sx; }

// K1: significant bits of a full 12-value block held in BW words (Terse.hpp:508-515, :551-560)
template <typename T>
TRPX_DEVICE u32 block_width12(const u32* w)
{
    typedef Pix<T> P;
    u32 s;
    bool nz;
    if (P::SZ == 8) {
        u64 m = 0;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            u64 v = (u64)w[2 * j] | ((u64)w[2 * j + 1] << 32);
            m |= P::SGN ? abs64(v) : v;
        }
        s = 64 - (u32)clz64(m);
        nz = m != 0;
    } else {
        u32 m = 0;
#pragma unroll
        for (int j = 0; j < P::BW; ++j)
            m |= !P::SGN ? w[j] : P::SZ == 1 ? vabs4(w[j]) : P::SZ == 2 ? vabs2(w[j]) : abs32(w[j]);
        if (P::SZ == 1) { m |= m >> 16; m |= m >> 8; m &= 0xffu; }
        if (P::SZ == 2) { m |= m >> 16; m &= 0xffffu; }
        s = 32 - (u32)clz32(m);
        nz = m != 0;
    }
    if (P::SGN && nz) s += 1;
    return s;
}

// K2: block header (Terse.hpp:517-535): value (LSB first) and length 1 / 4 / 6 / 12
TRPX_HD void block_header(u32 s, u32 prev, u32& hv, u32& hl)
{
    if (s == prev) { hv = 1; hl = 1; }
    else if (s < 7) { hv = s << 1; hl = 4; }
    else if (s < 10) { hv = (0x7u | ((s - 7) << 3)) << 1; hl = 6; }
    else { hv = (0x1Fu | ((s - 10) << 5)) << 1; hl = 12; }
}
// The explicit headers of all widths (0 .. 73) as a constant table, value | length << 16: the hot loop replaces
// the compare / select chain above by one indexed constant load and one select for "same width as before".
struct EncHdrTab { u32 v[80]; };
constexpr EncHdrTab make_enc_hdr_tab()
{
    EncHdrTab t{};
    for (u32 s = 0; s < 80; ++s) {
        u32 hv = 0, hl = 0;
        if (s < 7) { hv = s << 1; hl = 4; }
        else if (s < 10) { hv = (0x7u | ((s - 7) << 3)) << 1; hl = 6; }
        else { hv = (0x1Fu | ((s - 10) << 5)) << 1; hl = 12; }
        t.v[s] = hv | (hl << 16);
    }
    return t;
}
#ifndef TRPX_EMU
__constant__ EncHdrTab enc_hdr_tab = make_enc_hdr_tab();
#else
static const EncHdrTab enc_hdr_tab = make_enc_hdr_tab();
#endif
TRPX_DEVICE void block_header_fast(u32 s, u32 prev, u32& hv, u32& hl)      // s <= 73
{
    const u32 e = s == prev ? 0x00010001u : enc_hdr_tab.v[s];
    hv = e & 0xffffu;
    hl = e >> 16;
}

// ------------------------------------------------------------------ K4: per-thread bit sink
// Appends fields LSB-first (Bit_pointer.hpp:700-730) at a tile-relative bit offset.  Complete
// 32-bit words go to shared memory with plain stores, except the thread's FIRST word (shared with
// its predecessor) which stays in `head`, and the unfinished LAST word which stays in `acc`; both
// are resolved by merge_and_flush().
struct BitSink {
    u32* stg;               // (merge_and_flush reaches the head / tail words through it)
    saddr_t wp;             // shared-window address of the unfinished word
    u32 lo;                 // the unfinished word: nb < 32 valid bits
    u32 nb, w0, head;
    bool crossed;
    TRPX_DEVICE void init(u32* stg_, u32 off)
    {
        stg = stg_; w0 = off >> 5; wp = saddr(stg_) + w0 * 4; nb = off & 31; lo = 0; head = 0; crossed = false;
    }
    TRPX_DEVICE void put(u32 v, u32 n)     // n in [0, 32], v < 2^n
    {
        const u32 a0 = lo | (v << nb);
        const u32 a1 = funnel_l(v, 0u, nb);                     // v >> (32 - nb); 0 when nb == 0
        const u32 t = nb + n;                                   // < 64
        const bool c1 = t >= 32;
        sts_u32_if(c1 && crossed, wp, a0);
        head = c1 && !crossed ? a0 : head;
        crossed = crossed || c1;
        lo = c1 ? a1 : a0;
        wp += (t >> 3) & 4u;                                    // one word further when a word completed
        nb = t & 31;
    }
    // n in [0, 64], v < 2^n.  Same contract as put(); up to two words complete per call.  Straight-line
    // code (selects and predicated stores), so lanes with different widths stay converged.
    TRPX_DEVICE void put64(u64 v, u32 n)
    {
        const u32 v0 = (u32)v, v1 = (u32)(v >> 32);
        const u32 a0 = lo | (v0 << nb);                         // bits  0..31 of lo | v << nb
        const u32 a1 = funnel_l(v0, v1, nb);                    // bits 32..63
        const u32 a2 = funnel_l(v1, 0u, nb);                    // bits 64..95
        const u32 t = nb + n;                                   // < 96
        const bool c1 = t >= 32, c2 = t >= 64;
        sts_u32_if(c1 && crossed, wp, a0);
        head = c1 && !crossed ? a0 : head;
        sts_u32_if(c2, wp + 4, a1);
        crossed = crossed || c1;
        lo = c2 ? a2 : (c1 ? a1 : a0);
        wp += (t >> 3) & 12u;                                   // 4 bytes per completed word (0, 1 or 2)
        nb = t & 31;
    }
    TRPX_DEVICE void put_wide(u64 v, u32 s)   // low s bits of the sign-extended value, s in [1, 65]
    {
        u32 n0 = s < 32 ? s : 32;
        u32 lo = (u32)v;
        if (n0 < 32) lo &= (1u << n0) - 1;
        put(lo, n0);
        if (s > 32) {
            u32 n1 = s - 32 < 32 ? s - 32 : 32;
            u32 hi = (u32)(v >> 32);
            if (n1 < 32) hi &= (1u << n1) - 1;
            put(hi, n1);
        }
        if (s > 64) put((u32)(v >> 63) & 1u, 1);
    }
};

// value i of a block held in words, as a sign-extended 64-bit pattern
template <typename T>
TRPX_DEVICE u64 block_value(const u32* w, int i)
{
    typedef Pix<T> P;
    if (P::SZ == 1) { u32 b = (w[i >> 2] >> (8 * (i & 3))) & 0xffu; return P::SGN ? (u64)(i64)(int8_t)b : b; }
    if (P::SZ == 2) { u32 h = (w[i >> 1] >> (16 * (i & 1))) & 0xffffu; return P::SGN ? (u64)(i64)(int16_t)h : h; }
    if (P::SZ == 4) return P::SGN ? (u64)(i64)(int)w[i] : (u64)w[i];
    return (u64)w[2 * i] | ((u64)w[2 * i + 1] << 32);
}

template <typename T>
TRPX_DEVICE u64 block_value_dyn(const u32* w, u32 i)       // the same, for an index only known at run time (w in memory)
{
    typedef Pix<T> P;
    if (P::SZ == 1) { u32 b = (w[i >> 2] >> (8 * (i & 3))) & 0xffu; return P::SGN ? (u64)(i64)(int8_t)b : b; }
    if (P::SZ == 2) { u32 h = (w[i >> 1] >> (16 * (i & 1))) & 0xffffu; return P::SGN ? (u64)(i64)(int16_t)h : h; }
    if (P::SZ == 4) return P::SGN ? (u64)(i64)(int)w[i] : (u64)w[i];
    return (u64)w[2 * i] | ((u64)w[2 * i + 1] << 32);
}

// K4: the data of one block (cnt values of s bits).  Full blocks of the common widths are merged
// pairwise / quadwise in registers first (all fields of a block share s), so a 12-value block costs
// 3 (s <= 8) or 6 sink operations instead of 12.
template <typename T>
TRPX_DEVICE void pack_block12(BitSink& sk, const u32* w, u32 s, u32 cnt)
{
    typedef Pix<T> P;
    if (s == 0) return;
    if (cnt == 12) {
        if (P::SZ == 2 && s <= 16) {
            // two 16-bit halves -> one field of 2s bits with a single multiply-add:
            // lo + hi*2^16 + hi*(2^s - 2^16) == lo + hi*2^s; two such pairs -> one 64-bit put
            const u32 m = ((1u << s) - 1) * 0x00010001u;             // s <= 16: per-half mask (signed only)
            const u32 K = (1u << s) - 65536u;
            const u32 s2 = 2 * s;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const u32 x0 = P::SGN ? (w[2 * q] & m) : w[2 * q], x1 = P::SGN ? (w[2 * q + 1] & m) : w[2 * q + 1];
                const u32 p0 = x0 + (x0 >> 16) * K, p1 = x1 + (x1 >> 16) * K;
                sk.put64((u64)p0 | ((u64)p1 << s2), 2 * s2);
            }
            return;
        }
        if (P::SZ == 1 && s <= 8) {
            const u32 m = P::SGN ? ((1u << s) - 1) * 0x00010001u : 0x00ff00ffu;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                u32 a = w[q] & 0x00ff00ffu & m, b = (w[q] >> 8) & 0x00ff00ffu & m;
                u32 pr = a | (b << s);                               // two 16-bit lanes of 2s bits
                if (s == 8) sk.put(pr, 32);
                else sk.put((pr & 0xffffu) | ((pr >> 16) << (2 * s)), 4 * s);
            }
            return;
        }
        if (P::SZ == 4 && s <= 32) {
            const u32 m = s == 32 ? 0xffffffffu : (1u << s) - 1;
#pragma unroll
            for (int i = 0; i < 6; ++i)
                sk.put64((u64)(w[2 * i] & m) | ((u64)(w[2 * i + 1] & m) << s), 2 * s);
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i)
        if ((u32)i < cnt) sk.put_wide(block_value<T>(w, i), s);
}

// ------------------------------------------------------------------ K3: block-wide exclusive scan
// len -> tile-relative bit offset; warp shuffles + one shared round.  Contains ONE sync_block().
template <int NT>
TRPX_DEVICE void scan_lengths(u32 len, u32* sm_warp_tot, u32& off, u32& tile_bits, u32 named_bar = 0)
{
    const u32 t = tid(), lane = t & 31, warp = t >> 5;
    u32 incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 v = shfl_up(incl, d);
        if (lane >= (u32)d) incl += v;
    }
    if (lane == 31) sm_warp_tot[warp] = incl;
    if (named_bar) bar_sync(named_bar, NT); else sync_block();
    const u32 wt = lane < (u32)(NT / 32) ? sm_warp_tot[lane] : 0u;    // NT / 32 <= 32 warps
    const u32 base = warp_add(lane < warp ? wt : 0u);
    off = base + incl - len;
    tile_bits = warp_add(wt);
}

// Words that several warps may touch (each warp's first word, and the word after the tile's last bit)
// are accumulated with shared-memory atomics and therefore zeroed first.  Caller syncs afterwards.
template <int NT>
TRPX_DEVICE void zero_boundary_words(u32* stg, u32 off, u32 tile_bits)
{
    const u32 t = tid();
    if ((t & 31) == 0) stg[off >> 5] = 0;
    if (t == NT - 1) { stg[tile_bits >> 5] = 0; stg[(tile_bits >> 5) + 1] = 0; stg[(tile_bits >> 5) + 2] = 0; }   // partial last word + zero padding
}

// Resolve the partial words inside a warp without atomics.  OR of disjoint bit fields == ADD, so
// the bits carried into lane t's first word are a difference of two warp prefix sums of the lanes'
// unfinished last words: sum over lanes [p, t) where p is the last lane before t that completed a
// word.  Only the warp's first completed word and its outgoing tail can be shared with other warps;
// those two use shared-memory atomics (<= 2 per warp per tile).
TRPX_DEVICE void merge_and_flush(BitSink& sk, u32 end_off)
{
    const u32 lane = tid() & 31;
    const u32 tail = sk.lo;
    u32 incl = tail;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 v = shfl_up(incl, d);
        if (lane >= (u32)d) incl += v;
    }
    const u32 excl = incl - tail;
    const u32 cmask = ballot(sk.crossed);
    const u32 below = cmask & ((1u << lane) - 1);
    const u32 qp = shfl(excl, below ? 31 - clz32(below) : 0);
    const u32 carry = excl - (below ? qp : 0u);
    if (sk.crossed) {
        u32 word = sk.head | carry;
        if (below == 0) atomic_or(&sk.stg[sk.w0], word);
        else sk.stg[sk.w0] = word;
    }
    const u32 ql = shfl(excl, cmask ? 31 - clz32(cmask) : 0);
    if (lane == 31) {
        u32 out = incl - (cmask ? ql : 0u);
        if (out) atomic_or(&sk.stg[end_off >> 5], out);
    }
}

// ------------------------------------------------------------------ two-level look-back
// A flat decoupled look-back tops out near (32 descriptors / L2 round trip) ~ 32 tiles per
// microsecond -- far below the ~300 tiles/us this kernel has to sustain with 12 KB tiles.  So the chain
// has two levels.  Tiles are grouped into GROUPs of <= 64 consecutive tiles of ONE frame:
//   level 1  tdesc[tile] = VALID | bits.  A tile's offset inside its group is the sum of its <= 63
//            predecessors' bits: one coalesced read of the group's descriptors, no chain at all.
//   level 2  gdesc[group] = status | ends | value, the frame-aware look-back described at the top of
//            this file, but over groups (64x fewer links); the group's last tile publishes it.
constexpr u32 GROUP = 64;
constexpr u64 TD_VALID = 1ull << 63;

struct TileGeom { u64 frame, tif, group, gfirst_tile; u32 j, m; bool ends, gends; };
TRPX_DEVICE TileGeom tile_geom(const EncParams& p, u64 tile)
{
    TileGeom g;
    const u32 tpf = (u32)p.tiles_per_frame, fr = (u32)tile / tpf;   // n_tiles < 2^31 (plan)
    g.frame = fr;
    g.tif = (u32)tile - fr * tpf;
    const u64 gif = g.tif / GROUP;                         // group inside the frame
    g.group = g.frame * p.groups_per_frame + gif;
    g.j = (u32)(g.tif % GROUP);
    const u64 left = p.tiles_per_frame - gif * GROUP;
    g.m = left < GROUP ? (u32)left : GROUP;
    g.gfirst_tile = tile - g.j;
    g.ends = g.tif + 1 == p.tiles_per_frame;
    g.gends = gif + 1 == p.groups_per_frame;
    return g;
}

// Level 2: stream position (bits) at which `group` starts; executed by one whole warp.
TRPX_DEVICE u64 lookback_first_round(const u64* desc, u64 tile)   // issue round 0's load early
{
    const i64 idx = (i64)tile - 1 - (i64)(tid() & 31);
    return idx >= 0 ? ld_relaxed(&desc[idx]) : (ST_INCL << ST_SHIFT);
}
TRPX_DEVICE u64 lookback_start(const u64* desc, u64 tile, u64 d_first, u32* status)
{
    if (tile == 0) return 0;
    const u32 lane = tid() & 31;
    FrameFn g = fn_identity();
    i64 base = (i64)tile - 1;
    WaitClock wc;
    for (u32 round = 0;; ++round) {
        const i64 idx = base - (i64)lane;
        u64 d = d_first;
        u32 first_incl;
        for (u32 spins = 0;; ++spins) {
            if (round | spins) d = idx >= 0 ? ld_relaxed(&desc[idx]) : (ST_INCL << ST_SHIFT);   // "tile -1": position 0
            const u32 st = (u32)(d >> ST_SHIFT);
            const u32 incl_mask = ballot(st == ST_INCL);
            const u32 inval_mask = ballot(st == ST_INVALID);
            first_incl = incl_mask ? (u32)ffs32(incl_mask) - 1 : 32;
            const u32 needed = first_incl >= 32 ? 0xffffffffu : ((1u << first_incl) - 1);
            if ((inval_mask & needed) == 0) break;
            if (any_lane(wc.expired(status, spins, 21, 1))) return 0;   // never hang the device
            spin_hint();
        }
        for (u32 l = 0; l < first_incl; ++l) {            // nearest first: g <- g o f_l
            const u64 dl = shfl(d, (int)l);
            g = fn_after_tile(g, (dl & ENDS_BIT) != 0, dl & VAL_MASK);
        }
        if (first_incl < 32) return fn_apply(g, shfl(d, (int)first_incl) & VAL_MASK);
        base -= 32;
    }
}

// Both levels; executed by one whole warp (all lanes return the tile's start position P0).
TRPX_DEVICE u64 tile_start(const EncParams& p, u64 tile, const TileGeom& g, u32 tile_bits, bool publish)
{
    const u32 lane = tid() & 31;
    if (publish && lane == 0) st_relaxed(&p.tdesc[tile], TD_VALID | (u64)tile_bits);
    const u64 dg = lookback_first_round(p.gdesc, g.group);  // in flight together with the level-1 loads
    // level 1: bits of the tiles before me in my group
    const bool need0 = lane < g.j, need1 = lane + 32 < g.j;
    u64 d0 = 0, d1 = 0;
    WaitClock wc;
    for (u32 spins = 0;; ++spins) {
        if (need0 && !(d0 & TD_VALID)) d0 = ld_relaxed(&p.tdesc[g.gfirst_tile + lane]);
        if (need1 && !(d1 & TD_VALID)) d1 = ld_relaxed(&p.tdesc[g.gfirst_tile + 32 + lane]);
        const bool ok = (!need0 || (d0 & TD_VALID)) && (!need1 || (d1 & TD_VALID));
        if (all_lanes(ok)) break;
        if (any_lane(wc.expired(p.status, spins, 21, 2))) break;
        spin_hint();
    }
    const u32 before = warp_add((need0 ? (u32)d0 : 0u) + (need1 ? (u32)d1 : 0u));
    // level 2: the group's start; its last tile also publishes the group's aggregate and end position
    const bool last = g.j + 1 == g.m;
    const u64 gbits = (u64)before + tile_bits;
    if (last && lane == 0)
        st_relaxed(&p.gdesc[g.group], (ST_AGG << ST_SHIFT) | (g.gends ? ENDS_BIT : 0) | gbits);
    const u64 Pg = lookback_start(p.gdesc, g.group, dg, p.status);
    if (last && lane == 0 && (p.dbg_incl_stride == 0 || g.group % p.dbg_incl_stride == 0))
        st_relaxed(&p.gdesc[g.group], (ST_INCL << ST_SHIFT) | (g.gends ? align_frame(Pg + gbits) : Pg + gbits));
    return Pg + before;
}

// word i of a tile's output window: staging words shifted left by the start position's bit offset
// (stg[-1] is a zero word, and the word after the tile's last one is zeroed: no bounds checks)
TRPX_DEVICE u32 window_word(const u32* stg, u32 nstg, u32 i, u32 sh)
{
    (void)nstg;
    return funnel_l(stg[(int)i - 1], stg[i], sh);
}

// The word two neighbouring tiles share is stored by the LATER tile; the earlier one hands its bits of
// that word over through tails[].  One thread.  Returns the predecessor's bits of our first word.
TRPX_DEVICE u32 tail_handoff(const EncParams& p, u64 tile, const u32* stg, u32 tile_bits, u64 P0, u64 Pn, u32& tout, u64 tin_known = 0)
{
    const u32 nstg = (tile_bits + 31) >> 5;
    const u32 k = (u32)((Pn >> 5) - (P0 >> 5));            // complete words this tile owns
    const u32 sh = (u32)(P0 & 31);
    // Our bits of the word the NEXT tile starts in.  When we own a complete word (k >= 1) they do not
    // depend on our predecessor: publish before waiting, so the hand-off never chains across tiles.
    // (only bits below Pn belong to us; a frame end may leave Pn up to 8 bits past the staged data)
    tout = window_word(stg, nstg, k, sh) & ((1u << (Pn & 31)) - 1);
    if (k >= 1) st_relaxed(&p.tails[tile], TAIL_VALID | (u64)tout);
    u64 tin = tin_known;                                   // (the caller may have read the predecessor's word already)
    if (sh != 0 && !(tin & TAIL_VALID)) {                  // the word we start in is ours to store
        WaitClock wc;
        for (u32 spins = 0;; ++spins) {
            tin = ld_relaxed(&p.tails[tile - 1]);
            if (tin & TAIL_VALID) break;
#ifdef TRPX_EMU
            if (spins == (1u << 17)) fprintf(stderr, "emu: tail wait: bid %u warp %u tile %llu P0 %llu Pn %llu k %u\n", bid(), tid() >> 5, tile, P0, Pn, k);
#endif
            if (wc.expired(p.status, spins, 21, 3)) break;
            spin_hint();
        }
    }
    tin = sh != 0 ? tin & 0xffffffffull : 0;
    if (k == 0) { tout |= (u32)tin; st_relaxed(&p.tails[tile], TAIL_VALID | (u64)tout); }
    return (u32)tin;
}

// Store the words a tile owns, shifted into place; `rank` of `nthr` cooperating threads.
TRPX_DEVICE void store_tile(const EncParams& p, u64 tile, const TileGeom& g, const u32* stg, u32 tile_bits,
                            u64 P0, u64 Pn, u32 tail_in, u32 tout, u32 rank, u32 nthr)
{
    const u32 nstg = (tile_bits + 31) >> 5;
    const u64 W0 = P0 >> 5, Wn = Pn >> 5;
    const u32 sh = (u32)(P0 & 31);
    const u32 k = (u32)(Wn - W0);
    const bool fits = ((Pn + 7) >> 3) <= p.out_capacity;
    if (fits) {
        u32* outw = p.out_words + W0;
        if (rank == 0 && k) st_stream(outw, window_word(stg, nstg, 0, sh) | tail_in);   // the word shared with the previous tile
        for (u32 i = rank ? rank : nthr; i < k; i += nthr) st_stream(outw + i, window_word(stg, nstg, i, sh));
    } else if (rank == 0) {
        atomic_max(p.status, 2u);                          // TRPX_ERR_CAPACITY
    }
    if (rank == 0) {
        if (g.ends) p.frame_ends[g.frame] = Pn >> 3;
        if (tile + 1 == p.n_tiles && fits) {               // nobody follows: store the final bytes
            unsigned char* ob = (unsigned char*)p.out_words;
            for (u32 b = 0; b < (u32)((Pn >> 3) & 3); ++b) ob[Wn * 4 + b] = (unsigned char)(tout >> (8 * b));
        }
    }
}

// In-line variant (generic kernel): warp 0 resolves, then the whole CTA stores.
//   bc[0] = P0, bc[1] = tail_in, bc[2] = tail_out   (written by thread 0, read by all after the sync)
template <int NT>
TRPX_DEVICE void resolve_and_store(const EncParams& p, u64 tile, u32 tile_bits, const u32* stg, u64* bc)
{
    const u32 t = tid();
    const TileGeom g = tile_geom(p, tile);
    if (t < 32) {
        const u64 P0 = tile_start(p, tile, g, tile_bits, true);
        if (t == 0) {
            const u64 Pn = g.ends ? align_frame(P0 + tile_bits) : P0 + tile_bits;
            u32 tout;
            const u32 tin = tail_handoff(p, tile, stg, tile_bits, P0, Pn, tout);
            bc[0] = P0;
            bc[1] = tin;
            bc[2] = tout;
        }
    }
    sync_block();
    const u64 P0 = bc[0];
    const u64 Pn = g.ends ? align_frame(P0 + tile_bits) : P0 + tile_bits;
    store_tile(p, tile, g, stg, tile_bits, P0, Pn, (u32)bc[1], (u32)bc[2], t, NT);
}

// ------------------------------------------------------------------ shared-memory layout
constexpr int ENC_STAGES = 2;       // TMA pixel stages
constexpr int ENC_DEPTH = 6;        // tiles a CTA may have packed but not yet stored (mailbox entries)
constexpr int ENC_RESOLVERS = 2;    // resolver warps; resolver r serves iterations it % ENC_RESOLVERS == r
static_assert(ENC_DEPTH % ENC_RESOLVERS == 0, "a mailbox entry is always served by the same resolver");
static_assert(2 + 2 * ENC_DEPTH <= 16, "named barriers: 0 CTA, 1 workers, 2.. ready, 2+DEPTH.. packed");
constexpr int SM_BARS = 0;          // mbarriers, 8 bytes each: full[STAGES], ready[DEPTH], packed[DEPTH], resolved[DEPTH]
constexpr int SM_TICKETS = 192;     // 2 x ENC_STAGES u32: tile, tile-in-frame
constexpr int SM_WARP_TOT = 256;    // 32 u32
constexpr int SM_WARP_LAST = 384;   // 32 u32
constexpr int SM_BCAST = 512;       // 4 u64 (generic kernel)
constexpr int SM_MAX = 544;         // u32 running max width
constexpr int SM_MAIL = 576;        // ENC_DEPTH x {u64 tile, u64 P0, u32 bits, u32 tail_in, u32 tail_out, u32 ring word} (32 bytes each)
constexpr int SM_HEADER = 1024;
static_assert(8 * (ENC_STAGES + 3 * ENC_DEPTH) <= SM_TICKETS && SM_MAIL + 40 * ENC_DEPTH <= SM_HEADER, "shared-memory header layout");
constexpr u64 TILE_END = ~0ull;

template <typename T, int NT>
struct EncGeom {
    typedef Pix<T> P;
    static constexpr int TILE_BYTES = NT * P::UNIT_BYTES;
    static constexpr int TILE_BLOCKS = NT * P::BPU;
    static constexpr int STAGE_BYTES = ((P::UNIT_BYTES + TILE_BYTES + 127) / 128) * 128;   // halo + tile
    // Packed tiles wait in a ring of words until their stream position is known.  A tile takes
    // (bits / 32 + 1) words plus a zero word on either side (window_word), rounded to 4: ~0.7 K words
    // for a typical 512^2 u16 tile, WORST_WORDS when nothing compresses.  The ring always holds two
    // worst-case tiles; with typical data ENC_DEPTH tiles are in flight.
    static constexpr int WORST_WORDS = ((TILE_BLOCKS * P::MAXBITS + 31) / 32 + 1 + 3 + 3) / 4 * 4;
    static constexpr int RING_WORDS = (2 * WORST_WORDS <= 8192 || (P::SZ <= 4 && WORST_WORDS <= 8192)) ? 8192 : (2 * WORST_WORDS <= 16384 ? 16384 : 32768);   // a power of two
    static constexpr int SMEM_BYTES = SM_HEADER + ENC_STAGES * STAGE_BYTES + RING_WORDS * 4;
    static constexpr int THREADS = NT + 32 * ENC_RESOLVERS;   // worker warps + resolver warps
};

// ------------------------------------------------------------------ fast kernel: block == 12, 16-byte aligned frames
// Warp-specialised persistent CTA.
//   workers (NT threads)   wait for a TMA-staged tile, keep their 48 bytes in registers, compute widths /
//                          headers / lengths, scan, publish the tile's bit count, and pack the tile into
//                          the staging ring in tile-relative coordinates.  Packed tiles are stored --
//                          shifted into place, coalesced streaming stores -- up to ENC_DEPTH iterations
//                          later, by which time their stream position has long been resolved: workers
//                          never wait for a global-memory round trip.  Their blocking points are the TMA
//                          full-barrier (prefetched two tiles ahead) and a resolution that is really late.
//   resolvers (ENC_RESOLVERS warps, alternating tiles)
//                          find the tile's stream position with the two-level look-back and exchange the
//                          boundary word with the neighbour tile: pure latency, off the workers' path.
template <typename T, int NT>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(NT + 32 * ENC_RESOLVERS, (EncGeom<T, NT>::SMEM_BYTES > 75 * 1024 ? 2 : 3)) terse_encode_kernel(EncParams p)
{
    typedef Pix<T> P;
    typedef EncGeom<T, NT> G;
    TRPX_DYN_SMEM(sm);
    u64* bars = (u64*)(sm + SM_BARS);
    u64* bar_full = bars;
    u64* bar_resolved = bars + ENC_STAGES;              // (ready / packed are named barriers 2.. and 2+DEPTH..)
    u32* tickets = (u32*)(sm + SM_TICKETS);
    u32* vbases = (u32*)(sm + SM_TICKETS + 32);                         // [ENC_DEPTH] virtual ring offset of a pending tile
    volatile u64* pn64 = (volatile u64*)(sm + SM_MAIL + 32 * ENC_DEPTH);  // [ENC_DEPTH] end position of a resolved tile
    u32* sm_warp_tot = (u32*)(sm + SM_WARP_TOT);
    u32* sm_warp_last = (u32*)(sm + SM_WARP_LAST);
    u32* sm_max = (u32*)(sm + SM_MAX);
    volatile u64* mail64 = (volatile u64*)(sm + SM_MAIL);               // [e * 4 + {0: tile, 1: P0}]
    volatile u32* mail32 = (volatile u32*)(sm + SM_MAIL);               // [e * 8 + {4: bits, 5: tail_in, 6: tail_out, 7: ring word}]
    unsigned char* stages = sm + SM_HEADER;
    u32* ring = (u32*)(stages + ENC_STAGES * G::STAGE_BYTES);

    const u32 t = tid(), lane = t & 31, warp = t >> 5;
    const u64 frame_bytes = p.n_values * P::SZ;

    if (t == 0) {
        for (int s = 0; s < ENC_STAGES; ++s) mbar_init(&bar_full[s], 1);
        for (int e = 0; e < ENC_DEPTH; ++e) {
            mbar_init(&bar_resolved[e], 1);
        }
        mbar_init_fence();
        *sm_max = 0;
    }
    sync_block();

    if (t >= (u32)NT) {
        // ================================================================ resolver warp r
        for (u32 it = (t - NT) >> 5;; it += ENC_RESOLVERS) {
            const u32 e = it % ENC_DEPTH, use = it / ENC_DEPTH;
            bar_sync(2 + e, 64);                           // blocks in hardware until worker warp 0 has posted the tile (no spin)
            const u64 tile = mail64[e * 4];
            const u32 tile_bits = mail32[e * 8 + 4];
            if (tile == TILE_END) break;
            const u32* stg = ring + mail32[e * 8 + 7];
            const TileGeom g = tile_geom(p, tile);
            const u64 P0 = tile_start(p, tile, g, tile_bits, false);
            const u64 Pn = g.ends ? align_frame(P0 + tile_bits) : P0 + tile_bits;
            bar_sync(2 + ENC_DEPTH + e, NT + 32);          // all workers have packed: their staging stores are visible now
            if (lane == 0) {
                u32 tout;
                const u32 tin = tail_handoff(p, tile, stg, tile_bits, P0, Pn, tout);
                mail64[e * 4 + 1] = P0;
                mail32[e * 8 + 5] = tin;
                mail32[e * 8 + 6] = tout;
                pn64[e] = Pn;
                mbar_arrive(&bar_resolved[e]);
            }
            sync_warp();
        }
        return;
    }

    // ==================================================================== worker warps
    // thread 0 is also the TMA producer: take the next ticket, start that tile's bulk copy
    auto issue = [&](int s) {
        const u32 tk = atomic_add(p.ticket, 1u);
        tickets[s] = tk;
        if ((u64)tk < p.n_tiles) {
            const u32 tpf = (u32)p.tiles_per_frame;
            const u64 f = tk / tpf, tif = tk - (u32)f * tpf;
            tickets[ENC_STAGES + s] = (u32)tif;
            const u64 tile_off = tif * (u64)G::TILE_BYTES;
            u64 bytes = frame_bytes - tile_off;
            if (bytes > (u64)G::TILE_BYTES) bytes = G::TILE_BYTES;
            const unsigned char* src = (const unsigned char*)p.pixels + f * frame_bytes + tile_off;
            unsigned char* dst = stages + s * G::STAGE_BYTES + P::UNIT_BYTES;
            if (tif > 0) { src -= P::UNIT_BYTES; dst -= P::UNIT_BYTES; bytes += P::UNIT_BYTES; }   // halo: previous block
            mbar_arrive_expect_tx(&bar_full[s], (u32)bytes);
            bulk_g2s(dst, src, (u32)bytes, &bar_full[s]);
        }
    };
    // store the tile packed in iteration `j`, once resolved
    auto store_pending = [&](u32 j) {
        const u32 e = j % ENC_DEPTH;
        mbar_wait(&bar_resolved[e], (j / ENC_DEPTH) & 1, p.status, 6);
        const u64 ptile = mail64[e * 4];
        const u64 P0 = mail64[e * 4 + 1], Pn = pn64[e];
        const u32 pbits = mail32[e * 8 + 4];
        TileGeom g;
        if (t == 0) g = tile_geom(p, ptile);               // only rank 0 needs the frame bookkeeping
        else { g.frame = 0; g.ends = false; }
        store_tile(p, ptile, g, ring + mail32[e * 8 + 7], pbits, P0, Pn, mail32[e * 8 + 5], mail32[e * 8 + 6], t, NT);
    };
    if (t == 0)
        for (int s = 0; s < ENC_STAGES; ++s) issue(s);
    bar_sync(1, NT);

    // ring bookkeeping, identical in every worker thread: virtual word offsets that only grow; the
    // tiles of iterations [oldest, it) are packed but not stored and occupy [vtail, vhead)
    u32 vhead = 0, vtail = 0, oldest = 0;
    const u32 ring_words = p.dbg_ring_words ? p.dbg_ring_words : (u32)G::RING_WORDS;   // a power of two (tests vary it)
    const u32 ring_mask = ring_words - 1;
    u32 my_max = 0;
    u32 it = 0;
    for (;; ++it) {
        const int s = (int)(it % ENC_STAGES);
        const u32 e = it % ENC_DEPTH;
        const u64 tile = tickets[s];
        if (tile >= p.n_tiles) break;
        const u64 tif = tickets[ENC_STAGES + s];

        mbar_wait(&bar_full[s], (it / ENC_STAGES) & 1, p.status, 7);

        // ---- this thread's unit -> registers
        u32 w[P::UW];
        const unsigned char* tile_sm = stages + s * G::STAGE_BYTES + P::UNIT_BYTES;
        {
            const uint4* src = (const uint4*)(tile_sm + (size_t)t * P::UNIT_BYTES);
#pragma unroll
            for (int j = 0; j < P::UW / 4; ++j) {
                uint4 v = src[j];
                w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
            }
        }
        u32 nvalid = P::VPU;
        if (tif + 1 == p.tiles_per_frame) {                // the frame's last tile may be ragged
            const u64 tile_vals = p.n_values - tif * (u64)(G::TILE_BLOCKS * 12);
            const u64 my_first = (u64)t * P::VPU;
            nvalid = my_first >= tile_vals ? 0u : (tile_vals - my_first > (u64)P::VPU ? (u32)P::VPU : (u32)(tile_vals - my_first));
            if (nvalid < (u32)P::VPU) {                    // frame tail: wipe what is not ours
                const u32 vbytes = nvalid * P::SZ;
#pragma unroll
                for (int j = 0; j < P::UW; ++j) {
                    const u32 lo = 4u * j;
                    if (vbytes <= lo) w[j] = 0;
                    else if (vbytes < lo + 4) w[j] &= (1u << (8 * (vbytes - lo))) - 1;
                }
            }
        }

        // ---- K1: widths of my blocks
        u32 sb[P::BPU], cnt[P::BPU];
#pragma unroll
        for (int b = 0; b < P::BPU; ++b) {
            sb[b] = block_width12<T>(&w[b * P::BW]);
            const u32 v0 = 12u * b;
            cnt[b] = nvalid <= v0 ? 0u : (nvalid - v0 > 12u ? 12u : nvalid - v0);
            my_max = sb[b] > my_max ? sb[b] : my_max;
        }
        u32 prev0 = 0;                                     // width of the block before the tile
        if (t == 0 && tif > 0) {
            u32 h[P::BW];
            const u32* hs = (const u32*)(tile_sm - 4 * P::BW);
#pragma unroll
            for (int j = 0; j < P::BW; ++j) h[j] = hs[j];
            prev0 = block_width12<T>(h);
        }
        if (lane == 31) sm_warp_last[warp] = sb[P::BPU - 1];
        bar_sync(1, NT);                                   // A: stage `s` is free, warp_last visible
        if (t == 0) issue(s);

        // ---- K2: headers and lengths
        u32 prev = shfl_up(sb[P::BPU - 1], 1);
        if (lane == 0) prev = warp > 0 ? sm_warp_last[warp - 1] : prev0;
        u32 hv[P::BPU], hl[P::BPU], len = 0;
#pragma unroll
        for (int b = 0; b < P::BPU; ++b) {
            hv[b] = 0; hl[b] = 0;
            if (cnt[b]) {
                block_header_fast(sb[b], prev, hv[b], hl[b]);
                len += hl[b] + sb[b] * cnt[b];
                prev = sb[b];
            }
        }

        // ---- K3: offsets inside the tile
        u32 off, tile_bits;
        scan_lengths<NT>(len, sm_warp_tot, off, tile_bits, 1);   // bar B inside
        // publish the tile's bit count at once: other tiles' look-backs must never wait for our resolver
        if (t == 0) st_relaxed(&p.tdesc[tile], TD_VALID | (u64)tile_bits);

        // ---- room in the ring: physically contiguous, one zero word in front (window_word reads stg[-1])
        // words -1 .. (bits/32)+2: a frame end may push the window one word past the zero pad (tail_handoff)
        const u32 need = ((tile_bits >> 5) + 1 + 3 + 3) & ~3u;
        u32 vbase = vhead;
        if ((vbase & ring_mask) + need > ring_words) vbase += ring_words - (vbase & ring_mask);
        bool stored = false;
        while (it - oldest == (u32)ENC_DEPTH || (it > oldest && vbase + need - vtail > ring_words)) {
            store_pending(oldest);                         // blocks only if that resolution is really late
            ++oldest;
            vtail = oldest < it ? vbases[oldest % ENC_DEPTH] : vbase;   // virtual base of the next pending tile
            stored = true;
        }
        if (stored) bar_sync(1, NT);                       // E: freed ring words and mailbox entries are reusable
        u32* stg = ring + (vbase & ring_mask) + 1;
        if (warp == 0) {
            if (lane == 0) {
                mail64[e * 4] = tile;
                vbases[e] = vbase;
                mail32[e * 8 + 4] = tile_bits;
                mail32[e * 8 + 7] = (vbase & ring_mask) + 1;
                stg[-1] = 0;
            }
            sync_warp();
            bar_arrive(2 + e, 64);                         // the resolver may start its look-back
        }
        vhead = vbase + need;
        if (it == oldest) vtail = vbase;
        zero_boundary_words<NT>(stg, off, tile_bits);
        bar_sync(1, NT);                                   // C

        // ---- K4: pack into tile-relative staging
        BitSink sk;
        sk.init(stg, off);
#pragma unroll
        for (int b = 0; b < P::BPU; ++b)
            if (cnt[b]) {
                sk.put(hv[b], hl[b]);
                pack_block12<T>(sk, &w[b * P::BW], sb[b], cnt[b]);
            }
        merge_and_flush(sk, off + len);
        bar_arrive(2 + ENC_DEPTH + e, NT + 32);            // D: all NT workers arrive -> staging complete
        // Shared scratch reused by the next iteration is rewritten only after one of its barriers
        // A..C, which no worker passes before all have finished reading this iteration's values.
    }
    // drain: the tiles still waiting in the ring, oldest first
    for (; oldest < it; ++oldest) store_pending(oldest);
    bar_sync(1, NT);
    // tell the resolvers that there is nothing more: the next iteration each of them would serve
    if (warp == 0) {
        for (u32 k = 0; k < (u32)ENC_RESOLVERS; ++k) {
            if (lane == 0) mail64[((it + k) % ENC_DEPTH) * 4] = TILE_END;
            sync_warp();
            bar_arrive(2 + (it + k) % ENC_DEPTH, 64);
        }
    }
    my_max = warp_max(my_max);
    if (lane == 0 && my_max) atomic_max(sm_max, my_max);
    bar_sync(1, NT);
    if (t == 0 && *sm_max) atomic_max(p.prolix_bits, *sm_max);
}

// ------------------------------------------------------------------ generic kernel: any block size / alignment
// One thread per block, values read straight from global memory (two passes over the block's
// values, the second one hits L1/L2).  Same scan, look-back, staging and store as the fast kernel.
template <typename T, int NT>
struct GenGeom {
    static constexpr int STG_WORDS_MAX = (227 * 1024 - SM_HEADER) / 4 - 16;
};

template <typename T>
TRPX_DEVICE u64 load_value(const T* px, u64 i)
{
    return Pix<T>::SGN ? (u64)(i64)px[i] : (u64)px[i];
}

template <typename T, int NT>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(NT, 1) terse_encode_generic_kernel(EncParams p, u32 tile_blocks)
{
    typedef Pix<T> P;
    TRPX_DYN_SMEM(sm);
    u32* tickets = (u32*)(sm + SM_TICKETS);
    u32* sm_warp_tot = (u32*)(sm + SM_WARP_TOT);
    u32* sm_warp_last = (u32*)(sm + SM_WARP_LAST);
    u64* bc = (u64*)(sm + SM_BCAST);
    u32* sm_max = (u32*)(sm + SM_MAX);
    u32* stg = (u32*)(sm + SM_HEADER) + 4;                 // stg[-1] is a zero word (window_word)
    const u32 t = tid(), lane = t & 31, warp = t >> 5;
    const u64 tmask = P::W == 64 ? ~0ull : ((1ull << P::W) - 1);
    if (t == 0) { *sm_max = 0; stg[-1] = 0; }
    u32 my_max = 0;
    for (;;) {
        sync_block();                                      // previous tile fully stored; tickets reusable
        if (t == 0) tickets[0] = atomic_add(p.ticket, 1u);
        sync_block();
        const u64 tile = tickets[0];
        if (tile >= p.n_tiles) break;
        const u64 f = tile / p.tiles_per_frame, tif = tile % p.tiles_per_frame;
        const bool ends = tif + 1 == p.tiles_per_frame;
        const T* px = (const T*)p.pixels + f * p.n_values;
        const u64 blk = tif * tile_blocks + t;             // my block inside the frame
        const bool have = t < tile_blocks && blk < p.nblocks;
        const u64 from = blk * p.block;
        u32 cnt = 0;
        if (have) cnt = (u32)(p.n_values - from < (u64)p.block ? p.n_values - from : (u64)p.block);

        auto width_of = [&](u64 v0, u32 n) -> u32 {         // Terse.hpp:508-515, :551-560
            u64 m = 0;
            for (u32 i = 0; i < n; ++i) {
                u64 v = load_value<T>(px, v0 + i);
                m |= (P::SGN ? abs64(v) : v) & tmask;
            }
            u32 s = 64 - (u32)clz64(m);
            return (P::SGN && m) ? s + 1 : s;
        };
        u32 s = have ? width_of(from, cnt) : 0;
        my_max = s > my_max ? s : my_max;
        u32 prev0 = 0;
        if (t == 0 && tif > 0) prev0 = width_of(from - p.block, p.block);
        if (lane == 31) sm_warp_last[warp] = s;
        sync_block();
        u32 prev = shfl_up(s, 1);
        if (lane == 0) prev = warp > 0 ? sm_warp_last[warp - 1] : prev0;
        u32 hv = 0, hl = 0, len = 0;
        if (have) { block_header(s, prev, hv, hl); len = hl + s * cnt; }
        u32 off, tile_bits;
        scan_lengths<NT>(len, sm_warp_tot, off, tile_bits);
        zero_boundary_words<NT>(stg, off, tile_bits);
        sync_block();
        BitSink sk;
        sk.init(stg, off);
        if (have) {
            sk.put(hv, hl);
            if (s)
                for (u32 i = 0; i < cnt; ++i) sk.put_wide(load_value<T>(px, from + i), s);
        }
        merge_and_flush(sk, off + len);
        sync_block();
        resolve_and_store<NT>(p, tile, tile_bits, stg, bc);
    }
    my_max = warp_max(my_max);
    if (lane == 0 && my_max) atomic_max(sm_max, my_max);
    sync_block();
    if (t == 0 && *sm_max) atomic_max(p.prolix_bits, *sm_max);
}

}  // namespace trpx
