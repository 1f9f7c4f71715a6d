// prolix_decode.cuh -- PROLIX decoder for sm_100a.
//
// Replaces Terse::prolix(Iterator) (reference include/Terse.hpp:352-389), the header decode
// (:361-372), Bit_range::get_range (include/Bit_pointer.hpp:742-792) and f_find_terse_frame
// (Terse.hpp:562-585).
//
// The reference decodes serially because block b+1's header sits wherever block b ended.  Here the
// chain is broken in two stages:
//
//   P1  header-state resolution.  Every frame is cut into fixed-size byte segments.  One thread per
//       segment first walks the headers of the PREVIOUS segment speculatively ("warm-up": a walk
//       started at a wrong bit self-synchronises with the true chain, because an explicit header
//       resets the width and runs of '1' headers sit at a constant stride), then walks its own
//       segment for real, recording entry state, exit state and block count.  A cooperative kernel
//       then checks entry(j) == exit(j-1) for every segment and re-walks the ones that disagree until
//       a whole sweep changes nothing -- at that point every entry state is EXACT by induction from
//       the frame start, however adversarial the stream.  An exclusive scan of the block counts gives
//       each segment its first block index; a last walk writes one width byte per block plus the bit
//       position of every P2 tile.
//   P2  unpack.  With widths known, block offsets inside a tile are a prefix sum (the mirror of the
//       encoder's K3); each thread extracts whole blocks (uniform width), values are staged in
//       shared memory and leave with one TMA bulk store per tile.
#pragma once

#include "simt.cuh"

namespace trpx {

struct DecParams {
    const u32* payload;     // 16-byte aligned; readable up to the next multiple of 8 bytes
    u64 payload_bytes;
    u32 block;              // values per block
    u64 n_values;           // values per frame
    u64 n_frames;
    u64 nblocks;            // blocks per frame
    u32 last_cnt;           // values in a frame's last block
    u32 is_signed;
    const u64* frame_ends;  // [n_frames] end byte offsets
    // P1 segment tables
    u32 seg_bytes;          // segment size
    u32 warm_bytes;         // speculative warm-up length before the segment (clipped at the frame start)
    u64 max_segs;
    u64* seg_base;          // [n_frames + 1] first segment of each frame
    u32* seg_frame;         // [max_segs]
    u64* seg_entry;         // [max_segs] (pos << 8) | width, frame-relative bit of the first header
    u64* seg_exit;          // [max_segs]
    u32* seg_count;         // [max_segs] headers that start inside the segment
    u64* seg_b0;            // [max_segs] index of the segment's first block inside its frame
    u32* changed;           // [3] rotating sweep flags of the fix-up loop (zeroed)
    // P1 -> fused P2: exact checkpoints, one per SUB_BYTES of every segment (nullptr on the generic path)
    u64* ckpt;              // [max_segs * subs_per_seg]: first header at or after the sub-segment's first bit
    u32 subs_per_seg;
    u32 sub_shift;          // log2 of a sub-segment's size in BITS: 8 (32 bytes) for dense streams ... 5 (4 bytes) for sparse ones
    unsigned short* hdr_tab; // [4096] header look-up table in global memory (written by prolix_segments_kernel)
    u64* segd;              // [max_segs * 4] per segment: absolute bit of its start, absolute bit of its frame's end,
                            // index of its first block in the frame, frame | header count << 32 (written by the resolve kernel)
    // P1 -> P2
    unsigned char* widths;  // [n_frames * nblocks], zeroed
    u64* anchors;           // [n_frames * tiles_per_frame] frame-relative bit of each tile's first header
    u32 tile_blocks;
    u64 tiles_per_frame;
    void* out;
    u32* status;            // [1]
    u32 chain_mode;         // frame recovery: the payload is walked as ONE run of blocks (no frame structure known yet)
};

constexpr u32 DEC_MALFORMED = 4;   // == TRPX_ERR_MALFORMED

// ------------------------------------------------------------------ bit access
// >= 33 valid bits of the stream starting at absolute bit `abit`
TRPX_DEVICE u64 peek_bits(const u32* payload, u64 n_words, u64 abit)
{
    const u64 wi = abit >> 5;
    const u32 lo = wi < n_words ? payload[wi] : 0u;
    const u32 hi = wi + 1 < n_words ? payload[wi + 1] : 0u;
    return (((u64)hi << 32) | lo) >> (abit & 31);
}

// Header at the window's bit 0 (Terse.hpp:361-372): returns header length, updates s.
TRPX_DEVICE u32 decode_header(u64 win, u32& s)
{
    if (win & 1) return 1;
    s = (u32)(win >> 1) & 7;
    if (s != 7) return 4;
    s += (u32)(win >> 4) & 3;
    if (s != 10) return 6;
    s += (u32)(win >> 6) & 63;
    return 12;
}

// The same, without data-dependent branches: (length, width after the header) for carried width s.
TRPX_DEVICE void decode_header_bf(u32 win, u32 s, u32& hl, u32& s_new)
{
    const u32 s3 = (win >> 1) & 7, e2 = (win >> 4) & 3, e6 = (win >> 6) & 63;
    const bool p7 = s3 == 7, p10 = p7 && e2 == 3;
    const u32 sx = p10 ? 10 + e6 : s3 + (p7 ? e2 : 0u);
    const u32 hx = p10 ? 12u : (p7 ? 6u : 4u);
    const bool same = (win & 1) != 0;
    hl = same ? 1u : hx;
    s_new = same ? s : sx;
}

// Header look-up table in shared memory: 12 stream bits -> length | "same width" flag | new width.
constexpr u32 HDR_TAB_ENTRIES = 4096, HDR_TAB_BYTES = HDR_TAB_ENTRIES * 2;
constexpr u32 HDR_SAME = 0x80;
template <int NT>
TRPX_DEVICE void build_header_table(unsigned short* tab)      // caller syncs afterwards
{
    for (u32 i = tid(); i < HDR_TAB_ENTRIES; i += NT) {
        u32 hl, sx;
        decode_header_bf(i & ~1u, 0, hl, sx);                  // the explicit reading of these 12 bits
        tab[i] = (unsigned short)((i & 1) ? (1u | HDR_SAME) : (hl | (sx << 8)));
    }
}
template <int NT>
TRPX_DEVICE void copy_header_table(unsigned short* tab, const unsigned short* gtab)   // caller syncs afterwards
{
    const uint4* src = (const uint4*)gtab;
    uint4* dst = (uint4*)tab;
    for (u32 i = tid(); i < HDR_TAB_BYTES / 16; i += NT) dst[i] = src[i];
}
TRPX_DEVICE void lookup_header(const unsigned short* tab, u32 win, u32 s, u32& hl, u32& s_new)
{
    const u32 e = tab[win & (HDR_TAB_ENTRIES - 1)];
    hl = e & 15;
    s_new = (e & HDR_SAME) ? s : e >> 8;
}

TRPX_HD u64 pack_state(u64 pos, u32 s) { return (pos << 8) | (u64)s; }
TRPX_HD u64 state_pos(u64 st) { return st >> 8; }
TRPX_HD u32 state_s(u64 st) { return (u32)(st & 0xff); }

// A walker's view of the stream: two consecutive 64-bit words cached in registers, so that a header
// step usually needs no load at all and otherwise exactly one (the walk only moves forward).
struct StreamWindow {
    const u64* q;
    u64 n_q, qi, q0, q1;
    TRPX_DEVICE void init(const u32* payload, u64 n_words)
    {
        q = (const u64*)payload; n_q = (n_words + 1) >> 1; qi = ~0ull - 1; q0 = 0; q1 = 0;
    }
    TRPX_DEVICE u64 at(u64 i) const { return i < n_q ? q[i] : 0ull; }
    TRPX_DEVICE u64 peek(u64 abit)          // 64 valid bits starting at absolute bit `abit`
    {
        const u64 i = abit >> 6;
        const u32 sh = (u32)(abit & 63);
        if (i != qi) {
            q0 = (i == qi + 1) ? q1 : at(i);
            q1 = at(i + 1);
            qi = i;
        }
        return sh ? (q0 >> sh) | (q1 << (64 - sh)) : q0;
    }
};

// Checkpoint of a sub-segment (32 bytes of stream for diffraction-like data, down to 4 bytes for sparse data, so
// that a thread of the unpack kernel always owns about half a dozen blocks): the first header at or after its first bit, as
// (bit offset from the segment's start : 24 | width carried into that header : 8 | headers of the segment before it : 32).
constexpr u32 SUB_BYTES = 32, SUB_SHIFT_MAX = 8, SUB_SHIFT_MIN = 5;   // the largest sub-segment (32 bytes) sizes the shared-memory slice
TRPX_HD u64 pack_ckpt(u32 rel, u32 s, u32 n) { return ((u64)rel << 40) | ((u64)(s & 0xff) << 32) | (u64)n; }
TRPX_HD u32 ckpt_rel(u64 c) { return (u32)(c >> 40); }
TRPX_HD u32 ckpt_s(u64 c) { return (u32)(c >> 32) & 0xff; }
TRPX_HD u32 ckpt_n(u64 c) { return (u32)c; }

// Records the checkpoints of ONE segment while its headers are visited in order.  `rel` is the
// header's bit offset from the segment's start.
struct CkptSink {
    u64* row;               // this segment's checkpoints (nullptr: record nothing)
    u32 subs, next_m, sh;
    u32 next_rel;           // a header at rel >= next_rel opens a new sub-segment (0xffffffff: never)
    TRPX_DEVICE void init(u64* row_, u32 subs_, u32 sub_shift)
    {
        row = row_; subs = subs_; next_m = 0; sh = sub_shift;
        next_rel = row && subs ? 0u : 0xffffffffu;
    }
    TRPX_DEVICE void sync_rel() { next_rel = row && next_m < subs ? next_m << sh : 0xffffffffu; }   // after next_m was moved by hand
    TRPX_DEVICE void at(u32 rel, u32 s_prev, u32 n)          // a header starts at rel
    {
        if (rel < next_rel) return;                          // the common case: one compare
        const u32 m = rel >> sh;
        while (next_m <= m && next_m < subs) row[next_m++] = pack_ckpt(rel, s_prev, n);
        next_rel = next_m < subs ? next_m << sh : 0xffffffffu;
    }
    // The same for the walkers' hot loop, without a divergent branch in the common case (a header opens at
    // most ONE new sub-segment): a predicated 8-byte store and two selects.
    TRPX_DEVICE void at_fast(u32 rel, u32 s_prev, u32 n)
    {
        if (rel < next_rel) return;
        if ((rel >> sh) != next_m) { at(rel, s_prev, n); return; }   // skipped sub-segments: rare
        row[next_m] = pack_ckpt(rel, s_prev, n);
        ++next_m;
        next_rel = next_m < subs ? next_m << sh : 0xffffffffu;
    }
    TRPX_DEVICE void run(u32 rel, u32 n, u32 len)            // len one-bit headers (width 0) from rel
    {
        if (rel + len <= next_rel) return;
        at(rel, 0, n);
        while (next_m < subs && (next_m << sh) < rel + len) {   // boundaries inside the run are headers themselves
            const u32 r2 = next_m << sh;
            row[next_m++] = pack_ckpt(r2, 0, n + (r2 - rel));
        }
        next_rel = next_m < subs ? next_m << sh : 0xffffffffu;
    }
    TRPX_DEVICE void finish(u32 rel_exit, u32 s_exit, u32 n)  // sub-segments in which no header starts any more
    {
        if (!row) return;
        while (next_m < subs) row[next_m++] = pack_ckpt(rel_exit, s_exit, n);
        next_rel = 0xffffffffu;
    }
};

// Walk block headers from (pos, s) while pos < stop; all positions are bits relative to the frame,
// base_bit is the frame's first bit in the payload.  Returns the number of headers visited.
// Runs of '1' headers with width 0 (all-zero blocks, 1 bit each) are skipped a word at a time.
struct NoSink {
    TRPX_DEVICE void block(u64, u64, u32) {}
    TRPX_DEVICE void zeros(u64, u64, u64) {}
};
template <class Sink>
TRPX_DEVICE u64 walk_headers(const u32* payload, u64 n_words, u64 base_bit, u32 block, u64& pos, u32& s,
                             u64 stop, Sink& sink, CkptSink* ck = nullptr, u64 seg_r0 = 0)
{
    u64 n = 0;
    StreamWindow sw;
    sw.init(payload, n_words);
    while (pos < stop) {
        const u64 win = sw.peek(base_bit + pos);
        if (s == 0 && (win & 1)) {
            u64 run = (u64)ffs64(~win | (1ull << 63)) - 1;       // consecutive '1' headers (<= 63)
            if (run > stop - pos) run = stop - pos;
            sink.zeros(n, pos, run);
            if (ck) ck->run((u32)(pos - seg_r0), (u32)n, (u32)run);
            n += run;
            pos += run;
            continue;
        }
        if (ck) ck->at((u32)(pos - seg_r0), s, (u32)n);
        const u32 hl = decode_header(win, s);
        sink.block(n, pos, s);
        pos += hl + (u64)s * block;
        ++n;
    }
    if (ck) ck->finish((u32)(pos - seg_r0), s, (u32)n);
    return n;
}

// ------------------------------------------------------------------ warp-cooperative walkers
// The serial walk above pays one dependent global-memory round trip per 64 bits of stream.  The
// walkers below are the production path: 32 independent walks per warp (one per lane), the stream
// staged through shared memory in rounds.  Each lane fetches ITS OWN next 128 bytes with eight 16-byte
// loads (all in flight while the current round is walked: the prefetch lives in registers), then
// drops them into a [word][lane] table (pitch 32: a lane's words all live in bank `lane`), so that every
// LDS / STS of a walk step is bank-conflict free however far apart the lanes' positions are.  A round
// advances 28 words; the 4-word overlap lets a header that starts in word 30 still see its 12 bits.
constexpr u32 WALK_PITCH = 32;
constexpr u32 WALK_ROUND_WORDS = 32, WALK_ROUND_STRIDE = 28;
constexpr u32 WALK_BUF_WORDS = WALK_PITCH * WALK_ROUND_WORDS;

struct WalkLane {
    bool have;          // this lane walks something
    u64 cb;             // 16-byte aligned byte offset (in the payload) of the lane's word 0
    u32 q;              // position, bits relative to cb
    u32 s;              // width carried from the previous block
    u32 qA, qB;         // the first header at or after qA is the segment's entry; stop at the first one >= qB
    bool entered;
    u32 q_entry, s_entry;
    u32 n;              // headers counted since the entry (a segment holds < 2^23 bits)
};

TRPX_DEVICE void walk_fetch(const DecParams& p, const WalkLane& L, u32 round, uint4 (&pre)[8])
{
    const u64 byte0 = L.cb + (u64)round * (WALK_ROUND_STRIDE * 4);
    const bool want = L.have && (round + 1) * (WALK_ROUND_STRIDE * 32) + 128 > L.q;   // (rounds that end before the lane's position: nothing to read)
    const u64 safe_end = p.payload_bytes & ~15ull;          // 16-byte loads stay inside the payload buffer
    const unsigned char* base = (const unsigned char*)p.payload;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const u64 b = byte0 + 16u * k;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (want && b + 16 <= safe_end) {
            v = *(const uint4*)(base + b);
        } else if (want && b < p.payload_bytes) {            // ragged end: word by word, zero beyond
            const u64 n_words = (p.payload_bytes + 3) >> 2;
            const u64 wi = b >> 2;
            v.x = wi < n_words ? p.payload[wi] : 0u;
            v.y = wi + 1 < n_words ? p.payload[wi + 1] : 0u;
            v.z = wi + 2 < n_words ? p.payload[wi + 2] : 0u;
            v.w = wi + 3 < n_words ? p.payload[wi + 3] : 0u;
        }
        pre[k] = v;
    }
}

TRPX_DEVICE void walk_stage(u32* buf, const uint4 (&pre)[8])
{
    const u32 lane = tid() & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        buf[(4 * k + 0) * WALK_PITCH + lane] = pre[k].x;
        buf[(4 * k + 1) * WALK_PITCH + lane] = pre[k].y;
        buf[(4 * k + 2) * WALK_PITCH + lane] = pre[k].z;
        buf[(4 * k + 3) * WALK_PITCH + lane] = pre[k].w;
    }
}

// All 32 lanes of a warp call this together.  sink.block(k, q, s) / sink.zeros(k, q, run) see the k-th
// header after the entry at lane-relative bit q.  The common step is short and branch-light: two LDS
// + funnel shift for the window, one table look-up for the header, one compare for "anything
// special here?" (entering the segment, a new checkpoint, a run of empty blocks).
template <class Sink>
TRPX_DEVICE void warp_walk(const DecParams& p, u32* buf, const unsigned short* tab, WalkLane& L, Sink& sink, CkptSink& ck)
{
    const u32 lane = tid() & 31;
    if (!L.have) { L.qB = 0; L.q = 0xffffffffu; }
    const u32 my_rounds = L.have ? ((L.qB + 31) >> 5) / WALK_ROUND_STRIDE + 1 : 0u;
    const u32 first_round = L.have ? (L.q >> 5) / WALK_ROUND_STRIDE : 0xffffffffu;
    u32 r = warp_min_u32(first_round);
    const u32 r_end = warp_max(my_rounds);
    if (r >= r_end) return;
    // next position at which the slow path has work: the segment's first bit (afterwards only runs of empty blocks)
    u32 q_event = L.entered ? 0xffffffffu : L.qA;
    const u32 blk = p.block;
    uint4 pre[8];
    walk_fetch(p, L, r, pre);
    for (; r < r_end; ++r) {
        sync_warp();                                        // everybody is done reading the previous round
        walk_stage(buf, pre);
        sync_warp();
        if (r + 1 < r_end) walk_fetch(p, L, r + 1, pre);    // in flight while this round is walked
        const u32 w0 = r * WALK_ROUND_STRIDE;
        const u32 q_round = (w0 + 31) << 5;                 // headers below this bit are readable in this round
        const u32 qlim = L.qB < q_round ? L.qB : q_round;   // (lanes ahead of this round have q >= q_round)
        const u32* col = buf + lane - w0 * WALK_PITCH;
        for (;;) {
            if (!any_lane(L.q < qlim)) break;
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {             // two steps per vote
                if (L.q < qlim) {
                    const u32 wi = L.q >> 5;
                    const u32 win = funnel_r(col[wi * WALK_PITCH], col[wi * WALK_PITCH + WALK_PITCH], L.q & 31);
                    const bool isrun = L.s == 0 && (win & 1);
                    if (L.q >= q_event || isrun) {          // ---- slow path
                        if (!L.entered && L.q >= L.qA) { L.entered = true; L.q_entry = L.q; L.s_entry = L.s; L.n = 0; }
                        if (isrun) {                        // a run of '1' headers of empty blocks, 1 bit each
                            const u32 stop = L.entered ? L.qB : L.qA;
                            u32 run = (u32)ffs32(~win) - 1; // ffs32(0) == 0 -> 0xffffffff: all 32 bits set
                            run = run > 32u ? 32u : run;
                            run = run > stop - L.q ? stop - L.q : run;
                            if (L.entered) { sink.zeros(L.n, L.q, run); ck.run(L.q - L.qA, L.n, run); }
                            L.q += run;
                            L.n += run;
                        }
                        q_event = L.entered ? 0xffffffffu : L.qA;
                        if (isrun) continue;
                    }
                    u32 hl, s_new;
                    lookup_header(tab, win, L.s, hl, s_new);
                    if (L.entered) { ck.at_fast(L.q - L.qA, L.s, L.n); sink.block(L.n, L.q, s_new); }
                    L.q += hl + s_new * blk;
                    L.n += 1;
                    L.s = s_new;
                }
            }
        }
    }
}

// Header at the low bits of `win`, carried width s -> (header length, new width) for the walk loops: one look-up in
// the shared-memory table.  (Decoding the two common forms -- '1' and '0' + 3 bits -- with ALU operations instead,
// table only for the escape forms, measured exactly the same walk time: the table LDS is not what bounds a step.)
TRPX_DEVICE void walk_header(saddr_t tab_a, u32 win, u32 s, u32& hl, u32& s_new)
{
    const u32 e = lds_u16(tab_a + ((win & (HDR_TAB_ENTRIES - 1)) << 1));
    s_new = (e & HDR_SAME) ? s : e >> 8;
    hl = e & 15;
}

// The walk kernel's own loop: the steps of warp_walk with no sink, in two phases per round.  Phase A is the
// speculative warm-up (all lanes of a warp start it together and leave it within a round of each other): nothing
// but window -> table -> advance.  Phase B walks the segment proper: header count, and the checkpoint bookkeeping
// inlined (one compare per step against the stream position of the next sub-segment boundary).  Shared memory
// through window addresses; the block size folded in when it is 12.
template <bool B12>
TRPX_DEVICE void warp_walk_ckpt(const DecParams& p, u32* buf, const unsigned short* tab, WalkLane& L, CkptSink& ck)
{
    const u32 lane = tid() & 31;
    if (!L.have) { L.qA = 0; L.qB = 0; L.q = 0xffffffffu; }
    const u32 my_rounds = L.have ? ((L.qB + 31) >> 5) / WALK_ROUND_STRIDE + 1 : 0u;
    const u32 first_round = L.have ? (L.q >> 5) / WALK_ROUND_STRIDE : 0xffffffffu;
    u32 r = warp_min_u32(first_round);
    const u32 r_end = warp_max(my_rounds);
    if (r >= r_end) return;
    const u32 blk = B12 ? 12u : p.block;
    const saddr_t tab_a = saddr(tab);
    const saddr_t lane_a = saddr(buf) + lane * 4;
    const u32 sub = 1u << ck.sh;
    u32 q_ck = 0xffffffffu;                                 // position of the next sub-segment boundary; armed on entry
    uint4 pre[8];
    walk_fetch(p, L, r, pre);
    for (; r < r_end; ++r) {
        sync_warp();                                        // everybody is done reading the previous round
        walk_stage(buf, pre);
        sync_warp();
        if (r + 1 < r_end) walk_fetch(p, L, r + 1, pre);    // in flight while this round is walked
        const u32 w0 = r * WALK_ROUND_STRIDE;
        const u32 q_round = (w0 + 31) << 5;                 // headers below this bit are readable in this round
        const saddr_t col_a = lane_a - w0 * (WALK_PITCH * 4);
        // ---- phase A: warm-up, until the first header at or after the segment's first bit
        if (any_lane(!L.entered)) {
            const u32 limA = L.entered ? 0u : (L.qA < q_round ? L.qA : q_round);
            while (any_lane(L.q < limA)) {
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    if (L.q < limA) {
                        const saddr_t a = col_a + (L.q >> 5) * (WALK_PITCH * 4);
                        const u32 win = funnel_r(lds_u32(a), lds_u32_at<WALK_PITCH * 4>(a), L.q);
                        if (L.s == 0 && (win & 1)) {        // a run of one-bit headers of empty blocks: never past the segment's first bit
                            u32 run = (u32)ffs32(~win) - 1; // ffs32(0) == 0 -> 0xffffffff: all 32 bits set
                            run = run > 32u ? 32u : run;
                            run = run > L.qA - L.q ? L.qA - L.q : run;
                            L.q += run;
                        } else {
                            u32 hl;
                            walk_header(tab_a, win, L.s, hl, L.s);
                            L.q += hl + L.s * blk;
                        }
                    }
                }
            }
            if (!L.entered && L.q >= L.qA && L.have) {      // this header is the segment's entry
                L.entered = true; L.q_entry = L.q; L.s_entry = L.s; L.n = 0;
                ck.sync_rel();
                q_ck = ck.next_rel != 0xffffffffu ? L.qA + ck.next_rel : 0xffffffffu;
            }
        }
        // ---- phase B: the segment itself
        const u32 limB = !L.entered ? 0u : (L.qB < q_round ? L.qB : q_round);
        while (any_lane(L.q < limB)) {
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                if (L.q < limB) {
                    const saddr_t a = col_a + (L.q >> 5) * (WALK_PITCH * 4);
                    const u32 win = funnel_r(lds_u32(a), lds_u32_at<WALK_PITCH * 4>(a), L.q);
                    if (L.s == 0 && (win & 1)) {            // a run of one-bit headers of empty blocks
                        u32 run = (u32)ffs32(~win) - 1;
                        run = run > 32u ? 32u : run;
                        run = run > L.qB - L.q ? L.qB - L.q : run;
                        ck.sync_rel();
                        ck.run(L.q - L.qA, L.n, run);
                        q_ck = ck.next_rel != 0xffffffffu ? L.qA + ck.next_rel : 0xffffffffu;
                        L.q += run;
                        L.n += run;
                    } else {
                        u32 hl, s_new;
                        walk_header(tab_a, win, L.s, hl, s_new);
                        if (L.q >= q_ck) {                  // this header opens sub-segment next_m (and, rarely, more than one)
                            const u32 rel = L.q - L.qA;
                            do {
                                ck.row[ck.next_m] = pack_ckpt(rel, L.s, L.n);
                                ++ck.next_m;
                                q_ck = ck.next_m < ck.subs ? q_ck + sub : 0xffffffffu;
                            } while (L.q >= q_ck);
                        }
                        L.q += hl + s_new * blk;
                        L.n += 1;
                        L.s = s_new;
                    }
                }
            }
        }
    }
    if (L.entered) ck.sync_rel();
}

// ------------------------------------------------------------------ D0: segment table
// nseg(f) = ceil(bytes_f / seg_bytes); seg_base = exclusive scan; seg_frame[j] = f.  One CTA.
template <int NT>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(NT, 1) prolix_segments_kernel(DecParams p)
{
    TRPX_SHARED u64 sm_tot[NT / 32];
    const u32 t = tid(), lane = t & 31, warp = t >> 5;
    if (bid() == 0) build_header_table<NT>(p.hdr_tab);       // the walkers copy it into shared memory
    // thread t owns the frames [f0, f1): their segment counts, one block-wide scan of the per-thread sums, then
    // the table entries -- one round whatever the number of frames
    const u64 per = div_up(p.n_frames, (u64)NT);
    const u64 f0 = (u64)t * per < p.n_frames ? (u64)t * per : p.n_frames;
    const u64 f1 = f0 + per < p.n_frames ? f0 + per : p.n_frames;
    auto nseg_of = [&](u64 f) -> u64 {
        const u64 e = p.frame_ends[f], b = f ? p.frame_ends[f - 1] : 0;
        if (e <= b || e > p.payload_bytes) { atomic_max(p.status, DEC_MALFORMED); return 0; }
        return div_up(e - b, p.seg_bytes);
    };
    u64 mine = 0;
    for (u64 f = f0; f < f1; ++f) mine += nseg_of(f);
    u64 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 v = shfl_up(incl, d);
        if (lane >= (u32)d) incl += v;
    }
    if (lane == 31) sm_tot[warp] = incl;
    sync_block();
    u64 base = 0, total = 0;
    for (int i = 0; i < NT / 32; ++i) {
        const u64 v = sm_tot[i];
        if ((u32)i < warp) base += v;
        total += v;
    }
    // every CTA of the grid repeats the (cheap) scan; the table entries are shared out frame by frame.  A frame's
    // entries are written by the whole warp (the lanes serve each other's frames in turn): a stack of a few large
    // frames -- or the one "frame" of the G plan -- is thousands of entries behind a single thread otherwise.
    u64 run = base + incl - mine;
    const u64 cnt_mine = f1 > f0 ? f1 - f0 : 0;
    u32 cnt_max = (u32)(cnt_mine < 0xffffffffull ? cnt_mine : 0xffffffffull);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const u32 v = shfl_xor(cnt_max, d);
        cnt_max = v > cnt_max ? v : cnt_max;
    }
    for (u32 k = 0; k < cnt_max; ++k) {
        const bool have = k < cnt_mine;
        const u64 f = f0 + k;
        const u64 nseg = have ? nseg_of(f) : 0;
        const bool write = have && f % nblocks() == bid();
        if (write) p.seg_base[f] = run;
        u32 todo = ballot(write && nseg != 0);
        while (todo) {
            const int l = ffs32(todo) - 1;
            todo &= todo - 1;
            const u64 r = shfl(run, l), n = shfl(nseg, l);
            const u32 ff = shfl((u32)f, l);
            for (u64 i = lane; i < n && r + i < p.max_segs; i += 32) p.seg_frame[r + i] = ff;
        }
        run += nseg;
    }
    if (t == 0 && bid() == 0) p.seg_base[p.n_frames] = total > p.max_segs ? p.max_segs : total;
}

struct SegInfo { u64 frame, idx, base_bit, frame_bits, r0, r1; };
TRPX_DEVICE SegInfo seg_info(const DecParams& p, u64 j)
{
    SegInfo g;
    g.frame = p.seg_frame[j];
    g.idx = j - p.seg_base[g.frame];
    const u64 e = p.frame_ends[g.frame], b = g.frame ? p.frame_ends[g.frame - 1] : 0;
    g.base_bit = b * 8;
    g.frame_bits = (e - b) * 8;
    g.r0 = g.idx * (u64)p.seg_bytes * 8;
    g.r1 = g.r0 + (u64)p.seg_bytes * 8;
    if (g.r1 > g.frame_bits) g.r1 = g.frame_bits;
    return g;
}

// ------------------------------------------------------------------ D1: speculative + real walk
// One lane per segment.  The lane starts `warm_bytes` before its segment pretending a header sits
// right there (a wrong guess self-synchronises with the true chain: SURVEY 7.3; measured distance for
// 512^2 Poisson frames: median 0.3 KB, 99.9 % < 3 KB), records the state at the first header inside
// the segment (entry), the state at the first header past its end (exit) and the headers in between.
// (four 256-thread CTAs per SM = 64 registers per thread: the walk is a dependent chain per lane and is bound by how
// many chains are in flight, measured 1.17 -> 0.96 ms together with the shorter segments it allows)
#ifndef WALK_MIN_CTAS
#define WALK_MIN_CTAS 4
#endif
template <int NT>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(NT, WALK_MIN_CTAS) prolix_walk_kernel(DecParams p)
{
    TRPX_DYN_SMEM(sm);
    unsigned short* tab = (unsigned short*)sm;
    u32* buf = (u32*)(sm + HDR_TAB_BYTES) + (tid() >> 5) * WALK_BUF_WORDS;
    copy_header_table<NT>(tab, p.hdr_tab);
    sync_block();
    const u64 j = (u64)bid() * NT + tid();
    WalkLane L;
    L.have = j < p.seg_base[p.n_frames];
    L.cb = 0; L.q = 0; L.s = 0; L.qA = 0; L.qB = 0; L.entered = false; L.q_entry = 0; L.s_entry = 0; L.n = 0;
    u64 delta = 0;                                          // frame-relative bit = q + delta
    if (L.have) {
        const SegInfo g = seg_info(p, j);
        const u64 warm_bits = (u64)p.warm_bytes * 8;
        const u64 back = warm_bits < g.r0 ? warm_bits : g.r0;
        const u64 start_bit = g.base_bit + g.r0 - back;     // absolute; a multiple of 8
        // A frame's first segment has no warm-up (its start is exact).  Its lane still counts its rounds from
        // where a warm-up WOULD have started, so that all lanes of a warp reach their segments in the same round
        // (the walk loop has a warm-up phase and a segment phase per round; the lane idles through the first).
        const u64 seg_bit = g.base_bit + g.r0;
        const u64 origin_bit = seg_bit >= warm_bits ? seg_bit - warm_bits : 0;
        L.cb = (origin_bit >> 3) & ~15ull;
        L.q = (u32)(start_bit - L.cb * 8);
        delta = g.r0 - back - (u64)L.q;                     // (mod 2^64: the origin may lie before the frame)
        L.qA = (u32)(g.r0 - delta);
        L.qB = (u32)(g.r1 - delta);
    }
    CkptSink ck;
    ck.init(L.have && p.ckpt ? p.ckpt + j * p.subs_per_seg : nullptr, p.subs_per_seg, p.sub_shift);
    if (p.block == 12) warp_walk_ckpt<true>(p, buf, tab, L, ck);
    else warp_walk_ckpt<false>(p, buf, tab, L, ck);
    if (L.have) {
        if (!L.entered) { L.q_entry = L.q; L.s_entry = L.s; L.n = 0; }   // the warm-up jumped over the whole segment
        ck.finish(L.q - L.qA, L.s, (u32)L.n);
        p.seg_entry[j] = pack_state(L.q_entry + delta, L.s_entry);
        p.seg_exit[j] = pack_state(L.q + delta, L.s);
        p.seg_count[j] = L.n;
    }
}

// A warp's window on the stream, staged in shared memory with coalesced loads: dependent header steps then cost a
// shared-memory access each instead of a global-memory round trip (the fix-up below and the frame chain use it; all
// lanes of the warp execute those walks in lock step, with the same values).
constexpr u32 FC_CHUNK_WORDS = 1024;                       // stream staged per refill: 4 KB
struct ChainWin {                                          // one warp's window on the stream
    u32* chunk;
    u64 chunk_bit;                                         // absolute bit of chunk[0] (multiple of 128); ~0: nothing staged
    TRPX_DEVICE void need(const DecParams& p, u64 n_words, u64 abit, u32 span_bits)      // all lanes
    {
        if (chunk_bit != ~0ull && abit >= chunk_bit && abit - chunk_bit + span_bits + 64 <= (u64)FC_CHUNK_WORDS * 32) return;
        sync_warp();
        chunk_bit = abit & ~127ull;
        const u64 w0 = chunk_bit >> 5;
        for (u32 i = tid() & 31; i < FC_CHUNK_WORDS + 4; i += 32) chunk[i] = w0 + i < n_words ? p.payload[w0 + i] : 0u;
        sync_warp();
    }
    TRPX_DEVICE u32 peek(u64 abit) const                    // >= 32 - 0 valid bits from abit (inside the staged window)
    {
        const u32 q = (u32)(abit - chunk_bit);
        return funnel_r(chunk[q >> 5], chunk[(q >> 5) + 1], q & 31);
    }
};

// Fix-up of ONE segment whose speculative entry was wrong: re-walk it from the exact state (pos, s), but only
// until the new trajectory meets the one the walker recorded -- a checkpoint that agrees in (position, carried
// width) proves that both are identical from there on.  A walker that started wrong has locked onto the true
// chain long before its segment ends, so this costs a few hundred bytes instead of the whole segment: the later
// checkpoints keep their positions and only get a constant added to their header counts.  Returns false when the
// trajectories never met (then everything was rewritten and the exit state (pos, s) is new).
TRPX_DEVICE bool rewalk_until_merged(const DecParams& p, u64 n_words, const SegInfo& g, u64 j, u64& pos, u32& s, u32& count)
{
    u64* row = p.ckpt + j * p.subs_per_seg;
    const u32 subs = p.subs_per_seg, sh = p.sub_shift;
    StreamWindow sw;
    sw.init(p.payload, n_words);
    u32 next_m = 0, n = 0;
    while (pos < g.r1) {
        const u64 win = sw.peek(g.base_bit + pos);
        const u32 rel = (u32)(pos - g.r0);
        for (const u32 m = rel >> sh; next_m <= m && next_m < subs; ++next_m) {   // this header opens sub-segments next_m .. m
            const u64 old = row[next_m];
            if (ckpt_rel(old) == rel && ckpt_s(old) == (s & 0xff)) {
                const u32 delta = n - ckpt_n(old);
                if (delta)
                    for (u32 mm = next_m; mm < subs; ++mm) {
                        const u64 c = row[mm];
                        row[mm] = (c & ~0xffffffffull) | (u64)(u32)((u32)c + delta);
                    }
                count = p.seg_count[j] + delta;
                return true;
            }
            row[next_m] = pack_ckpt(rel, s, n);
        }
        if (s == 0 && (win & 1)) {                            // one-bit headers of empty blocks
            u64 run = (u64)ffs64(~win | (1ull << 63)) - 1;
            if (run > g.r1 - pos) run = g.r1 - pos;
            const u64 bound = g.r0 + ((u64)next_m << sh);     // the next boundary is a header of this run: stop on it
            if (next_m < subs && pos + run > bound) run = bound - pos;
            n += (u32)run;
            pos += run;
            continue;
        }
        const u32 hl = decode_header(win, s);
        pos += hl + (u64)s * p.block;
        ++n;
    }
    for (; next_m < subs; ++next_m) row[next_m] = pack_ckpt((u32)(pos - g.r0), s, n);
    count = n;
    return false;
}

// The same re-walk, executed by a WHOLE WARP in lock step (every lane computes the same values; lane 0 stores): the
// stream comes from the warp's shared-memory window, so a step costs ~a shared-memory access instead of an L2 round
// trip.  This is what makes short warm-ups affordable: a walker that arrived wrong is put right in microseconds.
constexpr u32 FIX_ROW_CACHE = 64;                         // checkpoints of the segment staged per refill (512 bytes)
TRPX_DEVICE bool rewalk_until_merged_warp(const DecParams& p, ChainWin& win, u64* row_cache, u64 n_words, const SegInfo& g, u64 j, u64& pos, u32& s, u32& count)
{
    const u32 lane = tid() & 31;
    u64* row = p.ckpt + j * p.subs_per_seg;
    u32 cache_base = 0xffffffffu;                            // first entry held in row_cache (a multiple of FIX_ROW_CACHE)
    // the recorded checkpoint of sub-segment m, through the warp's cache: one coalesced refill per 64 sub-segments
    // instead of a global-memory round trip every few header steps
    auto recorded = [&](u32 m) -> u64 {
        const u32 b = m & ~(FIX_ROW_CACHE - 1);
        if (b != cache_base) {
            sync_warp();
            for (u32 i = lane; i < FIX_ROW_CACHE; i += 32) row_cache[i] = b + i < p.subs_per_seg ? row[b + i] : 0ull;
            cache_base = b;
            sync_warp();
        }
        return row_cache[m - b];
    };
    const u32 subs = p.subs_per_seg, sh = p.sub_shift;
    const u32 span = 12 + p.block * 73 + 64;
    const bool windowed = (u64)span + 64 <= (u64)FC_CHUNK_WORDS * 32;       // (huge blocks: straight from global memory)
    u32 next_m = 0, n = 0;
    while (pos < g.r1) {
        u64 win64;
        if (windowed) {
            win.need(p, n_words, g.base_bit + pos, span);
            win64 = (u64)win.peek(g.base_bit + pos) | ((u64)win.peek(g.base_bit + pos + 32) << 32);
        } else {
            win64 = peek_bits(p.payload, n_words, g.base_bit + pos);
        }
        const u32 rel = (u32)(pos - g.r0);
        for (const u32 m = rel >> sh; next_m <= m && next_m < subs; ++next_m) {   // this header opens sub-segments next_m .. m
            const u64 old = recorded(next_m);
            sync_warp();                                      // every lane has read the old entry: it is rewritten below, either way
            if (ckpt_rel(old) == rel && ckpt_s(old) == (s & 0xff)) {
                const u32 delta = n - ckpt_n(old);
                if (delta)
                    for (u32 mm = next_m + lane; mm < subs; mm += 32) {          // (the lanes share this sweep of the row)
                        const u64 c = row[mm];
                        row[mm] = (c & ~0xffffffffull) | (u64)(u32)((u32)c + delta);
                    }
                count = p.seg_count[j] + delta;
                sync_warp();
                return true;
            }
            if (lane == 0) row[next_m] = pack_ckpt(rel, s, n);
        }
        if (s == 0 && (win64 & 1)) {                          // one-bit headers of empty blocks
            u64 run = (u64)ffs64(~win64 | (1ull << 63)) - 1;
            if (run > g.r1 - pos) run = g.r1 - pos;
            const u64 bound = g.r0 + ((u64)next_m << sh);     // the next boundary is a header of this run: stop on it
            if (next_m < subs && pos + run > bound) run = bound - pos;
            n += (u32)run;
            pos += run;
            continue;
        }
        const u32 hl = decode_header(win64, s);
        pos += hl + (u64)s * p.block;
        ++n;
    }
    if (lane == 0)
        for (; next_m < subs; ++next_m) row[next_m] = pack_ckpt((u32)(pos - g.r0), s, n);
    sync_warp();
    count = n;
    return false;
}

// ------------------------------------------------------------------ D2: verify / fix until stable, D3: scan
// Cooperative launch (whole grid resident): grid_sync() is cooperative_groups' grid barrier on the
// device and a block barrier in the single-CTA emulator run.
template <int NT>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(NT, 1) prolix_resolve_kernel(DecParams p)
{
    TRPX_SHARED u64 sm_tot[NT / 32];
    TRPX_SHARED u64 sm_run;
    TRPX_SHARED u32 sm_win[NT / 32][FC_CHUNK_WORDS + 4];      // one stream window per warp (the fix-up walks)
    TRPX_SHARED u64 sm_row[NT / 32][FIX_ROW_CACHE];           // ... and its window on the segment's recorded checkpoints
    const u32 t = tid(), lane = t & 31, warp = t >> 5;
    const u64 n_segs = p.seg_base[p.n_frames];
    const u64 n_words = (p.payload_bytes + 3) >> 2;
    const u64 gthreads = (u64)nblocks() * NT, gtid = (u64)bid() * NT + t;
    const u64 gwarps = gthreads / 32, gwarp = gtid / 32;
    u32 per_warp = 32;
    if (n_segs < gthreads) {
        per_warp = (u32)((n_segs + gwarps - 1) / gwarps);
        if (per_warp < 1) per_warp = 1;
    }
    NoSink ns;
    ChainWin win;
    win.chunk = sm_win[warp];
    // ---- fix-up sweeps: entry(j) must equal exit(j-1) (same position; same width unless the header
    // at that position is explicit).  Exact on exit: entry(0) is the frame start, and a sweep without
    // a single rewrite means every segment was walked from its predecessor's recorded exit.
    for (u32 sweep = 0;; ++sweep) {
        // three rotating flags: sweep k raises flag k%3 and clears flag (k+1)%3, which nobody reads
        // or raises until every CTA has passed this sweep's grid barrier
        u32* flag = &p.changed[sweep % 3];
        if (gtid == 0) p.changed[(sweep + 1) % 3] = 0;
        // a warp checks `per_warp` consecutive segments at a time (one per lane: 32, fewer when the call has fewer
        // segments than the grid has lanes) and then puts the wrong ones right ONE BY ONE, all lanes together
        // (rewalk_until_merged_warp) -- so a small call spreads its fix-ups over as many warps as it can
        for (u64 base = gwarp * per_warp; base < n_segs; base += gwarps * per_warp) {
            const u64 jl = base + lane;
            bool fix = false;
            u64 want = 0;
            if (lane < per_warp && jl < n_segs) {
                const SegInfo gl = seg_info(p, jl);
                if (gl.idx != 0) {
                    want = ld_relaxed(&p.seg_exit[jl - 1]);
                    const u64 have = p.seg_entry[jl];
                    bool same = want == have;
                    if (!same && state_pos(want) == state_pos(have))
                        same = (peek_bits(p.payload, n_words, gl.base_bit + state_pos(want)) & 1) == 0;
                    fix = !same;
                }
            }
            u32 todo = ballot(fix);
            while (todo) {
                const int l = ffs32(todo) - 1;
                todo &= todo - 1;
                const u64 j = base + (u64)l;
                const u64 w = shfl(want, l);
                const SegInfo g = seg_info(p, j);
                u64 pos = state_pos(w);
                u32 s = state_s(w);
                win.chunk_bit = ~0ull;
                sync_warp();
                if (lane == 0) p.seg_entry[j] = w;
                if (p.ckpt) {                               // fast path: stop as soon as the recorded trajectory is met
                    u32 count;
                    const bool merged = rewalk_until_merged_warp(p, win, sm_row[warp], n_words, g, j, pos, s, count);
                    if (lane == 0) {
                        p.seg_count[j] = count;
                        if (!merged) {                      // (merged: same exit as before, nothing downstream changes)
                            st_relaxed(&p.seg_exit[j], pack_state(pos, s));
                            atomic_or(flag, 1u);
                        }
                    }
                } else if (lane == 0) {
                    const u64 n = walk_headers(p.payload, n_words, g.base_bit, p.block, pos, s, g.r1, ns);
                    st_relaxed(&p.seg_exit[j], pack_state(pos, s));
                    p.seg_count[j] = n > 0xffffffffull ? 0xffffffffu : (u32)n;
                    atomic_or(flag, 1u);
                }
                sync_warp();
            }
        }
        grid_sync();
        const u32 any = ld_relaxed(flag);
        if (!any) break;
        if (sweep > (1u << 16)) {                           // (every CTA counts the same sweeps: all of them leave here together)
            if (gtid == 0) atomic_max(p.status, DEC_MALFORMED);
            break;
        }
    }
    // descriptors of table entries past the last segment say "nothing here" (the unpack grid covers max_segs)
    if (p.segd)
        for (u64 j = n_segs + gtid; j < p.max_segs; j += gthreads) {
            u64* d = p.segd + j * 4;
            d[0] = 0; d[1] = 0; d[2] = 0; d[3] = 0;
        }
    // ---- per-frame exclusive scan of the block counts (one CTA per frame, frames strided)
    for (u64 f = bid(); f < p.n_frames; f += nblocks()) {
        const u64 first = p.seg_base[f], nseg = p.seg_base[f + 1] - first;
        if (t == 0) sm_run = 0;
        sync_block();
        for (u64 i0 = 0; i0 < nseg; i0 += NT) {
            const u64 i = i0 + t;
            const u64 c = i < nseg ? p.seg_count[first + i] : 0;
            u64 incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 v = shfl_up(incl, d);
                if (lane >= (u32)d) incl += v;
            }
            if (lane == 31) sm_tot[warp] = incl;
            sync_block();
            u64 base = sm_run, total = 0;
            for (int k = 0; k < NT / 32; ++k) {
                u64 v = sm_tot[k];
                if ((u32)k < warp) base += v;
                total += v;
            }
            if (i < nseg) {
                p.seg_b0[first + i] = base + incl - c;
                if (p.segd) {
                    const SegInfo g = seg_info(p, first + i);
                    u64* d = p.segd + (first + i) * 4;
                    d[0] = g.base_bit + g.r0;
                    d[1] = g.base_bit + g.frame_bits;
                    d[2] = base + incl - c;
                    d[3] = (u64)(u32)f | (c << 32);
                }
            }
            sync_block();
            if (t == 0) sm_run += total;
            sync_block();
        }
        if (t == 0 && sm_run < p.nblocks && !p.chain_mode) atomic_max(p.status, DEC_MALFORMED);   // stream ends early
        sync_block();
    }
}

// ------------------------------------------------------------------ D4: emit widths + tile anchors
struct EmitSink {
    const DecParams* p;
    u64 frame, b0;          // frame index, first block index of the segment
    u32 pend, pend_mask;    // up to 4 widths waiting to leave as one 32-bit store
    u64 pend_word;          // word index of the pending group in the whole widths array
    u64 delta;              // frame-relative bit = walker position + delta
    u32 bad;
    TRPX_DEVICE void flush()
    {
        if (!pend_mask) return;
        if (pend_mask == 0xF) {
            ((u32*)p->widths)[pend_word] = pend;                  // group entirely ours
        } else {                                                  // shared with a neighbour: bytes
            for (int k = 0; k < 4; ++k)
                if (pend_mask >> k & 1) p->widths[pend_word * 4 + k] = (unsigned char)(pend >> (8 * k));
        }
        pend = 0; pend_mask = 0;
    }
    TRPX_DEVICE void note_tile(u64 b, u64 pos)
    {
        if (b % p->tile_blocks == 0) p->anchors[frame * p->tiles_per_frame + b / p->tile_blocks] = pos;
    }
    TRPX_DEVICE void block(u64 k, u64 q, u32 s)
    {
        const u64 pos = q + delta;
        const u64 b = b0 + k;
        if (b >= p->nblocks) return;                              // padding bits after the last block
        if (s > 65) bad = 1;
        note_tile(b, pos);
        if (s == 0) return;                                       // widths are pre-zeroed
        const u64 gi = frame * p->nblocks + b, wi = gi >> 2;
        if (pend_mask && wi != pend_word) flush();
        pend_word = wi;
        pend |= s << (8 * (u32)(gi & 3));
        pend_mask |= 1u << (gi & 3);
        if ((gi & 3) == 3) flush();
    }
    TRPX_DEVICE void zeros(u64 k, u64 q, u64 run)
    {
        // only tile anchors can fall inside a run of empty blocks
        const u64 pos = q + delta;
        u64 b = b0 + k;
        const u64 end = b + run < p->nblocks ? b + run : p->nblocks;
        u64 tb = div_up(b, p->tile_blocks) * p->tile_blocks;
        for (; tb < end; tb += p->tile_blocks)
            p->anchors[frame * p->tiles_per_frame + tb / p->tile_blocks] = pos + (tb - b);
    }
};

template <int NT>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(NT, 1) prolix_emit_kernel(DecParams p)
{
    TRPX_DYN_SMEM(sm);
    unsigned short* tab = (unsigned short*)sm;
    u32* buf = (u32*)(sm + HDR_TAB_BYTES) + (tid() >> 5) * WALK_BUF_WORDS;
    copy_header_table<NT>(tab, p.hdr_tab);
    sync_block();
    const u64 j = (u64)bid() * NT + tid();
    WalkLane L;
    L.have = j < p.seg_base[p.n_frames];
    L.cb = 0; L.q = 0; L.s = 0; L.qA = 0; L.qB = 0; L.entered = false; L.q_entry = 0; L.s_entry = 0; L.n = 0;
    EmitSink sink;
    sink.p = &p; sink.frame = 0; sink.b0 = 0; sink.pend = 0; sink.pend_mask = 0; sink.pend_word = 0; sink.bad = 0; sink.delta = 0;
    if (L.have) {
        const SegInfo g = seg_info(p, j);
        const u64 b0 = p.seg_b0[j];
        const u64 st = p.seg_entry[j];                      // exact after the resolve kernel
        if (b0 >= p.nblocks || state_pos(st) >= g.r1) {
            L.have = false;                                 // padding bits after the last block / no header starts here
        } else {
            const u64 start_bit = g.base_bit + state_pos(st);
            L.cb = (start_bit >> 3) & ~15ull;
            const u64 delta = state_pos(st) - (start_bit - L.cb * 8);
            L.q = (u32)(start_bit - L.cb * 8);
            L.s = state_s(st);
            L.qA = L.q;
            L.qB = (u32)(g.r1 - delta);
            sink.frame = g.frame; sink.b0 = b0; sink.delta = delta;
        }
    }
    CkptSink ck;
    ck.init(nullptr, 0, SUB_SHIFT_MAX);
    warp_walk(p, buf, tab, L, sink, ck);
    if (L.have) {
        sink.flush();
        if (sink.bad) atomic_max(p.status, DEC_MALFORMED);
    }
}

// ------------------------------------------------------------------ D5 (P2): unpack
struct BitSource {
    const u32* payload;
    u64 n_words, wi;
    u64 acc;
    u32 nb;
    TRPX_DEVICE void init(const u32* pl, u64 nw, u64 abit)
    {
        payload = pl; n_words = nw; wi = abit >> 5;
        const u32 sh = (u32)(abit & 31);
        acc = (u64)(wi < n_words ? ld_stream(&payload[wi]) : 0u) >> sh;
        nb = 32 - sh;
        ++wi;
    }
    TRPX_DEVICE u32 get(u32 n)      // n in [0, 32]
    {
        if (nb < n) {
            acc |= (u64)(wi < n_words ? ld_stream(&payload[wi]) : 0u) << nb;
            nb += 32;
            ++wi;
        }
        const u32 v = n == 32 ? (u32)acc : (u32)acc & ((1u << n) - 1);
        acc >>= n;
        nb -= n;
        return v;
    }
    TRPX_DEVICE u64 get_wide(u32 s)  // s in [1, 73]; returns the low 64 bits
    {
        u64 v = get(s < 32 ? s : 32);
        if (s > 32) v |= (u64)get(s - 32 < 32 ? s - 32 : 32) << 32;
        if (s > 64) (void)get(s - 64);                            // bits above 64 repeat the sign
        return v;
    }
};

// Bit_range::get_range semantics (Bit_pointer.hpp:742-792): sign-extend signed streams from bit
// s-1; a block wider than the output type is clamped to the type's range, else truncated.
template <typename O> struct IsFloat { static constexpr bool V = false; };
template <> struct IsFloat<float> { static constexpr bool V = true; };
template <> struct IsFloat<double> { static constexpr bool V = true; };

// Floating-point outputs go through a 64-bit integer and a double (Terse.hpp:379-383: `begin[i] =
// double(std::int64_t(bitr))`, unsigned streams through std::uint64_t), never clamped.
template <typename O, bool SGN>
TRPX_DEVICE O convert_value(u64 raw, u32 s)
{
    constexpr u32 WO = 8 * sizeof(O);
    constexpr bool OS = O(-1) < O(0);
    u64 v = raw;
    if (SGN && s < 64 && ((v >> (s - 1)) & 1)) v |= ~0ull << s;
    if (IsFloat<O>::V) return (O)(SGN ? (double)(i64)v : (double)v);
    if (s > WO) {
        if (!OS) {
            const u64 hi = WO == 64 ? ~0ull : ((1ull << WO) - 1);
            const u64 x = SGN ? ((i64)v < 0 ? 0ull : v) : v;
            v = x > hi ? hi : x;
        } else {
            const i64 hi = (i64)((1ull << (WO - 1)) - 1), lo = -hi - 1;
            const i64 x = SGN ? (i64)v : (v > 0x7fffffffffffffffull ? 0x7fffffffffffffffll : (i64)v);
            v = (u64)(x < lo ? lo : (x > hi ? hi : x));
        }
    }
    return (O)v;
}

constexpr int DEC_NT = 256;
constexpr int DEC_RB = 4;                               // blocks per thread in the tile scan
constexpr int DEC_TB = DEC_NT * DEC_RB;                 // blocks per P2 tile

TRPX_HD u32 header_len(u32 s, u32 prev) { return s == prev ? 1u : (s < 7 ? 4u : (s < 10 ? 6u : 12u)); }

// STAGED: block == 12, 16-byte aligned frames -> values leave through shared memory + TMA store.
// otherwise: each thread stores its block's values straight to global memory.
template <typename O, bool SGN, bool STAGED>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(DEC_NT, 1) prolix_unpack_kernel(DecParams p)
{
    constexpr int NT = DEC_NT;
    TRPX_DYN_SMEM(sm);
    u32* sm_warp_tot = (u32*)sm;                                  // 32 u32
    unsigned char* wsm = sm + 128;                                // DEC_TB + 1 widths (+ pad)
    u32* tp = (u32*)(sm + 128 + DEC_TB + 16);                     // DEC_TB data offsets (bits, tile-relative)
    O* stage = (O*)(sm + 128 + DEC_TB + 16 + DEC_TB * 4);         // DEC_TB * 12 values (STAGED only)
    const u32 t = tid(), lane = t & 31, warp = t >> 5;
    const u64 tile = bid();
    const u64 f = tile / p.tiles_per_frame, tif = tile % p.tiles_per_frame;
    const u64 b0 = tif * DEC_TB;
    const u64 nb_tile = p.nblocks - b0 < (u64)DEC_TB ? p.nblocks - b0 : (u64)DEC_TB;
    const unsigned char* wrow = p.widths + f * p.nblocks;
    const u64 n_words = (p.payload_bytes + 3) >> 2;
    const u64 base_bit = (f ? p.frame_ends[f - 1] : 0) * 8;
    const u64 frame_bits = (p.frame_ends[f] - (f ? p.frame_ends[f - 1] : 0)) * 8;
    const u64 anchor = p.anchors[tile];

    // widths of blocks b0-1 .. b0+nb_tile-1 -> wsm[0 .. nb_tile]
    for (u32 i = t; i <= nb_tile; i += NT) wsm[i] = (i == 0) ? (b0 ? wrow[b0 - 1] : 0) : wrow[b0 + i - 1];
    sync_block();

    // lengths of my DEC_RB consecutive blocks -> tile-relative data offsets
    u32 hl[DEC_RB], len = 0;
#pragma unroll
    for (int i = 0; i < DEC_RB; ++i) {
        const u32 k = t * DEC_RB + i;
        hl[i] = 0;
        if (k < nb_tile) {
            const u32 s = wsm[k + 1];
            const u32 cnt = (b0 + k + 1 == p.nblocks) ? p.last_cnt : p.block;
            hl[i] = header_len(s, wsm[k]);
            len += hl[i] + s * cnt;
        }
    }
    u32 incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 v = shfl_up(incl, d);
        if (lane >= (u32)d) incl += v;
    }
    if (lane == 31) sm_warp_tot[warp] = incl;
    sync_block();
    u32 off = incl - len, total = 0;
    for (int i = 0; i < NT / 32; ++i) {
        const u32 v = sm_warp_tot[i];
        if ((u32)i < warp) off += v;
        total += v;
    }
#pragma unroll
    for (int i = 0; i < DEC_RB; ++i) {
        const u32 k = t * DEC_RB + i;
        if (k < nb_tile) {
            const u32 s = wsm[k + 1];
            const u32 cnt = (b0 + k + 1 == p.nblocks) ? p.last_cnt : p.block;
            tp[k] = off + hl[i];
            off += hl[i] + s * cnt;
        }
    }
    if (t == 0 && anchor + total > frame_bits) atomic_max(p.status, DEC_MALFORMED);
    sync_block();

    // unpack: thread t takes blocks t, t+NT, ... (neighbouring lanes read neighbouring bits)
    O* outf = (O*)p.out + f * p.n_values;
    for (u32 k = t; k < nb_tile; k += NT) {
        const u32 s = wsm[k + 1];
        const u32 cnt = (b0 + k + 1 == p.nblocks) ? p.last_cnt : p.block;
        O* dst = STAGED ? stage + (size_t)k * 12 : outf + (b0 + k) * p.block;
        if (s == 0) {
            for (u32 i = 0; i < cnt; ++i) dst[i] = (O)0;
            continue;
        }
        BitSource src;
        src.init(p.payload, n_words, base_bit + anchor + tp[k]);
        if (s <= 32) {
            for (u32 i = 0; i < cnt; ++i) dst[i] = convert_value<O, SGN>(src.get(s), s);
        } else {
            for (u32 i = 0; i < cnt; ++i) dst[i] = convert_value<O, SGN>(src.get_wide(s), s);
        }
    }
    if (STAGED) {
        fence_async_smem();
        sync_block();
        if (t == 0) {
            const u64 v0 = b0 * 12;
            const u64 nv = p.n_values - v0 < (u64)DEC_TB * 12 ? p.n_values - v0 : (u64)DEC_TB * 12;
            bulk_s2g(outf + v0, stage, (u32)(nv * sizeof(O)));
            bulk_commit();
            bulk_wait_read0();
        }
    }
}

// ------------------------------------------------------------------ P2 (fast path): re-walk 32-byte sub-segments and unpack
// One thread per sub-segment: from its exact checkpoint it reads headers and values with ONE
// sequential bit reader over shared memory (the CTA's slice of the stream is staged there with one
// pad word per sub-segment, so that lanes 32 bytes apart hit different banks), and writes whole
// blocks into an output stage in block order.  The stage leaves with one TMA bulk store.  No widths
// array, no second header pass over global memory.
constexpr int UNP_NT = 256;                              // sub-segments (threads) per CTA: 8 KB of stream
constexpr u32 UNP_TAIL_WORDS = 64;                       // a block that starts in the slice ends inside this tail
constexpr u32 UNP_SPAN_WORDS = UNP_NT * (SUB_BYTES / 4) + UNP_TAIL_WORDS + 4;
// The slice is staged COLUMN-wise: 8-word (32-byte) columns, word r of column c at sp[r * UNP_CP + c].  With a
// pitch that is a multiple of 32 the bank of an access is its column, so lanes working in different
// columns never conflict, wherever they are inside their columns.  Rows 8..19 of a column repeat the
// words of the next columns (r = 8 + k: column c+1 word k; r = 16 + k: column c+2 word k), so a thread
// reads up to 20 consecutive words of the stream with one multiply-add per access and "next word" is
// always + UNP_CP.
constexpr u32 UNP_CP = 288, UNP_ROWS = 20;
constexpr u32 UNP_SPAN_SMEM_WORDS = UNP_ROWS * UNP_CP;
#ifndef UNP_STAGE_KB
#define UNP_STAGE_KB 40
#endif
#ifndef UNP_CTAS
#define UNP_CTAS 3          // resident CTAs per SM the unpack kernel is sized for (shared memory: stage + span + table)
#endif
constexpr u32 UNP_STAGE_BYTES = UNP_STAGE_KB * 1024;
constexpr u32 UNP_SM_TAB = 64;                           // byte offsets inside dynamic shared memory
constexpr u32 UNP_SM_SPAN = UNP_SM_TAB + HDR_TAB_BYTES;
constexpr u32 UNP_SM_STAGE = (UNP_SM_SPAN + UNP_SPAN_SMEM_WORDS * 4 + 127) / 128 * 128;
constexpr u32 UNP_SMEM_BYTES = UNP_SM_STAGE + UNP_STAGE_BYTES + 32;
static_assert((UNP_SPAN_WORDS + 7) / 8 <= UNP_CP, "columns of the staged slice");

TRPX_DEVICE u32 span_primary(u32 i) { return (i & 7) * UNP_CP + (i >> 3); }   // home of logical word i

struct SmemBits {                                        // sequential reader over the staged slice (any extent)
    const u32* sp;
    u64 acc;
    u32 nb, wi;
    TRPX_DEVICE void init(const u32* sp_, u32 bit)
    {
        sp = sp_; wi = bit >> 5;
        acc = (u64)sp[span_primary(wi)] >> (bit & 31);
        nb = 32 - (bit & 31);
        ++wi;
    }
    TRPX_DEVICE void fill()                              // afterwards nb >= 32
    {
        if (nb < 32) { acc |= (u64)sp[span_primary(wi)] << nb; nb += 32; ++wi; }
    }
    TRPX_DEVICE void skip(u32 n) { acc >>= n; nb -= n; } // n <= nb
    TRPX_DEVICE u32 get(u32 n)                           // n in [0, 32]
    {
        fill();
        const u32 v = n == 32 ? (u32)acc : (u32)acc & ((1u << n) - 1);
        skip(n);
        return v;
    }
    TRPX_DEVICE u64 get_wide(u32 s)                      // s in [1, 73]; returns the low 64 bits
    {
        u64 v = get(s < 32 ? s : 32);
        if (s > 32) v |= (u64)get(s - 32 < 32 ? s - 32 : 32) << 32;
        if (s > 64) (void)get(s - 64);
        return v;
    }
};

// 32 stream bits starting `rel` bits into the thread's column (rows 0..19: rel + 32 <= 640).  col_a is the
// shared-window address of the column's row 0: shift, multiply-add, two loads, one funnel shift (which takes
// rel mod 32 by itself).
TRPX_DEVICE u32 col_window(saddr_t col_a, u32 rel)
{
    const saddr_t a = col_a + (rel >> 5) * (UNP_CP * 4);
    return funnel_r(lds_u32(a), lds_u32_at<UNP_CP * 4>(a), rel);
}
TRPX_DEVICE void lookup_header_s(saddr_t tab_a, u32 win, u32 s, u32& hl, u32& s_new)
{
    const u32 e = lds_u16(tab_a + ((win & (HDR_TAB_ENTRIES - 1)) << 1));
    hl = e & 15;
    s_new = (e & HDR_SAME) ? s : e >> 8;
}

template <typename O> struct UnpCap { static constexpr u32 BLOCKS = UNP_STAGE_BYTES / (12 * sizeof(O)); };

template <u32 SO>
TRPX_DEVICE void sts_zero_block(saddr_t d)                // 12 values of SO bytes, aligned as the stores below
{
    if (SO == 1) { sts_u32(d, 0); sts_u32(d + 4, 0); sts_u32(d + 8, 0); }
    else if (SO == 2) { sts_v2(d, 0, 0); sts_v2(d + 8, 0, 0); sts_v2(d + 16, 0, 0); }
    else
#pragma unroll
        for (u32 i = 0; i < 12 * SO; i += 16) sts_v4(d + i, 0, 0, 0, 0);
}

// One full block (12 values of width s >= 1 starting at bit `pos`) -> dst (aligned for the vector
// stores used below).  Every field group is fetched straight from its own bit position: independent
// extractions, no serial bit-reader state.
template <typename O, bool SGN>
TRPX_DEVICE void unpack_block12(const u32* sp, saddr_t col_a, u32 rel, u32 pos, u32 s, saddr_t dst_a)
{
    constexpr u32 SO = sizeof(O);
    const bool in_rows = rel + 12 * s + 32 <= UNP_ROWS * 32;      // the whole block is inside the column's 20 rows
    if (SO == 2 && s <= 16 && rel + 8 * s + 96 <= UNP_ROWS * 32) {
        // Four fields (4s <= 64 bits) per 64-bit window, three windows per block -- the same code for every width, so
        // lanes with narrow and wide blocks stay converged.  Two fields of s bits become two 16-bit lanes with one
        // multiply-add (the encoder's trick reversed).
        const u32 m2 = low_mask(2 * s);
        const u32 K = 65536u - (1u << s);
        u32 o[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const u32 r = rel + 4 * s * i;
            const saddr_t a = col_a + (r >> 5) * (UNP_CP * 4);
            const u32 w0 = lds_u32(a), w1 = lds_u32_at<UNP_CP * 4>(a), w2 = lds_u32_at<UNP_CP * 8>(a);
            const u32 lo = funnel_r(w0, w1, r), hi = funnel_r(w1, w2, r);
            const u32 p0 = lo & m2, p1 = funnel_rc(lo, hi, 2 * s) & m2;
            o[2 * i] = p0 + (p0 >> s) * K;
            o[2 * i + 1] = p1 + (p1 >> s) * K;
        }
        if (SGN) {
            const u32 KS = (0xffffu << s) & 0xffffu;      // sign extension of a 16-bit lane
#pragma unroll
            for (int i = 0; i < 6; ++i) o[i] |= ((o[i] >> (s - 1)) & 0x00010001u) * KS;
        }
        sts_v2(dst_a, o[0], o[1]); sts_v2(dst_a + 8, o[2], o[3]); sts_v2(dst_a + 16, o[4], o[5]);
        return;
    }
    if (SO == 1 && s <= 8 && in_rows) {
        const u32 m4 = low_mask(4 * s);
        const u32 m2 = (1u << (2 * s)) - 1;
        const u32 K8 = 256u - (1u << s);
        const u32 KS = (0xffu << s) & 0xffu;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const u32 q = col_window(col_a, rel + 4 * s * i) & m4;
            const u32 p0 = q & m2, p1 = s == 8 ? (q >> 16) : (q >> (2 * s));
            u32 x = (p0 + (p0 >> s) * K8) | ((p1 + (p1 >> s) * K8) << 16);
            if (SGN) x |= ((x >> (s - 1)) & 0x01010101u) * KS;
            sts_u32(dst_a + 4 * i, x);
        }
        return;
    }
    if (SO == 4 && !IsFloat<O>::V && s <= 32 && in_rows) {
        const u32 m = low_mask(s);
        u32 o[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            u32 v = col_window(col_a, rel + s * i) & m;
            if (SGN && s < 32 && ((v >> (s - 1)) & 1)) v |= ~0u << s;
            o[i] = v;
        }
        sts_v4(dst_a, o[0], o[1], o[2], o[3]); sts_v4(dst_a + 16, o[4], o[5], o[6], o[7]); sts_v4(dst_a + 32, o[8], o[9], o[10], o[11]);
        return;
    }
    O* dst = (O*)saddr_to_ptr(dst_a);
    SmemBits br;
    br.init(sp, pos);
    for (u32 i = 0; i < 12; ++i) dst[i] = convert_value<O, SGN>(s <= 32 ? (u64)br.get(s) : br.get_wide(s), s);
}

TRPX_HD u32 umin3(u32 a, u32 b, u32 c) { const u32 m = a < b ? a : b; return m < c ? m : c; }

// What a thread needs to know about one slice (a CTA's 8 KB of stream); loaded two slices ahead.
struct SliceDesc {
    u64 seg_bit, frame_end_bit, b0;     // CTA-uniform: segment start (absolute bit), frame end, first block of the segment
    u32 frame, seg_cnt, h;
    u64 rowA, rowB;                     // checkpoints that bound the CTA's headers
    u64 c0, c1;                         // this thread's checkpoint and the next one
};
constexpr u32 UNP_CHUNKS = (UNP_SPAN_WORDS / 4 + UNP_NT - 1) / UNP_NT;    // 16-byte pieces of the slice per thread

TRPX_DEVICE SliceDesc load_slice_desc(const DecParams& p, u32 seg, u32 h)     // slice h of segment seg
{
    SliceDesc d;
    const u64 j = seg;
    d.h = h;
    const u64* sd = p.segd + j * 4;                                  // count 0 past the last segment
    // (opaque loads: these values are consumed two iterations later and must stay in flight until then)
    d.seg_bit = ldg_u64_opaque(sd); d.frame_end_bit = ldg_u64_opaque(sd + 1); d.b0 = ldg_u64_opaque(sd + 2);
    const u64 fc = ldg_u64_opaque(sd + 3);
    d.frame = (u32)fc; d.seg_cnt = (u32)(fc >> 32);
    const u64* row = p.ckpt + j * p.subs_per_seg;
    const u32 m = d.h * UNP_NT + tid();
    d.rowA = d.h * UNP_NT < p.subs_per_seg ? ldg_u64_opaque(row + d.h * UNP_NT) : ~0ull;
    d.rowB = (d.h + 1) * UNP_NT < p.subs_per_seg ? ldg_u64_opaque(row + (d.h + 1) * UNP_NT) : ~0ull;
    d.c0 = m < p.subs_per_seg ? ldg_u64_opaque(row + m) : ~0ull;
    d.c1 = m + 1 < p.subs_per_seg ? ldg_u64_opaque(row + m + 1) : ~0ull;
    return d;
}
TRPX_DEVICE u64 slice_a0(const SliceDesc& d, u32 sub_shift)      // 16-byte aligned byte offset of the slice's first word in the payload
{
    return ((d.seg_bit + (((u64)d.h * UNP_NT) << sub_shift)) >> 3) & ~15ull;
}
// (span_chunks: 16-byte pieces of this call's slice: UNP_NT sub-segments + tail)
TRPX_DEVICE void fetch_slice(const DecParams& p, u64 a0, u32 span_chunks, uint4 (&pre)[UNP_CHUNKS])
{
    const u64 safe_end = p.payload_bytes & ~15ull;
    const u64 n_words = (p.payload_bytes + 3) >> 2;
    const unsigned char* base = (const unsigned char*)p.payload;
#pragma unroll
    for (u32 q = 0; q < UNP_CHUNKS; ++q) {
        const u32 c = tid() + q * UNP_NT;
        const u64 b = a0 + 16ull * c;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < span_chunks) {
            if (b + 16 <= safe_end) {
                v = *(const uint4*)(base + b);
            } else if (b < p.payload_bytes) {
                const u64 wi = b >> 2;
                v.x = wi < n_words ? p.payload[wi] : 0u;
                v.y = wi + 1 < n_words ? p.payload[wi + 1] : 0u;
                v.z = wi + 2 < n_words ? p.payload[wi + 2] : 0u;
                v.w = wi + 3 < n_words ? p.payload[wi + 3] : 0u;
            }
        }
        pre[q] = v;
    }
}
TRPX_DEVICE void stage_slice(u32* span, u32 span_chunks, const uint4 (&pre)[UNP_CHUNKS])
{
#pragma unroll
    for (u32 q = 0; q < UNP_CHUNKS; ++q) {
        const u32 c = tid() + q * UNP_NT;
        if (c < span_chunks) {
            const u32 i = 4 * c, col = i >> 3, r0 = i & 7;           // 4 words of one column: rows r0 .. r0+3
            const u32 vv[4] = {pre[q].x, pre[q].y, pre[q].z, pre[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                span[(r0 + e) * UNP_CP + col] = vv[e];
                if (col >= 1) span[(8 + r0 + e) * UNP_CP + col - 1] = vv[e];
                if (col >= 2 && r0 == 0) span[(16 + e) * UNP_CP + col - 2] = vv[e];
            }
        }
    }
}

// Persistent CTAs: slice i is unpacked while the stream of slice i+1 and the descriptors of slice i+2 are in flight.
template <typename O, bool SGN>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(UNP_NT, UNP_CTAS) prolix_unpack_seg_kernel(DecParams p)
{
    constexpr u32 SO = sizeof(O);
    constexpr u32 CB = UnpCap<O>::BLOCKS;
    TRPX_DYN_SMEM(sm);
    unsigned short* tab = (unsigned short*)(sm + UNP_SM_TAB);
    u32* span = (u32*)(sm + UNP_SM_SPAN);
    unsigned char* stage = sm + UNP_SM_STAGE;
    const saddr_t tab_a = saddr(tab), span_a = saddr(span), stage_a = saddr(stage);
    const u32 t = tid();
    const u32 parts = (p.subs_per_seg + UNP_NT - 1) / UNP_NT;       // slices per segment
    const u64 total64 = p.max_segs * parts;
    const u32 total = total64 < 0xffffffffull ? (u32)total64 : 0xffffffffu;
    const u32 step = nblocks();
    u32 sl = bid();
    if (sl >= total) return;
    copy_header_table<UNP_NT>(tab, p.hdr_tab);

    // slices sl, sl + step, ... : (segment, slice-in-segment) pairs advanced incrementally, iterations counted down
    const u32 step_j = step / parts, step_h = step % parts;
    u32 left = (total - 1 - sl) / step;                             // iterations after this one
    u32 nj = sl / parts, nh = sl % parts;                           // the slice the next load_slice_desc() is for
    auto advance = [&]() { nj += step_j; nh += step_h; if (nh >= parts) { nh -= parts; ++nj; } };
    SliceDesc d0 = load_slice_desc(p, nj, nh), d1 = d0, d2 = d0;
    advance();
    if (left >= 1) { d1 = load_slice_desc(p, nj, nh); advance(); }
    uint4 pre[UNP_CHUNKS];
    const u32 sub_shift = p.sub_shift;
    const bool sparse = sub_shift < SUB_SHIFT_MAX;
    const u32 span_chunks = (((u32)UNP_NT << (sub_shift - 5)) + UNP_TAIL_WORDS + 4) / 4;   // words of UNP_NT sub-segments + tail
    const u32 pos_limit = (((u32)UNP_NT << (sub_shift - 3)) + 16) * 8;                    // no header is read past this bit
    fetch_slice(p, slice_a0(d0, sub_shift), span_chunks, pre);
    for (;;) {
        const bool have1 = left >= 1, have2 = left >= 2;
        if (have2) { d2 = load_slice_desc(p, nj, nh); advance(); }   // arrives during this iteration
        const u64 a0 = slice_a0(d0, sub_shift);
        stage_slice(span, span_chunks, pre);          // (nobody reads the span any more: every thread is past the barrier that followed the unpack loop)
        if (t == 0) bulk_wait_read0();                // the previous slice's bulk store has finished reading the output stage
        sync_block();
        if (have1) fetch_slice(p, slice_a0(d1, sub_shift), span_chunks, pre);               // in flight while this slice is unpacked

        // ---- headers of this CTA: [kA, kB) in segment-local numbering, clipped at the frame's last block
        const u32 kA = d0.rowA != ~0ull ? ckpt_n(d0.rowA) : d0.seg_cnt;
        u32 kB = d0.rowB != ~0ull ? ckpt_n(d0.rowB) : d0.seg_cnt;
        const bool live = d0.seg_cnt != 0 && d0.b0 < p.nblocks;
        if (live && (u64)kB > p.nblocks - d0.b0) kB = (u32)(p.nblocks - d0.b0);
        if (live && kA < kB) {
            // ---- my headers: [k, k_end), the first one at bit `pos` of the span
            u32 k = 0, k_end = 0, s = 0, pos = 0, cbase = 0;
            if (d0.c0 != ~0ull) {
                k = ckpt_n(d0.c0);
                k_end = d0.c1 != ~0ull ? ckpt_n(d0.c1) : d0.seg_cnt;
                if (k_end > kB) k_end = kB;
                s = ckpt_s(d0.c0);
                pos = (u32)(d0.seg_bit + ckpt_rel(d0.c0) - a0 * 8);
                cbase = pos & ~255u;                                 // first bit of the column my first header is in
            }
            const u32 k_begin = k;
            const saddr_t col_a = span_a + (cbase >> 8) * 4;         // my column's row 0
            const u64 b0 = d0.b0;
            const u32 k_last = p.nblocks - 1 - b0 < 0xffffffffull ? (u32)(p.nblocks - 1 - b0) : 0xffffffffu;   // the frame's (possibly ragged) last block
            const u32 frame_end_pos = (u32)((d0.frame_end_bit - a0 * 8 < 0xffffffffull) ? d0.frame_end_bit - a0 * 8 : 0xffffffffull);
            O* outf = (O*)p.out + (u64)d0.frame * p.n_values;
            for (u32 c0 = kA; c0 < kB; c0 += CB) {
                const u32 c1 = c0 + CB < kB ? c0 + CB : kB;
                // the chunk's first value in global memory; the stage mirrors its 16-byte phase
                const u64 v0 = (b0 + c0) * 12;
                unsigned char* gdst = (unsigned char*)(outf + v0);
                const u32 phase = (u32)((uintptr_t)gdst & 15);
                unsigned char* sbase = stage + phase;
                // full blocks (12 values) in a tight loop; the frame's ragged last block, if it is mine, afterwards
                const u32 stop_full = umin3(k_end, c1, k_last);
                saddr_t dst_a = stage_a + phase + (k - c0) * (12 * SO);   // (only used while k is inside the chunk)
                if (sparse) {
                    // Sparse streams (checkpoints closer than 32 bytes): the stage starts out all zero (one cooperative
                    // sweep), so empty blocks cost nothing and a run of their one-bit headers is skipped 32 at a time.
                    const u32 n16 = (phase + (c1 - c0) * (12 * SO) + 15) >> 4;
                    uint4* z = (uint4*)stage;
                    for (u32 i = t; i < n16; i += UNP_NT) z[i] = make_uint4(0, 0, 0, 0);
                    sync_block();
                    while (k < stop_full) {
                        const u32 win = col_window(col_a, pos - cbase);
                        if (s == 0 && (win & 1)) {
                            u32 run = (u32)ffs32(~win) - 1;          // ffs32(0) == 0 -> 0xffffffff: all 32 bits set
                            run = run > 32u ? 32u : run;
                            run = run > stop_full - k ? stop_full - k : run;
                            pos += run;
                            k += run;
                            dst_a += run * (12 * SO);
                        } else {
                            u32 hl;
                            lookup_header_s(tab_a, win, s, hl, s);
                            pos += hl;
                            if (s != 0) unpack_block12<O, SGN>(span, col_a, pos - cbase, pos, s, dst_a);
                            pos += s * 12;
                            ++k;
                            dst_a += 12 * SO;
                        }
                        if (k < stop_full && pos >= pos_limit) { atomic_max(p.status, DEC_MALFORMED); k = k_end; break; }
                    }
                } else {
                    while (k < stop_full) {
                        u32 hl;
                        lookup_header_s(tab_a, col_window(col_a, pos - cbase), s, hl, s);   // (a header starts < 384 bits into the column)
                        pos += hl;
                        if (s == 0) sts_zero_block<SO>(dst_a);
                        else unpack_block12<O, SGN>(span, col_a, pos - cbase, pos, s, dst_a);
                        pos += s * 12;
                        ++k;
                        dst_a += 12 * SO;
                        if (k < stop_full && pos >= pos_limit) { atomic_max(p.status, DEC_MALFORMED); k = k_end; break; }   // never read a header past the tail
                    }
                }
                if (k == k_last && k < k_end && k < c1) {           // the (possibly ragged) last block of the frame
                    u32 hl;
                    lookup_header_s(tab_a, col_window(col_a, pos - cbase), s, hl, s);
                    pos += hl;
                    O* dst = (O*)(sbase + (k - c0) * (12 * SO));
                    SmemBits br;
                    br.init(span, pos);
                    for (u32 i = 0; i < p.last_cnt; ++i)
                        dst[i] = s == 0 ? (O)0 : convert_value<O, SGN>(s <= 32 ? (u64)br.get(s) : br.get_wide(s), s);
                    pos += s * p.last_cnt;
                    ++k;
                }
                if (k == k_end && k_end > k_begin && pos > frame_end_pos) atomic_max(p.status, DEC_MALFORMED);   // my last block runs past the frame's end
                fence_async_smem();
                sync_block();
                // ---- store [v0, v1): unaligned head and tail bytes by hand, the 16-byte aligned middle as one bulk copy
                const u64 v1 = (b0 + c1) * 12 < p.n_values ? (b0 + c1) * 12 : p.n_values;
                const u32 bytes = (u32)((v1 - v0) * SO);
                u32 head = (16 - phase) & 15;
                if (head > bytes) head = bytes;
                const u32 mid = (bytes - head) & ~15u;
                for (u32 i = t; i < head; i += UNP_NT) gdst[i] = sbase[i];
                for (u32 i = head + mid + t; i < bytes; i += UNP_NT) gdst[i] = sbase[i];
                if (t == 0 && mid) {
                    bulk_s2g(gdst + head, sbase + head, mid);       // drains while the next slice is staged
                    bulk_commit();
                }
                if (c0 + CB < kB) {                                 // another chunk of this slice reuses the stage
                    if (t == 0) bulk_wait_read0();
                    sync_block();
                }
            }
        }
        if (!have1) break;
        --left;
        d0 = d1;
        d1 = d2;
    }
    if (t == 0) bulk_wait_read0();                                  // shared memory must outlive the last bulk store
}

// ------------------------------------------------------------------ frame sizes unknown
// The .trpx header stores only the total payload size (Terse.hpp:459), so a foreign multi-frame
// payload has to be walked to find where each frame ends (Terse.hpp:562-585).  Round-1 version: one
// thread walks the chain (correct for any stream; slow -- DESIGN.md lists the parallel replacement).
TRPX_KERNEL void prolix_single_frame_kernel(u64* frame_ends_out, u64 payload_bytes)
{
    if (bid() == 0 && tid() == 0) frame_ends_out[0] = payload_bytes;
}

// ------------------------------------------------------------------ frame sizes unknown, fast path
// The .trpx header stores only the total payload size (Terse.hpp:459), so the frame boundaries of a foreign stack have
// to be recovered from the stream (the reference walks every header of every earlier frame, Terse.hpp:562-585).  Here:
//   1. the speculative walkers + resolve kernel above parse the WHOLE payload as one run of full blocks, ignoring
//      frames ("G"): exact checkpoints (position, carried width, header count) every sub-segment, in parallel.
//   2. the true parse ("T") differs from G only near frame boundaries -- the last block of a frame may be ragged, the
//      frame is padded to whole bytes, the width restarts at 0 -- and G, being self-synchronising, is back on the true
//      chain a few hundred bytes later.  So ONE warp follows the frames: from a frame's start it walks T until T meets
//      one of G's checkpoints (same position, same carried width or an explicit header), then JUMPS: the frame's last
//      block is G's header number (G's count at the meeting point) + (blocks of the frame still to go), found by a
//      search over the segment counts and one checkpoint row, a handful of steps from there.
// Per frame that is ~100 header steps in shared memory and three look-ups instead of a walk over all its blocks
// (21 846 for a 512 x 512 frame).  The chain over frames itself stays serial: where frame f+1 starts is only known once
// frame f has been followed to its end.
// The chain's own stream window: 1 KB, refilled by ONE round of independent loads (a frame touches the stream in
// three places -- its start, the checkpoint before its last block, the last block -- and reads a few hundred
// bytes at each, so the refill latency, not its size, is what a frame costs).
constexpr u32 FW_WORDS = 256;
struct FrameWin {
    u32* chunk;
    u64 chunk_bit;                                         // absolute bit of chunk[0] (multiple of 128); ~0: nothing staged
    TRPX_DEVICE void need(const DecParams& p, u64 n_words, u64 abit, u32 span_bits)      // all lanes
    {
        if (chunk_bit != ~0ull && abit >= chunk_bit && abit - chunk_bit + span_bits + 64 <= (u64)FW_WORDS * 32) return;
        sync_warp();
        chunk_bit = abit & ~127ull;
        const u64 w0 = chunk_bit >> 5;
        const u32 lane = tid() & 31;
        u32 v[(FW_WORDS + 4 + 31) / 32];
#pragma unroll
        for (u32 k = 0; k < (FW_WORDS + 4 + 31) / 32; ++k) {
            const u32 i = lane + 32 * k;
            v[k] = (i < FW_WORDS + 4 && w0 + i < n_words) ? p.payload[w0 + i] : 0u;
        }
#pragma unroll
        for (u32 k = 0; k < (FW_WORDS + 4 + 31) / 32; ++k) {
            const u32 i = lane + 32 * k;
            if (i < FW_WORDS + 4) chunk[i] = v[k];
        }
        sync_warp();
    }
    TRPX_DEVICE u32 peek(u64 abit) const
    {
        const u32 q = (u32)(abit - chunk_bit);
        return funnel_r(chunk[q >> 5], chunk[(q >> 5) + 1], q & 31);
    }
};

// T: one warp follows the frames along G's checkpoints.  The chain is serial over frames (where frame f + 1 starts
// is only known once frame f has been followed to its end), so what counts is the number of DEPENDENT global-memory
// round trips per frame; every step below is one: (1) the stream window at the frame start together with the 32
// checkpoints that follow it, (2) 32 candidate segments for the frame's last header, starting where the previous
// frame's span says it will be, (3) that segment's whole checkpoint row, (4) the stream window at the checkpoint.
// `resume` (may be null): resume[0] = frames whose ends the speculative pass below has already written.
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(32, 1) prolix_frame_chain_kernel(DecParams p, u64* frame_ends_out, const u64* resume)
{
    TRPX_SHARED u32 chunk[FW_WORDS + 4];
    if (bid() != 0) return;
    const u64 f_first = resume ? resume[0] : 0;
    if (f_first >= p.n_frames) return;
    const u32 lane = tid() & 31;
    const u64 n_words = (p.payload_bytes + 3) >> 2;
    const u64 total_bits = p.payload_bytes * 8;
    const u64 n_segs = p.seg_base[1];                       // G: one "frame" = the whole payload
    const u64 seg_bits = (u64)p.seg_bytes * 8;
    const u32 subs = p.subs_per_seg, sh = p.sub_shift;
    const u64 n_ck = n_segs * subs;
    const u64 max_block_bits = 12 + (u64)p.block * 73;
    FrameWin win;
    win.chunk = chunk;
    win.chunk_bit = ~0ull;
    u64 P = f_first ? frame_ends_out[f_first - 1] * 8 : 0;  // absolute bit at which the frame starts (byte aligned)
    bool bad = max_block_bits + 64 > (u64)FW_WORDS * 32;    // (huge blocks: the caller uses the plain walker instead)
    u64 prev_span = 0;                                      // segments the previous frame's jump went forward
    for (u64 f = f_first; f < p.n_frames; ++f) {
        // ---- T from the frame start until it meets G (or the frame ends first: tiny frames)
        u64 pos = P;
        u32 s = 0;
        u64 c = 0;                                          // headers of this frame before `pos`
        bool met = false;
        u64 g_idx = 0;                                      // G's number of the header at `pos` when met
        u64 j = pos / seg_bits;                             // segment of `pos`, tracked incrementally
        u64 seg_first = j * seg_bits;
        u64 ckq0 = ~0ull, last_q = ~0ull;                   // flat index of lane 0's cached checkpoint / of the last one compared
        u64 ck_l = ~0ull;
        while (!bad && c + 1 < p.nblocks && !met) {
            if (pos >= total_bits) { bad = true; break; }
            while (pos >= seg_first + seg_bits) { ++j; seg_first += seg_bits; }
            // G's checkpoint of the sub-segment `pos` lies in names the first header at or after the sub-segment's
            // first bit: T has met G when that is exactly `pos`, with the same carried width
            const u32 m = (u32)((pos - seg_first) >> sh);
            const u64 q = j * subs + m;
            if (j < n_segs && m < subs && q != last_q) {
                last_q = q;
                if (ckq0 == ~0ull || q < ckq0 || q >= ckq0 + 32) {
                    ckq0 = q;
                    ck_l = q + lane < n_ck ? p.ckpt[q + lane] : ~0ull;
                }
                const u64 ck = shfl(ck_l, (int)(q - ckq0));
                win.need(p, n_words, pos, 64);
                const bool expl = (win.peek(pos) & 1) == 0;
                if (seg_first + ckpt_rel(ck) == pos && (ckpt_s(ck) == (s & 0xff) || expl)) {
                    met = true;
                    g_idx = p.seg_b0[j] + ckpt_n(ck);
                    break;
                }
            }
            win.need(p, n_words, pos, (u32)max_block_bits);
            const u32 hl = decode_header((u64)win.peek(pos), s);
            pos += hl + (u64)s * p.block;
            ++c;
        }
        // ---- jump to the frame's last block along G
        if (!bad && met && c + 1 < p.nblocks) {
            const u64 target = g_idx + (p.nblocks - 1 - c);  // G's number of the frame's last header
            const u64 j_from = j;
            u64 j0 = j + (prev_span > 24 ? prev_span - 16 : 0);
            for (;;) {                                      // the segment that holds header `target`: 32 candidates per round
                const u64 jj = j0 + lane;
                const u64 b0 = jj < n_segs ? p.seg_b0[jj] : ~0ull;
                const u32 cn = jj < n_segs ? p.seg_count[jj] : 0u;
                const bool in = jj < n_segs && b0 <= target && target < b0 + cn;
                const bool past = jj >= n_segs || b0 > target;
                const u32 hit = ballot(in), over = ballot(past);
                if (hit) { j = j0 + (u32)ffs32(hit) - 1; break; }
                if ((over & 1u) && j0 > j_from) { j0 = j_from; continue; }   // guessed too far: search from the start
                if (over) { bad = true; break; }            // the stream ends before the frame does
                j0 += 32;
            }
            if (!bad) {
                prev_span = j - j_from;
                seg_first = j * seg_bits;
                const u32 t_local = (u32)(target - p.seg_b0[j]);
                const u64* row = p.ckpt + j * subs;
                // the last checkpoint at or before header t_local (checkpoint counts never decrease along a row):
                // every lane reads its share of the row at once
                const u32 per = (subs + 31) / 32;
                u64 best = 0;
                bool any = false;
#pragma unroll 8
                for (u32 k = 0; k < per; ++k) {
                    const u32 i = lane * per + k;
                    if (i < subs) {
                        const u64 ck = row[i];
                        if (ckpt_n(ck) <= t_local) { best = ck; any = true; }
                    }
                }
                const u32 okm = ballot(any);
                if (!okm) {
                    bad = true;                             // (cannot happen: sub-segment 0 counts from 0)
                } else {
                    const u64 ck = shfl(best, 31 - clz32(okm));
                    pos = seg_first + ckpt_rel(ck);
                    s = ckpt_s(ck);
                    for (u32 k = ckpt_n(ck); k < t_local && !bad; ++k) {     // a few headers at most
                        if (pos >= total_bits) { bad = true; break; }
                        win.need(p, n_words, pos, (u32)max_block_bits);
                        const u32 hl = decode_header((u64)win.peek(pos), s);
                        pos += hl + (u64)s * p.block;
                    }
                    c = p.nblocks - 1;
                }
            }
        }
        // ---- the frame's last (possibly ragged) block, then the byte padding (Terse.hpp:547)
        if (!bad) {
            if (pos >= total_bits) bad = true;
            else {
                win.need(p, n_words, pos, 64);
                const u32 hl = decode_header((u64)win.peek(pos), s);
                pos += hl + (u64)s * p.last_cnt;
            }
        }
        u64 end = (P >> 3) + 1 + ((pos - P) >> 3);
        if (bad || end > p.payload_bytes) { if (lane == 0) atomic_max(p.status, DEC_MALFORMED); end = p.payload_bytes; bad = true; }
        if (lane == 0) frame_ends_out[f] = end;
        P = end * 8;
    }
}

// ---- The frame chain, in parallel ----------------------------------------------------------------------------
// "Frame f ends with G's header e" determines everything about frame f + 1 that the chain needs: the byte the
// frame ends at, and -- by walking T from the next byte boundary until it meets G -- the number F(e) of G's header
// that ends frame f + 1.  F does not depend on f, so it can be evaluated for MANY candidate headers at once, one
// thread each, and the chain itself shrinks to table look-ups: e[f + 1] = F(e[f]).  Frames have nblocks headers
// each and T and G disagree only over the few dozen headers after a frame boundary, so e[f + k] lies within a few
// dozen headers per frame of e[f] + k * nblocks: a batch evaluates F on the window of radius R0 + RS * sqrt(k) around
// that guess for k = 0 .. B - 1 (k = 0 is exact), then one thread follows the chain through the table as far as the
// windows hold (a miss just ends the batch early; the next batch is centred on the exact value again).  What a
// thread cannot decide -- frames too short for T to meet G, a stream that ends early -- is left to the serial
// chain kernel, which resumes where this one stopped.  spec[]: [0] frames done, [1] e of the next frame,
// [2] segment of the last header located (search hint), [3] segments per frame (hint), [4] stop, [5] headers of G per
// frame as the last batches saw it (G counts a few dozen headers more than T after every boundary), then the table
// (next header, end byte, segment) per candidate.
constexpr u32 SPEC_B = 64, SPEC_R0 = 32, SPEC_RS = 80;      // defaults: ~58,000 candidates per batch of 64 frames
constexpr u32 SPEC_MAX_STEPS = 256;                         // T steps before a candidate gives up (the chain thread then walks it out)
constexpr u64 SPEC_INVALID = ~0ull;
constexpr u32 SPEC_HDR_WORDS = 32;                          // u64 words before the table
constexpr int SPEC_NT = 256;
// window radius of the k-th frame of a batch: the disagreement between T and G adds up like a random walk
TRPX_HD u32 spec_radius(u32 k, u32 r0, u32 rs)
{
    const u64 v = (u64)rs * rs * k;
    u64 r = 0;
    for (u64 bit = 1ull << 31; bit; bit >>= 1)
        if ((r | bit) * (r | bit) <= v) r |= bit;
    return r0 + (u32)r;
}
TRPX_HD u64 spec_table_entries(u32 b, u32 r0, u32 rs)
{
    u64 n = 0;
    for (u32 k = 0; k < b; ++k) n += 2 * (u64)spec_radius(k, r0, rs) + 1;
    return n;
}
constexpr u32 SPEC_ENTRY_WORDS = 5;                         // next header, end byte, segment, where T ran out of steps: bit, width | headers << 8
TRPX_HD size_t spec_scratch_bytes(u32 b, u32 r0, u32 rs) { return (size_t)(SPEC_HDR_WORDS + SPEC_ENTRY_WORDS * spec_table_entries(b, r0, rs)) * 8; }

#ifdef TRPX_EMU
inline u64& emu_spec_followed() { static u64 n = 0; return n; }   // (tests: frames whose end the speculative pass produced)
#endif
struct StreamFifo {                                         // one thread's read-ahead on the stream: four 64-bit words in flight
    const u64* q;
    u64 n_q, qi, w0, w1, w2, w3;
    TRPX_DEVICE void init(const u32* payload, u64 n_words) { q = (const u64*)payload; n_q = (n_words + 1) >> 1; qi = ~0ull - 8; w0 = w1 = w2 = w3 = 0; }
    TRPX_DEVICE u64 at(u64 i) const { return i < n_q ? q[i] : 0ull; }
    TRPX_DEVICE u64 peek(u64 abit)                          // 64 valid bits starting at absolute bit `abit`
    {
        const u64 i = abit >> 6;
        const u32 sh = (u32)(abit & 63);
        if (i != qi) {
            if (i == qi + 1) { w0 = w1; w1 = w2; w2 = w3; w3 = at(i + 3); }
            else if (i == qi + 2) { w0 = w2; w1 = w3; w2 = at(i + 2); w3 = at(i + 3); }
            else { w0 = at(i); w1 = at(i + 1); w2 = at(i + 2); w3 = at(i + 3); }
            qi = i;
        }
        return sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
    }
};

// F for one candidate: G's header b ends a frame -> the byte that frame ends at (Terse.hpp:547), the header that
// ends the NEXT frame, and b's segment.  SPEC_INVALID where the answer is the serial chain's business.
// WARP: the 32 lanes of a warp evaluate 32 candidates together; the function then has NO early exits and every loop
// runs until the last lane is done with it (a ballot per iteration), so that the lanes stay converged -- lanes left
// to drift apart through the long T loop execute it one after another.  !WARP: a single thread on its own.
template <bool WARP>
TRPX_DEVICE void spec_twalk(const DecParams& p, u64 n_segs, u64 n_words, u64 total_bits, bool ok, u64 max_steps, u64& next, u64& pos_io,
                            u32& s_io, u64& c_io);
template <bool WARP>
TRPX_DEVICE void spec_eval(const DecParams& p, u64 n_segs, u64 n_words, u64 total_bits, u64 total_blocks, bool valid, u64 b, u64 hint,
                           u64 max_steps, u64& next, u64& endb, u64& jb, u64& cap_state, u64& cap_sc)
{
    next = SPEC_INVALID; endb = SPEC_INVALID; jb = 0;
    bool ok = valid && b < total_blocks;
    const u64 seg_bits = (u64)p.seg_bytes * 8;
    const u32 subs = p.subs_per_seg;
    // ---- b's segment: the last one whose first header number is <= b; gallop from the hint, then bisect
    u64 lo = 0, hi = 1;
    if (ok) {
        const u64 h = hint < n_segs ? hint : n_segs - 1;
        u64 st = 1;
        if (p.seg_b0[h] <= b) {
            lo = h;
            while (lo + st < n_segs && p.seg_b0[lo + st] <= b) { lo += st; st <<= 1; }
            hi = lo + st < n_segs ? lo + st : n_segs;
        } else {
            hi = h;
            while (hi > st && p.seg_b0[hi - st] > b) { hi -= st; st <<= 1; }
            lo = hi > st ? hi - st : 0;
        }
        while (hi - lo > 1) {
            const u64 mid = lo + ((hi - lo) >> 1);
            if (p.seg_b0[mid] <= b) lo = mid; else hi = mid;
        }
    }
    const u64 j = lo;
    jb = j;
    u32 t_local = 0;
    if (ok) {
        const u64 t64 = b - p.seg_b0[j];
        if (t64 >= p.seg_count[j]) ok = false;
        t_local = (u32)t64;
    }
    // ---- the last checkpoint at or before header t_local (counts never decrease along a row; row[0] counts from 0),
    // the few headers from there to header b, then b's own (possibly ragged) block and the byte padding
    u64 eb = 0;
    if (ok) {
        const u64* row = p.ckpt + j * subs;
        u32 clo = 0, chi = subs;
        while (chi - clo > 1) {
            const u32 mid = clo + ((chi - clo) >> 1);
            if (ckpt_n(row[mid]) <= t_local) clo = mid; else chi = mid;
        }
        const u64 ck0 = row[clo];
        u64 pos = j * seg_bits + ckpt_rel(ck0);
        u32 s = ckpt_s(ck0);
        ok = ckpt_n(ck0) <= t_local;
        StreamFifo fifo;
        fifo.init(p.payload, n_words);
        for (u32 k = ckpt_n(ck0); ok && k < t_local; ++k) {
            if (pos >= total_bits) { ok = false; break; }
            const u32 hl = decode_header(fifo.peek(pos), s);
            pos += hl + (u64)s * p.block;
        }
        if (ok && pos < total_bits) {
            const u32 hl = decode_header(fifo.peek(pos), s);
            pos += hl + (u64)s * p.last_cnt;
            eb = 1 + (pos >> 3);
            if (eb > p.payload_bytes) ok = false;
        } else {
            ok = false;
        }
    }
    if (ok) endb = eb;
    u64 cap_pos = eb * 8;
    u32 cap_s = 0;
    u64 cap_c = 0;
    spec_twalk<WARP>(p, n_segs, n_words, total_bits, ok, max_steps, next, cap_pos, cap_s, cap_c);
    cap_state = next == SPEC_INVALID && ok && cap_pos != SPEC_INVALID ? cap_pos : SPEC_INVALID;
    cap_sc = (u64)cap_s | (cap_c << 8);
}

// T from bit `pos` (carried width s, c headers of the frame behind it) until it meets G -> next = G's number of the
// frame's last header.  Out of steps: (pos, s, c) say where to go on; pos = SPEC_INVALID: T cannot meet G at all.
template <bool WARP>
TRPX_DEVICE void spec_twalk(const DecParams& p, u64 n_segs, u64 n_words, u64 total_bits, bool ok, u64 max_steps, u64& next, u64& pos_io,
                            u32& s_io, u64& c_io)
{
    const u64 seg_bits = (u64)p.seg_bytes * 8;
    const u32 subs = p.subs_per_seg, sh = p.sub_shift;
    const u64* q64 = (const u64*)p.payload;
    const u64 n_q = (n_words + 1) >> 1;
    const u64 n_ck = n_segs * subs;
    // ---- This loop is what a batch costs, so it is kept lean:
    // positions relative to `base` in 32 bits, the stream in four 64-bit registers (three of them read ahead), a
    // branch-free header decode, and G consulted only where T enters a new sub-segment (its checkpoint names the
    // first header there; the checkpoint after it is read ahead too).
    const u64 start = ok ? pos_io : 0;
    const u64 base = start & ~255ull;                        // a multiple of 64 and of every checkpoint spacing
    const u64 bw = base >> 6;
    const u64 c_lim = p.nblocks - 1;
    u32 rel = (u32)(start - base), wrel = rel >> 6;
    u64 w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    u64 j2 = 0, seg_first = 0;
    if (ok) {
        w0 = bw + wrel < n_q ? q64[bw + wrel] : 0ull;
        w1 = bw + wrel + 1 < n_q ? q64[bw + wrel + 1] : 0ull;
        w2 = bw + wrel + 2 < n_q ? q64[bw + wrel + 2] : 0ull;
        w3 = bw + wrel + 3 < n_q ? q64[bw + wrel + 3] : 0ull;
        j2 = start / seg_bits;
        seg_first = j2 * seg_bits;
    }
    // at or beyond next_chk, the next header is the first of its sub-segment (or not, where a walk resumes: then the
    // comparison with that sub-segment's checkpoint just fails)
    u32 next_chk = 0;
    u64 q_pref = ~0ull, ck_pref = 0;
    u32 s = s_io;
    u64 c = c_io;
    bool active = ok;
    bool dead = false;
    for (u64 it = 0; it < max_steps; ++it) {
        if (active) {
            const u64 at = base + rel;
            // a frame too short to meet G / the stream ends: the serial chain's business
            if (c >= c_lim || at >= total_bits || rel >= (1u << 31)) { active = false; dead = true; }
            const u32 shb = rel & 63;
            const u64 w = (w0 >> shb) | ((w1 << 1) << (63 - shb));
            if (active && rel >= next_chk) {
                while (at >= seg_first + seg_bits) { ++j2; seg_first += seg_bits; }
                const u64 q = at >> sh;                      // (segments are whole numbers of sub-segments)
                next_chk = (u32)(((q + 1) << sh) - base);
                if (q < n_ck) {
                    const u64 ck = q == q_pref ? ck_pref : p.ckpt[q];
                    q_pref = q + 1;
                    ck_pref = q + 1 < n_ck ? p.ckpt[q + 1] : ~0ull;
                    if (seg_first + ckpt_rel(ck) == at && (ckpt_s(ck) == (s & 0xff) || (w & 1) == 0)) {
                        next = p.seg_b0[j2] + ckpt_n(ck) + (c_lim - c);
                        active = false;
                    }
                }
            }
            if (active) {
                u32 adv;
                if (s == 0 && (w & 1)) {                     // one-bit headers of empty blocks: up to the next checkpoint in one step
                    u64 run = (u64)ffs64(~w | (1ull << 63)) - 1;
                    if (run > next_chk - rel) run = next_chk - rel;
                    if (run > c_lim - c) run = c_lim - c;
                    adv = (u32)run;
                    c += run;
                } else {
                    u32 hl, sn;
                    decode_header_bf((u32)w, s, hl, sn);
                    s = sn;
                    adv = hl + s * p.block;
                    ++c;
                }
                rel += adv;
                const u32 dw = (rel >> 6) - wrel;
                if (dw == 1) {
                    w0 = w1; w1 = w2; w2 = w3;
                    w3 = bw + wrel + 4 < n_q ? q64[bw + wrel + 4] : 0ull;
                    wrel += 1;
                } else if (dw) {
                    wrel += dw;
                    const u64 a0 = bw + wrel;
                    if (dw == 2) { w0 = w2; w1 = w3; }
                    else { w0 = a0 < n_q ? q64[a0] : 0ull; w1 = a0 + 1 < n_q ? q64[a0 + 1] : 0ull; }
                    w2 = a0 + 2 < n_q ? q64[a0 + 2] : 0ull;
                    w3 = a0 + 3 < n_q ? q64[a0 + 3] : 0ull;
                }
            }
        }
        if (WARP ? ballot(active) == 0u : !active) break;
    }
    pos_io = active && !dead ? base + rel : SPEC_INVALID;
    s_io = s;
    c_io = c;
}

template <int NT>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(NT, 2) prolix_frame_spec_kernel(DecParams p, u64* frame_ends_out, u64* spec, u32 B, u32 R0, u32 RS, u32 max_steps)
{
    const u32 t = tid();
    const u64 gthreads = (u64)nblocks() * NT, gtid = (u64)bid() * NT + t;
    const u64 n_words = (p.payload_bytes + 3) >> 2;
    const u64 total_bits = p.payload_bytes * 8;
    const u64 n_segs = p.seg_base[1];                       // G: one "frame" = the whole payload
    const u64 total_blocks = n_segs ? p.seg_b0[n_segs - 1] + p.seg_count[n_segs - 1] : 0;
    u64* tab = spec + SPEC_HDR_WORDS;
    TRPX_SHARED u32 win_off[SPEC_B + 1];                    // first table entry of window k
    TRPX_SHARED u32 win_r[SPEC_B];
    TRPX_SHARED u32 follow_idx[SPEC_B];                     // table entry the chain went through, per frame of the batch
    TRPX_SHARED u32 follow_done;
    if (t == 0) {
        u32 o = 0;
        for (u32 k = 0; k < B; ++k) { win_off[k] = o; win_r[k] = spec_radius(k, R0, RS); o += 2 * win_r[k] + 1; }
        win_off[B] = o;
    }
    sync_block();
    if (gtid == 0) {
        // frame 0: T from bit 0 IS G, so the frame's last header is G's header nblocks - 1
        st_relaxed(spec + 0, 0ull);
        st_relaxed(spec + 1, p.nblocks - 1);
        st_relaxed(spec + 2, SPEC_INVALID);
        st_relaxed(spec + 3, 0ull);
        st_relaxed(spec + 4, (n_segs == 0 || p.nblocks < 2) ? 1ull : 0ull);
        st_relaxed(spec + 5, p.nblocks);
    }
    grid_sync();
#ifdef TRPX_SPEC_DEBUG
    u64 dbg_batches = 0, dbg_retries = 0, dbg_eval = 0, dbg_follow = 0, dbg_sync = 0, dbg_t0 = 0;
#endif
    for (;;) {
#ifdef TRPX_SPEC_DEBUG
        dbg_t0 = clock64();
#endif
        const u64 f = ld_relaxed(spec + 0), a = ld_relaxed(spec + 1), j_last = ld_relaxed(spec + 2), spf = ld_relaxed(spec + 3);
        const u64 stride = ld_relaxed(spec + 5);             // headers of G per frame, as the last batches saw it (>= nblocks)
        if (f >= p.n_frames || ld_relaxed(spec + 4)) break;
        const u32 kmax = p.n_frames - f < B ? (u32)(p.n_frames - f) : B;
        const u64 total = win_off[kmax];
        for (u64 i0 = (u64)bid() * NT + (t & ~31u); i0 < total; i0 += gthreads) {     // (whole warps: see spec_eval)
            const u64 i = i0 + (t & 31);
            const bool in_table = i < total;
            u32 klo = 0, khi = kmax;                         // window k holds entry i
            while (khi - klo > 1) {
                const u32 mid = (klo + khi) >> 1;
                if (win_off[mid] <= i) klo = mid; else khi = mid;
            }
            const u32 k = klo, r = win_r[k];
            const u64 centre = a + (u64)k * stride;
            const u64 d = i - win_off[k];                    // 0 .. 2 r
            u64 next, endb, jb, cap_pos, cap_sc;
            const bool valid = in_table && centre + d >= r;
            const u64 b = valid ? centre + d - r : 0;
            u64 hint;
            if (j_last != SPEC_INVALID) hint = j_last + (u64)(k + 1) * spf;
            else hint = total_blocks ? (u64)((double)b / (double)total_blocks * (double)n_segs) : 0;
            spec_eval<true>(p, n_segs, n_words, total_bits, total_blocks, valid, b, hint, max_steps, next, endb, jb, cap_pos, cap_sc);
            if (in_table) {
                u64* en = tab + SPEC_ENTRY_WORDS * i;
                en[0] = next;
                en[1] = endb;
                en[2] = jb;
                en[3] = cap_pos;
                en[4] = cap_sc;
            }
        }
#ifdef TRPX_SPEC_DEBUG
        if (gtid == 0) { dbg_eval += clock64() - dbg_t0; dbg_t0 = clock64(); }
#endif
        grid_sync();
#ifdef TRPX_SPEC_DEBUG
        if (gtid == 0) { dbg_sync += clock64() - dbg_t0; dbg_t0 = clock64(); ++dbg_batches; }
#endif
        if (bid() == 0) {
            // one thread follows e -> F(e) through the windows (one dependent load per frame); the frame ends are
            // then gathered by the CTA's threads in parallel
            if (t == 0) {
                u64 e = a, done = 0, stop = 0;
                for (u32 k = 0; k < kmax; ++k) {
                    const u32 r = win_r[k];
                    const u64 centre = a + (u64)k * stride;
#ifdef TRPX_EMU
                    if ((e + r < centre || e > centre + r) && getenv("EMU_TRACE_SPEC")) fprintf(stderr, "[spec]   miss at k=%u: e=%llu centre=%llu r=%u nblocks=%llu\n", k, (unsigned long long)e, (unsigned long long)centre, r, (unsigned long long)p.nblocks);
#endif
                    if (e + r < centre || e > centre + r) break;        // outside this batch's window: next batch
                    const u32 i = win_off[k] + (u32)(e + r - centre);
                    const u64* en = tab + SPEC_ENTRY_WORDS * (u64)i;
                    u64 next = ld_relaxed(en);
                    if (next == SPEC_INVALID) {
                        if (ld_relaxed(en + 1) == SPEC_INVALID) { stop = 1; break; }     // not even this frame's end
                        u64 pos = ld_relaxed(en + 3);
                        if (f + k + 1 < p.n_frames && pos != SPEC_INVALID) {            // T ran out of steps: walk it out here, from where it stopped
                            const u64 sc = ld_relaxed(en + 4);
                            u32 s_t = (u32)(sc & 0xff);
                            u64 c_t = sc >> 8;
#ifdef TRPX_SPEC_DEBUG
                            ++dbg_retries;
#endif
                            spec_twalk<false>(p, n_segs, n_words, total_bits, true, ~0ull, next, pos, s_t, c_t);
                        }
                    }
                    follow_idx[k] = i;
                    done = k + 1;
                    if (next == SPEC_INVALID) { stop = f + done < p.n_frames ? 1 : 0; break; }
                    e = next;
                }
                follow_done = (u32)done;
                const u64 j_first = done ? ld_relaxed(tab + SPEC_ENTRY_WORDS * (u64)follow_idx[0] + 2) : 0;
                const u64 j_now = done ? ld_relaxed(tab + SPEC_ENTRY_WORDS * (u64)follow_idx[done - 1] + 2) : j_last;
#ifdef TRPX_EMU
                emu_spec_followed() += done;
                if (getenv("EMU_TRACE_SPEC")) fprintf(stderr, "[spec] batch at frame %llu: %llu of %u followed%s\n", (unsigned long long)f, (unsigned long long)done, kmax, stop ? ", stop" : "");
#endif
                st_relaxed(spec + 0, f + done);
                st_relaxed(spec + 1, e);
                st_relaxed(spec + 2, j_now);
                st_relaxed(spec + 3, done > 1 ? (j_now - j_first) / (done - 1) : spf);
                if (done >= 2 && !stop && e > a) {           // e - a headers of G over `done` frames
                    const u64 seen = (e - a + done / 2) / done;
                    st_relaxed(spec + 5, done >= 8 ? seen : (stride + seen) / 2);
                }
                st_relaxed(spec + 4, stop);
#ifdef TRPX_SPEC_DEBUG
                dbg_follow += clock64() - dbg_t0;
#endif
            }
            sync_block();
            for (u32 k = t; k < follow_done; k += NT) frame_ends_out[f + k] = ld_relaxed(tab + SPEC_ENTRY_WORDS * (u64)follow_idx[k] + 1);
            sync_block();
        }
        grid_sync();
    }
#ifdef TRPX_SPEC_DEBUG
    if (gtid == 0)
        printf("[spec] frames %llu of %llu, batches %llu, retries %llu; thread 0 cycles: eval %llu, waiting at the barrier %llu, follow %llu\n",
               (unsigned long long)ld_relaxed(spec + 0), (unsigned long long)p.n_frames, (unsigned long long)dbg_batches, (unsigned long long)dbg_retries,
               (unsigned long long)dbg_eval, (unsigned long long)dbg_sync, (unsigned long long)dbg_follow);
#endif
}

// The chain itself is serial, but one warp walks it co-operatively: the stream is staged in 16 KB
// chunks (coalesced loads), and one step probes the next 32 candidate header positions at once --
// inside a run of '1' headers ("same width as before", 71 % of the headers of a diffraction frame; all
// of them in empty regions) block k starts exactly k * (1 + block * s) bits further on, so lane k
// tests that bit and a ballot finds the first explicit header: a whole run costs one step.
constexpr u32 FF_CHUNK_WORDS = 4096;
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(32, 1) prolix_find_frames_kernel(DecParams p, u64* frame_ends_out)
{
    TRPX_SHARED u32 chunk[FF_CHUNK_WORDS + 4];
    if (bid() != 0) return;
    const u32 lane = tid() & 31;
    const u64 n_words = (p.payload_bytes + 3) >> 2;
    const u64 total_bits = p.payload_bytes * 8;
    const u64 max_stride = 1 + (u64)p.block * 73;
    const bool fits = 34 * max_stride + 64 < (u64)FF_CHUNK_WORDS * 32;     // 32 probes + one block always inside a chunk
    u64 P = 0;                                                    // absolute bit of the next header
    u64 chunk_bit = ~0ull;                                        // absolute bit of chunk[0] (multiple of 128)
    for (u64 f = 0; f < p.n_frames; ++f) {
        const u64 frame_bit = P;
        u32 s = 0;
        u64 left = p.nblocks;
        bool bad = false;
        while (left && !bad) {
            if (P >= total_bits) { bad = true; break; }
            if (!fits) {                                          // huge blocks: plain serial step from global memory
                const u64 win = peek_bits(p.payload, n_words, P);
                const u32 hl = decode_header(win, s);
                P += hl + (u64)s * (left == 1 ? p.last_cnt : p.block);
                --left;
                continue;
            }
            if (chunk_bit == ~0ull || P < chunk_bit || P - chunk_bit + 34 * max_stride + 64 > (u64)FF_CHUNK_WORDS * 32) {
                sync_warp();
                chunk_bit = P & ~127ull;
                const u64 w0 = chunk_bit >> 5;
                for (u32 i = lane; i < FF_CHUNK_WORDS + 4; i += 32) chunk[i] = w0 + i < n_words ? p.payload[w0 + i] : 0u;
                sync_warp();
            }
            const u32 q = (u32)(P - chunk_bit);
            const u32 stride = 1 + p.block * s;
            // lane k: is block k of a run of '1' headers still a '1' header?
            const u32 qk = q + lane * stride;
            const bool valid = (u64)lane < left;
            const u32 bit = (chunk[qk >> 5] >> (qk & 31)) & 1;
            const u32 expl = ballot(valid && bit == 0);
            const u32 nvalid = left < 32 ? (u32)left : 32u;
            const u32 j = expl ? (u32)ffs32(expl) - 1 : nvalid;    // leading '1' headers
            if (j) {
                const bool ends = (u64)j == left;                  // the run reaches the frame's last (possibly ragged) block
                P += (u64)(ends ? j - 1 : j) * stride + (ends ? 1 + (u64)s * p.last_cnt : 0);
                left -= j;
            }
            if (j < nvalid) {                                      // an explicit header at P
                const u32 q2 = (u32)(P - chunk_bit);
                const u32 win = funnel_r(chunk[q2 >> 5], chunk[(q2 >> 5) + 1], q2 & 31);
                const u32 hl = decode_header((u64)win, s);
                P += hl + (u64)s * (left == 1 ? p.last_cnt : p.block);
                --left;
            }
        }
        u64 end = (frame_bit >> 3) + 1 + ((P - frame_bit) >> 3);   // 1 + floor(bits / 8) bytes (Terse.hpp:547)
        if (bad || end > p.payload_bytes) { if (lane == 0) atomic_max(p.status, DEC_MALFORMED); end = p.payload_bytes; }
        if (lane == 0) frame_ends_out[f] = end;
        P = end * 8;
    }
}

}  // namespace trpx
