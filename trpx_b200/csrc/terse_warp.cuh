// terse_warp.cuh -- TERSE encoder fast path, round-2 design: WARP-AUTONOMOUS tiles, no CTA-wide barrier.
//
// Replaces Terse::f_compress (reference include/Terse.hpp:500-549), as terse_encode.cuh describes; this file holds
// the kernel that runs when block == 12 and frames are 16-byte aligned.  What changed against round 1 and why
// (profiles/r01_ncu_full_summary.txt): the round-1 kernel synchronised 6-8 worker warps four times per tile and spent
// 60 % of the ALU pipe's issue slots (LOP3 / SHF / IADD3 / SEL / ISETP run at half rate on sm_100) -- it was bound by
// barriers and ALU instructions, not by HBM.  Here:
//   * a WARP owns a tile (32 lanes x one unit of pixels): its own TMA stage and mbarrier, its own staging ring, an
//     in-warp shuffle scan.  Warps never wait for each other with a barrier; whatever a warp needs from somebody else
//     (a ticket, its stream position) was produced rounds earlier and is picked up from shared memory.
//   * a CTA is NW such warps working on NW CONSECUTIVE tiles (one "supertile", taken by ticket so that supertiles are
//     started in stream order), plus resolver warps that run the two-level decoupled look-back of terse_encode.cuh
//     once per supertile and hand every warp its start position.
//   * the bit sink keeps no "first word" state: a lane's first word is stored like any other and the bits of the
//     lanes before it are OR-ed in afterwards (one shuffle prefix per tile); left shifts are multiplications by a
//     power of two, which run on the otherwise idle FMA pipe.
#pragma once

#include "terse_encode.cuh"

#ifndef ENC_NW
#define ENC_NW 16
#endif

#ifdef TRPX_EMU
#define WTRACE(what) do { if (getenv("EMU_TRACE2") && lane == 0 && bid() == 0 && warp == 1) fprintf(stderr, "emu: w1 round %u oldest %u: %s\n", r, oldest, what); } while (0)
#else
#define WTRACE(what) do { } while (0)
#endif

namespace trpx {

#ifndef ENCW_NRES_N
#define ENCW_NRES_N 2
#endif
constexpr int ENCW_NRES = ENCW_NRES_N;   // resolver warps; resolver q serves rounds r % ENCW_NRES == q
constexpr int ENCW_D = 4;           // rounds a CTA may have in flight (slots of posted / resolved / bits / p0)
constexpr int ENCW_TD = 2 * ENCW_D; // ticket slots (workers drift by at most ENCW_D - 1 rounds)
constexpr int ENCW_MAXW = 24;       // layout bound on NW

// shared-memory header (bytes)
constexpr int WSM_FULL = 0;                                  // mbarrier full[NW]: the warp's TMA stage has landed
constexpr int WSM_POSTED = WSM_FULL + 8 * ENCW_MAXW;         // mbarrier posted[D]: all NW warps have posted their bit counts
constexpr int WSM_RESOLVED = WSM_POSTED + 8 * ENCW_D;        // mbarrier resolved[D]: start positions are in p0[]
constexpr int WSM_TICKETS = WSM_RESOLVED + 8 * ENCW_D;       // u64 tickets[TD]: (round + 1) << 32 | supertile
constexpr int WSM_MAX = WSM_TICKETS + 8 * ENCW_TD;           // u32 running max width
constexpr int WSM_REQ = WSM_MAX + 4;                         // u32: the highest round whose ticket has been requested
constexpr int WSM_BITS = 384;                                // u32 bits[D][NW]
constexpr int WSM_P0 = WSM_BITS + 4 * ENCW_D * ENCW_MAXW;    // u64 p0[D][NW]
constexpr int WSM_PEND = WSM_P0 + 8 * ENCW_D * ENCW_MAXW;    // u32 pend[NW][D][2]: virtual ring offset, supertile
constexpr int WSM_HEADER = 2304;
static_assert(WSM_REQ + 4 <= WSM_BITS && WSM_PEND + 8 * ENCW_D * ENCW_MAXW <= WSM_HEADER, "shared-memory header layout");

template <typename T, int NW>
struct WGeom {
    typedef Pix<T> P;
    static constexpr int WT_BYTES = 32 * P::UNIT_BYTES;                 // pixels of one warp tile
    static constexpr int WT_BLOCKS = 32 * P::BPU;
    static constexpr int HALO = (12 * P::SZ + 15) / 16 * 16;            // the block before the tile (its width decides the first header)
    static constexpr int STAGE_BYTES = (HALO + WT_BYTES + 127) / 128 * 128;
    static constexpr int WORST_WORDS = ((WT_BLOCKS * P::MAXBITS + 31) / 32 + 1 + 3 + 3) / 4 * 4;
    static constexpr int RING_WORDS = WORST_WORDS <= 1024 ? 1024 : WORST_WORDS <= 2048 ? 2048 : 4096;   // a power of two
    static constexpr int ST_BLOCKS = NW * WT_BLOCKS;                    // blocks of a supertile
    static constexpr int THREADS = 32 * (NW + ENCW_NRES);
    static constexpr int SMEM_BYTES = WSM_HEADER + NW * (STAGE_BYTES + RING_WORDS * 4);
    static_assert(NW <= ENCW_MAXW && WORST_WORDS <= 4096, "geometry");
};

// ------------------------------------------------------------------ K4: per-lane bit sink (warp-private ring)
// Appends fields LSB-first (Bit_pointer.hpp:700-730).  `lo` is the unfinished word (nb < 32 valid bits); every
// completed word is stored at once -- also the lane's first one, whose low bits belong to the lanes before it and
// are OR-ed in by warp_merge() afterwards.
struct WSink {
    saddr_t wp;             // shared-window address of the unfinished word
    u32 lo, nb;
    TRPX_DEVICE void init(saddr_t ring_a, u32 off) { wp = ring_a + (off >> 5) * 4; nb = off & 31; lo = 0; }
    TRPX_DEVICE void put(u32 v, u32 n)                       // n in [0, 32], v < 2^n
    {
        const u32 pw = 1u << nb;
        const u32 a0 = mad_lo(v, pw, lo);                       // lo | v << nb: disjoint bits, so OR == ADD (FMA pipe)
        const u32 a1 = funnel_l(v, 0u, nb);                     // v >> (32 - nb); 0 when nb == 0
        const u32 t = nb + n;
        const bool c1 = t >= 32;
        sts_u32_if(c1, wp, a0);
        lo = c1 ? a1 : a0;
        wp += c1 ? 4u : 0u;
        nb = t & 31;
    }
    TRPX_DEVICE void put64(u32 v0, u32 v1, u32 n)            // n in [0, 64], (v1:v0) < 2^n; up to two words complete
    {
        const u32 pw = 1u << nb;
        const u32 a0 = mad_lo(v0, pw, lo);
        const u32 a1 = funnel_l(v0, v1, nb);
        const u32 a2 = funnel_l(v1, 0u, nb);
        const u32 t = nb + n;                                   // < 96
        const bool c1 = t >= 32, c2 = t >= 64;
        sts_u32_if(c1, wp, a0);
        sts_u32_if(c2, wp + 4, a1);
        lo = c2 ? a2 : (c1 ? a1 : a0);
        wp += (t >> 3) & 12u;                                   // 4 bytes per completed word
        nb = t & 31;
    }
    TRPX_DEVICE void put_wide(u64 v, u32 s)                  // low s bits of the sign-extended value, s in [1, 65]
    {
        const u32 n0 = s < 32 ? s : 32;
        u32 x = (u32)v;
        if (n0 < 32) x &= (1u << n0) - 1;
        put(x, n0);
        if (s > 32) {
            const u32 n1 = s - 32 < 32 ? s - 32 : 32;
            u32 hi = (u32)(v >> 32);
            if (n1 < 32) hi &= (1u << n1) - 1;
            put(hi, n1);
        }
        if (s > 64) put((u32)(v >> 63) & 1u, 1);
    }
};

// header + data of one block (Terse.hpp:517-541).  Full blocks of the usual widths go as a few 64-bit fields built
// in registers (all fields of a block share s); the header rides in the first field whenever it fits.
template <typename T>
TRPX_DEVICE void pack_block_w(WSink& sk, const u32* w, u32 s, u32 cnt, u32 hv, u32 hl)
{
    typedef Pix<T> P;
    if (cnt == 12 && s != 0) {
        if (P::SZ == 2 && s <= 16) {
            // two 16-bit halves -> one field of 2s bits with a single multiply-add: lo + hi*2^16 + hi*(2^s - 2^16)
            const u32 m = ((1u << s) - 1) * 0x00010001u;                 // per-half mask (signed pixels only)
            const u32 K = (1u << s) - 65536u;
            const u32 s2 = 2 * s, pw2 = s2 < 32 ? 1u << s2 : 0u;
            u32 q0[3], q1[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const u32 x0 = P::SGN ? (w[2 * q] & m) : w[2 * q], x1 = P::SGN ? (w[2 * q + 1] & m) : w[2 * q + 1];
                const u32 p0 = mad_lo(x0 >> 16, K, x0), p1 = mad_lo(x1 >> 16, K, x1);
                q0[q] = mad_lo(p1, pw2, p0);                             // p0 | p1 << 2s (low word)
                q1[q] = funnel_l(p1, 0u, s2);                            // p1 >> (32 - 2s); s2 == 32 -> p1 (shift taken mod 32 ...
                if (s2 == 32) q1[q] = p1;                                // ... which would give 0)
            }
            if (hl + 2 * s2 <= 64) {                                     // header + first four values in one field
                const u32 f0 = mad_lo(q0[0], 1u << hl, hv), f1 = funnel_l(q0[0], q1[0], hl);
                sk.put64(f0, f1, hl + 2 * s2);
            } else {
                sk.put(hv, hl);
                sk.put64(q0[0], q1[0], 2 * s2);
            }
            sk.put64(q0[1], q1[1], 2 * s2);
            sk.put64(q0[2], q1[2], 2 * s2);
            return;
        }
        if (P::SZ == 1 && s <= 8) {
            const u32 m = P::SGN ? ((1u << s) - 1) * 0x00010001u : 0x00ff00ffu;
            u32 q[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const u32 a = w[i] & 0x00ff00ffu & m, b = (w[i] >> 8) & 0x00ff00ffu & m;
                const u32 pr = a | (b << s);                             // two 16-bit lanes of 2s bits
                q[i] = s == 8 ? pr : ((pr & 0xffffu) | ((pr >> 16) << (2 * s)));   // 4s bits
            }
            const u32 s4 = 4 * s;
            // header + first quad (<= 12 + 32 bits), then the other two quads (<= 64 bits)
            sk.put64(mad_lo(q[0], 1u << hl, hv), funnel_l(q[0], 0u, hl), hl + s4);
            const u32 g0 = s4 < 32 ? mad_lo(q[2], 1u << s4, q[1]) : q[1];
            const u32 g1 = s4 < 32 ? funnel_l(q[2], 0u, s4) : q[2];
            sk.put64(g0, g1, 2 * s4);
            return;
        }
        if (P::SZ == 4 && s <= 32) {
            const u32 m = s == 32 ? 0xffffffffu : (1u << s) - 1;
            const u32 pws = s < 32 ? 1u << s : 0u;
            sk.put(hv, hl);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const u32 x0 = w[2 * i] & m, x1 = w[2 * i + 1] & m;
                sk.put64(mad_lo(x1, pws, x0), s == 32 ? x1 : funnel_l(x1, 0u, s), 2 * s);
            }
            return;
        }
    }
    sk.put(hv, hl);
    if (s == 0) return;
#pragma unroll
    for (int i = 0; i < 12; ++i)
        if ((u32)i < cnt) sk.put_wide(block_value<T>(w, i), s);
}

// After the lanes of a warp have packed their units: OR the bits each lane left unfinished into the word of the
// lane that completed it (OR of disjoint fields == ADD, so the bits carried into lane t's first word are a difference
// of two warp prefix sums), store the tile's last partial word and two zero words behind it.  w0: index of the
// lane's first word; crossed: the lane completed at least one word.
TRPX_DEVICE void warp_merge(u32* stg, const WSink& sk, u32 w0, bool crossed, u32 tile_bits)
{
    const u32 lane = tid() & 31;
    const u32 tail = sk.lo;
    u32 incl = tail;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 v = shfl_up(incl, d);
        if (lane >= (u32)d) incl += v;
    }
    const u32 excl = incl - tail;
    const u32 cmask = ballot(crossed);
    const u32 below = cmask & ((1u << lane) - 1);
    const u32 qp = shfl(excl, below ? 31 - clz32(below) : 0);
    const u32 carry = excl - (below ? qp : 0u);
    if (crossed && carry) stg[w0] |= carry;                  // (only this lane touches word w0 now: see WSink)
    const u32 ql = shfl(excl, cmask ? 31 - clz32(cmask) : 0);
    if (lane == 31) {
        u32* e = stg + (tile_bits >> 5);
        e[0] = incl - (cmask ? ql : 0u);                     // the unfinished last word (zero bits above the data)
        e[1] = 0; e[2] = 0;                                  // a frame end may push the output window past it
    }
}

// ------------------------------------------------------------------ drain: ring -> payload
// The same hand-off as terse_encode.cuh's tail_handoff / store_tile, by ONE warp for its own tile `wt` (global
// index of the warp tile).  The word two neighbouring tiles share is stored by the later one.
TRPX_DEVICE void warp_drain(const EncParams& p, u64 wt, bool last_of_all, bool ends_frame, u64 frame, const u32* stg,
                            u32 tile_bits, u64 P0, u64 Pn, u64 tin_known)
{
    const u32 lane = tid() & 31;
    u32 tin = 0, tout = 0;
    if (lane == 0) tin = tail_handoff(p, wt, stg, tile_bits, P0, Pn, tout, tin_known);
    tin = shfl(tin, 0);
    tout = shfl(tout, 0);
    const u32 nstg = (tile_bits + 31) >> 5;
    const u64 W0 = P0 >> 5, Wn = Pn >> 5;
    const u32 sh = (u32)(P0 & 31);
    const u32 k = (u32)(Wn - W0);
    const bool fits = ((Pn + 7) >> 3) <= p.out_capacity;
    if (fits) {
        u32* outw = p.out_words + W0;
        if (lane == 0 && k) st_stream(outw, window_word(stg, nstg, 0, sh) | tin);   // the word shared with the previous tile
        for (u32 i = lane ? lane : 32u; i < k; i += 32) st_stream(outw + i, window_word(stg, nstg, i, sh));
    } else if (lane == 0) {
        atomic_max(p.status, 2u);                            // TRPX_ERR_CAPACITY
    }
    if (lane == 0) {
        if (ends_frame) p.frame_ends[frame] = Pn >> 3;
        if (last_of_all && fits) {                           // nobody follows: store the final bytes
            unsigned char* ob = (unsigned char*)p.out_words;
            for (u32 b = 0; b < (u32)((Pn >> 3) & 3); ++b) ob[Wn * 4 + b] = (unsigned char)(tout >> (8 * b));
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <typename T, int NW>
TRPX_KERNEL void TRPX_LAUNCH_BOUNDS(32 * (NW + ENCW_NRES), 1) terse_encode_warp_kernel(EncParams p)
{
    typedef Pix<T> P;
    typedef WGeom<T, NW> G;
    TRPX_DYN_SMEM(sm);
    u64* bar_full = (u64*)(sm + WSM_FULL);
    u64* bar_posted = (u64*)(sm + WSM_POSTED);
    u64* bar_resolved = (u64*)(sm + WSM_RESOLVED);
    volatile u64* tickets = (volatile u64*)(sm + WSM_TICKETS);
    u32* sm_max = (u32*)(sm + WSM_MAX);
    volatile u32* bits_s = (volatile u32*)(sm + WSM_BITS);
    volatile u64* p0_s = (volatile u64*)(sm + WSM_P0);
    const u32 t = tid(), lane = t & 31, warp = t >> 5;
    const u64 frame_bytes = p.n_values * P::SZ;
    const u32 spf = (u32)p.tiles_per_frame;                   // supertiles per frame
    const u32 wtpf = (u32)div_up(frame_bytes, (u64)G::WT_BYTES);   // warp tiles (with data) per frame
    const u64 n_super = p.n_tiles;

    if (t == 0) {
        for (int w = 0; w < NW; ++w) mbar_init(&bar_full[w], 1);
        for (int e = 0; e < ENCW_D; ++e) { mbar_init(&bar_posted[e], NW); mbar_init(&bar_resolved[e], 1); }
        for (int i = 0; i < ENCW_TD; ++i) tickets[i] = 0;
        for (u32 q = 0; q < 2; ++q) tickets[q] = ((u64)(q + 1) << 32) | atomic_add(p.ticket, 1u);
        *(u32*)(sm + WSM_REQ) = 1;
        mbar_init_fence();
        *sm_max = 0;
    }
    sync_block();

    // the supertile of round r (all lanes).  Tickets are taken just in time, by whichever worker warp reaches round
    // r - 2 first (no warp's progress depends on a particular other warp), one atomic per supertile, so that supertiles
    // are started in stream order.
    auto fetch_ticket = [&](u32 r) -> u32 {
        u64 v = 0;
        if (lane == 0) {
            WaitClock wc;
            for (u32 spins = 0;; ++spins) {
                v = tickets[r % ENCW_TD];
                if ((u32)(v >> 32) == r + 1) break;
                if (wc.expired(p.status, spins, 23, 4 | (warp << 4) | (r << 12))) { v = 0xffffffffu; break; }   // gives up: "nothing follows"
                spin_hint();
            }
        }
        return shfl((u32)v, 0);
    };

    if (warp >= (u32)NW) {
        // ================================================================ resolver warp q
        for (u32 r = warp - NW;; r += ENCW_NRES) {
            const u32 st = fetch_ticket(r);
            if ((u64)st >= n_super) break;
            const u32 e = r % ENCW_D;
            mbar_wait_sleep(&bar_posted[e], (r / ENCW_D) & 1, p.status, 5 | (warp << 4) | (r << 12));
            const u32 b = lane < (u32)NW ? bits_s[e * NW + lane] : 0u;
            u32 incl = b;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 v = shfl_up(incl, d);
                if (lane >= (u32)d) incl += v;
            }
            const u32 total = shfl(incl, 31);
            const TileGeom g = tile_geom(p, st);
            const u64 P0 = tile_start(p, st, g, total, true);
            // warps past the frame's last warp tile have no data: they sit at the (byte-aligned) end of the frame
            const u32 wl = g.ends ? wtpf - 1 - (u32)g.tif * NW : (u32)NW - 1;
            const u64 mine = lane <= wl ? P0 + (incl - b) : align_frame(P0 + total);
            if (lane < (u32)NW) p0_s[e * NW + lane] = mine;
            sync_warp();
            if (lane == 0) mbar_arrive(&bar_resolved[e]);
        }
        return;
    }

    // ==================================================================== worker warp
    unsigned char* stage = sm + WSM_HEADER + warp * (G::STAGE_BYTES + G::RING_WORDS * 4);
    u32* ring = (u32*)(stage + G::STAGE_BYTES);
    volatile u32* pend = (volatile u32*)(sm + WSM_PEND) + warp * (2 * ENCW_D);
    u32* req = (u32*)(sm + WSM_REQ);
    const u32 ring_words = p.dbg_ring_words ? p.dbg_ring_words : (u32)G::RING_WORDS;   // a power of two (tests shrink it)
    const u32 ring_mask = ring_words - 1;

    // lane 0: start the bulk copy of this warp's tile of supertile `st` (nothing to load past the frame's end)
    auto issue = [&](u32 st) {
        const u32 f = st / spf, sif = st - f * spf;
        const u32 wt = sif * NW + warp;
        if (wt < wtpf) {
            const u64 off = (u64)wt * G::WT_BYTES;
            u64 bytes = frame_bytes - off;
            if (bytes > (u64)G::WT_BYTES) bytes = G::WT_BYTES;
            const unsigned char* src = (const unsigned char*)p.pixels + (u64)f * frame_bytes + off;
            unsigned char* dst = stage + G::HALO;
            if (wt > 0) { src -= G::HALO; dst -= G::HALO; bytes += G::HALO; }
            mbar_arrive_expect_tx(&bar_full[warp], (u32)bytes);
            bulk_g2s(dst, src, (u32)bytes, &bar_full[warp]);
        }
    };
    // the tile this warp packed in round j, once the resolver has posted its start position
    struct Waiting { u64 P0, Pn, gwt; const u32* stg; u32 bits, frame; bool ends; };
    auto waiting = [&](u32 j) -> Waiting {
        const u32 e = j % ENCW_D;
        Waiting q;
        q.P0 = p0_s[e * NW + warp];
        q.bits = bits_s[e * NW + warp];
        const u32 vbase = pend[2 * e], st = pend[2 * e + 1];
        q.frame = st / spf;
        const u32 wt = (st - q.frame * spf) * NW + warp;
        q.ends = wt + 1 == wtpf;
        q.Pn = q.ends ? align_frame(q.P0 + q.bits) : q.P0 + q.bits;
        q.gwt = (u64)st * NW + warp;
        q.stg = ring + (vbase & ring_mask) + 1;
        return q;
    };
    // lane 0: hand our bits of the word the NEXT tile starts in over to it (tails[]), as soon as our position is known
    // -- nobody's store ever waits for more than that one word (a tile that owns no complete word passes its
    // predecessor's bits on and can only do so when it is stored)
    auto publish = [&](u32 j) {
        const Waiting q = waiting(j);
        const u32 k = (u32)((q.Pn >> 5) - (q.P0 >> 5));
        if (k >= 1) {
            const u32 tout = window_word(q.stg, 0, k, (u32)(q.P0 & 31)) & ((1u << (q.Pn & 31)) - 1);
            st_relaxed(&p.tails[q.gwt], TAIL_VALID | (u64)tout);
        }
    };
    // store the tile packed in round j (blocks until its start position and the boundary word before it are known)
    auto drain = [&](u32 j, u64 tin_known) {
        mbar_wait(&bar_resolved[j % ENCW_D], (j / ENCW_D) & 1, p.status, 6 | (warp << 4) | (j << 12));
        const Waiting q = waiting(j);
        warp_drain(p, q.gwt, q.gwt + 1 == n_super * NW, q.ends, q.frame, q.stg, q.bits, q.P0, q.Pn, tin_known);
    };

    u32 cur = fetch_ticket(0);
    if (lane == 0 && (u64)cur < n_super) issue(cur);
    u32 loads = 0, my_max = 0;
    u32 vhead = 0, vtail = 0;                                // ring: the waiting tiles occupy [vtail, vhead)
    u32 oldest = 0, pub = 0;                                 // rounds [oldest, r) wait to be stored; rounds < pub have handed their boundary word over
    u32 r = 0;
    for (; (u64)cur < n_super; ++r) {
        const u32 e = r % ENCW_D;
        // ---- (lane 0) the first warp to reach round r takes the ticket of round r + 2
        u32 tk_first = 0, tk_last = 0, tk_val = 0;
        if (lane == 0) {
            const u32 old = atomic_max(req, r + 2);
            if (old < r + 2) { tk_first = old + 1; tk_last = r + 2; tk_val = atomic_add(p.ticket, tk_last - tk_first + 1); }
        }
        const u32 f = cur / spf, sif = cur - f * spf;
        const u32 wt = sif * NW + warp;
        const bool have = wt < wtpf;
        WTRACE("start");

        // ---- this lane's unit -> registers
        u32 w[P::UW];
        u32 nvalid = 0;
        const unsigned char* tile_sm = stage + G::HALO;
        if (have) {
            mbar_wait(&bar_full[warp], loads & 1, p.status, 7 | (warp << 4) | (r << 12));
            ++loads;
            const uint4* src = (const uint4*)(tile_sm + (size_t)lane * P::UNIT_BYTES);
#pragma unroll
            for (int j = 0; j < P::UW / 4; ++j) {
                const uint4 v = src[j];
                w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
            }
            nvalid = P::VPU;
            if (wt + 1 == wtpf) {                              // the frame's last tile may be ragged
                const u64 tile_vals = p.n_values - (u64)wt * (G::WT_BLOCKS * 12);
                const u64 my_first = (u64)lane * P::VPU;
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
        } else {
#pragma unroll
            for (int j = 0; j < P::UW; ++j) w[j] = 0;
        }
        u32 prev0 = 0;                                       // width of the block before the tile
        if (lane == 0 && have && wt > 0) {
            u32 h[P::BW];
            const u32* hs = (const u32*)(tile_sm - 4 * P::BW);
#pragma unroll
            for (int j = 0; j < P::BW; ++j) h[j] = hs[j];
            prev0 = block_width12<T>(h);
        }
        // (lane 0) the boundary word of the tile before our oldest waiting one: in flight until the end of the round
        u64 tin_pref = 0;
        const u32 pref_round = oldest;                       // (a forced store further down may move `oldest` on)
        if (lane == 0 && oldest < r) {
            const u64 og = (u64)pend[2 * (oldest % ENCW_D) + 1] * NW + warp;
            if (og) tin_pref = ld_relaxed(&p.tails[og - 1]);
        }
        sync_warp();                                         // every lane has read the stage: it may be refilled
        WTRACE("loaded");
        if (lane == 0)                                       // post the ticket(s) taken above (more than one only after a give-up)
            for (u32 q = tk_first; q != 0 && q <= tk_last; ++q) tickets[q % ENCW_TD] = ((u64)(q + 1) << 32) | (tk_val + (q - tk_first));
        const u32 nxt = fetch_ticket(r + 1);
        WTRACE("ticket");
        if (lane == 0 && (u64)nxt < n_super) issue(nxt);

        // ---- K1 widths, K2 headers and lengths
        u32 sb[P::BPU], cnt[P::BPU];
#pragma unroll
        for (int b = 0; b < P::BPU; ++b) {
            sb[b] = block_width12<T>(&w[b * P::BW]);
            const u32 v0 = 12u * b;
            cnt[b] = nvalid <= v0 ? 0u : (nvalid - v0 > 12u ? 12u : nvalid - v0);
            my_max = sb[b] > my_max ? sb[b] : my_max;
        }
        WTRACE("widths");
        u32 prev = shfl_up(sb[P::BPU - 1], 1);
        if (lane == 0) prev = prev0;
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
        // ---- K3: offsets inside the tile (in-warp scan)
        u32 incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 v = shfl_up(incl, d);
            if (lane >= (u32)d) incl += v;
        }
        const u32 off = incl - len;
        const u32 tile_bits = shfl(incl, 31);
        WTRACE("scanned");

        // ---- room in the ring (physically contiguous, one zero word in front); the slot of round r - D must be free
        const u32 need = ((tile_bits >> 5) + 1 + 3 + 3) & ~3u;
        u32 vbase = vhead;
        if ((vbase & ring_mask) + need > ring_words) vbase += ring_words - (vbase & ring_mask);
        while (r - oldest == (u32)ENCW_D || (r > oldest && vbase + need - vtail > ring_words)) {
            drain(oldest, 0);
            ++oldest;
            vtail = oldest < r ? pend[2 * (oldest % ENCW_D)] : vbase;
        }
        if (pub < oldest) pub = oldest;
        if (r == oldest) vtail = vbase;
        WTRACE("room");
        sync_warp();                                         // (the drain's readers are done before the ring is rewritten)
        // publish the tile's bit count: the resolver sums the supertile and runs the look-back
        if (lane == 0) {
            bits_s[e * NW + warp] = tile_bits;
            pend[2 * e] = vbase;
            pend[2 * e + 1] = cur;
            mbar_arrive(&bar_posted[e]);
        }
        u32* stg = ring + (vbase & ring_mask) + 1;
        if (lane == 0) stg[-1] = 0;
        vhead = vbase + need;

        // ---- K4: pack into tile-relative staging
        WSink sk;
        sk.init(saddr(stg), off);
        const saddr_t wp0 = sk.wp;
#pragma unroll
        for (int b = 0; b < P::BPU; ++b)
            if (cnt[b]) pack_block_w<T>(sk, &w[b * P::BW], sb[b], cnt[b], hv[b], hl[b]);
        sync_warp();                                         // all plain stores of the warp before the merge's read-modify-write
        warp_merge(stg, sk, off >> 5, sk.wp != wp0, tile_bits);
        sync_warp();
        WTRACE("packed");

        // ---- whatever has been resolved meanwhile hands its boundary word over (never waits; lane 0 decides for the
        // warp: lanes that poll at different moments must not disagree) ...
        while (pub <= r) {
            u32 ok = 0;
            if (lane == 0 && mbar_test(&bar_resolved[pub % ENCW_D], (pub / ENCW_D) & 1)) { publish(pub); ok = 1; }
            if (!shfl(ok, 0)) break;
            ++pub;
        }
        // ... and the oldest waiting tile is stored if that takes no waiting at all: its position is known AND the tile
        // before it has handed its word over (read at the top of the round).  An early store must never block: a warp
        // waiting here for a tile of another CTA could not be relied on by the warps that wait for ITS earlier tiles.
#ifndef ENCW_NO_EARLY_DRAIN
        if (oldest < r && oldest == pref_round && oldest < pub) {
            u32 ready = 0;
            if (lane == 0) ready = (p0_s[(oldest % ENCW_D) * NW + warp] & 31) == 0 || (tin_pref & TAIL_VALID) != 0;
            if (shfl(ready, 0)) {
                drain(oldest, tin_pref);
                ++oldest;
                vtail = oldest <= r ? pend[2 * (oldest % ENCW_D)] : vhead;
            }
        }
#endif
        WTRACE("end");
        cur = nxt;
    }
    for (; oldest < r; ++oldest) drain(oldest, 0);
    my_max = warp_max(my_max);
    if (lane == 0 && my_max) atomic_max(sm_max, my_max);
    bar_sync(1, 32 * NW);
    if (t == 0 && *sm_max) atomic_max(p.prolix_bits, *sm_max);
}

}  // namespace trpx
