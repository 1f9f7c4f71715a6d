// emu.hpp -- TEST INFRASTRUCTURE ONLY: a tiny cooperative-fiber SIMT emulator.
//
// It lets the kernel sources under trpx_b200/csrc/ be compiled with plain g++ (-DTRPX_EMU) and run
// on the CPU so that `pytest -m "not gpu"` can check their index arithmetic, scans, look-back and
// bit packing against the oracle.  One ucontext fiber per CUDA thread; the blocks of a grid run one
// after another (so a look-back never has to wait), barriers and warp collectives are rendez-vous
// points between fibers.  It is never part of libtrpx_b200.so.
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

namespace emu {

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };

struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct alignas(8) uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

struct ThreadCtx { uint32_t tid, bid, block_dim, grid_dim; };

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    ThreadCtx tc{};
    int slot = 0;                         // which resident CTA the fiber belongs to
};

// One resident CTA: its barriers, warp exchange slots and shared memory.  Several CTAs of a grid are resident at
// once (EMU_CTAS, default 1) and their fibers are interleaved, so that look-backs, tickets and hand-offs BETWEEN
// CTAs run concurrently as on the device.
struct Block {
    uint32_t bar_count = 0, bar_gen = 0;
    uint32_t nbar_count[16] = {0}, nbar_gen[16] = {0};   // named barriers (bar.sync id, n)
    std::vector<uint32_t> wbar_count, wbar_gen;
    std::vector<uint64_t> xch;            // 32 slots per warp
    unsigned char* smem = nullptr;
    uint32_t alive = 0;
    bool active = false;
};

struct State {
    std::vector<Fiber> fibers;            // slot * block_dim + tid
    std::vector<Block> blocks;
    ucontext_t sched;
    int current = -1;
    const std::function<void()>* body = nullptr;
    uint64_t yields_without_progress = 0;
};

inline State& S() { static State s; return s; }
inline ThreadCtx& cur() { return S().fibers[S().current].tc; }
inline Block& blk() { return S().blocks[S().fibers[S().current].slot]; }
inline unsigned char* dyn_smem() { return blk().smem; }

inline void yield()
{
    State& s = S();
    if (++s.yields_without_progress > (1ull << 28)) { fprintf(stderr, "emu: livelock\n"); abort(); }
    swapcontext(&s.fibers[s.current].ctx, &s.sched);
}
inline void progress() { S().yields_without_progress = 0; }
inline void trap() { fprintf(stderr, "emu: trap() in block %u thread %u\n", cur().bid, cur().tid); abort(); }

inline void sync_block()
{
    Block& s = blk();
    uint32_t gen = s.bar_gen;
    if (++s.bar_count == cur().block_dim) { s.bar_count = 0; ++s.bar_gen; progress(); }
    else while (s.bar_gen == gen) yield();
}

inline void bar_sync(uint32_t id, uint32_t n)
{
    Block& s = blk();
    uint32_t gen = s.nbar_gen[id & 15];
    if (++s.nbar_count[id & 15] == n) { s.nbar_count[id & 15] = 0; ++s.nbar_gen[id & 15]; progress(); }
    else while (s.nbar_gen[id & 15] == gen) yield();
}

inline void bar_arrive(uint32_t id, uint32_t n)
{
    Block& s = blk();
    if (++s.nbar_count[id & 15] == n) { s.nbar_count[id & 15] = 0; ++s.nbar_gen[id & 15]; progress(); }
}

inline void sync_warp()
{
    Block& s = blk();
    uint32_t w = cur().tid >> 5;
    uint32_t lanes = cur().block_dim - w * 32 < 32 ? cur().block_dim - w * 32 : 32;
    uint32_t gen = s.wbar_gen[w];
    if (++s.wbar_count[w] == lanes) { s.wbar_count[w] = 0; ++s.wbar_gen[w]; progress(); }
    else while (s.wbar_gen[w] == gen) yield();
}

inline uint64_t shfl(uint64_t v, int src)
{
    Block& s = blk();
    uint32_t t = cur().tid, w = t >> 5;
    s.xch[w * 32 + (t & 31)] = v;
    sync_warp();
    uint64_t r = s.xch[w * 32 + (src & 31)];
    sync_warp();
    return r;
}
inline uint64_t shfl_up(uint64_t v, int d)
{
    int l = (int)(cur().tid & 31);
    return shfl(v, l - d >= 0 ? l - d : l);
}
inline uint64_t shfl_down(uint64_t v, int d)
{
    int l = (int)(cur().tid & 31);
    return shfl(v, l + d <= 31 ? l + d : l);
}
inline uint32_t ballot(bool p)
{
    Block& s = blk();
    uint32_t t = cur().tid, w = t >> 5;
    s.xch[w * 32 + (t & 31)] = p ? 1 : 0;
    sync_warp();
    uint32_t m = 0;
    uint32_t lanes = cur().block_dim - w * 32 < 32 ? cur().block_dim - w * 32 : 32;
    for (uint32_t i = 0; i < lanes; ++i) m |= (uint32_t)s.xch[w * 32 + i] << i;
    sync_warp();
    return m;
}
inline uint32_t warp_reduce(uint32_t v, int op)
{
    Block& s = blk();
    uint32_t t = cur().tid, w = t >> 5;
    s.xch[w * 32 + (t & 31)] = v;
    sync_warp();
    uint32_t lanes = cur().block_dim - w * 32 < 32 ? cur().block_dim - w * 32 : 32;
    uint32_t r = op == 0 ? 0 : 0;
    for (uint32_t i = 0; i < lanes; ++i) {
        uint32_t x = (uint32_t)s.xch[w * 32 + i];
        r = op == 0 ? (x > r ? x : r) : op == 1 ? (r | x) : (r + x);
    }
    sync_warp();
    return r;
}

// mbarrier emulation in the barrier's own 8 bytes
struct MBar { uint16_t expected, arrived; int32_t tx : 31; uint32_t phase : 1; };
static_assert(sizeof(MBar) == 8, "emulated mbarrier must fit the 8-byte hardware object");
inline void mbar_check(MBar* b)
{
    if (b->arrived >= b->expected && b->tx == 0) { b->phase ^= 1; b->arrived = 0; progress(); }
}
inline void mbar_init(unsigned long long* bar, uint32_t count)
{
    MBar* b = (MBar*)bar;
    memset(b, 0, sizeof(MBar));
    b->expected = (uint16_t)count;
}
inline void mbar_arrive(unsigned long long* bar)
{
    MBar* b = (MBar*)bar;
    b->arrived++;
    mbar_check(b);
}
inline void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    MBar* b = (MBar*)bar;
    b->tx += (int32_t)bytes;
    b->arrived++;
    mbar_check(b);
}
inline void mbar_wait(unsigned long long* bar, uint32_t parity, uint32_t code = 0)
{
    MBar* b = (MBar*)bar;
    uint64_t n = 0;
    while ((uint32_t)b->phase == (parity & 1)) {             // phase flips when the awaited phase completes
        if (++n == (1u << 16) && getenv("EMU_TRACE"))
            fprintf(stderr, "emu: long mbarrier wait: bid %u tid %u code %u (warp %u, round %u)\n", cur().bid, cur().tid, code & 15, (code >> 4) & 31, code >> 12);
        yield();
    }
}
inline bool mbar_test(unsigned long long* bar, uint32_t parity)
{
    MBar* b = (MBar*)bar;
    return (uint32_t)b->phase != (parity & 1);
}
inline void bulk_g2s(void* d, const void* s, uint32_t bytes, unsigned long long* bar)
{
    if (((uintptr_t)d | (uintptr_t)s | bytes) & 15) { fprintf(stderr, "emu: misaligned bulk_g2s\n"); abort(); }
    memcpy(d, s, bytes);
    MBar* b = (MBar*)bar;
    b->tx -= (int32_t)bytes;
    mbar_check(b);
}
inline void bulk_s2g(void* d, const void* s, uint32_t bytes)
{
    if (((uintptr_t)d | (uintptr_t)s | bytes) & 15) { fprintf(stderr, "emu: misaligned bulk_s2g\n"); abort(); }
    memcpy(d, s, bytes);
}

inline void fiber_entry()
{
    State& s = S();
    (*s.body)();
    s.fibers[s.current].done = true;
    progress();
    swapcontext(&s.fibers[s.current].ctx, &s.sched);
}

inline void run_grid(uint32_t grid, uint32_t block, size_t smem_bytes, const std::function<void()>& body)
{
    State& s = S();
    const size_t STACK = 128 * 1024;
    if (smem_bytes > 232448) { fprintf(stderr, "emu: %zu bytes of dynamic smem exceed 227 KB\n", smem_bytes); abort(); }
    const char* env = getenv("EMU_CTAS");
    uint32_t resident = env ? (uint32_t)atoi(env) : 1u;
    if (resident < 1) resident = 1;
    if (resident > grid) resident = grid;
    s.body = &body;
    if (s.blocks.size() < resident) s.blocks.resize(resident);
    for (uint32_t k = 0; k < resident; ++k)
        if (!s.blocks[k].smem) s.blocks[k].smem = (unsigned char*)aligned_alloc(1024, 256 * 1024);
    if (s.fibers.size() < (size_t)resident * block) {
        size_t old = s.fibers.size();
        s.fibers.resize((size_t)resident * block);
        for (size_t i = old; i < s.fibers.size(); ++i) s.fibers[i].stack = (char*)malloc(STACK);
    }
    const uint32_t nw = (block + 31) / 32;
    uint32_t next_block = 0, running = 0;
    auto start = [&](uint32_t slot, uint32_t b) {
        Block& B = s.blocks[slot];
        B.bar_count = 0; B.bar_gen = 0;
        memset(B.nbar_count, 0, sizeof(B.nbar_count)); memset(B.nbar_gen, 0, sizeof(B.nbar_gen));
        B.wbar_count.assign(nw, 0); B.wbar_gen.assign(nw, 0);
        B.xch.assign((size_t)nw * 32, 0);
        memset(B.smem, 0xA5, smem_bytes);                // uninitialised shared memory is garbage
        B.alive = block;
        B.active = true;
        for (uint32_t t = 0; t < block; ++t) {
            Fiber& f = s.fibers[(size_t)slot * block + t];
            f.done = false;
            f.slot = (int)slot;
            f.tc = ThreadCtx{t, b, block, grid};
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
    };
    for (uint32_t k = 0; k < resident; ++k) { s.blocks[k].active = false; }
    for (uint32_t k = 0; k < resident && next_block < grid; ++k) { start(k, next_block++); ++running; }
    s.yields_without_progress = 0;
    while (running) {
        for (uint32_t k = 0; k < resident; ++k) {
            Block& B = s.blocks[k];
            if (!B.active) continue;
            for (uint32_t t = 0; t < block; ++t) {
                Fiber& f = s.fibers[(size_t)k * block + t];
                if (f.done) continue;
                s.current = (int)((size_t)k * block + t);
                swapcontext(&s.sched, &f.ctx);
                if (f.done) --B.alive;
            }
            if (B.alive == 0) {                          // the CTA has retired: the next one of the grid takes its place
                B.active = false;
                --running;
                if (next_block < grid) { start(k, next_block++); ++running; }
            }
        }
    }
    s.current = -1;
}

// minimal runtime shims used by the launch sequences in codec_launch.cuh
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

}  // namespace emu
