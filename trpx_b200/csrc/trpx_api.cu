// trpx_api.cu -- the C ABI of libtrpx_b200.so (include/trpx_b200.h): contexts, lanes (stream +
// scratch + staging), the host-pointer pipelines (H2D -> kernels -> D2H over several lanes) and the
// device-pointer entry points.  All arithmetic of the codec is in terse_encode.cuh /
// prolix_decode.cuh; there is no host implementation of it anywhere in this library.
#include "../../include/trpx_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "codec_launch.cuh"

using namespace trpx;

namespace {

constexpr int N_LANES = 8;     // lanes of a context (stream + scratch + staging each)
constexpr int DEV_LANES = 3;   // lanes the *_device entry points may name (trpx_ctx_lanes)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

constexpr int MAX_STAGE_THREADS = 8;   // helper threads that stage pageable host memory (two pinned 8 MB slots each)

struct Lane {
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host flavours: payload D2H of the encoder, off the lane's main stream
    cudaEvent_t ev_drained = nullptr;     // ... recorded there once d_out has been copied out
    cudaEvent_t ev_done = nullptr;        // decoder: the lane's kernels have written d_out
    bool drain_pending = false;
    DevBuf enc_scratch, dec_scratch;   // look-back descriptors / P1 tables
    DevBuf d_in, d_out, d_ends;        // staging of the host-pointer flavours
    u32* d_small = nullptr;            // [0] prolix_bits, [1] status
    u32* h_small = nullptr;            // pinned mirror
    u64* h_ends = nullptr;             // pinned copy of frame ends
    size_t h_ends_cap = 0;
    // profiling (trpx_ctx_set_profiling): events dropped between the kernels of the last call
    cudaStream_t prof_stream = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<const char*> ev_names;
    size_t ev_used = 0;
};

// Progress of a trpx_encode_host call, readable from another host thread while the call runs: batches whose
// payload (and frame sizes) have landed in the caller's buffers, as a prefix of the stack.  Advanced by host
// functions enqueued behind each batch's payload D2H (cudaLaunchHostFunc), read through atomics only.
struct EncProgress;
struct EncProgressMark { EncProgress* pr; size_t batch; };
struct EncProgress {
    std::atomic<size_t> seq{0}, frames{0}, bytes{0};
    std::mutex m;
    std::vector<char> done;
    std::vector<size_t> cum_frames, cum_bytes;
    std::vector<EncProgressMark> marks;
    size_t next = 0;
};

void CUDART_CB enc_progress_cb(void* arg)
{
    EncProgressMark* mk = (EncProgressMark*)arg;
    EncProgress& pr = *mk->pr;
    std::lock_guard<std::mutex> g(pr.m);
    pr.done[mk->batch] = 1;
    while (pr.next < pr.done.size() && pr.done[pr.next]) {
        pr.bytes.store(pr.cum_bytes[pr.next], std::memory_order_relaxed);
        pr.frames.store(pr.cum_frames[pr.next], std::memory_order_release);
        ++pr.next;
    }
}

}  // namespace

struct trpx_ctx {
    int device = 0;
    int sm_count = 0;
    Lane lanes[N_LANES];
    u64 launches = 0;
    std::string last_error;
    std::mutex mu;
    u32 sub_shift = 0;                   // 0: chosen per call (TRPX_SUB_SHIFT overrides: 5..8)
    u32 seg_bytes = 0, warm_bytes = 0;   // 0: chosen per call from the stream's mean block size (TRPX_SEG_BYTES / TRPX_WARM_BYTES override)
    size_t batch_bytes = 128u << 20;   // raw pixel bytes per pipeline batch of the host flavours (TRPX_BATCH_MB)
    size_t enc_batch_bytes = 0;        // ... of trpx_encode_host alone (TRPX_ENC_BATCH_MB; 0: batch_bytes)
    int enc_lanes = 4, dec_lanes = 6;  // batches in flight in the host flavours (TRPX_ENC_LANES / TRPX_DEC_LANES)
    u64* h_call_ends = nullptr;        // pinned: frame ends of every batch of one trpx_decode_host call
    size_t h_call_ends_cap = 0;
    DevBuf d_call_status;              // one status word per batch of a trpx_decode_host call
    DevBuf d_foreign;                  // whole payload of a call that has to recover the frame boundaries first
    // staging of PAGEABLE caller memory (host flavours): helper threads copy chunks into pinned slots, the copy engine
    // takes them from there (the driver's own bounce path for pageable memory moved ~5 GB/s on the B200 hosts)
    uint8_t* stage_ring = nullptr;
    cudaEvent_t stage_ev[2 * MAX_STAGE_THREADS] = {};
    bool stage_used[2 * MAX_STAGE_THREADS] = {};
    int stage_threads = 3;             // TRPX_STAGE_THREADS (0: leave pageable copies to the driver)
    uint8_t* unstage_ring = nullptr;   // the same for downloads into pageable memory: pinned slots, one stream per helper
    cudaStream_t unstage_stream[MAX_STAGE_THREADS] = {};
    cudaEvent_t unstage_ev[2 * MAX_STAGE_THREADS] = {};
    EncProgress enc_progress;
    std::vector<u32> call_status;
    u32 coop_grid = 0;
    bool profiling = false;
};

namespace {

bool cuda_ok(trpx_ctx* c, cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    if (c) c->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}

bool ensure(trpx_ctx* c, DevBuf& b, size_t bytes)
{
    if (b.cap >= bytes) return true;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    if (!cuda_ok(c, cudaMalloc(&b.p, want), "cudaMalloc")) return false;
    b.cap = want;
    return true;
}

bool ensure_host_ends(trpx_ctx* c, Lane& l, size_t n)
{
    if (l.h_ends_cap >= n) return true;
    if (l.h_ends) cudaFreeHost(l.h_ends);
    l.h_ends = nullptr;
    l.h_ends_cap = 0;
    if (!cuda_ok(c, cudaMallocHost((void**)&l.h_ends, (n + 1024) * sizeof(u64)), "cudaMallocHost")) return false;
    l.h_ends_cap = n + 1024;
    return true;
}

u32 env_u32(const char* name, u32 dflt)
{
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return (u32)strtoul(v, nullptr, 10);
}

// P1 segment geometry.  A speculative walker needs ~70 blocks (median) to lock onto the true header chain and
// < ~700 in 99.9 % of the cases (measured on diffraction, sparse and dark-subtracted frames), whatever the block
// size; so the warm-up is sized in BLOCKS -- 600 of the stream's mean block size -- and a segment is 2.5 warm-ups.
// The ~1 % of walkers that still arrive wrong are re-walked by the resolve kernel, a warp each and from shared
// memory (never wrong, ~30 us each).  Round 2: 1200 blocks / 16 KB segments -> 600 blocks / 8 KB segments, with four
// walker CTAs per SM: decode 3.16 -> 3.0 ms per 10,000 frames (sweep in DESIGN.md).
// Checkpoint spacing: a thread of the unpack kernel owns the blocks whose headers start in one sub-segment; about
// six blocks per thread keeps its 256-thread slice within one output stage: 32 bytes for diffraction frames
// (~40 bits per block), down to 4 bytes for sparse counting data (~5 bits per block).
void walk_geometry(const trpx_ctx* c, u64 payload_bytes, u64 n_frames, u64 nblocks, u32& seg, u32& warm, u32& sub_shift)
{
    seg = c->seg_bytes;
    warm = c->warm_bytes;
    const double blocks = (double)n_frames * (double)nblocks;
    const double mean_bits = blocks > 0 ? 8.0 * (double)payload_bytes / blocks : 64.0;
    sub_shift = c->sub_shift;
    if (!sub_shift) {
        sub_shift = SUB_SHIFT_MAX;
        while (sub_shift > SUB_SHIFT_MIN && (double)(1u << sub_shift) > 9.0 * mean_bits) --sub_shift;
    }
    if (seg && warm) return;
    u64 w = (u64)(600.0 * mean_bits / 8.0);
    w = (w + 255) / 256 * 256;
    if (w < 512) w = 512;
    if (w > 131072) w = 131072;
    // a segment: twice the warm-up, in whole slices of the unpack kernel (256 sub-segments); sparse streams get
    // short segments in bytes -- the same ~2400 blocks -- and therefore enough walkers to fill the machine
    const u64 slice = (u64)256 << (sub_shift - 3);
    u64 sg = (5 * w / 2 + slice / 2) / slice * slice;        // (to the nearest whole slice)
    // A walker is one dependent chain over warm-up + segment, and a call whose payload gives fewer segments than
    // the machine has lanes for (a few big frames, one batch of a host call) is bound by the length of that chain,
    // not by throughput: then segments shrink towards one slice until ~12 warps of walkers per SM exist.
    const u64 lanes_wanted = (u64)c->sm_count * 12 * 32;
    const u64 fit = payload_bytes / lanes_wanted / slice * slice;
    if (sg > fit) sg = fit;
    if (sg < slice) sg = slice;
    // ... and half a slice when even whole slices leave fewer than four warps of walkers per SM (the share of one GPU when
    // a stack of a few large frames is sharded over eight): the unpack kernel's slices are then half empty, which costs
    // less than the walk gains.  Measured on 4 frames of 4148x4362 i32: 0.84 -> 0.72 ms per decode.
    if (slice >= 2048 && payload_bytes / sg < (u64)c->sm_count * 4 * 32) sg = slice / 2;
    // Small calls (a frame or a few: one batch of a few hundred KB) are pure latency: every walker is ONE dependent chain
    // over warm-up + segment, so both shrink to ~1 KB (a sixth of the usual warm-up; segments may then be shorter than
    // an unpack slice, whose spare threads idle).  More walkers arrive wrong and are re-walked by the resolve kernel, but
    // over 1 KB at most.  Measured, one 512x512 frame: 726 -> 200 us (tools/latency_probe.py).
    if (payload_bytes <= ((u64)8 << 20)) {
        u64 sw = 1024;                                          // the largest power of two <= w / 6, within [1 KB, 8 KB]
        while (sw < 8192 && 2 * sw <= w / 6) sw *= 2;
        if (payload_bytes > ((u64)256 << 10)) sw *= 2;
        w = sw;
        sg = sw;
    }
    if (!warm) warm = (u32)w;
    if (!seg) seg = (u32)sg;
}

u32 enc_ctas_per_sm(trpx_ctx* c, int dtype, const EncPlan& pl)
{
    int n = 0;
    const void* k = enc_kernel(dtype, pl.fast);
    if (pl.smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, (int)pl.threads, pl.smem);
    if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
    (void)c;
    return (u32)n;
}

void prof_mark(void* user, const char* name)
{
    Lane* l = (Lane*)user;
    if (l->ev_used == l->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        l->ev_pool.push_back(e);
        l->ev_names.push_back(name);
    }
    l->ev_names[l->ev_used] = name;
    cudaEventRecord(l->ev_pool[l->ev_used++], l->prof_stream);
}

Launcher make_launcher(trpx_ctx* c, cudaStream_t s, Lane* lane = nullptr)
{
    Launcher L;
    L.stream = s;
    L.sm_count = (u32)c->sm_count;
    L.launches = &c->launches;
    L.err = cudaSuccess;
    if (c->profiling && lane) {
        lane->prof_stream = s;
        lane->ev_used = 0;
        L.mark_fn = prof_mark;
        L.mark_user = lane;
    }
    return L;
}

// Host flavour of the encoder: (frame ends, prolix_bits, status) of a batch go to pinned host memory with plain
// stores over PCIe instead of a D2H copy, so that the host learns a batch's size without touching a copy engine.
__global__ void publish_results_kernel(const u64* d_ends, u64 n, const u32* d_small, u64* h_ends, u32* h_small)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) h_ends[i] = d_ends[i];
    if (blockIdx.x == 0 && threadIdx.x < 2) h_small[threadIdx.x] = d_small[threadIdx.x];
}

// device status words: 0, a TRPX_ERR_* code, or >= ST_INTERNAL when a bounded wait inside a kernel ran out (simt.cuh)
constexpr size_t STAGE_CHUNK = 8u << 20;

bool is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Host -> device copy on `stream`.  Pinned sources go straight to the copy engine.  Pageable sources are cut into
// 8 MB chunks that helper threads copy into pinned slots (two per thread) and enqueue from there: the host-side
// memcpy of one chunk overlaps the DMA of the others.  Returns when every chunk has been ENQUEUED.
cudaError_t h2d_async(trpx_ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t stream)
{
    const int T = c->stage_threads;
    if (T <= 0 || bytes < 2 * STAGE_CHUNK || !is_pageable(src)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
    if (!c->stage_ring) {
        if (cudaHostAlloc((void**)&c->stage_ring, 2 * (size_t)MAX_STAGE_THREADS * STAGE_CHUNK, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            c->stage_ring = nullptr;
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
        }
        for (cudaEvent_t& e : c->stage_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    }
    const size_t n_chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    std::vector<cudaError_t> err((size_t)T, cudaSuccess);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            cudaSetDevice(c->device);
            size_t k = 0;
            for (size_t i = (size_t)t; i < n_chunks; i += (size_t)T, ++k) {
                const int slot = t + T * (int)(k & 1);
                const size_t off = i * STAGE_CHUNK, n = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
                if (c->stage_used[slot]) cudaEventSynchronize(c->stage_ev[slot]);     // the slot's previous chunk has left
                memcpy(c->stage_ring + (size_t)slot * STAGE_CHUNK, (const uint8_t*)src + off, n);
                cudaError_t e = cudaMemcpyAsync((uint8_t*)dst + off, c->stage_ring + (size_t)slot * STAGE_CHUNK, n, cudaMemcpyHostToDevice, stream);
                if (e == cudaSuccess) e = cudaEventRecord(c->stage_ev[slot], stream);
                c->stage_used[slot] = true;
                if (e != cudaSuccess) { err[(size_t)t] = e; return; }
            }
        });
    for (auto& x : th) x.join();
    for (cudaError_t e : err)
        if (e != cudaSuccess) return e;
    return cudaSuccess;
}

// Device -> PAGEABLE host memory, blocking: helper threads pull 8 MB chunks into pinned slots on their own streams (two
// slots each: the DMA of a helper's next chunk runs while it copies the previous one out) and memcpy them to `dst`.
// The driver's own path for pageable destinations does the same with ONE thread.  `ready`: the event after which `src`
// may be read.  Falls back to a plain synchronous copy when the ring cannot be allocated.
cudaError_t d2h_staged(trpx_ctx* c, void* dst, const void* src, size_t bytes, cudaEvent_t ready)
{
    const int T = c->stage_threads;
    if (T > 0 && !c->unstage_ring) {
        if (cudaHostAlloc((void**)&c->unstage_ring, 2 * (size_t)MAX_STAGE_THREADS * STAGE_CHUNK, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            c->unstage_ring = nullptr;
        } else {
            for (cudaStream_t& st : c->unstage_stream) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            for (cudaEvent_t& e : c->unstage_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        }
    }
    if (T <= 0 || !c->unstage_ring) {
        cudaError_t e = cudaEventSynchronize(ready);
        return e != cudaSuccess ? e : cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
    }
    const size_t n_chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    std::vector<cudaError_t> err((size_t)T, cudaSuccess);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            cudaSetDevice(c->device);
            cudaStream_t st = c->unstage_stream[t];
            cudaError_t e = cudaStreamWaitEvent(st, ready, 0);
            auto issue = [&](size_t i, int k) {                   // chunk i -> this helper's slot k
                const size_t off = i * STAGE_CHUNK, n = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
                const int slot = t + T * k;
                cudaError_t r = cudaMemcpyAsync(c->unstage_ring + (size_t)slot * STAGE_CHUNK, (const uint8_t*)src + off, n, cudaMemcpyDeviceToHost, st);
                return r != cudaSuccess ? r : cudaEventRecord(c->unstage_ev[slot], st);
            };
            int k = 0;
            if (e == cudaSuccess && (size_t)t < n_chunks) e = issue((size_t)t, 0);
            for (size_t i = (size_t)t; i < n_chunks && e == cudaSuccess; i += (size_t)T, k ^= 1) {
                if (i + (size_t)T < n_chunks) e = issue(i + (size_t)T, k ^ 1);
                const int slot = t + T * k;
                const cudaError_t w = cudaEventSynchronize(c->unstage_ev[slot]);
                if (w != cudaSuccess) { e = w; break; }
                const size_t off = i * STAGE_CHUNK, n = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
                memcpy((uint8_t*)dst + off, c->unstage_ring + (size_t)slot * STAGE_CHUNK, n);
            }
            if (e != cudaSuccess) cudaStreamSynchronize(st);     // nothing of this call stays in flight
            err[(size_t)t] = e;
        });
    for (auto& x : th) x.join();
    for (cudaError_t e : err)
        if (e != cudaSuccess) return e;
    return cudaSuccess;
}

int status_of_device_word(u32 w) { return w == 0 ? TRPX_OK : w >= ST_INTERNAL ? TRPX_ERR_CUDA : (int)w; }
void note_device_word(trpx_ctx* c, u32 w)
{
    if (w >= ST_INTERNAL && c) c->last_error = "a bounded wait inside a kernel ran out (internal code " + std::to_string(w - ST_INTERNAL) + ")";
}

}  // namespace

extern "C" {

int trpx_abi_version(void) { return TRPX_ABI_VERSION; }

const char* trpx_strerror(int s)
{
    switch (s) {
    case TRPX_OK: return "ok";
    case TRPX_ERR_BAD_ARG: return "bad argument";
    case TRPX_ERR_CAPACITY: return "output buffer too small";
    case TRPX_ERR_CUDA: return "CUDA error";
    case TRPX_ERR_MALFORMED: return "malformed TERSE payload";
    case TRPX_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
    case TRPX_ERR_NOMEM: return "out of memory";
    case TRPX_ALREADY: return "already pinned";
    default: return "unknown status";
    }
}

size_t trpx_dtype_size(int dtype) { return dtype_size(dtype); }
int trpx_dtype_is_signed(int dtype) { return dtype_signed(dtype) ? 1 : 0; }

size_t trpx_max_compressed_bytes(size_t n_values, int dtype, unsigned block, size_t n_frames)
{
    const size_t sz = dtype_is_pixel(dtype) ? dtype_size(dtype) : 0;
    if (!sz || !block) return 0;
    const size_t w = 8 * sz + (dtype_signed(dtype) ? 1 : 0);
    const size_t nblocks = (n_values + block - 1) / block;
    const size_t per_frame = (12 * nblocks + n_values * w + 7) / 8 + 1;
    return (per_frame * n_frames + 15 + 16) / 16 * 16;
}

int trpx_ctx_create(int device, trpx_ctx** out)
{
    if (!out) return TRPX_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return TRPX_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) return TRPX_ERR_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return TRPX_ERR_NO_DEVICE; }
    trpx_ctx* c = new trpx_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (!cuda_ok(c, cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) { delete c; return TRPX_ERR_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) { delete c; return TRPX_ERR_NO_DEVICE; }
    for (int i = 0; i < N_LANES; ++i) {
        Lane& l = c->lanes[i];
        if (!cuda_ok(c, cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
            !cuda_ok(c, cudaStreamCreateWithFlags(&l.copy_stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
            !cuda_ok(c, cudaEventCreateWithFlags(&l.ev_drained, cudaEventDisableTiming), "cudaEventCreate") ||
            !cuda_ok(c, cudaEventCreateWithFlags(&l.ev_done, cudaEventDisableTiming), "cudaEventCreate") ||
            !cuda_ok(c, cudaMalloc((void**)&l.d_small, 64), "cudaMalloc") ||
            !cuda_ok(c, cudaMallocHost((void**)&l.h_small, 64), "cudaMallocHost")) {
            trpx_ctx_destroy(c);
            return TRPX_ERR_NOMEM;
        }
    }
    c->seg_bytes = env_u32("TRPX_SEG_BYTES", 0);
    c->warm_bytes = env_u32("TRPX_WARM_BYTES", 0);
    c->sub_shift = env_u32("TRPX_SUB_SHIFT", 0);
    c->batch_bytes = (size_t)env_u32("TRPX_BATCH_MB", (u32)(c->batch_bytes >> 20)) << 20;
    c->enc_batch_bytes = getenv("TRPX_ENC_BATCH_MB") ? (size_t)env_u32("TRPX_ENC_BATCH_MB", 0) << 20 : c->batch_bytes;
    c->enc_lanes = (int)env_u32("TRPX_ENC_LANES", (u32)c->enc_lanes);
    c->dec_lanes = (int)env_u32("TRPX_DEC_LANES", (u32)c->dec_lanes);
    {   // half the host's hardware threads, at most six: host memory saturates there (1 / 2 / 3 / 5 / 8 threads moved
        // 1 GB into pageable memory in 141 / 81 / 63 / 53 / 52 ms on the bench host)
        const unsigned hw = std::thread::hardware_concurrency();
        if (hw) c->stage_threads = hw / 2 < 1 ? 1 : hw / 2 > 6 ? 6 : (int)(hw / 2);
    }
    c->stage_threads = (int)env_u32("TRPX_STAGE_THREADS", (u32)c->stage_threads);
    if (c->stage_threads > MAX_STAGE_THREADS) c->stage_threads = MAX_STAGE_THREADS;
    if (c->enc_lanes < 1) c->enc_lanes = 1;
    if (c->enc_lanes > N_LANES) c->enc_lanes = N_LANES;
    if (c->dec_lanes < 1) c->dec_lanes = 1;
    if (c->dec_lanes > N_LANES) c->dec_lanes = N_LANES;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)prolix_resolve_kernel<RESOLVE_NT>, RESOLVE_NT, 0) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    if (occ > 4) occ = 4;
    c->coop_grid = (u32)(c->sm_count * occ);
    *out = c;
    return TRPX_OK;
}

void trpx_ctx_destroy(trpx_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < N_LANES; ++i) {
        Lane& l = c->lanes[i];
        if (l.stream) { cudaStreamSynchronize(l.stream); cudaStreamDestroy(l.stream); }
        if (l.copy_stream) { cudaStreamSynchronize(l.copy_stream); cudaStreamDestroy(l.copy_stream); }
        if (l.ev_drained) cudaEventDestroy(l.ev_drained);
        if (l.ev_done) cudaEventDestroy(l.ev_done);
        DevBuf* bufs[] = {&l.enc_scratch, &l.dec_scratch, &l.d_in, &l.d_out, &l.d_ends};
        for (DevBuf* b : bufs)
            if (b->p) cudaFree(b->p);
        if (l.d_small) cudaFree(l.d_small);
        if (l.h_small) cudaFreeHost(l.h_small);
        if (l.h_ends) cudaFreeHost(l.h_ends);
        for (cudaEvent_t e : l.ev_pool) cudaEventDestroy(e);
    }
    if (c->h_call_ends) cudaFreeHost(c->h_call_ends);
    if (c->d_call_status.p) cudaFree(c->d_call_status.p);
    if (c->d_foreign.p) cudaFree(c->d_foreign.p);
    if (c->stage_ring) cudaFreeHost(c->stage_ring);
    if (c->unstage_ring) {
        cudaFreeHost(c->unstage_ring);
        for (cudaStream_t st : c->unstage_stream)
            if (st) cudaStreamDestroy(st);
        for (cudaEvent_t e : c->unstage_ev)
            if (e) cudaEventDestroy(e);
    }
    for (cudaEvent_t e : c->stage_ev)
        if (e) cudaEventDestroy(e);
    delete c;
}

int trpx_ctx_device(const trpx_ctx* c) { return c ? c->device : -1; }
const char* trpx_last_error(const trpx_ctx* c) { return c ? c->last_error.c_str() : ""; }
int trpx_ctx_lanes(const trpx_ctx* c) { return c ? DEV_LANES : 0; }
uint64_t trpx_ctx_launch_count(const trpx_ctx* c) { return c ? c->launches : 0; }
size_t trpx_ctx_scratch_bytes(const trpx_ctx* c)
{
    size_t t = 0;
    if (c)
        for (int i = 0; i < N_LANES; ++i) {
            const Lane& l = c->lanes[i];
            t += l.enc_scratch.cap + l.dec_scratch.cap + l.d_in.cap + l.d_out.cap + l.d_ends.cap;
        }
    return t;
}

int trpx_ctx_encode_progress(trpx_ctx* c, size_t* call_seq, size_t* frames_done, size_t* payload_bytes_done)
{
    if (!c) return TRPX_ERR_BAD_ARG;
    EncProgress& pr = c->enc_progress;
    std::lock_guard<std::mutex> g(pr.m);                    // (the marks advance under the same lock: a consistent pair)
    if (call_seq) *call_seq = pr.seq.load(std::memory_order_acquire);
    if (frames_done) *frames_done = pr.frames.load(std::memory_order_acquire);
    if (payload_bytes_done) *payload_bytes_done = pr.bytes.load(std::memory_order_acquire);
    return TRPX_OK;
}

int trpx_ctx_set_profiling(trpx_ctx* c, int on)
{
    if (!c) return TRPX_ERR_BAD_ARG;
    c->profiling = on != 0;
    return TRPX_OK;
}

int trpx_ctx_last_kernel_times(trpx_ctx* c, int lane, const char** names, float* ms, int cap)
{
    if (!c || lane < 0 || lane >= DEV_LANES) return 0;
    Lane& l = c->lanes[lane];
    int n = 0;
    for (size_t i = 1; i < l.ev_used && n < cap; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, l.ev_pool[i - 1], l.ev_pool[i]) != cudaSuccess) { cudaGetLastError(); break; }
        if (names) names[n] = l.ev_names[i];
        if (ms) ms[n] = t;
        ++n;
    }
    return n;
}

// ------------------------------------------------------------------------------ TERSE, device pointers
int trpx_encode_device(trpx_ctx* c, int lane, const void* d_pixels, int dtype, size_t n_values,
                       size_t n_frames, unsigned block, uint8_t* d_out, size_t out_capacity,
                       uint64_t* d_frame_ends, uint32_t* d_prolix_bits, uint32_t* d_status, void* stream)
{
    if (!c) return TRPX_ERR_BAD_ARG;
    if (lane < 0 || lane >= DEV_LANES || !d_pixels || !d_out || !d_frame_ends || !d_prolix_bits || !d_status ||
        !dtype_is_pixel(dtype) || !block || !n_values || !n_frames)
        return TRPX_ERR_BAD_ARG;
    if (((uintptr_t)d_out & 15) || ((uintptr_t)d_pixels & (dtype_size(dtype) - 1))) return TRPX_ERR_BAD_ARG;
    cudaSetDevice(c->device);
    EncPlan pl = enc_plan(dtype, d_pixels, n_values, n_frames, block);
    if (!pl.ok) { c->last_error = "unsupported geometry (block too large or too many tiles)"; return TRPX_ERR_BAD_ARG; }
    Lane& l = c->lanes[lane];
    if (!ensure(c, l.enc_scratch, pl.scratch_bytes)) return TRPX_ERR_NOMEM;
    Launcher L = make_launcher(c, (cudaStream_t)stream, &l);
    encode_async(L, dtype, d_pixels, n_values, n_frames, block, d_out, out_capacity, (u64*)d_frame_ends,
                 d_prolix_bits, d_status, l.enc_scratch.p, pl, enc_ctas_per_sm(c, dtype, pl));
    if (!cuda_ok(c, L.err, "encode launch")) return TRPX_ERR_CUDA;
    return TRPX_OK;
}

// ------------------------------------------------------------------------------ PROLIX, device pointers
int trpx_decode_device(trpx_ctx* c, int lane, const uint8_t* d_payload, size_t payload_bytes, int is_signed,
                       unsigned block, size_t n_values, size_t n_frames, const uint64_t* d_frame_ends,
                       uint64_t* d_frame_ends_out, void* d_out, int out_dtype, uint32_t* d_status, void* stream)
{
    if (!c) return TRPX_ERR_BAD_ARG;
    if (lane < 0 || lane >= DEV_LANES || !d_payload || !payload_bytes || !d_out || !d_status ||
        !dtype_size(out_dtype) || !block || !n_values || !n_frames)
        return TRPX_ERR_BAD_ARG;
    if (is_signed && !dtype_signed(out_dtype)) return TRPX_ERR_BAD_ARG;      // Terse.hpp:356-357
    if (((uintptr_t)d_payload & 15) || ((uintptr_t)d_out & (dtype_size(out_dtype) - 1))) return TRPX_ERR_BAD_ARG;
    cudaSetDevice(c->device);
    u32 seg, warm, sub_shift;
    walk_geometry(c, payload_bytes, n_frames, (n_values + block - 1) / block, seg, warm, sub_shift);
    DecPlan pl = dec_plan(out_dtype, payload_bytes, n_values, n_frames, block, d_out, seg, warm, sub_shift);
    if (!pl.ok) return TRPX_ERR_BAD_ARG;
    Lane& l = c->lanes[lane];
    if (!ensure(c, l.dec_scratch, dec_scratch_need(pl, out_dtype, payload_bytes, block, n_frames, d_frame_ends == nullptr))) return TRPX_ERR_NOMEM;
    Launcher L = make_launcher(c, (cudaStream_t)stream, &l);
    decode_async(L, d_payload, payload_bytes, is_signed != 0, block, n_values, n_frames, (const u64*)d_frame_ends,
                 (u64*)d_frame_ends_out, d_out, out_dtype, d_status, l.dec_scratch.p, pl, c->coop_grid);
    if (!cuda_ok(c, L.err, "decode launch")) return TRPX_ERR_CUDA;
    return TRPX_OK;
}

// ------------------------------------------------------------------------------ TERSE, host pointers
// Frames are cut into batches of ~batch_bytes; batch b runs on lane b % enc_lanes: H2D, encode, then a small
// kernel stores (frame ends, prolix_bits, status) straight into pinned host memory -- no copy engine, so these few
// bytes never queue behind another context's bulk D2H traffic.  Once a batch's size is known its payload goes
// D2H to its final place in `out` on the lane's copy stream; the host never waits for that copy before the call's
// end (the lane's next kernel waits for it on the device), so the H2D engine is never left idle.
int trpx_encode_host(trpx_ctx* c, const void* pixels, int dtype, size_t n_values, size_t n_frames, unsigned block,
                     uint8_t* out, size_t out_capacity, size_t* frame_bytes, size_t* total_bytes,
                     unsigned* prolix_bits)
{
    if (!c) return TRPX_ERR_BAD_ARG;
    const size_t sz = dtype_is_pixel(dtype) ? dtype_size(dtype) : 0;
    if (!pixels || !out || !sz || !block || !n_values || !n_frames) return TRPX_ERR_BAD_ARG;
    std::lock_guard<std::mutex> guard(c->mu);
    cudaSetDevice(c->device);
    const size_t frame_raw = n_values * sz;
    size_t fpb = c->enc_batch_bytes / (frame_raw ? frame_raw : 1);   // frames per batch
    if (fpb < 1) fpb = 1;
    if (fpb > n_frames) fpb = n_frames;
    const size_t n_batches = (n_frames + fpb - 1) / fpb;
    const int nl = c->enc_lanes;

    struct Pending { size_t f0, nf, batch; bool active; };
    Pending pend[N_LANES] = {};
    size_t out_off = 0;
    unsigned pb_max = 0;
    int rc = TRPX_OK;
    EncProgress& pr = c->enc_progress;
    {
        std::lock_guard<std::mutex> g(pr.m);
        pr.done.assign(n_batches, 0);
        pr.cum_frames.assign(n_batches, 0);
        pr.cum_bytes.assign(n_batches, 0);
        pr.marks.resize(n_batches);
        pr.next = 0;
        pr.frames.store(0, std::memory_order_relaxed);
        pr.bytes.store(0, std::memory_order_relaxed);
        pr.seq.fetch_add(1, std::memory_order_release);
    }

    auto finish = [&](int li) -> int {
        Lane& l = c->lanes[li];
        Pending& q = pend[li];
        if (!q.active) return TRPX_OK;
        q.active = false;
        if (!cuda_ok(c, cudaStreamSynchronize(l.stream), "encode batch")) return TRPX_ERR_CUDA;
        if (l.h_small[1] != 0) { note_device_word(c, l.h_small[1]); return status_of_device_word(l.h_small[1]); }
        const size_t bytes = (size_t)l.h_ends[q.nf - 1];
        if (out_off + bytes > out_capacity) return TRPX_ERR_CAPACITY;
        if (!cuda_ok(c, cudaMemcpyAsync(out + out_off, l.d_out.p, bytes, cudaMemcpyDeviceToHost, l.copy_stream), "D2H payload") ||
            !cuda_ok(c, cudaEventRecord(l.ev_drained, l.copy_stream), "event record"))
            return TRPX_ERR_CUDA;
        l.drain_pending = true;
        if (frame_bytes)
            for (size_t i = 0; i < q.nf; ++i) frame_bytes[q.f0 + i] = (size_t)(l.h_ends[i] - (i ? l.h_ends[i - 1] : 0));
        if (l.h_small[0] > pb_max) pb_max = l.h_small[0];
        out_off += bytes;
        pr.cum_frames[q.batch] = q.f0 + q.nf;               // (frame_bytes[] above is written before the mark can fire)
        pr.cum_bytes[q.batch] = out_off;
        pr.marks[q.batch] = EncProgressMark{&pr, q.batch};
        cudaLaunchHostFunc(l.copy_stream, enc_progress_cb, &pr.marks[q.batch]);
        return TRPX_OK;
    };

    for (size_t b = 0; b < n_batches && rc == TRPX_OK; ++b) {
        const int li = (int)(b % nl);
        Lane& l = c->lanes[li];
        rc = finish(li);
        if (rc != TRPX_OK) break;
        const size_t f0 = b * fpb, nf = (f0 + fpb <= n_frames) ? fpb : n_frames - f0;
        const size_t cap = trpx_max_compressed_bytes(n_values, dtype, block, nf);
        if (!ensure(c, l.d_in, nf * frame_raw + 16) || !ensure(c, l.d_out, cap) || !ensure(c, l.d_ends, nf * 8) ||
            !ensure_host_ends(c, l, nf)) { rc = TRPX_ERR_NOMEM; break; }
        EncPlan pl = enc_plan(dtype, l.d_in.p, n_values, nf, block);
        if (!pl.ok) { c->last_error = "unsupported geometry (block too large or too many tiles)"; rc = TRPX_ERR_BAD_ARG; break; }
        if (!ensure(c, l.enc_scratch, pl.scratch_bytes)) { rc = TRPX_ERR_NOMEM; break; }
        if (!cuda_ok(c, h2d_async(c, l.d_in.p, (const uint8_t*)pixels + f0 * frame_raw, nf * frame_raw, l.stream), "H2D pixels")) { rc = TRPX_ERR_CUDA; break; }
        if (l.drain_pending) {                                   // d_out still holds the lane's previous payload
            cudaStreamWaitEvent(l.stream, l.ev_drained, 0);
            l.drain_pending = false;
        }
        Launcher L = make_launcher(c, l.stream);
        encode_async(L, dtype, l.d_in.p, n_values, nf, block, l.d_out.p, cap, (u64*)l.d_ends.p, l.d_small, l.d_small + 1,
                     l.enc_scratch.p, pl, enc_ctas_per_sm(c, dtype, pl));
        if (!cuda_ok(c, L.err, "encode launch")) { rc = TRPX_ERR_CUDA; break; }
        L.err = launch(publish_results_kernel, (u32)((nf + 255) / 256 < 64 ? (nf + 255) / 256 : 64), 256u, 0, l.stream,
                       (const u64*)l.d_ends.p, (u64)nf, (const u32*)l.d_small, l.h_ends, l.h_small);
        L.count("publish_results");
        if (!cuda_ok(c, L.err, "publish launch")) { rc = TRPX_ERR_CUDA; break; }
        pend[li] = Pending{f0, nf, b, true};
    }
    // drain in batch order
    for (int k = 0; k < nl; ++k) {
        const int li = (int)((n_batches + (size_t)k) % (size_t)nl);
        int r = finish(li);
        if (rc == TRPX_OK) rc = r;
    }
    for (int i = 0; i < nl; ++i) {
        Lane& l = c->lanes[i];
        if (rc != TRPX_OK) cudaStreamSynchronize(l.stream);
        if (!cuda_ok(c, cudaStreamSynchronize(l.copy_stream), "D2H payload") && rc == TRPX_OK) rc = TRPX_ERR_CUDA;
        l.drain_pending = false;
    }
    if (rc != TRPX_OK) return rc;
    if (total_bytes) *total_bytes = out_off;
    if (prolix_bits) *prolix_bits = pb_max;
    return TRPX_OK;
}

// ------------------------------------------------------------------------------ PROLIX, host pointers
int trpx_decode_host(trpx_ctx* c, const uint8_t* payload, size_t payload_bytes, int is_signed, unsigned block,
                     size_t n_values, size_t total_frames, size_t first_frame, size_t n_frames,
                     const size_t* frame_bytes, size_t* frame_bytes_out, void* out, int out_dtype)
{
    if (!c) return TRPX_ERR_BAD_ARG;
    const size_t so = dtype_size(out_dtype);
    if (!payload || !payload_bytes || !out || !so || !block || !n_values || !total_frames || !n_frames ||
        first_frame + n_frames > total_frames)
        return TRPX_ERR_BAD_ARG;
    if (is_signed && !dtype_signed(out_dtype)) return TRPX_ERR_BAD_ARG;      // Terse.hpp:356-357
    std::lock_guard<std::mutex> guard(c->mu);
    cudaSetDevice(c->device);

    // absolute end offsets of every frame
    std::vector<u64> ends(total_frames);
    const uint8_t* resident = nullptr;                          // device copy of the whole payload, when one exists
    if (frame_bytes) {
        u64 acc = 0;
        for (size_t f = 0; f < total_frames; ++f) { acc += frame_bytes[f]; ends[f] = acc; }
        if (acc > payload_bytes) return TRPX_ERR_MALFORMED;
    } else if (total_frames == 1) {
        ends[0] = payload_bytes;
    } else {
        // the container does not store frame boundaries (Terse.hpp:459, :562-585): recover them on the device.  The
        // payload is uploaded ONCE: it stays resident and the batches below take their slabs from it device-to-device.
        Lane& l = c->lanes[0];
        u32 seg, warm, sub_shift;
        const size_t nblocks = (n_values + block - 1) / block;
        walk_geometry(c, payload_bytes, total_frames, nblocks, seg, warm, sub_shift);
        const DecPlan gpl = dec_plan(out_dtype, payload_bytes, block, 1, block, nullptr, seg, warm, sub_shift, true);
        if (!gpl.ok) return TRPX_ERR_BAD_ARG;   // (the batches below know their frame sizes: only this pass needs the G tables)
        if (!ensure(c, c->d_foreign, payload_bytes + 32) || !ensure(c, l.d_ends, total_frames * 8) || !ensure(c, l.dec_scratch, gpl.scratch_bytes))
            return TRPX_ERR_NOMEM;
        if (!cuda_ok(c, cudaMemsetAsync((uint8_t*)c->d_foreign.p + (payload_bytes & ~(size_t)15), 0, 32, l.stream), "memset") ||
            !cuda_ok(c, h2d_async(c, c->d_foreign.p, payload, payload_bytes, l.stream), "H2D payload"))
            return TRPX_ERR_CUDA;
        resident = (const uint8_t*)c->d_foreign.p;
        cudaMemsetAsync(l.d_small, 0, 8, l.stream);
        Launcher L = make_launcher(c, l.stream);
        find_frames_async(L, c->d_foreign.p, payload_bytes, block, nblocks, (u32)(n_values - (nblocks - 1) * block), total_frames,
                          (u64*)l.d_ends.p, l.d_small + 1, l.dec_scratch.p, gpl, c->coop_grid);
        if (!cuda_ok(c, L.err, "find frames launch")) return TRPX_ERR_CUDA;
        cudaMemcpyAsync(ends.data(), l.d_ends.p, total_frames * 8, cudaMemcpyDeviceToHost, l.stream);
        cudaMemcpyAsync(l.h_small, l.d_small, 8, cudaMemcpyDeviceToHost, l.stream);
        if (!cuda_ok(c, cudaStreamSynchronize(l.stream), "find frames")) return TRPX_ERR_CUDA;
        if (l.h_small[1] != 0) { note_device_word(c, l.h_small[1]); return status_of_device_word(l.h_small[1]); }
    }
    if (frame_bytes_out)
        for (size_t f = 0; f < total_frames; ++f) frame_bytes_out[f] = (size_t)(ends[f] - (f ? ends[f - 1] : 0));

    // Batches of ~batch_bytes of output; batch b runs on lane b % dec_lanes: H2D of its payload slab and frame ends,
    // the kernels, D2H of the pixels.  Nothing here waits on the host: every batch has its own slice of the pinned
    // frame-end table and its own status word, buffers are reused in stream order, and the call drains once at the end.
    const size_t frame_raw = n_values * so;
    size_t fpb = c->batch_bytes / (frame_raw ? frame_raw : 1);
    if (fpb < 1) fpb = 1;
    if (fpb > n_frames) fpb = n_frames;
    const size_t n_batches = (n_frames + fpb - 1) / fpb;
    const int nl = c->dec_lanes;
    int rc = TRPX_OK;
    if (c->h_call_ends_cap < n_frames) {
        if (c->h_call_ends) cudaFreeHost(c->h_call_ends);
        c->h_call_ends = nullptr;
        c->h_call_ends_cap = 0;
        if (!cuda_ok(c, cudaMallocHost((void**)&c->h_call_ends, (n_frames + 1024) * sizeof(u64)), "cudaMallocHost")) return TRPX_ERR_NOMEM;
        c->h_call_ends_cap = n_frames + 1024;
    }
    if (!ensure(c, c->d_call_status, n_batches * sizeof(u32))) return TRPX_ERR_NOMEM;
    c->call_status.assign(n_batches, 0);
    size_t issued = 0;
    // A pageable destination is filled by helper threads (d2h_staged), one batch behind the batch being issued: the GPU
    // works on batch b while the host drains batch b - 1, and a lane's buffer is free again as soon as its drain returns.
    const bool staged_out = c->stage_threads > 0 && n_frames * frame_raw >= 2 * STAGE_CHUNK && is_pageable(out);
    auto drain_staged = [&](size_t b) {
        Lane& l = c->lanes[b % (size_t)nl];
        const size_t nf = (b * fpb + fpb <= n_frames) ? fpb : n_frames - b * fpb;
        if (!cuda_ok(c, d2h_staged(c, (uint8_t*)out + b * fpb * frame_raw, l.d_out.p, nf * frame_raw, l.ev_done), "D2H pixels (staged)") && rc == TRPX_OK)
            rc = TRPX_ERR_CUDA;
    };

    for (size_t b = 0; b < n_batches && rc == TRPX_OK; ++b) {
        const int li = (int)(b % nl);
        Lane& l = c->lanes[li];
        const size_t f0 = first_frame + b * fpb, nf = (b * fpb + fpb <= n_frames) ? fpb : n_frames - b * fpb;
        const u64 slab0 = f0 ? ends[f0 - 1] : 0, slab1 = ends[f0 + nf - 1];
        const size_t slab = (size_t)(slab1 - slab0);
        if (slab1 > payload_bytes || slab1 <= slab0) { rc = TRPX_ERR_MALFORMED; break; }
        if (!ensure(c, l.d_in, slab + 32) || !ensure(c, l.d_out, nf * frame_raw + 16) || !ensure(c, l.d_ends, nf * 8)) {
            rc = TRPX_ERR_NOMEM;
            break;
        }
        u64* h_ends = c->h_call_ends + b * fpb;
        for (size_t i = 0; i < nf; ++i) h_ends[i] = ends[f0 + i] - slab0;
        u32 seg, warm, sub_shift;
        walk_geometry(c, slab, nf, (n_values + block - 1) / block, seg, warm, sub_shift);
        DecPlan pl = dec_plan(out_dtype, slab, n_values, nf, block, l.d_out.p, seg, warm, sub_shift);
        if (!pl.ok) { rc = TRPX_ERR_BAD_ARG; break; }
        if (!ensure(c, l.dec_scratch, pl.scratch_bytes)) { rc = TRPX_ERR_NOMEM; break; }
        cudaMemsetAsync((uint8_t*)l.d_in.p + (slab & ~(size_t)15), 0, 32, l.stream);   // defined bytes after the slab
        if (!cuda_ok(c, resident ? cudaMemcpyAsync(l.d_in.p, resident + slab0, slab, cudaMemcpyDeviceToDevice, l.stream)
                                 : h2d_async(c, l.d_in.p, payload + slab0, slab, l.stream), "payload slab") ||
            !cuda_ok(c, cudaMemcpyAsync(l.d_ends.p, h_ends, nf * 8, cudaMemcpyHostToDevice, l.stream), "H2D ends")) {
            rc = TRPX_ERR_CUDA;
            break;
        }
        if (l.drain_pending) {                                   // d_out still holds the lane's previous batch
            cudaStreamWaitEvent(l.stream, l.ev_drained, 0);
            l.drain_pending = false;
        }
        Launcher L = make_launcher(c, l.stream);
        decode_async(L, l.d_in.p, slab, is_signed != 0, block, n_values, nf, (const u64*)l.d_ends.p, nullptr, l.d_out.p,
                     out_dtype, (u32*)c->d_call_status.p + b, l.dec_scratch.p, pl, c->coop_grid);
        if (!cuda_ok(c, L.err, "decode launch")) { rc = TRPX_ERR_CUDA; break; }
        // the pixels leave on the lane's copy stream: the lane's next payload upload does not wait for them
        cudaEventRecord(l.ev_done, l.stream);
        issued = b + 1;
        if (staged_out) {
            if (nl < 2) drain_staged(b);
            else if (b >= 1) drain_staged(b - 1);
            continue;
        }
        cudaStreamWaitEvent(l.copy_stream, l.ev_done, 0);
        cudaMemcpyAsync((uint8_t*)out + b * fpb * frame_raw, l.d_out.p, nf * frame_raw, cudaMemcpyDeviceToHost, l.copy_stream);
        cudaEventRecord(l.ev_drained, l.copy_stream);
        l.drain_pending = true;
    }
    if (staged_out && nl >= 2 && issued >= 1 && rc == TRPX_OK) drain_staged(issued - 1);
    for (int li = 0; li < nl; ++li) {
        Lane& l = c->lanes[li];
        if (!cuda_ok(c, cudaStreamSynchronize(l.stream), "decode batch") && rc == TRPX_OK) rc = TRPX_ERR_CUDA;
        if (!cuda_ok(c, cudaStreamSynchronize(l.copy_stream), "D2H pixels") && rc == TRPX_OK) rc = TRPX_ERR_CUDA;
        l.drain_pending = false;
    }
    if (rc == TRPX_OK && issued) {
        if (!cuda_ok(c, cudaMemcpy(c->call_status.data(), c->d_call_status.p, issued * sizeof(u32), cudaMemcpyDeviceToHost), "D2H status"))
            return TRPX_ERR_CUDA;
        for (size_t b = 0; b < issued; ++b)
            if (c->call_status[b] != 0) { note_device_word(c, c->call_status[b]); return status_of_device_word(c->call_status[b]); }
    }
    return rc;
}

// ------------------------------------------------------------------------------ pinned host memory
int trpx_host_pin(void* p, size_t bytes)
{
    if (!p || !bytes) return TRPX_ERR_BAD_ARG;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return TRPX_ALREADY; }
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? TRPX_ERR_NO_DEVICE : TRPX_ERR_NOMEM; }
    return TRPX_OK;
}

int trpx_host_unpin(void* p)
{
    if (!p) return TRPX_ERR_BAD_ARG;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return TRPX_ERR_BAD_ARG; }
    return TRPX_OK;
}

void* trpx_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (!bytes || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void trpx_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"

// ------------------------------------------------------------------------------ several devices: frame shards
struct trpx_pool {
    std::vector<trpx_ctx*> ctx;
    std::string last_error;
};

namespace {

// frames [lo, hi) of shard i out of g: contiguous ranges, the first (n % g) shards one frame longer
void shard_range(size_t n, size_t g, size_t i, size_t& lo, size_t& hi)
{
    const size_t q = n / g, r = n % g;
    lo = i * q + (i < r ? i : r);
    hi = lo + q + (i < r ? 1 : 0);
}

}  // namespace

extern "C" {

int trpx_pool_create(const int* devices, int n_devices, trpx_pool** out)
{
    if (!out) return TRPX_ERR_BAD_ARG;
    *out = nullptr;
    std::vector<int> devs;
    if (devices) {
        if (n_devices <= 0) return TRPX_ERR_BAD_ARG;
        devs.assign(devices, devices + n_devices);
    } else {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return TRPX_ERR_NO_DEVICE; }
        for (int d = 0; d < n; ++d) devs.push_back(d);
    }
    trpx_pool* p = new trpx_pool();
    for (int d : devs) {
        trpx_ctx* c = nullptr;
        const int rc = trpx_ctx_create(d, &c);
        if (rc != TRPX_OK) { trpx_pool_destroy(p); return rc; }
        p->ctx.push_back(c);
    }
    *out = p;
    return TRPX_OK;
}

void trpx_pool_destroy(trpx_pool* p)
{
    if (!p) return;
    for (trpx_ctx* c : p->ctx) trpx_ctx_destroy(c);
    delete p;
}

int trpx_pool_size(const trpx_pool* p) { return p ? (int)p->ctx.size() : 0; }
int trpx_pool_device(const trpx_pool* p, int i) { return p && i >= 0 && i < (int)p->ctx.size() ? p->ctx[i]->device : -1; }
const char* trpx_pool_last_error(const trpx_pool* p) { return p ? p->last_error.c_str() : ""; }

int trpx_pool_encode_host(trpx_pool* p, const void* pixels, int dtype, size_t n_values, size_t n_frames, unsigned block,
                          uint8_t* out, size_t out_capacity, size_t* frame_bytes, size_t* total_bytes, unsigned* prolix_bits)
{
    if (!p || p->ctx.empty()) return TRPX_ERR_BAD_ARG;
    const size_t sz = dtype_is_pixel(dtype) ? dtype_size(dtype) : 0;
    if (!pixels || !out || !sz || !block || !n_values || !n_frames) return TRPX_ERR_BAD_ARG;
    const size_t g = p->ctx.size() < n_frames ? p->ctx.size() : n_frames;
    if (g == 1) return trpx_encode_host(p->ctx[0], pixels, dtype, n_values, n_frames, block, out, out_capacity, frame_bytes, total_bytes, prolix_bits);
    // every shard encodes into its own slab (worst-case capacity, pinned when the host grants it); the slabs are then
    // concatenated in shard order -- the only cross-shard step, and it runs on the host
    struct Shard { size_t lo = 0, hi = 0, total = 0, cap = 0; unsigned pb = 0; int rc = TRPX_OK; uint8_t* slab = nullptr; bool pinned = false; std::vector<size_t> fb; };
    std::vector<Shard> sh(g);
    std::vector<std::thread> th;
    for (size_t i = 0; i < g; ++i) {
        Shard& s = sh[i];
        shard_range(n_frames, g, i, s.lo, s.hi);
        s.cap = trpx_max_compressed_bytes(n_values, dtype, block, s.hi - s.lo);
        s.fb.resize(s.hi - s.lo);
        if (i == 0 && s.cap <= out_capacity) s.slab = out;                         // the first slab is already in place
        else if ((s.slab = (uint8_t*)trpx_host_alloc(s.cap)) != nullptr) s.pinned = true;
        else s.slab = (uint8_t*)malloc(s.cap);
        if (!s.slab) s.rc = TRPX_ERR_NOMEM;
    }
    for (size_t i = 0; i < g; ++i)
        th.emplace_back([&, i] {
            Shard& s = sh[i];
            if (s.rc != TRPX_OK) return;
            s.rc = trpx_encode_host(p->ctx[i], (const uint8_t*)pixels + s.lo * n_values * sz, dtype, n_values, s.hi - s.lo, block,
                                    s.slab, s.cap, s.fb.data(), &s.total, &s.pb);
        });
    for (auto& t : th) t.join();
    int rc = TRPX_OK;
    size_t off = 0;
    unsigned pb = 0;
    for (size_t i = 0; i < g && rc == TRPX_OK; ++i) {
        Shard& s = sh[i];
        if (s.rc != TRPX_OK) { rc = s.rc; p->last_error = "device " + std::to_string(p->ctx[i]->device) + ": " + p->ctx[i]->last_error; break; }
        if (off + s.total > out_capacity) { rc = TRPX_ERR_CAPACITY; break; }
        if (s.slab != out + off) memmove(out + off, s.slab, s.total);
        if (frame_bytes) memcpy(frame_bytes + s.lo, s.fb.data(), s.fb.size() * sizeof(size_t));
        off += s.total;
        if (s.pb > pb) pb = s.pb;
    }
    for (size_t i = 0; i < g; ++i) {
        Shard& s = sh[i];
        if (s.slab && s.slab != out) { if (s.pinned) trpx_host_free(s.slab); else free(s.slab); }
    }
    if (rc != TRPX_OK) return rc;
    if (total_bytes) *total_bytes = off;
    if (prolix_bits) *prolix_bits = pb;
    return TRPX_OK;
}

int trpx_pool_decode_host(trpx_pool* p, const uint8_t* payload, size_t payload_bytes, int is_signed, unsigned block,
                          size_t n_values, size_t total_frames, size_t first_frame, size_t n_frames, const size_t* frame_bytes,
                          size_t* frame_bytes_out, void* out, int out_dtype)
{
    if (!p || p->ctx.empty()) return TRPX_ERR_BAD_ARG;
    const size_t so = dtype_size(out_dtype);
    if (!payload || !payload_bytes || !out || !so || !block || !n_values || !total_frames || !n_frames || first_frame + n_frames > total_frames)
        return TRPX_ERR_BAD_ARG;
    const size_t g = p->ctx.size() < n_frames ? p->ctx.size() : n_frames;
    std::vector<size_t> recovered;
    if (!frame_bytes && total_frames > 1) {
        // the container does not store frame boundaries (Terse.hpp:459): the first device recovers them, together with
        // its own share of the frames; the other shards then know where their slabs start
        recovered.resize(total_frames);
        size_t lo, hi;
        shard_range(n_frames, g, 0, lo, hi);
        const int rc = trpx_decode_host(p->ctx[0], payload, payload_bytes, is_signed, block, n_values, total_frames, first_frame + lo, hi - lo,
                                        nullptr, recovered.data(), out, out_dtype);
        if (rc != TRPX_OK) { p->last_error = p->ctx[0]->last_error; return rc; }
        frame_bytes = recovered.data();
        if (frame_bytes_out) memcpy(frame_bytes_out, recovered.data(), total_frames * sizeof(size_t));
        if (g == 1) return TRPX_OK;
    } else if (g == 1) {
        return trpx_decode_host(p->ctx[0], payload, payload_bytes, is_signed, block, n_values, total_frames, first_frame, n_frames, frame_bytes,
                                frame_bytes_out, out, out_dtype);
    } else if (frame_bytes_out && frame_bytes) {
        memcpy(frame_bytes_out, frame_bytes, total_frames * sizeof(size_t));
    }
    std::vector<int> rcs(g, TRPX_OK);
    std::vector<std::thread> th;
    for (size_t i = recovered.empty() ? 0 : 1; i < g; ++i)
        th.emplace_back([&, i] {
            size_t lo, hi;
            shard_range(n_frames, g, i, lo, hi);
            rcs[i] = trpx_decode_host(p->ctx[i], payload, payload_bytes, is_signed, block, n_values, total_frames, first_frame + lo, hi - lo,
                                      frame_bytes, nullptr, (uint8_t*)out + lo * n_values * so, out_dtype);
        });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < g; ++i)
        if (rcs[i] != TRPX_OK) { p->last_error = "device " + std::to_string(p->ctx[i]->device) + ": " + p->ctx[i]->last_error; return rcs[i]; }
    return TRPX_OK;
}

}  // extern "C"
