/*
 * trpx_b200.h -- C ABI of the B200-native TERSE/PROLIX codec (libtrpx_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path.  senikm/trpx has no FFI layer of its
 * own -- the path lives inside the header-only class jpa::Terse -- so each entry point below names the
 * reference member it replaces (paths relative to the reference repo):
 *
 *   trpx_encode_host / trpx_encode_device   <- Terse::f_compress            include/Terse.hpp:500-549
 *                                              (+ f_highest_set_bit :551-560, header emit :517-535,
 *                                               Bit_range::append_range  include/Bit_pointer.hpp:700-730,
 *                                               d_prolix_bits running max :516, size rule :547)
 *   trpx_decode_host / trpx_decode_device   <- Terse::prolix(Iterator)      include/Terse.hpp:352-389
 *                                              (+ Bit_range::get_range   include/Bit_pointer.hpp:742-792,
 *                                               f_find_terse_frame :562-585)
 *   trpx_max_compressed_bytes               <- worst-case buffer rule      include/Terse.hpp:502-504
 *
 * The host C++ class include/trpx/Terse.hpp (same public surface as jpa::Terse, Terse.hpp:228-474)
 * sits on top of this ABI; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; the caller owns every buffer; nothing
 * throws across the boundary; every function returns a TRPX_* status (0 = ok).  There is NO CPU
 * fallback: without a usable CUDA device every compute entry point returns TRPX_ERR_NO_DEVICE.
 */
#ifndef TRPX_B200_H
#define TRPX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRPX_ABI_VERSION 2

/* status codes */
enum {
    TRPX_OK = 0,
    TRPX_ERR_BAD_ARG = 1,     /* null pointer, unknown dtype, block == 0, misaligned device buffer ... */
    TRPX_ERR_CAPACITY = 2,    /* output buffer too small for the compressed payload */
    TRPX_ERR_CUDA = 3,        /* a CUDA runtime call or kernel failed (see trpx_last_error) */
    TRPX_ERR_MALFORMED = 4,   /* payload runs past its end / block counts do not add up */
    TRPX_ERR_NO_DEVICE = 5,   /* no CUDA device: this library has no CPU path */
    TRPX_ERR_NOMEM = 6,       /* device or pinned-host allocation failed */
    TRPX_ALREADY = 7          /* trpx_host_pin: the range is pinned already (not an error; do not unpin it) */
};

/* pixel types: std::is_signed_v of the iterator's value_type decides the stream's signedness
 * (Terse.hpp:265, :294) */
enum {
    TRPX_U8 = 0, TRPX_U16 = 1, TRPX_U32 = 2, TRPX_U64 = 3,
    TRPX_I8 = 4, TRPX_I16 = 5, TRPX_I32 = 6, TRPX_I64 = 7,
    /* output types of the DECODER only (the encoder returns TRPX_ERR_BAD_ARG for them): every value goes
     * through a 64-bit integer and a double, as in Terse::prolix for floating-point iterators
     * (Terse.hpp:379-383); signed and unsigned streams are both accepted, nothing is clamped */
    TRPX_F32 = 8, TRPX_F64 = 9
};

typedef struct trpx_ctx trpx_ctx;   /* one per GPU / per host thread: streams, scratch, pinned staging */

/* ---- context ---------------------------------------------------------------------------- */

/* Create a context on CUDA device `device`.  Fails with TRPX_ERR_NO_DEVICE when CUDA is unusable. */
int trpx_ctx_create(int device, trpx_ctx** ctx);
void trpx_ctx_destroy(trpx_ctx* ctx);
int trpx_ctx_device(const trpx_ctx* ctx);
/* Text of the last CUDA / library error seen by this context ("" if none). */
const char* trpx_last_error(const trpx_ctx* ctx);
const char* trpx_strerror(int status);
int trpx_abi_version(void);

/* ---- sizes ------------------------------------------------------------------------------ */

size_t trpx_dtype_size(int dtype);                       /* 1, 2, 4, 8 (TRPX_F32: 4, TRPX_F64: 8); 0 for an unknown code */
int trpx_dtype_is_signed(int dtype);
/* Capacity that always suffices for n_frames frames of n_values each (multiple of 16 bytes).
 * Replaces the reference's worst-case resize (Terse.hpp:502-504), with the App. C7 under-count fixed. */
size_t trpx_max_compressed_bytes(size_t n_values, int dtype, unsigned block, size_t n_frames);

/* ---- TERSE: encode (replaces Terse::f_compress, Terse.hpp:500-549) ------------------------ */

/* Host-pointer flavour: pixels (n_frames x n_values, contiguous, frame-major) -> payload.
 * The payload is byte-identical to the concatenation of the reference's per-frame payloads.
 *   out / out_capacity   caller buffer; trpx_max_compressed_bytes() always suffices
 *   frame_bytes          optional [n_frames]: payload bytes of each frame (1 + floor(bits/8), :547)
 *   total_bytes          payload size (the XML memory_size)
 *   prolix_bits          max block width over all frames (Terse.hpp:516; the XML prolix_bits)   */
int trpx_encode_host(trpx_ctx* ctx, const void* pixels, int dtype, size_t n_values, size_t n_frames,
                     unsigned block, uint8_t* out, size_t out_capacity, size_t* frame_bytes,
                     size_t* total_bytes, unsigned* prolix_bits);

/* Progress of the trpx_encode_host call that is running (or ran last) on `ctx`; callable from ANOTHER host thread
 * while that call runs (it takes no lock the call holds).  The first *frames_done frames are complete: their
 * payload -- *payload_bytes_done bytes from the start of `out` -- and their frame_bytes[] entries have landed in
 * the caller's buffers, so a consumer (a file writer, a decoder on another context) can start on them while the
 * rest of the stack is still being uploaded and encoded.  *call_seq counts the trpx_encode_host calls started on
 * the context: a consumer reads it before it launches the call and ignores progress until it has changed. */
int trpx_ctx_encode_progress(trpx_ctx* ctx, size_t* call_seq, size_t* frames_done, size_t* payload_bytes_done);

/* Device-pointer flavour: everything resident in HBM, asynchronous on `stream` (a cudaStream_t).
 *   d_pixels       16-byte aligned
 *   d_out          16-byte aligned, out_capacity >= trpx_max_compressed_bytes() recommended
 *   d_frame_ends   [n_frames] uint64: END byte offset of each frame inside d_out (inclusive prefix
 *                  sum of the frame sizes; d_frame_ends[n_frames-1] is the payload size)
 *   d_prolix_bits  [1] uint32
 *   d_status       [1] uint32: TRPX_OK or TRPX_ERR_CAPACITY, valid once the stream has drained
 * Concurrent calls on one context must use distinct `lane`s (0 .. trpx_ctx_lanes()-1): a lane owns
 * the launch's scratch (look-back descriptors). */
int trpx_encode_device(trpx_ctx* ctx, int lane, const void* d_pixels, int dtype, size_t n_values,
                       size_t n_frames, unsigned block, uint8_t* d_out, size_t out_capacity,
                       uint64_t* d_frame_ends, uint32_t* d_prolix_bits, uint32_t* d_status,
                       void* stream);

/* ---- PROLIX: decode (replaces Terse::prolix, Terse.hpp:352-389) --------------------------- */

/* Host-pointer flavour.  `payload` holds total_frames frames back to back; frames
 * [first_frame, first_frame + n_frames) are decoded into `out` (n_frames x n_values of out_dtype).
 *   frame_bytes      optional [total_frames] sizes (as returned by trpx_encode_host); when NULL the
 *                    frame boundaries are recovered from the stream (the .trpx header does not
 *                    store them, Terse.hpp:562-585)
 *   frame_bytes_out  optional [total_frames]: receives the recovered sizes
 * Conversion to out_dtype follows Bit_range::get_range (Bit_pointer.hpp:742-792): values are
 * sign-extended from bit s-1 for signed streams; a block wider than the output type is clamped to
 * the type's range; otherwise the value is truncated.  A signed stream into an unsigned type is
 * TRPX_ERR_BAD_ARG (the reference asserts, Terse.hpp:356-357).
 * Deliberate difference: sign extension follows the STREAM's signedness.  The reference sign-extends whenever the
 * OUTPUT type is signed (Bit_pointer.hpp:784-789), which turns unsigned 5 in a 3-bit block into -3; here an unsigned
 * stream decoded into int16 / int32 / int64 keeps its values (tests/test_gpu_parity.py pins this).
 * `payload` / `d_payload` may be read up to 16 bytes past payload_bytes rounded up to 16 (the device flavour needs
 * that much capacity behind the payload; the host flavours stage and zero-pad it themselves); those bytes never
 * influence the result. */
int trpx_decode_host(trpx_ctx* ctx, const uint8_t* payload, size_t payload_bytes, int is_signed,
                     unsigned block, size_t n_values, size_t total_frames, size_t first_frame,
                     size_t n_frames, const size_t* frame_bytes, size_t* frame_bytes_out, void* out,
                     int out_dtype);

/* Device-pointer flavour, asynchronous on `stream`.
 *   d_payload      16-byte aligned payload of n_frames frames
 *   d_frame_ends   [n_frames] uint64 END byte offsets (as written by trpx_encode_device), or NULL
 *                  to recover them from the stream (slower: frames are then resolved one after
 *                  another); when NULL and d_frame_ends_out != NULL the recovered ends are stored
 *   d_out          16-byte aligned, n_frames x n_values of out_dtype
 *   d_status       [1] uint32: TRPX_OK or TRPX_ERR_MALFORMED once the stream has drained */
int trpx_decode_device(trpx_ctx* ctx, int lane, const uint8_t* d_payload, size_t payload_bytes,
                       int is_signed, unsigned block, size_t n_values, size_t n_frames,
                       const uint64_t* d_frame_ends, uint64_t* d_frame_ends_out, void* d_out,
                       int out_dtype, uint32_t* d_status, void* stream);

/* ---- several GPUs of one box: ONE stack sharded by frame (Terse.hpp:25-26, :290-302, :505) -- */

/* Frames are independent: `prevbits` restarts per frame (Terse.hpp:505), frames are byte-aligned (:547), and the
 * only scalars a stack shares are prolix_bits (a max, :516) and memory_size (a sum, :459).  A pool owns one context
 * and one host thread per device; a call cuts the frames into contiguous ranges, one per device, runs the
 * single-device host-pointer pipeline on each range concurrently, and concatenates the per-device slabs, frame sizes
 * and max(prolix_bits) on the host.  There is no collective and no device-to-device traffic.  The results are
 * byte-identical to the single-device calls. */
typedef struct trpx_pool trpx_pool;
/* devices == NULL: every visible CUDA device (n_devices ignored).  TRPX_ERR_NO_DEVICE without a usable GPU. */
int trpx_pool_create(const int* devices, int n_devices, trpx_pool** pool);
void trpx_pool_destroy(trpx_pool* pool);
int trpx_pool_size(const trpx_pool* pool);                    /* number of devices */
int trpx_pool_device(const trpx_pool* pool, int i);           /* CUDA ordinal of shard i */
const char* trpx_pool_last_error(const trpx_pool* pool);
/* Same contract as trpx_encode_host / trpx_decode_host.  trpx_pool_decode_host needs the frame sizes to cut the
 * payload (frame_bytes != NULL); without them the first device recovers them from the stream first. */
int trpx_pool_encode_host(trpx_pool* pool, const void* pixels, int dtype, size_t n_values, size_t n_frames,
                          unsigned block, uint8_t* out, size_t out_capacity, size_t* frame_bytes,
                          size_t* total_bytes, unsigned* prolix_bits);
int trpx_pool_decode_host(trpx_pool* pool, const uint8_t* payload, size_t payload_bytes, int is_signed,
                          unsigned block, size_t n_values, size_t total_frames, size_t first_frame,
                          size_t n_frames, const size_t* frame_bytes, size_t* frame_bytes_out, void* out,
                          int out_dtype);

/* ---- pinned host memory for the host-pointer flavours ------------------------------------- */

/* The host-pointer calls copy with cudaMemcpyAsync; from PAGEABLE memory the driver stages every copy through its
 * own bounce buffer (about a third of the link rate, and the call blocks).  A caller that owns long-lived buffers
 * pins them once: trpx_host_pin page-locks an existing range (cudaHostRegister), trpx_host_alloc returns pinned
 * memory.  Both are optional; every entry point accepts pageable pointers. */
int trpx_host_pin(void* p, size_t bytes);
int trpx_host_unpin(void* p);
void* trpx_host_alloc(size_t bytes);
void trpx_host_free(void* p);

/* ---- introspection (used by bench.py / tests) --------------------------------------------- */

int trpx_ctx_lanes(const trpx_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t trpx_ctx_launch_count(const trpx_ctx* ctx);
/* Bytes of device scratch currently held by the context. */
size_t trpx_ctx_scratch_bytes(const trpx_ctx* ctx);
/* Profiling: when on, the *_device entry points drop a CUDA event on the caller's stream between
 * their kernels.  Once that stream has drained, trpx_ctx_last_kernel_times() returns the device
 * time (ms) of each kernel of the last call on `lane` (names[i]: static strings such as
 * "terse_encode", "prolix_walk", "prolix_unpack"); the return value is the number of entries. */
int trpx_ctx_set_profiling(trpx_ctx* ctx, int on);
int trpx_ctx_last_kernel_times(trpx_ctx* ctx, int lane, const char** names, float* ms, int cap);

#ifdef __cplusplus
}
#endif
#endif /* TRPX_B200_H */
