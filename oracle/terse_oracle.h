/*
 * terse_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference TERSE/PROLIX algorithm
 * (senikm/trpx: include/Terse.hpp:500-560 encode, :352-389 decode, :454-474 header;
 * include/Bit_pointer.hpp:700-730 pack, :742-792 unpack).  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA path.
 * Nothing in the product (trpx_b200/, include/) may include, link or call it.
 *
 * Parity status: PINNED -- checked in tests/test_oracle.py against
 *   (1) the known-answer vectors of SURVEY.md App. B (tests/golden/kat_small.json,
 *       kat_large.json; generated from the reference headers),
 *   (2) the live reference build oracle/_ref/libtrpx_ref.so when present.
 */
#ifndef TRPX_ORACLE_H
#define TRPX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pixel type codes (shared with include/trpx_b200.h; restated here so the oracle is standalone) */
enum {
    ORC_U8 = 0, ORC_U16 = 1, ORC_U32 = 2, ORC_U64 = 3,
    ORC_I8 = 4, ORC_I16 = 5, ORC_I32 = 6, ORC_I64 = 7
};

/* bytes per pixel of a type code, 0 if invalid */
size_t orc_dtype_size(int dtype);

/* worst-case payload bytes of ONE frame (App. C7-safe: ceil((12*nblocks + n*(W+1))/8) + 1) */
size_t orc_max_frame_bytes(size_t n, int dtype, unsigned block);

/* Encode one frame (Terse.hpp:500-549).  `out` must hold orc_max_frame_bytes() zeroed-or-not bytes
 * (the function clears what it uses).  *prolix_bits is raised to the max block width seen
 * (Terse.hpp:516).  Returns the frame's payload size: 1 + floor(total_bits/8) (Terse.hpp:547). */
size_t orc_encode_frame(const void* pixels, int dtype, size_t n, unsigned block,
                        uint8_t* out, unsigned* prolix_bits);

/* Encode a stack: frames concatenated ([probed]).  per_frame_bytes may be NULL. */
size_t orc_encode_stack(const void* pixels, int dtype, size_t n, size_t n_frames, unsigned block,
                        uint8_t* out, size_t* per_frame_bytes, unsigned* prolix_bits);

/* Decode one frame starting at `in` (byte aligned) (Terse.hpp:352-389, App. A "Decode").
 * Values are sign-extended from bit s-1 iff is_signed (stored signedness), then converted to
 * out_dtype: a block whose width s exceeds the output type's bits is clamped to the type's range
 * (Bit_pointer.hpp:747-763), otherwise the value is truncated (C cast).
 * Returns bytes consumed (1 + floor(bits/8)), or 0 if the stream runs past in_bytes. */
size_t orc_decode_frame(const uint8_t* in, size_t in_bytes, int is_signed, unsigned block,
                        size_t n, void* out, int out_dtype);

/* Per-block widths of one frame (used to check the GPU header-resolution pass).
 * widths must hold ceil(n/block) bytes.  Returns bytes consumed or 0 on overrun. */
size_t orc_frame_widths(const uint8_t* in, size_t in_bytes, unsigned block, size_t n,
                        uint8_t* widths);

/* The XML header exactly as Terse.hpp:454-470 prints it.  dims may be NULL (n_dims = 0).
 * Returns the number of characters written (excluding NUL), or 0 if buf is too small. */
size_t orc_header(char* buf, size_t buf_size, unsigned prolix_bits, int is_signed, unsigned block,
                  size_t memory_size, size_t number_of_values, const size_t* dims, size_t n_dims,
                  size_t number_of_frames);

/* FNV-1a-64 (SURVEY App. B checksum of payloads) */
uint64_t orc_fnv1a64(const uint8_t* p, size_t n);

/* ---- reproducible synthetic inputs (SURVEY App. B generator and §8d frames) ---- */

/* App. B KAT generator: fills n values of dtype from seed S. */
void orc_kat_fill(void* out, int dtype, size_t n, uint64_t seed);

/* Diffraction frame: Poisson(lambda) background + n_peaks Gaussian Bragg peaks (sigma in [1,2] px,
 * amplitude log-uniform in [amp_lo, amp_hi], Poisson-sampled), clipped to the type's max.
 * For signed dtypes the frame is "dark-subtracted": Poisson(lambda) - round(lambda) + round(N(0,2^2))
 *  and peaks are ignored.  splitmix64/inverse-CDF, seed = base + frame index. */
void orc_synth_frame(void* out, int dtype, size_t width, size_t height, double lambda,
                     unsigned n_peaks, double amp_lo, double amp_hi, uint64_t seed);

/* FNV-1a-64 of each frame of a payload; ends[] are the frames' end byte offsets */
void orc_fnv64_frames(const uint8_t* payload, const uint64_t* ends, size_t n_frames, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif
