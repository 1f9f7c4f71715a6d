/*
 * terse_oracle.c -- TEST INFRASTRUCTURE ONLY (see terse_oracle.h).
 *
 * CPU restatement of the TERSE/PROLIX bitstream, written from the format description
 *  and following, function by function:
 *   orc_encode_frame   <- Terse::f_compress            include/Terse.hpp:500-549
 *   block_width        <- OR-reduce + f_highest_set_bit include/Terse.hpp:508-515, :551-560
 *   put_header         <- header emit                  include/Terse.hpp:517-535
 *   put_bits           <- Bit_range::append_range      include/Bit_pointer.hpp:700-730
 *   orc_decode_frame   <- Terse::prolix(Iterator)      include/Terse.hpp:352-389
 *   get_bits / convert <- Bit_range::get_range         include/Bit_pointer.hpp:742-792
 *   orc_header         <- Terse::write                 include/Terse.hpp:454-470
 * It is deliberately simple (byte-wise bit I/O) -- it is the checker, never the product.
 */
#include "terse_oracle.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

static int dtype_signed(int dtype) { return dtype >= ORC_I8; }

size_t orc_dtype_size(int dtype)
{
    switch (dtype) {
    case ORC_U8: case ORC_I8: return 1;
    case ORC_U16: case ORC_I16: return 2;
    case ORC_U32: case ORC_I32: return 4;
    case ORC_U64: case ORC_I64: return 8;
    default: return 0;
    }
}

size_t orc_max_frame_bytes(size_t n, int dtype, unsigned block)
{
    size_t w = 8 * orc_dtype_size(dtype) + (dtype_signed(dtype) ? 1 : 0);   /* signed T_MIN needs W+1 bits */
    size_t nblocks = (n + block - 1) / block;
    size_t bits = 12 * nblocks + n * w;
    return (bits + 7) / 8 + 1;
}

/* value i of a frame as (sign-extended) 64-bit pattern */
static int64_t load_val(const void* px, int dtype, size_t i)
{
    switch (dtype) {
    case ORC_U8:  return (int64_t)((const uint8_t*)px)[i];
    case ORC_U16: return (int64_t)((const uint16_t*)px)[i];
    case ORC_U32: return (int64_t)((const uint32_t*)px)[i];
    case ORC_U64: return (int64_t)((const uint64_t*)px)[i];
    case ORC_I8:  return (int64_t)((const int8_t*)px)[i];
    case ORC_I16: return (int64_t)((const int16_t*)px)[i];
    case ORC_I32: return (int64_t)((const int32_t*)px)[i];
    default:      return ((const int64_t*)px)[i];
    }
}

static unsigned bitlen64(uint64_t m)
{
    unsigned r = 0;
    for (; m; m >>= 1) ++r;                          /* Terse.hpp:556-558 */
    return r;
}

/* Terse.hpp:508-515 + :551-560.  The reference accumulates |v| in T itself, so the magnitude is
 * taken modulo 2^W (T_MIN -> 2^(W-1) -> width W+1). */
static unsigned block_width(const void* px, int dtype, size_t from, size_t to)
{
    unsigned W = (unsigned)(8 * orc_dtype_size(dtype));
    uint64_t tmask = (W == 64) ? ~0ull : ((1ull << W) - 1);
    uint64_t m = 0;
    for (size_t i = from; i < to; ++i) {
        int64_t v = load_val(px, dtype, i);
        if (dtype_signed(dtype)) {
            uint64_t mag = v < 0 ? (0ull - (uint64_t)v) : (uint64_t)v;
            m |= mag & tmask;
        } else {
            m |= (uint64_t)v & tmask;
        }
    }
    unsigned s = bitlen64(m);
    if (dtype_signed(dtype) && m != 0) s += 1;
    return s;
}

typedef struct { uint8_t* p; uint64_t pos; } bitw;

/* LSB-first little-endian bit stream over bytes (Bit_pointer.hpp:438, :490, :700-730) */
static void put_bits(bitw* w, uint64_t v, unsigned n)
{
    while (n) {
        unsigned off = (unsigned)(w->pos & 7);
        unsigned k = 8 - off;
        if (k > n) k = n;
        w->p[w->pos >> 3] |= (uint8_t)((v & ((1u << k) - 1)) << off);
        v >>= k;
        n -= k;
        w->pos += k;
    }
}

/* Terse.hpp:517-535: '1' | 0+3 bits | 0 111 + 2 bits | 0 111 11 + 6 bits */
static void put_header(bitw* w, unsigned s, unsigned* prev)
{
    if (s == *prev) { put_bits(w, 1, 1); return; }
    if (s < 7)        put_bits(w, (uint64_t)s << 1, 4);
    else if (s < 10)  put_bits(w, (uint64_t)(0x7u | ((s - 7) << 3)) << 1, 6);
    else              put_bits(w, (uint64_t)(0x1Fu | ((s - 10) << 5)) << 1, 12);
    *prev = s;
}

size_t orc_encode_frame(const void* pixels, int dtype, size_t n, unsigned block,
                        uint8_t* out, unsigned* prolix_bits)
{
    memset(out, 0, orc_max_frame_bytes(n, dtype, block));
    bitw w = { out, 0 };
    unsigned prev = 0;                                 /* Terse.hpp:505 */
    for (size_t from = 0; from < n; from += block) {
        size_t to = from + block < n ? from + block : n;
        unsigned s = block_width(pixels, dtype, from, to);
        if (prolix_bits && s > *prolix_bits) *prolix_bits = s;   /* Terse.hpp:516 */
        put_header(&w, s, &prev);
        if (s == 0) continue;
        for (size_t i = from; i < to; ++i) {
            int64_t v = load_val(pixels, dtype, i);
            if (s <= 64) {
                uint64_t mask = (s == 64) ? ~0ull : ((1ull << s) - 1);
                put_bits(&w, (uint64_t)v & mask, s);   /* two's-complement truncation, App. A step 3 */
            } else {                                   /* s == 65: int64 minimum */
                put_bits(&w, (uint64_t)v, 64);
                put_bits(&w, v < 0 ? 1u : 0u, 1);
            }
        }
    }
    return 1 + (size_t)(w.pos >> 3);                   /* Terse.hpp:547 */
}

size_t orc_encode_stack(const void* pixels, int dtype, size_t n, size_t n_frames, unsigned block,
                        uint8_t* out, size_t* per_frame_bytes, unsigned* prolix_bits)
{
    size_t sz = orc_dtype_size(dtype), total = 0;
    for (size_t f = 0; f < n_frames; ++f) {
        size_t b = orc_encode_frame((const uint8_t*)pixels + f * n * sz, dtype, n, block,
                                    out + total, prolix_bits);
        if (per_frame_bytes) per_frame_bytes[f] = b;
        total += b;
    }
    return total;
}

typedef struct { const uint8_t* p; uint64_t pos; uint64_t end; int overrun; } bitr;

static uint64_t get_bits(bitr* r, unsigned n)
{
    uint64_t v = 0;
    unsigned got = 0;
    if (r->pos + n > r->end) { r->overrun = 1; r->pos += n; return 0; }
    while (got < n) {
        unsigned off = (unsigned)(r->pos & 7);
        unsigned k = 8 - off;
        if (k > n - got) k = n - got;
        uint64_t bits = ((uint64_t)r->p[r->pos >> 3] >> off) & ((1u << k) - 1);
        v |= bits << got;
        got += k;
        r->pos += k;
    }
    return v;
}

/* Terse.hpp:361-372 */
static unsigned get_header(bitr* r, unsigned s)
{
    if (get_bits(r, 1) == 0) {
        s = (unsigned)get_bits(r, 3);
        if (s == 7) {
            s += (unsigned)get_bits(r, 2);
            if (s == 10) s += (unsigned)get_bits(r, 6);
        }
    }
    return s;
}

static void store_val(void* out, int out_dtype, size_t i, int64_t sv, uint64_t uv, int is_signed,
                      unsigned s)
{
    unsigned Wo = (unsigned)(8 * orc_dtype_size(out_dtype));
    if (s > Wo) {                                      /* clamp path, Bit_pointer.hpp:747-763 */
        if (!dtype_signed(out_dtype)) {
            uint64_t hi = (Wo == 64) ? ~0ull : ((1ull << Wo) - 1);
            uint64_t x = is_signed ? (sv < 0 ? 0 : (uint64_t)sv) : uv;
            uv = x > hi ? hi : x;
            sv = (int64_t)uv;
        } else {
            int64_t lo = -(int64_t)(1ull << (Wo - 1)), hi = (int64_t)((1ull << (Wo - 1)) - 1);
            int64_t x = is_signed ? sv : (uv > (uint64_t)INT64_MAX ? INT64_MAX : (int64_t)uv);
            sv = x < lo ? lo : (x > hi ? hi : x);
            uv = (uint64_t)sv;
        }
    }
    switch (out_dtype) {
    case ORC_U8:  ((uint8_t*)out)[i]  = (uint8_t)uv; break;
    case ORC_U16: ((uint16_t*)out)[i] = (uint16_t)uv; break;
    case ORC_U32: ((uint32_t*)out)[i] = (uint32_t)uv; break;
    case ORC_U64: ((uint64_t*)out)[i] = uv; break;
    case ORC_I8:  ((int8_t*)out)[i]   = (int8_t)sv; break;
    case ORC_I16: ((int16_t*)out)[i]  = (int16_t)sv; break;
    case ORC_I32: ((int32_t*)out)[i]  = (int32_t)sv; break;
    default:      ((int64_t*)out)[i]  = sv; break;
    }
}

size_t orc_decode_frame(const uint8_t* in, size_t in_bytes, int is_signed, unsigned block,
                        size_t n, void* out, int out_dtype)
{
    bitr r = { in, 0, (uint64_t)in_bytes * 8, 0 };
    unsigned s = 0;                                    /* Terse.hpp:359 */
    for (size_t from = 0; from < n; from += block) {
        size_t to = from + block < n ? from + block : n;
        s = get_header(&r, s);
        for (size_t i = from; i < to; ++i) {
            uint64_t uv = 0;
            int64_t sv = 0;
            if (s > 0) {
                if (s <= 64) {
                    uv = get_bits(&r, s);
                    if (is_signed && s < 64 && (uv >> (s - 1)) & 1)   /* Bit_pointer.hpp:784-789 */
                        uv |= ~0ull << s;
                } else {
                    uv = get_bits(&r, 64);
                    (void)get_bits(&r, s - 64);        /* bits above 64 only repeat the sign */
                }
                sv = (int64_t)uv;
            }
            store_val(out, out_dtype, i, sv, uv, is_signed, s);
        }
        if (r.overrun) return 0;
    }
    size_t used = 1 + (size_t)(r.pos >> 3);
    return used <= in_bytes ? used : 0;
}

size_t orc_frame_widths(const uint8_t* in, size_t in_bytes, unsigned block, size_t n,
                        uint8_t* widths)
{
    bitr r = { in, 0, (uint64_t)in_bytes * 8, 0 };
    unsigned s = 0;
    size_t b = 0;
    for (size_t from = 0; from < n; from += block, ++b) {
        size_t cnt = from + block < n ? block : n - from;
        s = get_header(&r, s);
        widths[b] = (uint8_t)s;
        r.pos += (uint64_t)s * cnt;
        if (r.overrun || r.pos > r.end) return 0;
    }
    size_t used = 1 + (size_t)(r.pos >> 3);
    return used <= in_bytes ? used : 0;
}

size_t orc_header(char* buf, size_t buf_size, unsigned prolix_bits, int is_signed, unsigned block,
                  size_t memory_size, size_t number_of_values, const size_t* dims, size_t n_dims,
                  size_t number_of_frames)
{
    size_t k = 0;
    int w = snprintf(buf, buf_size,
                     "<Terse prolix_bits=\"%u\" signed=\"%d\" block=\"%u\" memory_size=\"%zu\""
                     " number_of_values=\"%zu\"",
                     prolix_bits, is_signed ? 1 : 0, block, memory_size, number_of_values);
    if (w < 0 || (size_t)w >= buf_size) return 0;
    k = (size_t)w;
    if (n_dims) {
        w = snprintf(buf + k, buf_size - k, " dimensions=\"");
        if (w < 0 || (size_t)w >= buf_size - k) return 0;
        k += (size_t)w;
        for (size_t i = 0; i < n_dims; ++i) {
            w = snprintf(buf + k, buf_size - k, i + 1 == n_dims ? "%zu\"" : "%zu ", dims[i]);
            if (w < 0 || (size_t)w >= buf_size - k) return 0;
            k += (size_t)w;
        }
    }
    w = snprintf(buf + k, buf_size - k, " number_of_frames=\"%zu\"/>", number_of_frames);
    if (w < 0 || (size_t)w >= buf_size - k) return 0;
    return k + (size_t)w;
}

uint64_t orc_fnv1a64(const uint8_t* p, size_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

/* ------------------------------------------------------------------ synthetic inputs */

static uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static void store_raw(void* out, int dtype, size_t i, int64_t v)
{
    switch (dtype) {
    case ORC_U8:  ((uint8_t*)out)[i]  = (uint8_t)v; break;
    case ORC_U16: ((uint16_t*)out)[i] = (uint16_t)v; break;
    case ORC_U32: ((uint32_t*)out)[i] = (uint32_t)v; break;
    case ORC_U64: ((uint64_t*)out)[i] = (uint64_t)v; break;
    case ORC_I8:  ((int8_t*)out)[i]   = (int8_t)v; break;
    case ORC_I16: ((int16_t*)out)[i]  = (int16_t)v; break;
    case ORC_I32: ((int32_t*)out)[i]  = (int32_t)v; break;
    default:      ((int64_t*)out)[i]  = v; break;
    }
}

/* SURVEY App. B generator */
void orc_kat_fill(void* out, int dtype, size_t n, uint64_t S)
{
    unsigned W = (unsigned)(8 * orc_dtype_size(dtype));
    for (size_t i = 0; i < n; ++i) {
        uint64_t hb = splitmix64(S * 0x10001ull + i / 12);
        unsigned sh = (unsigned)(hb % (W + 1));
        uint64_t h = splitmix64(S ^ ((uint64_t)i * 0x9E37ull + 1));
        if (!dtype_signed(dtype)) {
            uint64_t m = (W == 64) ? ~0ull : ((1ull << W) - 1);
            uint64_t v = (sh == W) ? 0 : (h & m) >> sh;
            store_raw(out, dtype, i, (int64_t)v);
        } else {
            uint64_t m = (1ull << (W - 2)) - 1;
            uint64_t mag = (sh >= W - 2) ? 0 : (h & m) >> sh;
            store_raw(out, dtype, i, (h >> 63) ? -(int64_t)mag : (int64_t)mag);
        }
    }
}

typedef struct { uint64_t s; } rng_t;
static uint64_t rng_next(rng_t* r) { r->s += 0x9E3779B97F4A7C15ull; return splitmix64(r->s - 0x9E3779B97F4A7C15ull); }
static double rng_uniform(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* Knuth / inverse-CDF Poisson for small lambda, normal approximation above 64 */
static uint64_t rng_poisson(rng_t* r, double lambda)
{
    if (lambda <= 0) return 0;
    if (lambda < 64.0) {
        double p = exp(-lambda), cdf = p, u = rng_uniform(r);
        uint64_t k = 0;
        while (u > cdf && k < 1000) { ++k; p *= lambda / (double)k; cdf += p; }
        return k;
    }
    double u1 = rng_uniform(r), u2 = rng_uniform(r);
    if (u1 < 1e-300) u1 = 1e-300;
    double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    double v = lambda + sqrt(lambda) * z + 0.5;
    return v < 0 ? 0 : (uint64_t)v;
}

static double rng_normal(rng_t* r)
{
    double u1 = rng_uniform(r), u2 = rng_uniform(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

void orc_synth_frame(void* out, int dtype, size_t width, size_t height, double lambda,
                     unsigned n_peaks, double amp_lo, double amp_hi, uint64_t seed)
{
    size_t n = width * height;
    unsigned W = (unsigned)(8 * orc_dtype_size(dtype));
    rng_t r = { splitmix64(seed) };
    if (dtype_signed(dtype)) {
        int64_t dark = (int64_t)floor(lambda + 0.5);
        for (size_t i = 0; i < n; ++i) {
            int64_t v = (int64_t)rng_poisson(&r, lambda) - dark + (int64_t)floor(2.0 * rng_normal(&r) + 0.5);
            store_raw(out, dtype, i, v);
        }
        return;
    }
    uint64_t vmax = (W == 64) ? ~0ull : ((1ull << W) - 1);
    for (size_t i = 0; i < n; ++i) store_raw(out, dtype, i, (int64_t)rng_poisson(&r, lambda));
    for (unsigned p = 0; p < n_peaks; ++p) {
        double cx = rng_uniform(&r) * (double)width, cy = rng_uniform(&r) * (double)height;
        double sigma = 1.0 + rng_uniform(&r);
        double amp = amp_lo * exp(rng_uniform(&r) * log(amp_hi / amp_lo));
        long x0 = (long)floor(cx - 4 * sigma), x1 = (long)ceil(cx + 4 * sigma);
        long y0 = (long)floor(cy - 4 * sigma), y1 = (long)ceil(cy + 4 * sigma);
        for (long y = y0; y <= y1; ++y) {
            if (y < 0 || y >= (long)height) continue;
            for (long x = x0; x <= x1; ++x) {
                if (x < 0 || x >= (long)width) continue;
                double d2 = ((double)x + 0.5 - cx) * ((double)x + 0.5 - cx) +
                            ((double)y + 0.5 - cy) * ((double)y + 0.5 - cy);
                double mean = amp * exp(-d2 / (2 * sigma * sigma));
                if (mean < 1e-3) continue;
                size_t i = (size_t)y * width + (size_t)x;
                uint64_t cur = (uint64_t)load_val(out, dtype, i);
                if (W < 64) cur &= vmax;
                uint64_t add = rng_poisson(&r, mean);
                uint64_t v = cur + add;
                if (v > vmax || v < cur) v = vmax;
                store_raw(out, dtype, i, (int64_t)v);
            }
        }
    }
}

/* FNV-1a-64 of each frame of a payload (frame f = bytes [ends[f-1], ends[f])): the checker's side of the full-scale
 * byte-identity test against the reference's per-frame digests (oracle/ref_shim.cpp: ref_frame_digests). */
void orc_fnv64_frames(const uint8_t* payload, const uint64_t* ends, size_t n_frames, uint64_t* out)
{
    uint64_t b = 0;
    for (size_t f = 0; f < n_frames; ++f) {
        uint64_t x = 0xcbf29ce484222325ull;
        for (uint64_t i = b; i < ends[f]; ++i) { x ^= payload[i]; x *= 0x100000001b3ull; }
        out[f] = x;
        b = ends[f];
    }
}
