/* TEST INFRASTRUCTURE — part of the parity oracle, never linked into the product library.
 *
 * From-spec restatement of libfor (frame-of-reference bit packing), the third-party C library the reference uses
 * for every compressed integer container (include/array_base.h:12 METADATA_OVERHEAD 5 = u32 base + u8 bits).
 * libfor is pinned by the reference at github.com/cruppstahl/libfor @ 49611808d08d4e47116aa2a3ddcabeb418f405f7
 * (cmake/For.cmake:3, WORKSPACE:165-170) and is absent from /root/reference and from this machine.
 *
 * Published format restated here: header [u32 base (min value)][u8 bits], then (value - base) packed LSB-first
 * with `bits` bits per value. libfor packs groups of 32/16/8 values byte-aligned and the <8 remainder as a
 * plain bit stream; because 32*b, 16*b and 8*b are all multiples of 8, that is byte-for-byte a contiguous LSB-first
 * bit stream, which is what this file writes. Only round-trip behaviour is observable through the reference's
 * sorted_array/array classes; no reference test asserts raw bytes, so "parity" here means the
 * reference's own container tests pass on top of this port (tests/test_oracle_ref.py).
 */
#include <stdint.h>
#include <string.h>

#define FOR_HDR 5u

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline void wr32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }

static inline uint32_t bits_required(uint32_t v) { return v == 0 ? 0u : 32u - (uint32_t)__builtin_clz(v); }

uint32_t for_compressed_size_bits(uint32_t length, uint32_t bits) {
    return (uint32_t)(((uint64_t)length * bits + 7) / 8);
}

static inline uint32_t get_bits(const uint8_t* data, uint32_t bits, uint32_t index) {
    if (bits == 0) return 0;
    uint64_t bitpos = (uint64_t)index * bits;
    const uint8_t* p = data + (bitpos >> 3);
    uint32_t shift = (uint32_t)(bitpos & 7);
    uint64_t w = 0;
    uint32_t nbytes = (shift + bits + 7) / 8;     /* <= 5 */
    for (uint32_t i = 0; i < nbytes; i++) w |= (uint64_t)p[i] << (8 * i);
    uint64_t mask = bits == 32 ? 0xffffffffull : ((1ull << bits) - 1);
    return (uint32_t)((w >> shift) & mask);
}

static inline void put_bits(uint8_t* data, uint32_t bits, uint32_t index, uint32_t v) {
    if (bits == 0) return;
    uint64_t bitpos = (uint64_t)index * bits;
    uint8_t* p = data + (bitpos >> 3);
    uint32_t shift = (uint32_t)(bitpos & 7);
    uint64_t mask = (bits == 32 ? 0xffffffffull : ((1ull << bits) - 1)) << shift;
    uint64_t val = ((uint64_t)v << shift) & mask;
    uint32_t nbytes = (shift + bits + 7) / 8;
    for (uint32_t i = 0; i < nbytes; i++) {
        uint8_t m = (uint8_t)(mask >> (8 * i));
        p[i] = (uint8_t)((p[i] & ~m) | (uint8_t)(val >> (8 * i)));
    }
}

static uint32_t compress_with(const uint32_t* in, uint8_t* out, uint32_t length, uint32_t base, uint32_t bits) {
    wr32(out, base);
    out[4] = (uint8_t)bits;
    uint32_t nbytes = for_compressed_size_bits(length, bits);
    memset(out + FOR_HDR, 0, nbytes);
    for (uint32_t i = 0; i < length; i++) put_bits(out + FOR_HDR, bits, i, in[i] - base);
    return FOR_HDR + nbytes;
}

uint32_t for_compress_unsorted(const uint32_t* in, uint8_t* out, uint32_t length) {
    if (length == 0) { wr32(out, 0); out[4] = 0; return FOR_HDR; }
    uint32_t m = in[0], M = in[0];
    for (uint32_t i = 1; i < length; i++) { if (in[i] < m) m = in[i]; if (in[i] > M) M = in[i]; }
    return compress_with(in, out, length, m, bits_required(M - m));
}

uint32_t for_compress_sorted(const uint32_t* in, uint8_t* out, uint32_t length) {
    if (length == 0) { wr32(out, 0); out[4] = 0; return FOR_HDR; }
    return compress_with(in, out, length, in[0], bits_required(in[length - 1] - in[0]));
}

uint32_t for_uncompress(const uint8_t* in, uint32_t* out, uint32_t length) {
    uint32_t base = rd32(in), bits = in[4];
    for (uint32_t i = 0; i < length; i++) out[i] = base + get_bits(in + FOR_HDR, bits, i);
    return FOR_HDR + for_compressed_size_bits(length, bits);
}

uint32_t for_select_bits(const uint8_t* in, uint32_t base, uint32_t bits, uint32_t index) {
    return base + get_bits(in, bits, index);
}

uint32_t for_select(const uint8_t* in, uint32_t index) {
    return for_select_bits(in + FOR_HDR, rd32(in), in[4], index);
}

uint32_t for_linear_search(const uint8_t* in, uint32_t length, uint32_t value) {
    uint32_t base = rd32(in), bits = in[4];
    for (uint32_t i = 0; i < length; i++) if (base + get_bits(in + FOR_HDR, bits, i) == value) return i;
    return length;
}

/* first index whose value is >= `value`; *actual = that value. If all values are smaller, returns length-1 with
 * *actual = last value (libfor's documented behaviour, relied on by src/sorted_array.cpp:98-116). */
uint32_t for_lower_bound_search(const uint8_t* in, uint32_t length, uint32_t value, uint32_t* actual) {
    uint32_t base = rd32(in), bits = in[4];
    if (length == 0) { *actual = 0; return 0; }
    uint32_t lo = 0, hi = length;               /* invariant: [0,lo) < value, [hi,len) >= value */
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        uint32_t v = base + get_bits(in + FOR_HDR, bits, mid);
        if (v < value) lo = mid + 1; else hi = mid;
    }
    if (lo >= length) lo = length - 1;
    *actual = base + get_bits(in + FOR_HDR, bits, lo);
    return lo;
}

static uint32_t append_generic(uint8_t* in, uint32_t length, uint32_t value) {
    uint32_t base = rd32(in), bits = in[4];
    if (length == 0) {
        wr32(in, value); in[4] = 0;
        return FOR_HDR;
    }
    /* current max must be recovered to decide the new width */
    uint32_t M = base, m = base;
    int fits = value >= base && bits_required(value - base) <= bits;
    if (fits) {
        uint32_t old_bytes = for_compressed_size_bits(length, bits);
        uint32_t new_bytes = for_compressed_size_bits(length + 1, bits);
        if (new_bytes > old_bytes) memset(in + FOR_HDR + old_bytes, 0, new_bytes - old_bytes);
        put_bits(in + FOR_HDR, bits, length, value - base);
        return FOR_HDR + new_bytes;
    }
    /* re-encode: widen in place, walking from the last element to the first so nothing is overwritten early
       when the base is unchanged; when the base drops we go through the same backwards walk because new offsets
       (v - newbase) still need >= as many bits per slot. */
    for (uint32_t i = 0; i < length; i++) { uint32_t v = base + get_bits(in + FOR_HDR, bits, i); if (v > M) M = v; }
    if (value > M) M = value;
    if (value < m) m = value;
    uint32_t nbits = bits_required(M - m);
    uint32_t new_bytes = for_compressed_size_bits(length + 1, nbits);
    uint32_t old_bytes = for_compressed_size_bits(length, bits);
    if (new_bytes > old_bytes) memset(in + FOR_HDR + old_bytes, 0, new_bytes - old_bytes);
    /* nbits >= bits always here, so slot i of the new stream starts at or after slot i of the old one */
    for (uint32_t k = length; k-- > 0;) {
        uint32_t v = base + get_bits(in + FOR_HDR, bits, k);
        /* clear target then write */
        put_bits(in + FOR_HDR, nbits, k, v - m);
    }
    put_bits(in + FOR_HDR, nbits, length, value - m);
    wr32(in, m); in[4] = (uint8_t)nbits;
    return FOR_HDR + new_bytes;
}

uint32_t for_append_unsorted(uint8_t* in, uint32_t length, uint32_t value) { return append_generic(in, length, value); }
uint32_t for_append_sorted(uint8_t* in, uint32_t length, uint32_t value) { return append_generic(in, length, value); }
