// `.mofreak` text rows, byte for byte what MoFREAKUtilities::writeMoFREAKFeaturesToFile produces
// (MoFREAKUtilities.cpp:691-719): `x y frame scale motion_x motion_y a0..a7 m0..m7`, every token followed by one
// space, then '\n'; floats through the default ostream operator<< (== printf "%g", 6 significant digits); motion_x
// and motion_y are always 0 (:476-477).
//
// "%g" of a float is reproduced with integer arithmetic only.  A finite float is m * 2^e exactly (m < 2^24); for
// 1e-4 <= |v| < 1e6 the six significant digits are round-half-even(m * 10^(5-d) * 2^e) with d = floor(log10 |v|),
// which fits 64 bits (m * 10^9 < 2^54, and 2^e only shifts right there), so the result is the correctly rounded
// decimal that glibc prints.  Values outside that range (may) need exponent notation; the formatter reports them
// and the caller formats such rows on the host.  Keypoint coordinates and sizes never leave the range.
// Also compiles as plain C++ for the CPU harness in tests/.
#pragma once
#include <stdint.h>

#include "geom.h"

namespace mfk {

// TEXT_ROW_MAX = 128 (geom.h): 3 floats <= 12 chars, frame <= 11, "0 0", 16 bytes <= 3 chars, 22 spaces, '\n' -> 120

struct TextRow { float x, y; int32_t frame; float scale; uint8_t appearance[8]; uint8_t motion[8]; };   // == mfk_row

struct CountSink {                       // first pass: lengths only
    int n = 0;
    __host__ __device__ void put(char) { n++; }
};
struct BufSink {                         // second pass: bytes
    char* p; int n = 0;
    __host__ __device__ explicit BufSink(char* out) : p(out) {}
    __host__ __device__ void put(char c) { p[n++] = c; }
};

template <class Sink>
__host__ __device__ inline void put_uint(Sink& s, uint32_t v) {
    char d[10];
    int n = 0;
    do { d[n++] = (char)('0' + v % 10u); v /= 10u; } while (v);
    while (n) s.put(d[--n]);
}

template <class Sink>
__host__ __device__ inline void put_int(Sink& s, int32_t v) {
    if (v < 0) { s.put('-'); put_uint(s, 0u - (uint32_t)v); }
    else put_uint(s, (uint32_t)v);
}

// printf("%g", (double)v); false if v needs exponent notation (or is not finite)
template <class Sink>
__host__ __device__ inline bool put_float_g(Sink& s, float v) {
    uint32_t bits;
#ifdef __CUDA_ARCH__
    bits = __float_as_uint(v);
#else
    { union { float f; uint32_t u; } pun; pun.f = v; bits = pun.u; }
#endif
    if (bits >> 31) s.put('-');
    const uint32_t mag = bits & 0x7fffffffu;
    if (mag == 0) { s.put('0'); return true; }
    const int be = (int)(mag >> 23);
    // 1e-4 <= |v| < 1e6 as bit patterns: 0x38d1b718 is the smallest float >= 1e-4 (1e-4f itself lies just below), 0x49742400 == 1e6f
    if (mag < 0x38d1b718u || mag >= 0x49742400u || be == 0) return false;
    const uint64_t m = (uint64_t)((mag & 0x7fffffu) | 0x800000u);
    const int sh = 150 - be;                              // |v| = m >> sh, 4 <= sh <= 37
    // d = floor(log10 |v|): compare m * 2^-sh with 10^k exactly, as integers: |v| >= 10^k  <=>  m >= 10^k << sh (k >= 0),
    // m * 10^-k >= 1 << sh (k < 0)
    int d = -4;
    {
        const uint64_t one = (uint64_t)1 << sh;
        const uint64_t p10[6] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull};
        if (m >= one) { d = 0; for (int k = 1; k <= 5; k++) if (m >= p10[k] * one) d = k; }   // 10^5 << 37 < 2^54
        else { for (int k = 1; k <= 3; k++) if (m * p10[4 - k] >= one) { d = -4 + k; } }      // m * 10^(3..1) >= 2^sh -> d = -3..-1
    }
    // six digits N = round-half-even(m * 10^(5-d) / 2^sh)
    uint64_t scale10 = 1;
    for (int k = 0; k < 5 - d; k++) scale10 *= 10ull;
    const uint64_t prod = m * scale10;
    uint64_t N = prod >> sh;
    const uint64_t rem = prod & (((uint64_t)1 << sh) - 1), half = (uint64_t)1 << (sh - 1);
    if (rem > half || (rem == half && (N & 1))) N++;
    if (N == 1000000ull) { N = 100000ull; d++; }
    if (d > 5) return false;                              // rounds up to 1e+06
    char dig[6];
    uint32_t t = (uint32_t)N;
    for (int k = 5; k >= 0; k--) { dig[k] = (char)('0' + t % 10u); t /= 10u; }
    int last = 5;                                         // %g strips trailing zeros (and a bare '.')
    while (last > 0 && dig[last] == '0') last--;
    if (d >= 0) {
        for (int k = 0; k <= d; k++) s.put(dig[k]);       // zeros inside the integer part are kept: k <= d is always emitted
        if (last > d) { s.put('.'); for (int k = d + 1; k <= last; k++) s.put(dig[k]); }
    } else {
        s.put('0'); s.put('.');
        for (int k = 0; k < -d - 1; k++) s.put('0');
        for (int k = 0; k <= last; k++) s.put(dig[k]);
    }
    return true;
}

// one row; returns false if a float of the row needs exponent notation
template <class Sink>
__host__ __device__ inline bool put_row(Sink& s, const TextRow& r) {
    bool ok = put_float_g(s, r.x);       s.put(' ');
    ok &= put_float_g(s, r.y);           s.put(' ');
    put_int(s, r.frame);                 s.put(' ');
    ok &= put_float_g(s, r.scale);       s.put(' ');
    s.put('0'); s.put(' '); s.put('0'); s.put(' ');
    for (int i = 0; i < 8; i++) { put_uint(s, r.appearance[i]); s.put(' '); }
    for (int i = 0; i < 8; i++) { put_uint(s, r.motion[i]); s.put(' '); }
    s.put('\n');
    return ok;
}

}  // namespace mfk
