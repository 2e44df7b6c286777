// Per-pixel device functions of the dense stage (pyramid samplers, packed OAST-9/16 score).  Also compiles as
// plain C++ for the CPU harness in tests/ (which supplies scalar stand-ins for the packed min/max intrinsics).
#pragma once
#include "geom.h"

namespace mfk {

__device__ __forceinline__ int avg_up(int a, int b) { return (a + b + 1) >> 1; }   // _mm_avg_epu8

// brisk.cpp:1840-1972: paired 16-byte blocks use avg-of-avg (both rounding up); an odd trailing block a
// truncating pair mean of the vertical averages (:1917-1942); the (cols % 16) / 2 leftover outputs a 4-tap
// truncating mean over columns k, k+1 counted from the end of the blocks (:1958-1964).
// src: rows of `pitch` bytes, `w` of them image columns
__device__ __forceinline__ int half_px(const uint8_t* __restrict__ src, int pitch, int w, int ox, int oy) {
    const uint8_t* up = src + (2 * oy) * pitch;
    const uint8_t* lo = up + pitch;
    const int blocks = w >> 4;
    const int paired_out = (blocks >> 1) * 16;
    if (ox < paired_out)
        return avg_up(avg_up(up[2 * ox], lo[2 * ox]), avg_up(up[2 * ox + 1], lo[2 * ox + 1]));
    int base = 2 * paired_out, o = ox - paired_out;
    if (blocks & 1) {
        if (o < 8)
            return (avg_up(up[base + 2 * o], lo[base + 2 * o]) + avg_up(up[base + 2 * o + 1], lo[base + 2 * o + 1])) >> 1;
        base += 16; o -= 8;
    }
    return (up[base + o] + up[base + o + 1] + lo[base + o] + lo[base + o + 1]) >> 2;
}

// brisk.cpp:1974-2065: blocks of 15 source columns give 10 outputs through avg(avg(a,b),a) chains
// (vertical, then horizontal); the fifth triple takes its middle sample from byte 12 (:1982-1983);
// trailing triples use (4A + 2(B+C) + D) / 9 (:2036-2051).
__device__ __forceinline__ int twothird_px(const uint8_t* __restrict__ src, int w, int ox, int oy) {
    const int rt = oy >> 1;
    const bool lower = oy & 1;
    const uint8_t* r2 = src + (3 * rt + 1) * w;
    const uint8_t* re = lower ? r2 + w : r2 - w;          // the row this output leans on
    const int blocks = w / 15;
    if (ox < 10 * blocks) {
        const int i = ox / 10, j = ox - 10 * i, t = j >> 1;
        const int c_end = 15 * i + 3 * t + ((j & 1) ? 2 : 0);
        const int c_mid = 15 * i + ((t == 4) ? 12 : 3 * t + 1);
        const int e = avg_up(avg_up(re[c_end], r2[c_end]), re[c_end]);
        const int m = avg_up(avg_up(re[c_mid], r2[c_mid]), re[c_mid]);
        return avg_up(avg_up(e, m), e);
    }
    const int o = ox - 10 * blocks;
    const int sc = 15 * blocks + 3 * (o >> 1);
    if (!(o & 1)) return (4 * re[sc] + 2 * (re[sc + 1] + r2[sc]) + r2[sc + 1]) / 9;
    return (4 * re[sc + 2] + 2 * (re[sc + 1] + r2[sc + 2]) + r2[sc + 1]) / 9;
}

// OAST-9/16 closed-form score from the 16 ring words and the centre word, each packed as v | (255 - v) << 16
// (see k_scores): 16 + 16 packed 3-input minima build the 16 arc minima, 8 packed 3-input maxima reduce them.
__device__ __forceinline__ int oast_score_packed(const uint32_t r[16], const uint32_t centre) {
    uint32_t m3[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m3[i] = __vimin3_u16x2(r[i], r[(i + 1) & 15], r[(i + 2) & 15]);
    uint32_t best = 0;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const uint32_t a = __vimin3_u16x2(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]);
        const uint32_t b = __vimin3_u16x2(m3[i + 1], m3[(i + 4) & 15], m3[(i + 7) & 15]);
        best = __vimax3_u16x2(best, a, b);
    }
    const int p = (int)(centre & 0xffffu);
    const int brighter = (int)(best & 0xffffu) - p;
    const int darker = (int)(best >> 16) - (255 - p);
    const int s = max(brighter, darker) - 1;
    return min(254, max(0, s));
}


// Upper bound of the OAST-9/16 score from the 8 antipodal ring pairs: an arc of 9 contiguous ring pixels contains at
// least one pixel of every pair (j, j+8), so  min_arc v - p <= min_j max(v_j, v_{j+8}) - p  (and the mirrored statement
// for the darker test).  In the packed (v, 255 - v) form both bounds come out of 8 packed maxima and 4 packed 3-input
// minima.  A bound <= 1 proves S = 0 without evaluating the 16 arcs.
__device__ __forceinline__ int oast_score_bound_packed(const uint32_t r[16], const uint32_t centre) {
    uint32_t m[8];
#pragma unroll
    for (int j = 0; j < 8; j++) m[j] = __vimax3_u16x2(r[j], r[j + 8], r[j]);
    const uint32_t a = __vimin3_u16x2(m[0], m[1], m[2]), b = __vimin3_u16x2(m[3], m[4], m[5]);
    const uint32_t c = __vimin3_u16x2(m[6], m[7], a);
    const uint32_t lo = __vimin3_u16x2(b, c, c);
    const int p = (int)(centre & 0xffffu);
    const int brighter = (int)(lo & 0xffffu) - p;
    const int darker = (int)(lo >> 16) - (255 - p);
    return max(brighter, darker) - 1;
}

// The same bound for TWO horizontally adjacent centres at once: r[i] and `centre` hold the 16-bit pixels of the left
// centre in the low half and of the right centre in the high half.  Both polarities need their own chain here (8 packed
// max + 8 packed min, then 4 + 4 three-input reductions), which is the same 12 packed operations per pixel as the
// (v, 255 - v) form, but every operand covers two centres, so the shared-memory loads per pixel halve.
// Returns, per half, bound + 256 (the bias keeps the packed subtractions free of borrows between the halves).
__device__ __forceinline__ uint32_t oast_bound_pair(const uint32_t r[16], const uint32_t centre) {
    uint32_t mx[8], mn[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { mx[j] = __vmaxu2(r[j], r[j + 8]); mn[j] = __vminu2(r[j], r[j + 8]); }
    const uint32_t a = __vimin3_u16x2(mx[0], mx[1], mx[2]), b = __vimin3_u16x2(mx[3], mx[4], mx[5]);
    const uint32_t lo = __vminu2(b, __vimin3_u16x2(mx[6], mx[7], a));            // min over the pairs of the pair maxima
    const uint32_t c = __vimax3_u16x2(mn[0], mn[1], mn[2]), d = __vimax3_u16x2(mn[3], mn[4], mn[5]);
    const uint32_t hi = __vmaxu2(d, __vimax3_u16x2(mn[6], mn[7], c));            // max over the pairs of the pair minima
    // brighter: lo - p - 1, darker: p - hi - 1, each half in [-256, 254]; + 256 keeps every half non-negative
    const uint32_t brighter = lo + 0x00ff00ffu - centre;
    const uint32_t darker = centre + 0x00ff00ffu - hi;
    return __vmaxu2(brighter, darker);
}

}  // namespace mfk
