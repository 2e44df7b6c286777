// Device-side BRISK scale-space arithmetic: what BriskScaleSpace does per candidate in reference
// src/MoFREAK/brisk.cpp — getScoreMaxAbove/Below (:1106-1416), subpixel2D (:1535-1644), refine1D*
// (:1418-1533), refine3D (:937-1103), the float score sampler (:1711-1725) — on dense closed-form
// score maps instead of the reference's lazily memoised ones.
//
// The float results are part of the bit-exact contract (they become x, y, scale in the .mofreak rows and
// decide the FREAK scale index), so this file is compiled with -fmad=false and every expression keeps
// the reference's precision and association: float where it computes in float, double where a double
// literal promotes the expression.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "geom.h"

namespace mfk {

// One layer of one pair's score map, with the optional "seen" side-effect map of getAgastScore's
// write-back (only meaningful for reads issued by getScoreMaxAbove, see kernels_sparse.cu).
struct ScoreView {
    const uint8_t* S;
    uint8_t* seen;     // nullptr: pure read
    int w, h;

    // BriskLayer::getAgastScore(int, int, 1), brisk.cpp:1691-1700
    __device__ __forceinline__ int at(int x, int y) const {
        if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
        const int i = y * w + x;
        if (seen) seen[i] = 1;
        return S[i];
    }
    // BriskLayer::getAgastScore(float, float, 1, scale <= 1), brisk.cpp:1711-1725
    __device__ __forceinline__ float atf(float xf, float yf) const {
        const int x = int(xf);
        const float rx1 = xf - float(x);
        const float rx = 1.0f - rx1;
        const int y = int(yf);
        const float ry1 = yf - float(y);
        const float ry = 1.0f - ry1;
        const int s00 = at(x, y), s10 = at(x + 1, y), s01 = at(x, y + 1), s11 = at(x + 1, y + 1);
        const float v = rx * ry * s00 + rx1 * ry * s10 + rx * ry1 * s01 + rx1 * ry1 * s11;
        return (float)(unsigned char)v;
    }
};

// Unaligned little-endian reads built from aligned 32-bit loads (the maps carry slack past their end, context.cu).
__device__ __forceinline__ uint32_t load4(const uint8_t* p) {
#ifdef __CUDA_ARCH__
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    return __funnelshift_r(q[0], q[1], (unsigned)(a & 3) * 8);
#else
    uint32_t v; memcpy(&v, p, 4); return v;
#endif
}
// bytes p[0..7] as two little-endian words
__device__ __forceinline__ void load8(const uint8_t* p, uint32_t& lo, uint32_t& hi) {
#ifdef __CUDA_ARCH__
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t q0 = q[0], q1 = q[1], q2 = q[2];
    const unsigned sh = (unsigned)(a & 3) * 8;
    lo = __funnelshift_r(q0, q1, sh);
    hi = __funnelshift_r(q1, q2, sh);
#else
    memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
#endif
}
// bytes p[0..3] in `first4`, p[4] returned
__device__ __forceinline__ uint32_t load5(const uint8_t* p, uint32_t& first4) {
#ifdef __CUDA_ARCH__
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t lo = q[0], hi = q[1];
    const unsigned sh = (unsigned)(a & 3) * 8;
    first4 = __funnelshift_r(lo, hi, sh);
    return (hi >> sh) & 0xffu;
#else
    memcpy(&first4, p, 4); return p[4];
#endif
}

// A 4x4 window of a neighbouring layer's score map held in registers.  The footprint of a getScoreMaxAbove /
// getScoreMaxBelow scan, including the 3x3 refinement around its integer maximum and the tie-break reads of
// the below scan, stays inside 4 x 4 cells: footprint widths are 4/6, 6/8, 8/6 and 6/4 of a cell, so per axis the
// walk reads [int(a_1), int(a1)+1] when the footprint crosses an integer and [int(a_1)-1, int(a_1)+2] when it does
// not (then the integer maximum can sit on either side, brisk.cpp:1139-1157); window_origin() picks the start.  Loading it once with 8 aligned loads replaces the
// ~50 dependent byte loads of the cell-by-cell walk; `touched` records which cells the walk read, i.e. the
// write-back side effect of the reference's lazy getAgastScore.
struct Window4 {
    uint32_t row[4];
    int x0, y0, w, h;
    uint32_t touched;

    __device__ __forceinline__ void load(const uint8_t* S, int w_, int h_, int x0_, int y0_) {
        w = w_; h = h_; x0 = x0_; y0 = y0_; touched = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) row[r] = load4(S + (y0 + r) * w + x0);
    }
    __device__ __forceinline__ int at(int x, int y) {
        if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
        const int r = y - y0, c = x - x0;
#ifndef __CUDA_ARCH__
        if (r < 0 || r > 3 || c < 0 || c > 3) { fprintf(stderr, "Window4 violated: x=%d y=%d x0=%d y0=%d w=%d h=%d\n", x, y, x0, y0, w, h); abort(); }   // the CPU harness checks the 4x4 claim on every read
#endif
        touched |= 1u << (r * 4 + c);
        const uint32_t v = r == 0 ? row[0] : (r == 1 ? row[1] : (r == 2 ? row[2] : row[3]));
        return (int)((v >> (8 * c)) & 0xffu);
    }
    __device__ __forceinline__ float atf(float xf, float yf) {
        const int x = int(xf);
        const float rx1 = xf - float(x);
        const float rx = 1.0f - rx1;
        const int y = int(yf);
        const float ry1 = yf - float(y);
        const float ry = 1.0f - ry1;
        const int s00 = at(x, y), s10 = at(x + 1, y), s01 = at(x, y + 1), s11 = at(x + 1, y + 1);
        const float v = rx * ry * s00 + rx1 * ry * s10 + rx * ry1 * s01 + rx1 * ry1 * s11;
        return (float)(unsigned char)v;
    }
    __device__ __forceinline__ void flush_marks(uint8_t* seen) const {
#pragma unroll
        for (int b = 0; b < 16; b++)
            if (touched & (1u << b)) seen[(y0 + (b >> 2)) * w + x0 + (b & 3)] = 1;
    }
};

// brisk.cpp:1535-1644, including the delta_y = delta_x quirk of the out-of-range branch (:1627-1636)
__device__ inline float subpixel2D(const int s_0_0, const int s_0_1, const int s_0_2, const int s_1_0,
                                   const int s_1_1, const int s_1_2, const int s_2_0, const int s_2_1,
                                   const int s_2_2, float& delta_x, float& delta_y) {
    const int tmp1 = s_0_0 + s_0_2 - 2 * s_1_1 + s_2_0 + s_2_2;
    const int coeff1 = 3 * (tmp1 + s_0_1 - ((s_1_0 + s_1_2) << 1) + s_2_1);
    const int coeff2 = 3 * (tmp1 - ((s_0_1 + s_2_1) << 1) + s_1_0 + s_1_2);
    const int tmp2 = s_0_2 - s_2_0;
    const int tmp3 = (s_0_0 + tmp2 - s_2_2);
    const int tmp4 = tmp3 - 2 * tmp2;
    const int coeff3 = -3 * (tmp3 + s_0_1 - s_2_1);
    const int coeff4 = -3 * (tmp4 + s_1_0 - s_1_2);
    const int coeff5 = (s_0_0 - s_0_2 - s_2_0 + s_2_2) * 4;
    const int coeff6 = -(s_0_0 + s_0_2 - ((s_1_0 + s_0_1 + s_1_2 + s_2_1) << 1) - 5 * s_1_1 + s_2_0 + s_2_2) * 2;

    const int H_det = 4 * coeff1 * coeff2 - coeff5 * coeff5;
    if (H_det == 0) {
        delta_x = 0.0f;
        delta_y = 0.0f;
        return (float)((double)float(coeff6) / 18.0);
    }
    if (!(H_det > 0 && coeff1 < 0)) {
        int tmp_max = coeff3 + coeff4 + coeff5;
        delta_x = 1.0f; delta_y = 1.0f;
        int tmp = -coeff3 + coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = 1.0f; }
        tmp = coeff3 - coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = 1.0f; delta_y = -1.0f; }
        tmp = -coeff3 - coeff4 + coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = -1.0f; }
        return (float)((double)float(tmp_max + coeff1 + coeff2 + coeff6) / 18.0);
    }
    delta_x = float(2 * coeff2 * coeff3 - coeff4 * coeff5) / float(-H_det);
    delta_y = float(2 * coeff1 * coeff4 - coeff3 * coeff5) / float(-H_det);
    bool tx = false, tx_ = false, ty = false, ty_ = false;
    if (delta_x > 1.0f) tx = true;
    else if (delta_x < -1.0f) tx_ = true;
    if (delta_y > 1.0f) ty = true;
    if (delta_y < -1.0f) ty_ = true;

    if (tx || tx_ || ty || ty_) {
        float delta_x1 = 0.0f, delta_x2 = 0.0f, delta_y1 = 0.0f, delta_y2 = 0.0f;
        if (tx) {
            delta_x1 = 1.0f;
            delta_y1 = -float(coeff4 + coeff5) / float(2 * coeff2);
            if (delta_y1 > 1.0f) delta_y1 = 1.0f; else if (delta_y1 < -1.0f) delta_y1 = -1.0f;
        } else if (tx_) {
            delta_x1 = -1.0f;
            delta_y1 = -float(coeff4 - coeff5) / float(2 * coeff2);
            if (delta_y1 > 1.0f) delta_y1 = 1.0f; else if (delta_y1 < -1.0f) delta_y1 = -1.0f;
        }
        if (ty) {
            delta_y2 = 1.0f;
            delta_x2 = -float(coeff3 + coeff5) / float(2 * coeff1);
            if (delta_x2 > 1.0f) delta_x2 = 1.0f; else if (delta_x2 < -1.0f) delta_x2 = -1.0f;
        } else if (ty_) {
            delta_y2 = -1.0f;
            delta_x2 = -float(coeff3 - coeff5) / float(2 * coeff1);
            if (delta_x2 > 1.0f) delta_x2 = 1.0f; else if (delta_x2 < -1.0f) delta_x2 = -1.0f;
        }
        const float q1 = coeff1 * delta_x1 * delta_x1 + coeff2 * delta_y1 * delta_y1 + coeff3 * delta_x1 +
                         coeff4 * delta_y1 + coeff5 * delta_x1 * delta_y1 + coeff6;
        const float q2 = coeff1 * delta_x2 * delta_x2 + coeff2 * delta_y2 * delta_y2 + coeff3 * delta_x2 +
                         coeff4 * delta_y2 + coeff5 * delta_x2 * delta_y2 + coeff6;
        const float max1 = (float)((double)q1 / 18.0);
        const float max2 = (float)((double)q2 / 18.0);
        if (max1 > max2) { delta_x = delta_x1; delta_y = delta_x1; return max1; }
        delta_x = delta_x2; delta_y = delta_x2; return max2;
    }
    const float q = coeff1 * delta_x * delta_x + coeff2 * delta_y * delta_y + coeff3 * delta_x + coeff4 * delta_y +
                    coeff5 * delta_x * delta_y + coeff6;
    return (float)((double)q / 18.0);
}

// The three parabola fits over relative scale (brisk.cpp:1418-1533).  kind 0: refine1D (octave layers
// 2, 4), 1: refine1D_1 (intra layers), 2: refine1D_2 (layer 0).
__device__ inline float refine1D_fit(const int kind, const float s_05, const float s0, const float s05, float& max) {
    int a0, a1, a2, b0, b1, b2, c0, c1, c2;
    double lo, hi, den;
    if (kind == 0)      { a0 = 16; a1 = -24; a2 = 8;  b0 = -40; b1 = 54; b2 = -14; c0 = 24; c1 = -27; c2 = 6; lo = 0.75; hi = 1.5; den = 3072.0; }
    else if (kind == 1) { a0 = 9;  a1 = -18; a2 = 9;  b0 = -21; b1 = 36; b2 = -15; c0 = 12; c1 = -16; c2 = 6;
                          lo = 0.6666666666666666666666666667; hi = 1.3333333333333333333333333333; den = 2048.0; }
    else                { a0 = 2;  a1 = -4;  a2 = 2;  b0 = -5;  b1 = 8;  b2 = -3;  c0 = 3;  c1 = -3;  c2 = 1; lo = 0.7; hi = 1.5; den = 1024.0; }
    const int i_05 = int(1024.0 * (double)s_05 + 0.5);
    const int i0 = int(1024.0 * (double)s0 + 0.5);
    const int i05 = int(1024.0 * (double)s05 + 0.5);
    const int qa = a0 * i_05 + a1 * i0 + a2 * i05;
    if (qa >= 0) {
        if (s0 >= s_05 && s0 >= s05) { max = s0; return 1.0f; }
        if (s_05 >= s0 && s_05 >= s05) { max = s_05; return (float)lo; }
        if (s05 >= s0 && s05 >= s_05) { max = s05; return (float)hi; }
    }
    const int qb = b0 * i_05 + b1 * i0 + b2 * i05;
    float ret_val = -float(qb) / float(2 * qa);
    if ((double)ret_val < lo) ret_val = (float)lo;
    else if ((double)ret_val > hi) ret_val = (float)hi;
    const int qc = c0 * i_05 + c1 * i0 + c2 * i05;
    max = float(qc) + float(qa) * ret_val * ret_val + float(qb) * ret_val;
    max = (float)((double)max / den);
    return ret_val;
}

struct Scan { bool ismax; float max, dx, dy; };

// Sampling footprint in the neighbouring layer: getScoreMaxAbove (below == false, brisk.cpp:1122-1135) /
// getScoreMaxBelow (below == true, :1263-1275) of candidate (x_layer, y_layer) on layer `layer`.
struct Footprint { float x_1, x1, y_1, y1; };

__device__ __forceinline__ int window_origin(float a_1, float a1) { return int(a1) > int(a_1) ? int(a_1) : int(a_1) - 1; }

__device__ __forceinline__ Footprint scan_footprint(const bool below, const int layer, const int x_layer, const int y_layer) {
    float x_1, x1, y_1, y1;
    if (!below) {
        if (layer % 2 == 0) {
            x_1 = (float)((double)float(4 * x_layer - 1 - 2) / 6.0); x1 = (float)((double)float(4 * x_layer - 1 + 2) / 6.0);
            y_1 = (float)((double)float(4 * y_layer - 1 - 2) / 6.0); y1 = (float)((double)float(4 * y_layer - 1 + 2) / 6.0);
        } else {
            x_1 = float(6 * x_layer - 1 - 3) / 8.0f; x1 = float(6 * x_layer - 1 + 3) / 8.0f;
            y_1 = float(6 * y_layer - 1 - 3) / 8.0f; y1 = float(6 * y_layer - 1 + 3) / 8.0f;
        }
    } else {
        if (layer % 2 == 0) {
            x_1 = (float)((double)float(8 * x_layer + 1 - 4) / 6.0); x1 = (float)((double)float(8 * x_layer + 1 + 4) / 6.0);
            y_1 = (float)((double)float(8 * y_layer + 1 - 4) / 6.0); y1 = (float)((double)float(8 * y_layer + 1 + 4) / 6.0);
        } else {
            x_1 = (float)((double)float(6 * x_layer + 1 - 3) / 4.0); x1 = (float)((double)float(6 * x_layer + 1 + 3) / 4.0);
            y_1 = (float)((double)float(6 * y_layer + 1 - 3) / 4.0); y1 = (float)((double)float(6 * y_layer + 1 + 3) / 4.0);
        }
    }
    Footprint f; f.x_1 = x_1; f.x1 = x1; f.y_1 = y_1; f.y1 = y1;
    return f;
}

// The scan itself over a loaded window: getScoreMaxAbove (:1138-1249) / getScoreMaxBelow (:1281-1416).
__device__ inline Scan scan_neighbour(Window4& N, const Footprint fp, const bool below, const int layer, const int x_layer,
                                      const int y_layer, const int threshold) {
    Scan R; R.ismax = false; R.max = 0.0f; R.dx = 0.0f; R.dy = 0.0f;
    const float x_1 = fp.x_1, x1 = fp.x1, y_1 = fp.y_1, y1 = fp.y1;
    const int xa = int(x_1 + 1), xb = int(x1);
    const int ya = int(y_1 + 1), yb = int(y1);
    int max_x = xa, max_y = ya;
    const float thr = (float)threshold;
    float tmp;
    float max = N.atf(x_1, y_1);
    // first (fractional) row: leave as soon as a sample beats the candidate
    if (max > thr) return R;
    for (int x = xa; x <= xb; x++) {
        tmp = N.atf(float(x), y_1);
        if (tmp > thr) return R;
        if (tmp > max) { max = tmp; max_x = x; }
    }
    tmp = N.atf(x1, y_1);
    if (tmp > thr) return R;
    if (tmp > max) { max = tmp; max_x = xb; }
    // integer rows
    for (int y = ya; y <= yb; y++) {
        tmp = N.atf(x_1, float(y));
        if (tmp > thr) return R;
        if (tmp > max) { max = tmp; max_x = xa; max_y = y; }
        for (int x = xa; x <= xb; x++) {
            tmp = (float)N.at(x, y);
            if (tmp > thr) return R;
            if (below && tmp == max) {
                const int t1 = 2 * (N.at(x - 1, y) + N.at(x + 1, y) + N.at(x, y + 1) + N.at(x, y - 1)) +
                               (N.at(x + 1, y + 1) + N.at(x - 1, y + 1) + N.at(x + 1, y - 1) + N.at(x - 1, y - 1));
                const int t2 = 2 * (N.at(max_x - 1, max_y) + N.at(max_x + 1, max_y) + N.at(max_x, max_y + 1) +
                                    N.at(max_x, max_y - 1)) +
                               (N.at(max_x + 1, max_y + 1) + N.at(max_x - 1, max_y + 1) +
                                N.at(max_x + 1, max_y - 1) + N.at(max_x - 1, max_y - 1));
                if (t1 > t2) { max_x = x; max_y = y; }
            }
            if (tmp > max) { max = tmp; max_x = x; max_y = y; }
        }
        tmp = N.atf(x1, float(y));
        if (tmp > thr) return R;
        if (tmp > max) { max = tmp; max_x = xb; max_y = y; }
    }
    // last (fractional) row: no early exit
    tmp = N.atf(x_1, y1);
    if (tmp > max) { max = tmp; max_x = xa; max_y = yb; }
    for (int x = xa; x <= xb; x++) {
        tmp = N.atf(float(x), y1);
        if (tmp > max) { max = tmp; max_x = x; max_y = yb; }
    }
    tmp = N.atf(x1, y1);
    if (tmp > max) { max = tmp; max_x = xb; max_y = yb; }

    const int s_0_0 = N.at(max_x - 1, max_y - 1), s_1_0 = N.at(max_x, max_y - 1), s_2_0 = N.at(max_x + 1, max_y - 1);
    const int s_2_1 = N.at(max_x + 1, max_y), s_1_1 = N.at(max_x, max_y), s_0_1 = N.at(max_x - 1, max_y);
    const int s_0_2 = N.at(max_x - 1, max_y + 1), s_1_2 = N.at(max_x, max_y + 1), s_2_2 = N.at(max_x + 1, max_y + 1);
    float dx_1, dy_1;
    const float refined_max = subpixel2D(s_0_0, s_0_1, s_0_2, s_1_0, s_1_1, s_1_2, s_2_0, s_2_1, s_2_2, dx_1, dy_1);
    const float real_x = float(max_x) + dx_1;
    const float real_y = float(max_y) + dy_1;
    bool returnrefined = true;
    float dx, dy;
    if (!below) {
        if (layer % 2 == 0) {
            dx = (real_x * 6.0f + 1.0f) / 4.0f - float(x_layer);
            dy = (real_y * 6.0f + 1.0f) / 4.0f - float(y_layer);
        } else {
            dx = (float)(((double)real_x * 8.0 + 1.0) / 6.0 - (double)float(x_layer));
            dy = (float)(((double)real_y * 8.0 + 1.0) / 6.0 - (double)float(y_layer));
        }
    } else {
        if (layer % 2 == 0) {
            dx = (float)(((double)real_x * 6.0 + 1.0) / 8.0 - (double)float(x_layer));
            dy = (float)(((double)real_y * 6.0 + 1.0) / 8.0 - (double)float(y_layer));
        } else {
            dx = (float)(((double)real_x * 4.0 - 1.0) / 6.0 - (double)float(x_layer));
            dy = (float)(((double)real_y * 4.0 - 1.0) / 6.0 - (double)float(y_layer));
        }
    }
    if (dx > 1.0f) { dx = 1.0f; returnrefined = false; }
    if (dx < -1.0f) { dx = -1.0f; returnrefined = false; }
    if (dy > 1.0f) { dy = 1.0f; returnrefined = false; }
    if (dy < -1.0f) { dy = -1.0f; returnrefined = false; }
    R.ismax = true;
    R.dx = dx; R.dy = dy;
    R.max = returnrefined ? fmaxf(refined_max, max) : max;
    return R;
}

// AGAST-5/8 closed-form scores with b0 = 0 (agast5_8_nms.cc:42; ring offsets agast5_8.h:65-75) at the nine positions
// (x+dx, y+dy), dx, dy in -1..1, out[(dy+1)*3 + (dx+1)] — what refine3D reads on layer 0 through
// BriskLayer::getAgastScore_5_8 (brisk.cpp:959-986, 1702-1709).  The 5x5 image window is loaded once (10 aligned
// loads) and every pixel is packed as (v, 255 - v), so one 3-input packed minimum serves the brighter and the
// darker test at once, as in the 9/16 score.  Requires 3 <= x < w-3, 3 <= y < h-3 (true for every candidate), which
// also keeps all nine positions inside getAgastScore_5_8's 2-px border.
// (A scalar max(best, max(mn, -mx)) formulation is miscompiled by nvcc 12.9 for sm_100a: ptxas folds the negation
// into a VIMNMX3 operand and drops it; see DESIGN.md "toolchain notes" and tests/repro/.)
__device__ inline void agast58_patch9(const uint8_t* img, int w, int x, int y, int out[9]) {
    uint32_t X[5][5];
#pragma unroll
    for (int r = 0; r < 5; r++) {
        uint32_t first4;
        const uint32_t fifth = load5(img + (y - 2 + r) * w + (x - 2), first4);
#pragma unroll
        for (int c = 0; c < 4; c++) { const uint32_t v = (first4 >> (8 * c)) & 0xffu; X[r][c] = v | ((255u - v) << 16); }
        X[r][4] = fifth | ((255u - fifth) << 16);
    }
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
            const int cy = dy + 1, cx = dx + 1;
            const uint32_t r[8] = {X[cy][cx - 1], X[cy - 1][cx - 1], X[cy - 1][cx], X[cy - 1][cx + 1],
                                   X[cy][cx + 1], X[cy + 1][cx + 1], X[cy + 1][cx], X[cy + 1][cx - 1]};
            uint32_t best = 0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
                uint32_t m = __vimin3_u16x2(r[s], r[(s + 1) & 7], r[(s + 2) & 7]);
                m = __vimin3_u16x2(m, r[(s + 3) & 7], r[(s + 4) & 7]);
                best = __vimax3_u16x2(best, m, m);
            }
            const int c = (int)(X[cy][cx] & 0xffffu);
            const int brighter = (int)(best & 0xffffu) - c;
            const int darker = (int)(best >> 16) - (255 - c);
            const int sc = (brighter > darker ? brighter : darker) - 1;
            out[dy * 3 + dx] = sc < 0 ? 0 : (sc > 254 ? 254 : sc);
        }
}

}  // namespace mfk
