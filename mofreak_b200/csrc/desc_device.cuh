// Scalar device functions of the descriptor stage (FREAK orientation, receptive-field means).  Also compiles
// as plain C++ for the CPU harness in tests/.  Compiled with -fmad=false on the device.
#pragma once
#include <math.h>
#include <string.h>

#include "kernels.h"

namespace mfk {

__device__ __forceinline__ int float_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}

// atan2f exactly as glibc 2.39's libm evaluates it in float arithmetic (the Sun fdlibm scheme: argument
// reduction against atan(0.5), atan(1), atan(1.5), atan(inf) and an odd/even split degree-11 polynomial).
// The host's atan2f is what the reference's cv::FREAK would call on this machine; reproducing its operation
// sequence with IEEE float ops gives the identical bit pattern for every input (checked against libm for
// every integer pair in [-700, 700]^2 and random pairs by tests/test_device_source_cpu.py through the CPU harness).
__device__ __forceinline__ float atanf_fdlibm(float x) {
    // atan(0.5), atan(1), atan(1.5), atan(inf): high and low parts, picked by selects (an indexed array would live in local memory)
    const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
    const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                          9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                          4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    const int hx = float_bits(x);
    const int ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c800000) return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const float ahi = id == 0 ? hi0 : (id == 1 ? hi1 : (id == 2 ? hi2 : hi3));
    const float alo = id == 0 ? lo0 : (id == 1 ? lo1 : (id == 2 ? lo2 : lo3));
    const float r = ahi - ((x * (s1 + s2) - alo) - x);
    return hx < 0 ? -r : r;
}

__device__ __forceinline__ float atan2f_fdlibm(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int hx = float_bits(x), hy = float_bits(y);
    const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (hx == 0x3f800000) return atanf_fdlibm(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_fdlibm(fabsf(y / x));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

// FREAK::meanIntensity: box mean around one receptive field from the (W+1)-wide integral image.
// The reference rounds the box edges as int(float + 0.5) / int(float + 1.5) evaluated in double.  For a non-negative
// float t that equals trunc(t) + (frac(t) >= 0.5) resp. trunc(t) + 1 + (frac(t) >= 0.5), and frac(t) = t - trunc(t) is
// exact in float (Sterbenz), so no double arithmetic is needed; t >= 1 holds for every kept keypoint because FREAK's
// border filter keeps the whole pattern inside the image.  The truncating integer division is done with a float
// reciprocal estimate corrected by one step (sums are < 2^24, so the estimate is within 1).
__device__ __forceinline__ int round_half_up_nonneg(float t, int bias) {
    const int i = (int)t;
    return i + bias + ((t - (float)i) >= 0.5f ? 1 : 0);
}

__device__ __forceinline__ int mean_intensity(const int32_t* __restrict__ I, int W1, const float px, const float py, const float radius,
                                              float kx, float ky) {
    const float xf = px + kx;
    const float yf = py + ky;
    const int x_left = round_half_up_nonneg(xf - radius, 0);
    const int y_top = round_half_up_nonneg(yf - radius, 0);
    const int x_right = round_half_up_nonneg(xf + radius, 1);
    const int y_bottom = round_half_up_nonneg(yf + radius, 1);
#ifdef __CUDA_ARCH__
    // read-only global loads spelled out: the kernel keeps the pair's base pointer opaque, which would otherwise make these generic
    const int v = __ldg(I + (y_bottom * W1 + x_right)) - __ldg(I + (y_bottom * W1 + x_left)) + __ldg(I + (y_top * W1 + x_left)) -
                  __ldg(I + (y_top * W1 + x_right));
#else
    const int v = I[y_bottom * W1 + x_right] - I[y_bottom * W1 + x_left] + I[y_top * W1 + x_left] - I[y_top * W1 + x_right];
#endif
    const int area = (x_right - x_left) * (y_bottom - y_top);
#ifdef __CUDA_ARCH__
    float rcp;                                  // one MUFU.RCP: 1 <= area < 2^17 needs none of __fdividef's range scaling
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"((float)area));
    int q = (int)((float)v * rcp);
#else
    int q = (int)((float)v / (float)area);
#endif
    const int r = v - q * area;
    q += (r >= area) - (r < 0);
    return q & 0xff;
}

__device__ __forceinline__ int mean_intensity(const int32_t* __restrict__ I, int W1, const FreakPoint& P, float kx, float ky) {
    return mean_intensity(I, W1, P.x, P.y, P.sigma, kx, ky);
}

// ---- cv::resize(ROI s x s -> 19 x 19, INTER_LINEAR, 8-bit) as OpenCV's fixed-point bilinear path computes it ----
constexpr int PATCH = 19;

struct ResizeTap {          // for destination index d along one axis
    int16_t ofs;            // floor of the source coordinate (before clamping; used for rows)
    int16_t by0, by1;       // row weights (fraction kept at the borders, indices clamped independently)
    int16_t xofs;           // column index after the border collapse
    int16_t ax0, ax1;       // column weights (an out-of-range tap collapses onto the border pixel with weight 1)
};

__device__ __forceinline__ int round_half_even(float v) {
#ifdef __CUDA_ARCH__
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}

__device__ __forceinline__ ResizeTap resize_tap(int s, int d) {
    ResizeTap t;
    const double scale = (double)s / PATCH;
    float f = (float)((d + 0.5) * scale - 0.5);
    int i = (int)floorf(f);
    f -= (float)i;
    t.ofs = (int16_t)i;
    t.by0 = (int16_t)round_half_even((1.f - f) * 2048.f);
    t.by1 = (int16_t)round_half_even(f * 2048.f);
    float fx = f;
    if (i < 0) { i = 0; fx = 0.f; }
    if (i >= s - 1) { i = s - 1; fx = 0.f; }
    t.xofs = (int16_t)i;
    t.ax0 = (int16_t)round_half_even((1.f - fx) * 2048.f);
    t.ax1 = (int16_t)round_half_even(fx * 2048.f);
    return t;
}

// one destination pixel: 11-bit coefficients, vertical pass ((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2 >> 2
__device__ __forceinline__ int resize_px(const uint8_t* __restrict__ roi, int stride, int s, const ResizeTap ty, const ResizeTap tx) {
    const int y0 = min(max((int)ty.ofs, 0), s - 1), y1 = min(max((int)ty.ofs + 1, 0), s - 1);
    const int x0 = tx.xofs, x1 = min(x0 + 1, s - 1);
    const int R0 = roi[y0 * stride + x0] * tx.ax0 + roi[y0 * stride + x1] * tx.ax1;
    const int R1 = roi[y1 * stride + x0] * tx.ax0 + roi[y1 * stride + x1] * tx.ax1;
    const int v = (((ty.by0 * (R0 >> 4)) >> 16) + ((ty.by1 * (R1 >> 4)) >> 16) + 2) >> 2;
    return min(max(v, 0), 255);
}

// The taps of one destination row / column with everything that does not depend on the other axis folded in once per
// keypoint (clamps, the row offset in bytes, the column shift of a region staged from an aligned address): the per-pixel
// work is then four byte reads and the fixed-point blend.  Same arithmetic as resize_px.
struct RowTap { int32_t o0, o1, b0, b1; };
struct ColTap { int32_t x0, x1, a0, a1; };

__device__ __forceinline__ RowTap row_tap(const ResizeTap t, int s, int stride) {
    RowTap r;
    r.o0 = min(max((int)t.ofs, 0), s - 1) * stride;
    r.o1 = min(max((int)t.ofs + 1, 0), s - 1) * stride;
    r.b0 = t.by0; r.b1 = t.by1;
    return r;
}
__device__ __forceinline__ ColTap col_tap(const ResizeTap t, int s, int shift) {
    ColTap c;
    c.x0 = t.xofs + shift;
    c.x1 = min((int)t.xofs + 1, s - 1) + shift;
    c.a0 = t.ax0; c.a1 = t.ax1;
    return c;
}
__device__ __forceinline__ int resize_px_flat(const uint8_t* __restrict__ roi, const RowTap ty, const ColTap tx) {
    const int R0 = roi[ty.o0 + tx.x0] * tx.a0 + roi[ty.o0 + tx.x1] * tx.a1;
    const int R1 = roi[ty.o1 + tx.x0] * tx.a0 + roi[ty.o1 + tx.x1] * tx.a1;
    const int v = (((ty.b0 * (R0 >> 4)) >> 16) + ((ty.b1 * (R1 >> 4)) >> 16) + 2) >> 2;
    return min(max(v, 0), 255);
}

// SSD of motion test (centre c, shift k) on the two 19x19 patches.  Each "3x3 patch" is 9 CONSECUTIVE bytes of
// the row-major buffer starting at the patch's top-left corner: the reference walks ROI .data linearly
// (MoFREAKUtilities.cpp:79-88), and that is part of the output contract.  Centres :307-316, shifts :55-70.
__device__ __forceinline__ int mip_ssd(const uint8_t* cur19, const uint8_t* prev19, int c, int k) {
    // centre and shift tables as nibbles of one word each (indexable without a stack array): CX, CY, SX + 4, SY + 4
    const int cx = (int)((0xDDD99555u >> (4 * c)) & 15u), cy = (int)((0xD95D5D95u >> (4 * c)) & 15u);
    const int sx = (int)((0x14787410u >> (4 * k)) & 15u) - 4, sy = (int)((0x10147874u >> (4 * k)) & 15u) - 4;
    const uint8_t* p = cur19 + (cy - 1) * PATCH + (cx - 1);
    const uint8_t* q = prev19 + (cy + sy - 1) * PATCH + (cx + sx - 1);
#ifdef __CUDA_ARCH__
    // 9 bytes from any byte offset = three aligned words and two funnel shifts per patch (the buffers are word aligned and padded);
    // |p - q| per byte (VABSDIFF4), then the dot product of the differences with themselves (dp4a): exact, 19 instead of 36 instructions
    const uint32_t pa = (uint32_t)(uintptr_t)p, qa = (uint32_t)(uintptr_t)q;
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(p - (pa & 3u));
    const uint32_t* qw = reinterpret_cast<const uint32_t*>(q - (qa & 3u));
    const uint32_t ps = (pa & 3u) * 8u, qs = (qa & 3u) * 8u;
    const uint32_t p0 = pw[0], p1 = pw[1], p2 = pw[2], q0 = qw[0], q1 = qw[1], q2 = qw[2];
    const uint32_t d0 = __vabsdiffu4(__funnelshift_r(p0, p1, ps), __funnelshift_r(q0, q1, qs));
    const uint32_t d1 = __vabsdiffu4(__funnelshift_r(p1, p2, ps), __funnelshift_r(q1, q2, qs));
    const uint32_t d2 = __vabsdiffu4((p2 >> ps) & 0xffu, (q2 >> qs) & 0xffu);
    return (int)__dp4a(d2, d2, __dp4a(d1, d1, __dp4a(d0, d0, 0u)));
#else
    int ssd = 0;
    for (int i = 0; i < 9; i++) { const int d = (int)p[i] - (int)q[i]; ssd += d * d; }
    return ssd;
#endif
}

// FREAK orientation bin from the two weighted gradient sums (cv::FREAK::computeImpl): the angle in degrees is
// rounded to float, scaled in double, and truncated after adding 0.5
__device__ __forceinline__ int freak_theta_index(int d0, int d1) {
    const float angle = (float)((double)atan2f_fdlibm((float)d1, (float)d0) * (180.0 / 3.1415926535897932384626433832795));
    int theta = int((double)(256.0f * angle) * (1 / 360.0) + 0.5);
    if (theta < 0) theta += FREAK_ORIENTATIONS;
    if (theta >= FREAK_ORIENTATIONS) theta -= FREAK_ORIENTATIONS;
    return theta;
}

}  // namespace mfk
