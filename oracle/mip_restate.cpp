// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restatement of the motion-interchange-pattern bytes: reference
// src/MoFREAK/MoFREAKUtilities.cpp:46-99 (motionInterchangePattern) and :288-325
// (extractMotionByMotionInterchangePatterns), with cv::resize(INTER_LINEAR, 8-bit) modelled after
// OpenCV's fixed-point bilinear path (imgproc/src/imgwarp.cpp; un-vendored).  The resize model is
// pinned against cv2.resize of the OpenCV 4.13 in this image by tests/test_oracle_resize.py.
#include "oracle.h"

#include <math.h>
#include <string.h>

namespace orc {

// s x s -> 19 x 19.  11-bit coefficients; the horizontal pass keeps 32-bit products, the vertical pass
// is ((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2 >> 2.  Along x an out-of-range tap collapses onto the
// border pixel with weight 1; along y the two row indices are clamped independently and the
// fraction is kept.
void resize_linear_u8_19(const uint8_t* src, int src_stride, int s, uint8_t dst[361]) {
    const int D = 19;
    const double scale = (double)s / D;
    int xofs[D], yofs[D];
    short ax[D][2], by[D][2];
    for (int d = 0; d < D; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int i = (int)floorf(f);
        f -= i;
        // y keeps the fraction
        yofs[d] = i;
        by[d][0] = (short)lrintf((1.f - f) * 2048);
        by[d][1] = (short)lrintf(f * 2048);
        // x collapses at the borders
        float fx = f;
        int ix = i;
        if (ix < 0) { ix = 0; fx = 0; }
        if (ix >= s - 1) { ix = s - 1; fx = 0; }
        xofs[d] = ix;
        ax[d][0] = (short)lrintf((1.f - fx) * 2048);
        ax[d][1] = (short)lrintf(fx * 2048);
    }
    for (int dy = 0; dy < D; dy++) {
        int y0 = yofs[dy], y1 = yofs[dy] + 1;
        if (y0 < 0) y0 = 0; if (y0 > s - 1) y0 = s - 1;
        if (y1 < 0) y1 = 0; if (y1 > s - 1) y1 = s - 1;
        const uint8_t* r0 = src + (size_t)y0 * src_stride;
        const uint8_t* r1 = src + (size_t)y1 * src_stride;
        for (int dx = 0; dx < D; dx++) {
            const int x0 = xofs[dx], x1 = (x0 + 1 < s) ? x0 + 1 : s - 1;
            const int R0 = r0[x0] * ax[dx][0] + r0[x1] * ax[dx][1];
            const int R1 = r1[x0] * ax[dx][0] + r1[x1] * ax[dx][1];
            int v = (((by[dy][0] * (R0 >> 4)) >> 16) + ((by[dy][1] * (R1 >> 4)) >> 16) + 2) >> 2;
            if (v < 0) v = 0; if (v > 255) v = 255;
            dst[dy * D + dx] = (uint8_t)v;
        }
    }
}

// One motion byte around patch centre (x, y) of the 19x19 patches.  The reference builds 3x3 ROI
// headers but then walks `.data` linearly, so each "patch" is 9 CONSECUTIVE bytes of the row-major
// 19x19 buffer starting at the ROI's top-left corner (MoFREAKUtilities.cpp:79-88).
unsigned mip_byte(const uint8_t cur19[361], const uint8_t prev19[361], int x, int y) {
    static const int SHIFT[8][2] = {{-4, 0}, {-3, 3}, {0, 4}, {3, 3}, {4, 0}, {3, -3}, {0, -4}, {-3, -3}};
    const int THETA = 288;
    const uint8_t* p = cur19 + (y - 1) * 19 + (x - 1);
    unsigned descriptor = 0;
    for (int k = 0; k < 8; k++) {
        const uint8_t* q = prev19 + (y + SHIFT[k][1] - 1) * 19 + (x + SHIFT[k][0] - 1);
        int ssd = 0;
        for (int i = 0; i < 9; i++) {
            const int d = (int)p[i] - (int)q[i];
            ssd += d * d;      // pow((float)d, 2) is exact for |d| <= 255
        }
        if (ssd > THETA) descriptor |= (1u << k);
    }
    return descriptor;
}

void mip_extract(const uint8_t* cur, const uint8_t* prev, int w, int h, float size, int x, int y,
                 uint8_t motion[8]) {
    static const int CENTRE[8][2] = {{5, 5}, {5, 9}, {5, 13}, {9, 5}, {9, 13}, {13, 5}, {13, 9}, {13, 13}};
    const int tl_x = x - (int)size / 2;
    const int tl_y = y - (int)size / 2;
    const int side = (int)ceil(size);
    (void)h;
    uint8_t a[361], b[361];
    resize_linear_u8_19(cur + (size_t)tl_y * w + tl_x, w, side, a);
    resize_linear_u8_19(prev + (size_t)tl_y * w + tl_x, w, side, b);
    for (int k = 0; k < 8; k++) motion[k] = (uint8_t)mip_byte(a, b, CENTRE[k][0], CENTRE[k][1]);
}

}  // namespace orc
