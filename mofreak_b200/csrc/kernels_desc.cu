// Descriptor stage, one warp per output row:
//   appearance  cv::FREAK::compute on the difference image, first 8 of its 64 bytes
//               (OpenCV 2.4.2 freak.cpp, un-vendored; call site MoFREAKUtilities.cpp:427-428, 453-456)
//   motion      MoFREAKUtilities::extractMotionByMotionInterchangePatterns / motionInterchangePattern
//               (MoFREAKUtilities.cpp:288-325, 46-99) with cv::resize(INTER_LINEAR) to 19x19
//   row         MoFREAKFeature assembly, MoFREAKUtilities.cpp:447-480
// Lanes own receptive fields (43 box means from the integral image), orientation pairs, pair tests and SSD
// tests; bits are packed with __ballot_sync, whose lane order is laid out so that the ballot words ARE the
// descriptor bytes.  Compiled with -fmad=false: the orientation angle must round exactly as on the host.
#include "kernels.h"

namespace mfk {

namespace {

struct RowOut { float x, y; int32_t frame; float scale; uint32_t app[2]; uint32_t mot[2]; };   // == mfk_row

// atan2f exactly as glibc 2.39's libm evaluates it in float arithmetic (the Sun fdlibm scheme: argument
// reduction against atan(0.5), atan(1), atan(1.5), atan(inf) and an odd/even split degree-11 polynomial).
// The host's atan2f is what the reference's cv::FREAK would call on this machine; reproducing its operation
// sequence with IEEE float ops gives the identical bit pattern for every input (checked against libm for
// every integer pair in [-700, 700]^2 and 2e7 random pairs by tests/test_host_tables.py's sibling in oracle).
__device__ __forceinline__ float atanf_fdlibm(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                          9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                          4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    const int hx = __float_as_int(x);
    const int ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c800000) return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
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
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return hx < 0 ? -r : r;
}

__device__ __forceinline__ float atan2f_fdlibm(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int hx = __float_as_int(x), hy = __float_as_int(y);
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

// FREAK::meanIntensity: box mean around one receptive field from the (W+1)-wide integral image
__device__ __forceinline__ int mean_intensity(const int32_t* __restrict__ I, int W1, const FreakPoint P, float kx, float ky) {
    const float xf = P.x + kx;
    const float yf = P.y + ky;
    const float radius = P.sigma;
    const int x_left = int((double)(xf - radius) + 0.5);
    const int y_top = int((double)(yf - radius) + 0.5);
    const int x_right = int((double)(xf + radius) + 1.5);
    const int y_bottom = int((double)(yf + radius) + 1.5);
    int v = __ldg(I + y_bottom * W1 + x_right) - __ldg(I + y_bottom * W1 + x_left) + __ldg(I + y_top * W1 + x_left) -
            __ldg(I + y_top * W1 + x_right);
    v = v / ((x_right - x_left) * (y_bottom - y_top));
    return v & 0xff;
}

constexpr int DW = 8;                 // warps per block
constexpr int PATCH = 19, PATCH_PX = 361, PATCH_PAD = 368;

struct WarpScratch {
    uint8_t cur[PATCH_PAD], prev[PATCH_PAD];
    uint8_t field[2][44];             // receptive-field means: unrotated, rotated
    int16_t ofs[PATCH];               // source index per destination index (same for x and y before clamping)
    int16_t ax[PATCH][2], by[PATCH][2];
    int16_t xofs[PATCH];
};

// cv::resize(ROI s x s -> 19 x 19, INTER_LINEAR) for both frames into scratch (fixed-point bilinear of OpenCV's
// imgwarp.cpp: 11-bit coefficients, vertical pass ((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2 >> 2)
__device__ __forceinline__ void resize_roi_pair(WarpScratch& ws, const uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev,
                                                int W, int tl_x, int tl_y, int s, int lane) {
    if (lane < PATCH) {
        const double scale = (double)s / PATCH;
        float f = (float)((lane + 0.5) * scale - 0.5);
        int i = (int)floorf(f);
        f -= (float)i;
        ws.ofs[lane] = (int16_t)i;
        ws.by[lane][0] = (int16_t)__float2int_rn((1.f - f) * 2048.f);
        ws.by[lane][1] = (int16_t)__float2int_rn(f * 2048.f);
        float fx = f;
        if (i < 0) { i = 0; fx = 0.f; }
        if (i >= s - 1) { i = s - 1; fx = 0.f; }
        ws.xofs[lane] = (int16_t)i;
        ws.ax[lane][0] = (int16_t)__float2int_rn((1.f - fx) * 2048.f);
        ws.ax[lane][1] = (int16_t)__float2int_rn(fx * 2048.f);
    }
    __syncwarp();
    const uint8_t* c0 = cur + (size_t)tl_y * W + tl_x;
    const uint8_t* p0 = prev + (size_t)tl_y * W + tl_x;
    for (int o = lane; o < PATCH_PX; o += 32) {
        const int dy = o / PATCH, dx = o - dy * PATCH;
        const int y0 = min(max((int)ws.ofs[dy], 0), s - 1), y1 = min(max((int)ws.ofs[dy] + 1, 0), s - 1);
        const int x0 = ws.xofs[dx], x1 = min(x0 + 1, s - 1);
        const int a0 = ws.ax[dx][0], a1 = ws.ax[dx][1], b0 = ws.by[dy][0], b1 = ws.by[dy][1];
        {
            const int R0 = c0[y0 * W + x0] * a0 + c0[y0 * W + x1] * a1;
            const int R1 = c0[y1 * W + x0] * a0 + c0[y1 * W + x1] * a1;
            const int v = (((b0 * (R0 >> 4)) >> 16) + ((b1 * (R1 >> 4)) >> 16) + 2) >> 2;
            ws.cur[o] = (uint8_t)min(max(v, 0), 255);
        }
        {
            const int R0 = p0[y0 * W + x0] * a0 + p0[y0 * W + x1] * a1;
            const int R1 = p0[y1 * W + x0] * a0 + p0[y1 * W + x1] * a1;
            const int v = (((b0 * (R0 >> 4)) >> 16) + ((b1 * (R1 >> 4)) >> 16) + 2) >> 2;
            ws.prev[o] = (uint8_t)min(max(v, 0), 255);
        }
    }
    __syncwarp();
}

// 64 SSD tests -> 8 motion bytes.  Test index t = 8*centre + shift sits on lane t % 32 of ballot t / 32, so the
// two ballot words are motion bytes 0..3 and 4..7.  Each "3x3 patch" is 9 CONSECUTIVE bytes of the 19x19
// buffer starting at the patch's top-left corner — the reference walks ROI .data linearly
// (MoFREAKUtilities.cpp:79-88) and that is part of the output contract.
__device__ __forceinline__ void mip_bits(const WarpScratch& ws, int lane, uint32_t out[2]) {
    const int CX[8] = {5, 5, 5, 9, 9, 13, 13, 13};
    const int CY[8] = {5, 9, 13, 5, 13, 5, 9, 13};
    const int SX[8] = {-4, -3, 0, 3, 4, 3, 0, -3};
    const int SY[8] = {0, 3, 4, 3, 0, -3, -4, -3};
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int c = 4 * r + (lane >> 3), k = lane & 7;
        const uint8_t* p = ws.cur + (CY[c] - 1) * PATCH + (CX[c] - 1);
        const uint8_t* q = ws.prev + (CY[c] + SY[k] - 1) * PATCH + (CX[c] + SX[k] - 1);
        int ssd = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) { const int d = (int)p[i] - (int)q[i]; ssd += d * d; }
        out[r] = __ballot_sync(0xffffffffu, ssd > 288);
    }
}

}  // namespace

__global__ void __launch_bounds__(DW * 32)
k_describe(const __grid_constant__ Geom g, const uint8_t* __restrict__ store, int store_frames, PairList pl,
           const int32_t* __restrict__ integral, CandBuf cb, DeviceTables tb, RowOut* __restrict__ rows) {
    __shared__ WarpScratch scratch[DW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpScratch& ws = scratch[wib];
    const uint32_t n_rows = cb.keep_scan[cb.counters[3]];
    const uint32_t rows_base = cb.counters[1];
    const int W1 = g.W + 1;
    const size_t ipx = (size_t)W1 * (g.H + 1);
    for (uint32_t r = blockIdx.x * DW + wib; r < n_rows; r += gridDim.x * DW) {
        if (rows_base + r >= (uint32_t)g.row_cap) break;            // overflow is flagged by k_advance
        const uint32_t c = cb.row_src[r];
        const int f = cb.meta[c] >> 3;
        const float kx = cb.kx[c], ky = cb.ky[c], ksize = cb.ksize[c];
        const int sidx = cb.scale_idx[c];
        const int32_t* I = integral + (size_t)f * ipx;

        // --- FREAK: orientation from the unrotated pattern ---
        const FreakPoint* P0 = tb.points + (size_t)sidx * FREAK_ORIENTATIONS * FREAK_POINTS;
        for (int p = lane; p < FREAK_POINTS; p += 32) ws.field[0][p] = (uint8_t)mean_intensity(I, W1, P0[p], kx, ky);
        __syncwarp();
        int d0 = 0, d1 = 0;
        for (int q = lane; q < FREAK_ORIENT_PAIRS; q += 32) {
            const OrientPair op = tb.orient[q];
            const int delta = (int)ws.field[0][op.i] - (int)ws.field[0][op.j];
            d0 += delta * op.wdx / 2048;
            d1 += delta * op.wdy / 2048;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
        const float angle = (float)((double)atan2f_fdlibm((float)d1, (float)d0) * (180.0 / 3.1415926535897932384626433832795));
        int theta = int((double)(256.0f * angle) * (1 / 360.0) + 0.5);
        if (theta < 0) theta += FREAK_ORIENTATIONS;
        if (theta >= FREAK_ORIENTATIONS) theta -= FREAK_ORIENTATIONS;

        // --- FREAK: the 64 pair tests the reference keeps ---
        const FreakPoint* PT = P0 + (size_t)theta * FREAK_POINTS;
        for (int p = lane; p < FREAK_POINTS; p += 32) ws.field[1][p] = (uint8_t)mean_intensity(I, W1, PT[p], kx, ky);
        __syncwarp();
        uint32_t app[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const UsedPair up = tb.used[32 * h + lane];                 // byte (32h+lane)/8, bit lane%8
            app[h] = __ballot_sync(0xffffffffu, ws.field[1][up.i] >= ws.field[1][up.j]);
        }

        // --- MIP on the gray frames ---
        const uint8_t* cur = store + (size_t)(pl.cur_slot[f] % store_frames) * g.L[0].px;
        const uint8_t* prev = store + (size_t)(pl.prev_slot[f] % store_frames) * g.L[0].px;
        const int xi = (int)kx, yi = (int)ky;                          // float -> int at the call, :460
        const int half = (int)ksize / 2;
        const int side = (int)ceilf(ksize);
        resize_roi_pair(ws, cur, prev, g.W, xi - half, yi - half, side, lane);
        uint32_t mot[2];
        mip_bits(ws, lane, mot);

        if (lane == 0) {
            RowOut o;
            o.x = kx; o.y = ky; o.frame = pl.label[f]; o.scale = ksize;
            o.app[0] = app[0]; o.app[1] = app[1]; o.mot[0] = mot[0]; o.mot[1] = mot[1];
            uint4* dst = reinterpret_cast<uint4*>(rows + rows_base + r);
            const uint4* src = reinterpret_cast<const uint4*>(&o);
            dst[0] = src[0]; dst[1] = src[1];
        }
        __syncwarp();
    }
}

void launch_describe(cudaStream_t st, const Geom& g, const uint8_t* store, int store_frames, PairList pl, const uint8_t*,
                     const int32_t* integral, CandBuf cb, DeviceTables tb, void* rows) {
    k_describe<<<148 * 4, DW * 32, 0, st>>>(g, store, store_frames, pl, integral, cb, tb, (RowOut*)rows);
}

// MIP bytes for caller-supplied keypoints on one frame pair (the MotionInterchangePatterns facade).
__global__ void __launch_bounds__(DW * 32)
k_mip_only(int W, int H, const uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev, const float* __restrict__ kp, int n,
           uint8_t* __restrict__ motion, uint32_t* __restrict__ err) {
    __shared__ WarpScratch scratch[DW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpScratch& ws = scratch[wib];
    for (int r = blockIdx.x * DW + wib; r < n; r += gridDim.x * DW) {
        const float kx = kp[3 * r], ky = kp[3 * r + 1], ksize = kp[3 * r + 2];
        const int xi = (int)kx, yi = (int)ky, half = (int)ksize / 2, side = (int)ceilf(ksize);
        const int tl_x = xi - half, tl_y = yi - half;
        // cv::Mat::operator()(Rect) asserts the ROI lies inside the image; report instead of reading outside
        if (side < 1 || tl_x < 0 || tl_y < 0 || tl_x + side > W || tl_y + side > H) {
            if (lane == 0) { atomicOr(err, 1u); for (int i = 0; i < 8; i++) motion[8 * r + i] = 0; }
            continue;
        }
        resize_roi_pair(ws, cur, prev, W, tl_x, tl_y, side, lane);
        uint32_t mot[2];
        mip_bits(ws, lane, mot);
        if (lane == 0) {
            for (int i = 0; i < 4; i++) { motion[8 * r + i] = (uint8_t)(mot[0] >> (8 * i)); motion[8 * r + 4 + i] = (uint8_t)(mot[1] >> (8 * i)); }
        }
        __syncwarp();
    }
}

void launch_mip_only(cudaStream_t st, int W, int H, const uint8_t* cur, const uint8_t* prev, const float* kp_xys, int n,
                     uint8_t* motion, uint32_t* err_flag) {
    if (n <= 0) return;
    const int blocks = min((n + DW - 1) / DW, 148 * 4);
    k_mip_only<<<blocks, DW * 32, 0, st>>>(W, H, cur, prev, kp_xys, n, motion, err_flag);
}

}  // namespace mfk
