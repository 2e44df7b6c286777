// Descriptor stage, one warp per output row:
//   appearance  cv::FREAK::compute on the difference image, first 8 of its 64 bytes
//               (OpenCV 2.4.2 freak.cpp, un-vendored; call site MoFREAKUtilities.cpp:427-428, 453-456)
//   motion      MoFREAKUtilities::extractMotionByMotionInterchangePatterns / motionInterchangePattern
//               (MoFREAKUtilities.cpp:288-325, 46-99) with cv::resize(INTER_LINEAR) to 19x19
//   row         MoFREAKFeature assembly, MoFREAKUtilities.cpp:447-480
// Lanes own receptive fields (43 box means from the integral image), orientation pairs, pair tests and SSD
// tests; bits are packed with __ballot_sync, whose lane order is laid out so that the ballot words ARE the
// descriptor bytes.  Compiled with -fmad=false: the orientation angle must round exactly as on the host.
#include <stddef.h>
#include <stdlib.h>

#include "brisk_device.cuh"
#include "desc_device.cuh"

namespace mfk {

namespace {

struct RowOut { float x, y; int32_t frame; float scale; uint32_t app[2]; uint32_t mot[2]; };   // == mfk_row

constexpr int DW = 8;                 // warps per block
constexpr int PATCH_PAD = 368;        // 19 x 19 = 361 patch bytes, padded

constexpr int ROI_MAX = 32;           // regions of interest up to this side are staged in shared memory
constexpr int ROI_WORDS = 9;          // aligned 32-bit words that cover 32 columns from any byte offset
constexpr int ROI_STRIDE = 4 * ROI_WORDS + 4;    // bytes per staged row (40: keeps rows 8-byte aligned)

// one receptive field of the pattern table: 16 bytes, one 128-bit read-only load
__device__ __forceinline__ float4 load_point(const FreakPoint* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct WarpScratch {
    alignas(16) uint8_t roi[2][ROI_MAX * ROI_STRIDE];
    alignas(16) RowTap rtap[PATCH];   // per destination row / column of the 19 x 19 patch
    alignas(16) ColTap ctap[PATCH];
    alignas(4) uint8_t cur[PATCH_PAD], prev[PATCH_PAD];     // word aligned and padded: mip_ssd reads them as aligned words
    uint8_t field[2][44];             // receptive-field means: unrotated, rotated
};

// Region of interest (side s <= ROI_MAX at tl_x, tl_y) of both gray frames -> shared memory, as the aligned words that
// cover it: staged column = image column - (tl_x & ~3).  With a frame width that is a multiple of 4 the words are aligned
// in global memory too and travel with cp.async, so the copy runs under the FREAK arithmetic of the same row; wait with
// roi_ready().  Other widths take the synchronous path (unaligned reads assembled by load4).
__device__ __forceinline__ void roi_stage(WarpScratch& ws, const uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev,
                                          int W, int tl_x, int tl_y, int s, int lane, bool async) {
    const int ax = tl_x & ~3;
    const int nw = ((tl_x & 3) + s + 3) >> 2;                       // words per row, <= 9
    // four rows per pass: lane = (row within the pass) * 8 + word, so that the lanes of a row read neighbouring words;
    // a ninth word (only for sides >= 30 at an odd offset) takes one extra pass with a lane per row
    const int rl = lane >> 3, k = lane & 7;
    const size_t org = (size_t)tl_y * W + ax;
    const uint8_t* c = cur + org + (size_t)rl * W + 4 * k;
    const uint8_t* p = prev + org + (size_t)rl * W + 4 * k;
    uint8_t* s0 = &ws.roi[0][rl * ROI_STRIDE + 4 * k];
    uint8_t* s1 = &ws.roi[1][rl * ROI_STRIDE + 4 * k];
    const bool col = k < nw;
    // two copies of the loop: one branch on `async` inside it becomes predicated code that issues both bodies
    if (async) {
        if (col)
            for (int r = rl; r < s; r += 4) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(s0)), "l"(c) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(s1)), "l"(p) : "memory");
                c += (size_t)4 * W; p += (size_t)4 * W; s0 += 4 * ROI_STRIDE; s1 += 4 * ROI_STRIDE;
            }
        if (nw == 9 && lane < s) {
            const size_t off = org + (size_t)lane * W + 32;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(&ws.roi[0][lane * ROI_STRIDE + 32])), "l"(cur + off));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(&ws.roi[1][lane * ROI_STRIDE + 32])), "l"(prev + off));
        }
    } else {
        if (col)
            for (int r = rl; r < s; r += 4) {
                *reinterpret_cast<uint32_t*>(s0) = load4(c);
                *reinterpret_cast<uint32_t*>(s1) = load4(p);
                c += (size_t)4 * W; p += (size_t)4 * W; s0 += 4 * ROI_STRIDE; s1 += 4 * ROI_STRIDE;
            }
        if (nw == 9 && lane < s) {
            const size_t off = org + (size_t)lane * W + 32;
            *reinterpret_cast<uint32_t*>(&ws.roi[0][lane * ROI_STRIDE + 32]) = load4(cur + off);
            *reinterpret_cast<uint32_t*>(&ws.roi[1][lane * ROI_STRIDE + 32]) = load4(prev + off);
        }
    }
    if (async) asm volatile("cp.async.commit_group;");
}

__device__ __forceinline__ void roi_ready(bool async) {
    if (async) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}

// cv::resize(ROI s x s -> 19 x 19, INTER_LINEAR) for both frames into scratch (fixed-point bilinear of OpenCV's
// imgwarp.cpp: 11-bit coefficients, vertical pass ((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2 >> 2).  Regions up to
// ROI_MAX must have been staged (roi_stage + roi_ready); larger ones are read from global memory.
__device__ __forceinline__ void resize_roi_pair(WarpScratch& ws, const uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev,
                                                int W, int tl_x, int tl_y, int s, int lane, const DeviceTables& tb) {
    const bool staged = s <= ROI_MAX;
    if (lane < PATCH) {
        const ResizeTap t = resize_tap(s, lane);
        ws.rtap[lane] = row_tap(t, s, staged ? ROI_STRIDE : W);
        ws.ctap[lane] = col_tap(t, s, staged ? (tl_x & 3) : 0);
    }
    __syncwarp();
    // only the 276 of the 2 x 361 patch pixels that the SSD tests read are interpolated, in the order of tb.mip_plan, whose
    // entries carry the byte offsets of the pixel's destination and of its row / column taps.  Two copies of the loop so
    // that the staged one reads through shared-memory loads (a pointer that may be either space costs generic loads)
    uint8_t* const patch = ws.cur;                                  // cur and prev are adjacent: prev = cur + PATCH_PAD
    const char* const rt = reinterpret_cast<const char*>(ws.rtap);
    const char* const ct = reinterpret_cast<const char*>(ws.ctap);
    if (staged) {
        for (int i = lane; i < tb.n_mip_needed; i += 32) {
            const uint32_t e = __ldg(tb.mip_plan + i);
            const int4 ry = *reinterpret_cast<const int4*>(rt + ((e >> 19) & 0x1ffu));      // 128-bit shared loads (the taps are 16-byte aligned)
            const int4 cx = *reinterpret_cast<const int4*>(ct + ((e >> 10) & 0x1ffu));
            const RowTap ty{ry.x, ry.y, ry.z, ry.w};
            const ColTap tx{cx.x, cx.y, cx.z, cx.w};
            patch[e & 0x3ffu] = (uint8_t)resize_px_flat(ws.roi[0] + (e >> 28) * (ROI_MAX * ROI_STRIDE), ty, tx);
        }
    } else {
        const uint8_t* src0 = cur + (size_t)tl_y * W + tl_x;
        const uint8_t* src1 = prev + (size_t)tl_y * W + tl_x;
        for (int i = lane; i < tb.n_mip_needed; i += 32) {
            const uint32_t e = __ldg(tb.mip_plan + i);
            const int4 ry = *reinterpret_cast<const int4*>(rt + ((e >> 19) & 0x1ffu));      // 128-bit shared loads (the taps are 16-byte aligned)
            const int4 cx = *reinterpret_cast<const int4*>(ct + ((e >> 10) & 0x1ffu));
            const RowTap ty{ry.x, ry.y, ry.z, ry.w};
            const ColTap tx{cx.x, cx.y, cx.z, cx.w};
            patch[e & 0x3ffu] = (uint8_t)resize_px_flat((e >> 28) ? src1 : src0, ty, tx);
        }
    }
    __syncwarp();
}

// 64 SSD tests -> 8 motion bytes.  Test index t = 8*centre + shift sits on lane t % 32 of ballot t / 32, so the
// two ballot words are motion bytes 0..3 and 4..7.  Each "3x3 patch" is 9 CONSECUTIVE bytes of the 19x19
// buffer starting at the patch's top-left corner — the reference walks ROI .data linearly
// (MoFREAKUtilities.cpp:79-88) and that is part of the output contract.
__device__ __forceinline__ void mip_bits(const WarpScratch& ws, int lane, uint32_t out[2]) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int ssd = mip_ssd(ws.cur, ws.prev, 4 * r + (lane >> 3), lane & 7);
        out[r] = __ballot_sync(0xffffffffu, ssd > 288);
    }
}

}  // namespace

__global__ void __launch_bounds__(DW * 32, 5)       // 5 blocks per SM at 48 registers (forcing 6 at 40 measured 6 % slower, twice)
k_describe(const __grid_constant__ Geom g, const uint8_t* __restrict__ store, int store_frames, PairList pl,
           const int32_t* __restrict__ integral, CandBuf cb, DeviceTables tb, RowOut* __restrict__ rows) {
    __shared__ WarpScratch scratch[DW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpScratch& ws = scratch[wib];
    const uint32_t n_rows = cb.keep_scan[cb.live[0]];
    const uint32_t rows_base = cb.live[3];                         // saved by the advance step (k_scatter_rows), which already ran for this batch
    const int W1 = g.W + 1;
    const size_t ipx = (size_t)W1 * (g.H + 1);
    // rows differ a lot in cost (region staged or not, pattern scale), so warps draw row tickets instead of striding; the
    // next ticket is drawn before the current row is processed, which hides the atomic's latency
    uint32_t r = blockIdx.x * DW + wib;                            // first ticket is implicit
    const uint32_t first_dynamic = gridDim.x * DW;
    for (;;) {
        if (r >= n_rows) break;
        uint32_t next = 0;
        if (lane == 0) {              // plain atom: the compiler wraps atomicAdd in a warp-aggregation sequence (~35 instructions per row)
            uint32_t t;
            asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(cb.live + 2) : "memory");
            next = first_dynamic + t;
        }
        if (rows_base + r >= (uint32_t)g.row_cap) break;            // overflow is flagged by the advance step
        const uint32_t c = cb.row_src[r];
        const int f = cb.meta[c] >> 3;
        const float kx = cb.kx[c], ky = cb.ky[c], ksize = cb.ksize[c];
        const int sidx = cb.scale_idx[c];
        const int32_t* I = integral + (size_t)f * ipx;
        asm volatile("" : "+l"(I));     // keep the pair's base in a register pair: folded into every gather it costs 64-bit adds with carries

        // --- MIP geometry first: the two regions of interest start travelling to shared memory now and are needed last ---
        const uint8_t* cur = store + (size_t)pl.cur_slot[f] * g.L[0].px;
        const uint8_t* prev = store + (size_t)pl.prev_slot[f] * g.L[0].px;
        const int xi = (int)kx, yi = (int)ky;                          // float -> int at the call, :460
        const int half = (int)ksize / 2;
        const int side = (int)ceilf(ksize);
        const bool async = (g.W & 3) == 0;
        if (side <= ROI_MAX) roi_stage(ws, cur, prev, g.W, xi - half, yi - half, side, lane, async);

        // --- FREAK: orientation from the unrotated pattern ---
        const FreakPoint* P0 = tb.points + (size_t)sidx * FREAK_ORIENTATIONS * FREAK_POINTS;
        for (int p = lane; p < FREAK_POINTS; p += 32) { const float4 P = load_point(P0 + p); ws.field[0][p] = (uint8_t)mean_intensity(I, W1, P.x, P.y, P.z, kx, ky); }
        __syncwarp();
        int d0 = 0, d1 = 0;
        for (int q = lane; q < FREAK_ORIENT_PAIRS; q += 32) {
            const OrientPair op = tb.orient[q];
            const int delta = (int)ws.field[0][op.i] - (int)ws.field[0][op.j];
            d0 += delta * op.wdx / 2048;
            d1 += delta * op.wdy / 2048;
        }
        d0 = __reduce_add_sync(0xffffffffu, d0);                       // REDUX: one instruction per warp sum
        d1 = __reduce_add_sync(0xffffffffu, d1);
        const int theta = freak_theta_index(d0, d1);

        // --- FREAK: the 64 pair tests the reference keeps ---
        const FreakPoint* PT = P0 + (size_t)theta * FREAK_POINTS;
        for (int p = lane; p < FREAK_POINTS; p += 32) { const float4 P = load_point(PT + p); ws.field[1][p] = (uint8_t)mean_intensity(I, W1, P.x, P.y, P.z, kx, ky); }
        __syncwarp();
        uint32_t app[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const UsedPair up = tb.used[32 * h + lane];                 // byte (32h+lane)/8, bit lane%8
            app[h] = __ballot_sync(0xffffffffu, ws.field[1][up.i] >= ws.field[1][up.j]);
        }

        // --- MIP on the gray frames ---
        if (side <= ROI_MAX) roi_ready(async);
        resize_roi_pair(ws, cur, prev, g.W, xi - half, yi - half, side, lane, tb);
        uint32_t mot[2];
        mip_bits(ws, lane, mot);

        if (lane == 0) {                                            // one row = two 16-byte stores, built in registers
            uint4* dst = reinterpret_cast<uint4*>(rows + rows_base + r);
            dst[0] = make_uint4(__float_as_uint(kx), __float_as_uint(ky), (uint32_t)pl.label[f], __float_as_uint(ksize));
            dst[1] = make_uint4(app[0], app[1], mot[0], mot[1]);
        }
        __syncwarp();                                               // scratch is reused by the next row
        r = __shfl_sync(0xffffffffu, next, 0);
    }
}

static_assert(offsetof(WarpScratch, prev) == offsetof(WarpScratch, cur) + PATCH_PAD, "the patch buffers are adjacent");
static_assert(2 * PATCH_PAD <= 1024 && 16 * PATCH <= 512, "mip_plan field widths");

uint32_t describe_patch_offset(int fr, int dx, int dy) { return (uint32_t)(fr * PATCH_PAD + dy * PATCH + dx); }

void launch_describe(cudaStream_t st, const Geom& g, const uint8_t* store, int store_frames, PairList pl, const uint8_t*,
                     const int32_t* integral, CandBuf cb, DeviceTables tb, void* rows) {
    // exactly one resident wave: tickets balance the rows, so a partial second wave would only add a tail (cached per device:
    // a process may hold contexts on several GPUs)
    static int blocks_of[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& blocks = blocks_of[dev & 63];
    if (!blocks) {
        int per_sm = 0, sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_describe, DW * 32, 0);
        if (const char* e = getenv("MFK_DESC_BLOCKS")) { const int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }   // measurements
        blocks = max(1, per_sm) * max(1, sms);
    }
    k_describe<<<blocks, DW * 32, 0, st>>>(g, store, store_frames, pl, integral, cb, tb, (RowOut*)rows);
}

// MIP bytes for caller-supplied keypoints on one frame pair (the MotionInterchangePatterns facade).
__global__ void __launch_bounds__(DW * 32)
k_mip_only(int W, int H, const uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev, const float* __restrict__ kp, int n,
           uint8_t* __restrict__ motion, uint32_t* __restrict__ err, DeviceTables tb) {
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
        if (side <= ROI_MAX) { roi_stage(ws, cur, prev, W, tl_x, tl_y, side, lane, false); roi_ready(false); }
        resize_roi_pair(ws, cur, prev, W, tl_x, tl_y, side, lane, tb);
        uint32_t mot[2];
        mip_bits(ws, lane, mot);
        if (lane == 0) {
            for (int i = 0; i < 4; i++) { motion[8 * r + i] = (uint8_t)(mot[0] >> (8 * i)); motion[8 * r + 4 + i] = (uint8_t)(mot[1] >> (8 * i)); }
        }
        __syncwarp();
    }
}

void launch_mip_only(cudaStream_t st, int W, int H, const uint8_t* cur, const uint8_t* prev, const float* kp_xys, int n,
                     uint8_t* motion, uint32_t* err_flag, DeviceTables tb) {
    if (n <= 0) return;
    const int blocks = min((n + DW - 1) / DW, 148 * 4);
    k_mip_only<<<blocks, DW * 32, 0, st>>>(W, H, cur, prev, kp_xys, n, motion, err_flag, tb);
}

}  // namespace mfk
