// Dense (per-pixel) stage of the extraction path, sm_100a:
//   absdiff        cv::absdiff(current, prev)                       MoFREAKUtilities.cpp:413-414
//   pyramid        BriskLayer::halfsample / twothirdsample          brisk.cpp:1840-2065, driven as :572-588
//   scores         OastDetector9_16::detect + cornerScore           oast9_16.cc:46, oast9_16_nms.cc:42
//                  (closed form of the segment test, every pixel)   + per-row hit counts
//   candidates     ordered (pair, layer, y, x) list of hits         getAgastPoints order, brisk.cpp:1678-1690
//   integral       cv::integral(diff)                               inside cv::FREAK::compute
//   exclusive scan utility used for the ordered compactions
//
// All of this is byte/integer work bound by memory and integer-issue throughput; no stage is a dense
// contraction, so tensor cores are not involved.
#include "brisk_device.cuh"
#include "dense_device.cuh"
#include <cuda.h>

#include "kernels.h"
#include "tma_device.cuh"

namespace mfk {

// ---------------------------------------------------------------------------------------------------
// |a - b|
// ---------------------------------------------------------------------------------------------------
__global__ void k_absdiff(const uint8_t* __restrict__ store, int store_frames, const int32_t* __restrict__ cur_slot,
                          const int32_t* __restrict__ prev_slot, uint8_t* __restrict__ img, int px, int vec_ok) {
    const int f = blockIdx.y;
    const uint8_t* a = store + (size_t)cur_slot[f] * px;            // slots arrive reduced modulo the store size (launch_set_pairs)
    const uint8_t* b = store + (size_t)prev_slot[f] * px;
    uint8_t* d = img + (size_t)f * px;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    if (vec_ok) {
        const uint4* a4 = reinterpret_cast<const uint4*>(a);
        const uint4* b4 = reinterpret_cast<const uint4*>(b);
        uint4* d4 = reinterpret_cast<uint4*>(d);
        for (int i = tid; i < px / 16; i += nthreads) {
            const uint4 u = __ldg(a4 + i), v = __ldg(b4 + i);
            uint4 r;
            r.x = __vabsdiffu4(u.x, v.x); r.y = __vabsdiffu4(u.y, v.y);
            r.z = __vabsdiffu4(u.z, v.z); r.w = __vabsdiffu4(u.w, v.w);
            d4[i] = r;
        }
    } else {
        for (int i = tid; i < px; i += nthreads) d[i] = (uint8_t)abs((int)a[i] - (int)b[i]);
    }
}

// ---------------------------------------------------------------------------------------------------
// frame ingest: interleaved BGR -> gray, cv::cvtColor(CV_BGR2GRAY) on 8-bit data
// (call sites MoFREAKUtilities.cpp:395, 410).  OpenCV's fixed-point form: (B*cb + G*cg + R*cr + half) >> bits with
// 14-bit coefficients 1868 / 9617 / 4899 in OpenCV 2.4 (the reference's version) and 15-bit 3735 / 19235 / 9798 from 3.x on.
// ---------------------------------------------------------------------------------------------------
__global__ void k_bgr2gray(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, size_t n_px, int cb, int cg, int cr, int bits) {
    const size_t quad = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // 4 pixels = 12 source bytes = 3 aligned words
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int half = 1 << (bits - 1);
    for (size_t q = quad; q * 4 < n_px; q += stride) {
        if (q * 4 + 4 <= n_px) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(bgr) + q * 3;
            const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];            // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
            const int y0 = ((int)(w0 & 0xff) * cb + (int)((w0 >> 8) & 0xff) * cg + (int)((w0 >> 16) & 0xff) * cr + half) >> bits;
            const int y1 = ((int)(w0 >> 24) * cb + (int)(w1 & 0xff) * cg + (int)((w1 >> 8) & 0xff) * cr + half) >> bits;
            const int y2 = ((int)((w1 >> 16) & 0xff) * cb + (int)(w1 >> 24) * cg + (int)(w2 & 0xff) * cr + half) >> bits;
            const int y3 = ((int)((w2 >> 8) & 0xff) * cb + (int)((w2 >> 16) & 0xff) * cg + (int)(w2 >> 24) * cr + half) >> bits;
            reinterpret_cast<uint32_t*>(gray)[q] = (uint32_t)y0 | ((uint32_t)y1 << 8) | ((uint32_t)y2 << 16) | ((uint32_t)y3 << 24);
        } else {
            for (size_t p = q * 4; p < n_px; p++)
                gray[p] = (uint8_t)(((int)bgr[3 * p] * cb + (int)bgr[3 * p + 1] * cg + (int)bgr[3 * p + 2] * cr + half) >> bits);
        }
    }
}

__global__ void k_bgr2gray_bytes(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, size_t n_px, int cb, int cg, int cr, int bits) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int half = 1 << (bits - 1);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += stride)
        gray[p] = (uint8_t)(((int)bgr[3 * p] * cb + (int)bgr[3 * p + 1] * cg + (int)bgr[3 * p + 2] * cr + half) >> bits);
}

// word path when both pointers are 4-byte aligned (always, for frame sizes that are a multiple of 4 pixels), else bytes
void launch_bgr2gray(cudaStream_t st, const uint8_t* bgr, uint8_t* gray, size_t n_px, int coeff_bits) {
    if (n_px == 0) return;
    const bool v24 = coeff_bits == 14;
    const int cb = v24 ? 1868 : 3735, cg = v24 ? 9617 : 19235, cr = v24 ? 4899 : 9798, bits = v24 ? 14 : 15;
    if ((((uintptr_t)bgr | (uintptr_t)gray) & 3) == 0) {
        const size_t quads = (n_px + 3) / 4;
        k_bgr2gray<<<(unsigned)min((quads + 255) / 256, (size_t)148 * 16), 256, 0, st>>>(bgr, gray, n_px, cb, cg, cr, bits);
    } else {
        k_bgr2gray_bytes<<<(unsigned)min((n_px + 255) / 256, (size_t)148 * 16), 256, 0, st>>>(bgr, gray, n_px, cb, cg, cr, bits);
    }
}

// ---------------------------------------------------------------------------------------------------
// pair descriptions: the host arrays ride in the kernel parameter block, 256 pairs (3 KB) per launch
// ---------------------------------------------------------------------------------------------------
struct PairChunk { int32_t cur[256], prev[256], label[256]; };

__global__ void k_set_pairs(const __grid_constant__ PairChunk ch, int n, int32_t* cur, int32_t* prev, int32_t* label) {
    const int i = threadIdx.x;
    if (i < n) { cur[i] = ch.cur[i]; prev[i] = ch.prev[i]; label[i] = ch.label[i]; }
}

// (slots are reduced modulo the store size here, once per pair: the kernels index the store with them as they are)
int launch_set_pairs(cudaStream_t st, const int32_t* cur, const int32_t* prev, const int32_t* label, int n, int store_frames, PairList dst) {
    int launches = 0;
    PairChunk ch;
    for (int done = 0; done < n; done += 256, launches++) {
        const int m = min(256, n - done);
        for (int i = 0; i < m; i++) { ch.cur[i] = cur[done + i] % store_frames; ch.prev[i] = prev[done + i] % store_frames; }
        memcpy(ch.label, label + done, (size_t)m * 4);
        k_set_pairs<<<1, 256, 0, st>>>(ch, m, const_cast<int32_t*>(dst.cur_slot) + done, const_cast<int32_t*>(dst.prev_slot) + done,
                                       const_cast<int32_t*>(dst.label) + done);
    }
    return launches;
}

void launch_absdiff(cudaStream_t st, const Geom& g, const uint8_t* store, int store_frames, PairList pl, uint8_t* img) {
    const int px = g.L[0].px;
    const int vec_ok = (px % 16 == 0);     // slots and layer-0 images then start 16-byte aligned
    const int work = vec_ok ? px / 16 : px;
    dim3 grid((unsigned)min((work + 255) / 256, 1024), (unsigned)g.F);
    k_absdiff<<<grid, 256, 0, st>>>(store, store_frames, pl.cur_slot, pl.prev_slot, img, px, vec_ok);
}

// ---------------------------------------------------------------------------------------------------
// pyramid samplers, one destination pixel per thread
// ---------------------------------------------------------------------------------------------------
// half sampling, 4 destination pixels per thread: inside the paired 16-byte blocks the rule is the same for every
// column (avg of avg, all rounding up), so two unaligned 8-byte reads give four outputs; the few columns of the odd
// trailing block and of the leftover use the per-pixel function.
__global__ void k_half(const uint8_t* __restrict__ src_base, uint8_t* __restrict__ dst_base, int sw, int spx, int dw, int dh, int dpx) {
    const int f = blockIdx.z;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y * blockDim.y + threadIdx.y;
    const int paired_out = ((sw >> 4) >> 1) * 16;                  // columns that follow the uniform rule (a multiple of 16)
    const int quads = paired_out >> 2;
    if (oy >= dh) return;
    const uint8_t* src = src_base + (size_t)f * spx;
    uint8_t* row = dst_base + (size_t)f * dpx + oy * dw;
    if (t < quads) {
        const int ox0 = 4 * t;
        uint32_t u0, u1, l0, l1;
        load8(src + (2 * oy) * sw + 2 * ox0, u0, u1);
        load8(src + (2 * oy + 1) * sw + 2 * ox0, l0, l1);
        const uint32_t v0 = __vavgu4(u0, l0), v1 = __vavgu4(u1, l1);      // vertical averages of 8 columns, rounding up
        // horizontal: avg(v[2j], v[2j+1]) — pair the even and the odd bytes
        const uint32_t ev = __byte_perm(v0, v1, 0x6420), od = __byte_perm(v0, v1, 0x7531);
        const uint32_t r = __vavgu4(ev, od);
        uint8_t* dst = row + ox0;
        if (((uintptr_t)dst & 3) == 0) *reinterpret_cast<uint32_t*>(dst) = r;
        else { dst[0] = (uint8_t)r; dst[1] = (uint8_t)(r >> 8); dst[2] = (uint8_t)(r >> 16); dst[3] = (uint8_t)(r >> 24); }
    } else {
        // the columns of the odd trailing block and of the leftover follow per-column rules: ONE of them per thread, so
        // that the slow path runs once in the row's last warp instead of four times in a row
        const int ox = paired_out + (t - quads);
        if (ox < dw) row[ox] = (uint8_t)half_px(src, sw, sw, ox, oy);
    }
}

// two-third sampling, one 3x3 source block -> 2x2 destination block per thread (9 loads for 4 outputs)
__global__ void k_twothird(const uint8_t* __restrict__ src_base, uint8_t* __restrict__ dst_base, int sw, int spx, int dw, int dh, int dpx) {
    const int f = blockIdx.z;
    const int tc = blockIdx.x * blockDim.x + threadIdx.x;       // column triple
    const int tr = blockIdx.y * blockDim.y + threadIdx.y;       // row triple
    if (2 * tc >= dw || 2 * tr >= dh) return;
    const uint8_t* r1 = src_base + (size_t)f * spx + (3 * tr) * sw + 3 * tc;
    const uint8_t* r2 = r1 + sw;
    const uint8_t* r3 = r2 + sw;
    uint8_t* d1 = dst_base + (size_t)f * dpx + (2 * tr) * dw + 2 * tc;
    uint8_t* d2 = d1 + dw;
    const int a0 = r1[0], a1 = r1[1], a2 = r1[2], b0 = r2[0], b1 = r2[1], b2 = r2[2], c0 = r3[0], c1 = r3[1], c2 = r3[2];
    const int blocks = sw / 15;
    if (tc < 5 * blocks) {
        // inside a 15-column block: avg(avg(a,b),a) chains, vertically then horizontally; the fifth triple takes its
        // middle sample from the block's byte 12, i.e. from its own first column (brisk.cpp:1982-1983)
        const bool fifth = (tc % 5) == 4;
        const int u0 = avg_up(avg_up(a0, b0), a0), u2 = avg_up(avg_up(a2, b2), a2);
        const int l0 = avg_up(avg_up(c0, b0), c0), l2 = avg_up(avg_up(c2, b2), c2);
        const int um = fifth ? u0 : avg_up(avg_up(a1, b1), a1);
        const int lm = fifth ? l0 : avg_up(avg_up(c1, b1), c1);
        d1[0] = (uint8_t)avg_up(avg_up(u0, um), u0); d1[1] = (uint8_t)avg_up(avg_up(u2, um), u2);
        d2[0] = (uint8_t)avg_up(avg_up(l0, lm), l0); d2[1] = (uint8_t)avg_up(avg_up(l2, lm), l2);
    } else {
        // trailing triples: (4A + 2(B+C) + D) / 9 (brisk.cpp:2036-2051)
        d1[0] = (uint8_t)((4 * a0 + 2 * (a1 + b0) + b1) / 9); d1[1] = (uint8_t)((4 * a2 + 2 * (a1 + b2) + b1) / 9);
        d2[0] = (uint8_t)((4 * c0 + 2 * (c1 + b0) + b1) / 9); d2[1] = (uint8_t)((4 * c2 + 2 * (c1 + b2) + b1) / 9);
    }
}

void launch_pyramid(cudaStream_t st, const Geom& g, uint8_t* img) {
    // L1 = 2/3 L0; L2 = 1/2 L0; L3 = 1/2 L1; L4 = 1/2 L2; L5 = 1/2 L3   (brisk.cpp:578-587)
    static const int SRC[6] = {-1, 0, 0, 1, 2, 3};
    for (int l = 1; l < N_LAYERS; l++) {
        const LayerGeom& D = g.L[l];
        const LayerGeom& S = g.L[SRC[l]];
        dim3 block(32, 8);
        if (l == 1) {
            dim3 grid((D.w / 2 + 31) / 32, (D.h / 2 + 7) / 8, g.F);
            k_twothird<<<grid, block, 0, st>>>(img + S.base, img + D.base, S.w, S.px, D.w, D.h, D.px);
        } else {
            const int paired_out = ((S.w >> 4) >> 1) * 16;               // one thread per 4 uniform columns, then one per tail column
            dim3 grid((paired_out / 4 + (D.w - paired_out) + 31) / 32, (D.h + 7) / 8, g.F);
            k_half<<<grid, block, 0, st>>>(img + S.base, img + D.base, S.w, S.px, D.w, D.h, D.px);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// OAST-9/16 score map + per-row hit masks
// ---------------------------------------------------------------------------------------------------
// S_full(p) = clamp(max over the 16 arcs of 9 contiguous ring pixels of min(v_i - p) resp. min(p - v_i), minus
// 1, to [0, 254]) — the value cornerScore's bisection converges to from b0 = 0 (oast9_16_nms.cc:42-2115), and
// detect() at threshold b fires exactly where S_full >= b (oast9_16.cc:46).
// One launch covers all six layers of all pairs: grid = (tile of the pair over all layers, pair) of 64x32 tiles; a block of
// 256 threads stages the tile plus its 3-px halo (one TMA tensor copy where the layer's rows are 16-byte aligned, see the
// kernel) and scores it in two steps:
//   1. every pixel: the antipodal-pair BOUND  S <= max(min_j max(v_j, v_j+8) - p, p - max_j min(v_j, v_j+8)) - 1  (an arc of 9
//      contains one pixel of each antipodal pair).  bound <= 0 proves S = 0: most pixels of a difference image.  This
//      step runs on PAIRS of horizontally adjacent centres packed into the two 16-bit halves of a word
//      (VIMNMX.U16x2): the tile is staged as 16-bit pixels twice, once aligned to even columns (E) and once to odd
//      columns (O), so that the ring pixels of BOTH centres at any ring offset are ONE aligned 32-bit shared load -
//      8.5 loads and 12 packed min/max per pixel (17 and 12 with one centre per word).
//   2. the survivors (~20 % on noisy difference images), queued per warp so that full warps score them: 17 16-bit loads,
//      each widened to (v, 255 - v) in the halves of a word, so that ONE 3-input packed minimum (VIMNMX3.U16x2) serves the
//      brighter and the darker test at once: 16 + 16 packed min3 build the 16 arc minima, 8 packed max3 reduce them.
// Scores land in a shared 64x32 byte tile that leaves with 64-bit stores (zeros included: no clearing of S); hits
// (S >= 30) that do not lose the 8-neighbour test inside the tile leave as one 64-bit mask per tile row.
constexpr int SC_TW = 64, SC_TH = 32;                 // output tile
constexpr int SC_SW = SC_TW + 8, SC_SH = SC_TH + 6;   // staged: columns x0-4 .. x0+67, rows y0-3 .. y0+34
constexpr int SC_PW = SC_SW / 2;                      // 16-bit pixel pairs (words) per staged row
constexpr int SC_HITS = 256;                          // listed hits per tile (an eighth of its pixels)

struct ScoreTiles {                         // blockIdx.x = tile of the pair, layers one after the other, row-major inside a layer
    int across[N_LAYERS];                   // tile columns of each layer
    int first[N_LAYERS + 1];                // first tile of each layer; [N_LAYERS] = tiles per pair
    uint32_t magic[N_LAYERS];               // floor(2^32 / across) + 1 (0 for across = 1): tile / across by one multiply-high
};

// 8 pixels (two little-endian words w0, w1 and the byte after them) -> their four even-aligned and four odd-aligned
// 16-bit pixel pairs
__device__ __forceinline__ void widen8(uint32_t w0, uint32_t w1, uint32_t w2, uint4& even, uint4& odd) {
    const uint32_t f0 = __funnelshift_r(w0, w1, 8), f1 = __funnelshift_r(w1, w2, 8);
    even.x = __byte_perm(w0, 0, 0x4140); even.y = __byte_perm(w0, 0, 0x4342);
    even.z = __byte_perm(w1, 0, 0x4140); even.w = __byte_perm(w1, 0, 0x4342);
    odd.x = __byte_perm(f0, 0, 0x4140); odd.y = __byte_perm(f0, 0, 0x4342);
    odd.z = __byte_perm(f1, 0, 0x4140); odd.w = __byte_perm(f1, 0, 0x4342);
}

constexpr int SC_RAW = 96;                            // bytes per raw staged row of the tensor copy: columns x0-16 .. x0+79
constexpr int SC_BOX_X = 12;                          // tensor copy: the box starts at column x0-16 (box starts must be 16-byte aligned in memory:
                                                      // measured, an unaligned inner coordinate faults as an illegal instruction), i.e. 12 bytes before staged column 0

struct ScoreMaps {                          // one 2-D tensor map per layer whose rows are 16-byte aligned (see launch_scores)
    CUtensorMap map[N_LAYERS];
    int use[N_LAYERS];
};
static_assert(sizeof(ScoreMaps) == sizeof(ScoreMapCache::blob), "ScoreMapCache::blob holds a ScoreMaps");

__global__ void __launch_bounds__(256, 8)          // 8 blocks per SM: 32 registers, 18 KB of shared memory each
k_scores(const __grid_constant__ Geom g, const __grid_constant__ ScoreTiles tiles, const __grid_constant__ ScoreMaps tmaps,
         const uint8_t* __restrict__ img_all, uint8_t* __restrict__ S_all, uint8_t* __restrict__ status_all,
         unsigned long long* __restrict__ hitmask, int mask_words) {
    __shared__ __align__(16) uint32_t E[SC_SH][SC_PW];      // E[r][k] = pixels (2k, 2k+1) of staged row r, 16 bits each
    __shared__ __align__(16) uint32_t O[SC_SH][SC_PW];      // O[r][k] = pixels (2k+1, 2k+2)
    // the raw rows as the TMA engine delivers them; the survivor queues of step 2 reuse the space once the rows are widened
    __shared__ __align__(128) uint8_t raw_or_queue[SC_TW * SC_TH * 2];
    static_assert(SC_SH * SC_RAW <= SC_TW * SC_TH * 2, "raw rows must fit the queue space");
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned long long row_hits[SC_TH];      // bit x of word y: pixel (x0 + x, y0 + y) is a detector hit
    __shared__ uint16_t hit_list[SC_HITS];              // the same hits as (tile row << 6) | tile column, in no order; hits beyond the
    __shared__ int n_hits;                              // list (a tile full of corners) simply stay candidates
    __shared__ __align__(16) uint8_t score_tile[SC_TH][SC_TW];
    uint8_t* const raw = raw_or_queue;
    uint16_t* const hard = reinterpret_cast<uint16_t*>(raw_or_queue);
    // grid = (tiles of a pair, pairs): exactly the live tiles (a grid of tile columns x tile rows of all layers launched 35 %
    // of its blocks beyond the narrower layers' last column)
    const int f = blockIdx.y;
    int l = 0;
#pragma unroll
    for (int k = 1; k < N_LAYERS; k++) l += ((int)blockIdx.x >= tiles.first[k]);
    const int t = (int)blockIdx.x - tiles.first[l];
    const int trow = tiles.magic[l] ? (int)__umulhi((uint32_t)t, tiles.magic[l]) : t;
    const int tcol = t - trow * tiles.across[l];
    const LayerGeom& L = g.L[l];
    const int x0 = tcol * SC_TW, y0 = trow * SC_TH;
    const int w = L.w, h = L.h;
    const uint8_t* img = img_all + L.base + (size_t)f * L.px;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int lane = threadIdx.x, wrp = threadIdx.y;
    // ---- stage: staged row r = image row y0-3+r, staged column c = image column x0-4+c.
    //  * Layers whose rows are 16-byte aligned in memory (width and base multiples of 16: layers 0, 1 and 3 of 720x576, 82 % of
    //    the pyramid) are 2-D tensors [F * h][w] to the TMA engine: ONE tensor copy (cp.async.bulk.tensor.2d, SASS UTMALDG)
    //    brings the 38 x 96-byte box that starts at column x0-16 (a box must start on a 16-byte boundary of memory: measured, an
    //    unaligned inner coordinate faults as an illegal instruction); coordinates outside the tensor are filled with zeros,
    //    rows above / below the pair's image are the neighbouring pair's rows.  The threads then widen the raw bytes from
    //    shared memory.
    //  * Any other layer: the threads read the rows themselves (aligned words where the layer allows it; one bulk copy per
    //    row, 38 per tile, measured slower than these loads: 3.02 vs 2.88 ms per 1024 pairs), rows outside the image repeat
    //    the border row and columns outside read as 0.
    // What lands outside the image - left of column 0, right of column w-1, above row 0, below row h-1 - is never looked at:
    // those are only ever ring pixels of border centres, which are not scored. ----
    const bool boxed = tmaps.use[l] != 0;
    if (boxed && tid == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, SC_SH * SC_RAW);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(smem_addr(raw)), "l"(reinterpret_cast<uint64_t>(&tmaps.map[l])), "r"(x0 - 4 - SC_BOX_X), "r"(f * h + y0 - 3),
                        "r"(smem_addr(&bar)) : "memory");
    }
    if (tid < SC_TH) row_hits[tid] = 0ull;
    if (tid == SC_TH) n_hits = 0;
    reinterpret_cast<unsigned long long*>(&score_tile[0][0])[tid] = 0ull;          // 256 threads x 8 bytes = the tile
    // item (r, j) = the 8 pixels from staged column 8j of row r -> four even-aligned and four odd-aligned 16-bit pairs.
    // 38 rows x 9 items, at most two per thread and no index division: the first 32 rows x 8 items are one item per thread,
    // the remaining 6 rows x 8 items and the ninth item of all 38 rows go to the first 86 threads
    if (boxed) {
        __syncthreads();                               // the initialised barrier is visible to the waiters
        mbar_wait(&bar, 0);
        auto widen_item = [&](const int r, const int j) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(raw + r * SC_RAW + SC_BOX_X) + 2 * j;      // words 3 + 2j .. 5 + 2j of 24
            uint4 ev, od;
            widen8(q[0], q[1], q[2], ev, od);
            *reinterpret_cast<uint4*>(&E[r][4 * j]) = ev;
            *reinterpret_cast<uint4*>(&O[r][4 * j]) = od;
        };
        widen_item(tid >> 3, tid & 7);
        if (tid < (SC_SH - 32) * 8) widen_item(32 + (tid >> 3), tid & 7);
        else if (tid < (SC_SH - 32) * 8 + SC_SH) widen_item(tid - (SC_SH - 32) * 8, 8);
    } else {
        const bool word_rows = ((((uintptr_t)img) | (uintptr_t)w) & 3) == 0;      // every staged item then starts on a word of the image
        auto stage_item = [&](const int r, const int j) {
            const int gy = min(max(y0 - 3 + r, 0), h - 1);
            const int gx = x0 - 4 + 8 * j;
            const uint8_t* p = img + gy * w + gx;
            uint32_t w0, w1, w2;
            if (word_rows && gx >= 0 && gx + 12 <= w) {                         // interior of the row: three aligned words
                const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
                w0 = q[0]; w1 = q[1]; w2 = q[2];
            } else if (gx >= w) {                                                // right of the image: nothing to read
                w0 = w1 = w2 = 0u;
            } else {
                w0 = (gx >= 0) ? load4(p) : 0u;
                w1 = (gx + 4 < w) ? load4(p + 4) : 0u;                          // gx + 4 >= 0 always
                w2 = (gx + 8 < w) ? load4(p + 8) : 0u;
            }
            uint4 ev, od;
            widen8(w0, w1, w2, ev, od);
            *reinterpret_cast<uint4*>(&E[r][4 * j]) = ev;
            *reinterpret_cast<uint4*>(&O[r][4 * j]) = od;
        };
        stage_item(tid >> 3, tid & 7);
        if (tid < (SC_SH - 32) * 8) stage_item(32 + (tid >> 3), tid & 7);
        else if (tid < (SC_SH - 32) * 8 + SC_SH) stage_item(tid - (SC_SH - 32) * 8, 8);
    }
    __syncthreads();
    // ---- step 1: a thread owns the centre pairs (x0 + 2 lane, +1) of rows y0 + ty + 8 jy; staged column of the pair's first
    // pixel is 4 + 2 lane = word 2 + lane of E.
    // Each warp queues its own survivors (no block-wide scan or barrier) in (pixel slot, lane) order, so that queue neighbours are
    // tile-row neighbours and the ring loads of step 2 stay nearly free of bank conflicts (lane-major order measured 57 M
    // conflict wavefronts per launch instead of 32 M): one ballot per pixel slot, a survivor's place is the slot's base plus the
    // survivors on lower lanes (POPC runs on the otherwise idle XU pipe; a byte-packed shuffle scan of all eight slots at once
    // cost 120 instructions per warp, this 70).
    const int xa = x0 + 2 * lane;
    const bool in_lo = xa >= 3 && xa < w - 3, in_hi = xa + 1 >= 3 && xa + 1 < w - 3;
    uint16_t* queue = hard + wrp * (SC_TW * SC_TH / 8);
    const uint32_t below = (1u << lane) - 1u;
    int nh = 0;
#pragma unroll
    for (int jy = 0; jy < SC_TH / 8; jy++) {
        const int ly = wrp + 8 * jy;
        const int y = y0 + ly;
        if (!(y >= 3 && y < h - 3)) continue;                      // the same for the whole warp
        uint32_t both = 0x01000100u;                               // halves: bound + 256 of the two centres
        if (in_lo || in_hi) {
            const uint32_t* e = &E[ly + 3][lane + 2];
            const uint32_t* o = &O[ly + 3][lane + 2];
            // ring in circle order (oast9_16.h:214-232); offset dx even -> E[k + dx/2], odd -> O[k + (dx-1)/2]
            uint32_t r[16];
            r[0] = o[-2];                 r[1] = o[-SC_PW - 2];         r[2] = e[-2 * SC_PW - 1];     r[3] = o[-3 * SC_PW - 1];
            r[4] = e[-3 * SC_PW];         r[5] = o[-3 * SC_PW];         r[6] = e[-2 * SC_PW + 1];     r[7] = o[-SC_PW + 1];
            r[8] = o[1];                  r[9] = o[SC_PW + 1];          r[10] = e[2 * SC_PW + 1];     r[11] = o[3 * SC_PW];
            r[12] = e[3 * SC_PW];         r[13] = o[3 * SC_PW - 1];     r[14] = e[2 * SC_PW - 1];     r[15] = o[SC_PW - 2];
            both = oast_bound_pair(r, e[0]);
        }
        const bool s_lo = in_lo && (both & 0xffffu) > 256u, s_hi = in_hi && (both >> 16) > 256u;    // these pixels need the full score
        const uint32_t b_lo = __ballot_sync(0xffffffffu, s_lo), b_hi = __ballot_sync(0xffffffffu, s_hi);
        const uint32_t at = (uint32_t)((ly << 6) | (2 * lane));    // (tile row << 6) | tile column
        if (s_lo) queue[nh + __popc(b_lo & below)] = (uint16_t)at;
        nh += __popc(b_lo);
        if (s_hi) queue[nh + __popc(b_hi & below)] = (uint16_t)(at + 1u);
        nh += __popc(b_hi);
    }
    __syncwarp();
    // ---- step 2: full closed-form score of the queued pixels ----
    const uint16_t* px16 = reinterpret_cast<const uint16_t*>(&E[0][0]);       // E as 16-bit pixels: staged (row, column)
    for (int i = lane; i < nh; i += 32) {
        const int ly = queue[i] >> 6, lx = queue[i] & 63;
        const uint16_t* c = px16 + (ly + 3) * SC_SW + lx + 4;
        uint32_t r[16];
#define MFK_W(v) (0x00ff0000u - (uint32_t)(v) * 65535u)                       /* v | (255 - v) << 16 */
        r[0] = MFK_W(c[-3]);                 r[1] = MFK_W(c[-SC_SW - 3]);         r[2] = MFK_W(c[-2 * SC_SW - 2]);     r[3] = MFK_W(c[-3 * SC_SW - 1]);
        r[4] = MFK_W(c[-3 * SC_SW]);         r[5] = MFK_W(c[-3 * SC_SW + 1]);     r[6] = MFK_W(c[-2 * SC_SW + 2]);     r[7] = MFK_W(c[-SC_SW + 3]);
        r[8] = MFK_W(c[3]);                  r[9] = MFK_W(c[SC_SW + 3]);          r[10] = MFK_W(c[2 * SC_SW + 2]);     r[11] = MFK_W(c[3 * SC_SW + 1]);
        r[12] = MFK_W(c[3 * SC_SW]);         r[13] = MFK_W(c[3 * SC_SW - 1]);     r[14] = MFK_W(c[2 * SC_SW - 2]);     r[15] = MFK_W(c[SC_SW - 3]);
        const int sc = oast_score_packed(r, MFK_W(c[0]));
#undef MFK_W
        score_tile[ly][lx] = (uint8_t)sc;
        if (sc >= BRISK_THRESHOLD) {
            atomicOr(reinterpret_cast<unsigned*>(&row_hits[ly]) + (lx >> 5), 1u << (lx & 31));   // native 32-bit shared atomic
            const int k = atomicAdd(&n_hits, 1);                   // the tile's hits, for the 8-neighbour test below (a few dozen)
            if (k < SC_HITS) hit_list[k] = queue[i];
        }
    }
    __syncthreads();
    // ---- hits that cannot be maxima leave here: a hit with a strictly larger score among its 8 neighbours fails isMax2D
    // (brisk.cpp:838-869) whatever else happens - k_pa_test would only give it status 0 and drop it - and for a hit inside the
    // tile's 1-pixel frame all 8 neighbours are in the shared score tile.  Three of four hits go this way (measured on the
    // synthetic streams: 72 % lose, 11 % tie, 17 % win), so the candidate list, its fill, the 8-neighbour kernel and the second
    // compaction shrink to a third.  Ties and hits on the tile's frame stay candidates. ----
    // The hits were listed by step 2, so a few threads test one hit each (walking the hit masks instead, a thread per 8 pixels,
    // made every warp run the test's 40 instructions for its one or two hits: + 17 % on the kernel).
    if (status_all != nullptr) {
        const int n_listed = min(n_hits, SC_HITS);
        for (int k = tid; k < n_listed; k += 256) {
            const int ly = hit_list[k] >> 6, lx = hit_list[k] & 63;
            if (ly < 1 || ly > SC_TH - 2 || lx < 1 || lx > SC_TW - 2) continue;
            const int c = score_tile[ly][lx];
            const int up = max(max((int)score_tile[ly - 1][lx - 1], (int)score_tile[ly - 1][lx]), (int)score_tile[ly - 1][lx + 1]);
            const int dn = max(max((int)score_tile[ly + 1][lx - 1], (int)score_tile[ly + 1][lx]), (int)score_tile[ly + 1][lx + 1]);
            const int mx = max(max(up, dn), max((int)score_tile[ly][lx - 1], (int)score_tile[ly][lx + 1]));
            if (mx > c) {
                atomicAnd(reinterpret_cast<unsigned*>(&row_hits[ly]) + (lx >> 5), ~(1u << (lx & 31)));
                status_all[L.base + (size_t)f * L.px + (size_t)(y0 + ly) * w + x0 + lx] = 0;         // as k_pa_test leaves a loser
            }
        }
    }
    // the score tile, 8 bytes per thread: one 64-bit store where the layer's rows allow it, else two words, else bytes
    {
        const int ry = tid >> 3, cx = (tid & 7) * 8;
        const int y = y0 + ry;
        if (y < h) {
            uint8_t* dst = S_all + L.base + (size_t)f * L.px + (size_t)y * w + x0 + cx;
            const uint2 v = *reinterpret_cast<const uint2*>(&score_tile[ry][cx]);
            const uint32_t a = (uint32_t)(uintptr_t)dst;                        // odd frame sizes shift the layer bases and rows
            if (x0 + cx + 8 <= w && (a & 3u) == 0) {
                if ((a & 7u) == 0) *reinterpret_cast<uint2*>(dst) = v;
                else { reinterpret_cast<uint32_t*>(dst)[0] = v.x; reinterpret_cast<uint32_t*>(dst)[1] = v.y; }
            } else {
                for (int j = 0; j < 8; j++) if (x0 + cx + j < w) dst[j] = score_tile[ry][cx + j];
            }
        }
    }
    __syncthreads();
    // one 64-bit word per (layer row, tile column); every word of the tile is written, so the map needs no clearing
    if (tid < SC_TH && y0 + tid < h)
        hitmask[((size_t)f * g.rows + L.row0 + y0 + tid) * mask_words + tcol] = row_hits[tid];
}

int score_mask_words(int frame_width) { return (frame_width + SC_TW - 1) / SC_TW; }
// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup: the library does not link libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (getenv("MFK_SCORES_NO_TENSOR_MAP")) return (EncodeTiledFn) nullptr;       // A/B measurements: thread loads everywhere
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (EncodeTiledFn) nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// The tensor maps describe the lane's pyramid buffer for one batch size (the layer bases depend on F): rebuilt when either
// changes, i.e. practically once per lane.  A layer without 16-byte aligned rows (or a driver without the entry point) keeps
// use = 0 and is staged by the threads' own loads.
static void refresh_score_maps(ScoreMapCache& cache, const Geom& g, const uint8_t* img) {
    if (cache.F == g.F && cache.img == img) return;
    ScoreMaps& sm = *reinterpret_cast<ScoreMaps*>(cache.blob);
    memset(&sm, 0, sizeof(sm));
    EncodeTiledFn enc = encode_tiled();
    for (int l = 0; l < N_LAYERS && enc; l++) {
        const LayerGeom& L = g.L[l];
        const uintptr_t base = (uintptr_t)(img + L.base);
        if ((L.w & 15) || (base & 15) || L.w < 16) continue;
        const cuuint64_t dims[2] = {(cuuint64_t)L.w, (cuuint64_t)L.h * (cuuint64_t)g.F};
        const cuuint64_t strides[1] = {(cuuint64_t)L.w};
        const cuuint32_t box[2] = {(cuuint32_t)SC_RAW, (cuuint32_t)SC_SH};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult rc = enc(&sm.map[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        sm.use[l] = rc == CUDA_SUCCESS;
    }
    cache.F = g.F; cache.img = img;
}

// status: the candidate status map when hits that lose the 8-neighbour test inside their tile may leave at once (they get
// status 0 here), nullptr to keep every hit a candidate
void launch_scores(cudaStream_t st, const Geom& g, const uint8_t* img, uint8_t* S, uint8_t* status, unsigned long long* hitmask,
                   ScoreMapCache& cache) {
    ScoreTiles tiles;
    int n = 0;
    for (int l = 0; l < N_LAYERS; l++) {
        const int across = (g.L[l].w + SC_TW - 1) / SC_TW, down = (g.L[l].h + SC_TH - 1) / SC_TH;
        tiles.across[l] = across;
        tiles.first[l] = n;
        tiles.magic[l] = across <= 1 ? 0u : (uint32_t)(0x100000000ull / (uint32_t)across) + 1u;      // exact for tile * across < 2^32
        n += across * down;
    }
    tiles.first[N_LAYERS] = n;
    refresh_score_maps(cache, g, img);
    // gridDim.y <= 65535 pairs per launch sequence (context.cu caps max_pairs far below)
    k_scores<<<dim3(n, g.F), dim3(32, 8), 0, st>>>(g, tiles, *reinterpret_cast<const ScoreMaps*>(cache.blob), img, S, status, hitmask,
                                                               score_mask_words(g.W));
}

// ---------------------------------------------------------------------------------------------------
// ordered candidate list: one warp per (pair, layer row), hits written at the row's scanned offset in x order
// ---------------------------------------------------------------------------------------------------
// hits per (pair, layer row) from the hit masks: the input of the exclusive scan that orders the candidate list
__global__ void k_row_counts(Geom g, const unsigned long long* __restrict__ hitmask, int mask_words, uint32_t* __restrict__ row_count) {
    const int rowi = blockIdx.x * blockDim.x + threadIdx.x;
    const int total_rows = g.F * g.rows;
    if (rowi > total_rows) return;
    if (rowi == total_rows) { row_count[rowi] = 0; return; }
    const int rr = rowi % g.rows;
    int l = 0;
#pragma unroll
    for (int k = 1; k < N_LAYERS; k++) l += (rr >= g.L[k].row0);
    const int words = (g.L[l].w + SC_TW - 1) / SC_TW;
    const unsigned long long* m = hitmask + (size_t)rowi * mask_words;
    int n = 0;
    for (int k = 0; k < words; k++) n += __popcll(m[k]);
    row_count[rowi] = (uint32_t)n;
}

void launch_row_counts(cudaStream_t st, const Geom& g, const unsigned long long* hitmask, uint32_t* row_count) {
    const int n = g.F * g.rows + 1;
    k_row_counts<<<(n + 255) / 256, 256, 0, st>>>(g, hitmask, score_mask_words(g.W), row_count);
}

// ---------------------------------------------------------------------------------------------------
// ordered candidate list: one THREAD per (pair, layer) row walks the row's hit-mask words and writes the hits in x order at
// the row's scanned offset.  (One warp per row with a shuffle scan of the word counts cost 40x the instructions: a row has
// ~8 hits in <= 12 words, far too little for 32 lanes.)
// ---------------------------------------------------------------------------------------------------
__global__ void k_fill_candidates(Geom g, const unsigned long long* __restrict__ hitmask, int mask_words, CandBuf cb) {
    const int rowi = blockIdx.x * blockDim.x + threadIdx.x;
    const int total_rows = g.F * g.rows;
    if (rowi == 0 && cb.row_count[total_rows] > (uint32_t)g.cand_cap) atomicOr(&cb.live[1], 1u);
    if (rowi >= total_rows) return;
    uint32_t at = cb.row_count[rowi];
    const uint32_t end = cb.row_count[rowi + 1];
    if (at == end) return;
    const int f = rowi / g.rows, rr = rowi - f * g.rows;
    int l = 0;
#pragma unroll
    for (int k = 1; k < N_LAYERS; k++) l += (rr >= g.L[k].row0);
    const uint32_t y16 = (uint32_t)(rr - g.L[l].row0) << 16;
    const uint32_t meta = ((uint32_t)f << 3) | (uint32_t)l;
    const int words = (g.L[l].w + SC_TW - 1) / SC_TW;
    const unsigned long long* m = hitmask + (size_t)rowi * mask_words;
    const uint32_t cap = (uint32_t)g.cand_cap;
    for (int k = 0; k < words && at < end; k++) {
        unsigned long long bits = m[k];
        while (bits) {
            const int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            if (at < cap) {
                cb.xy[at] = (uint32_t)(k * SC_TW + b) | y16;
                cb.meta[at] = meta;
                cb.flags[at] = 0;
            }
            at++;
        }
    }
}

void launch_fill_candidates(cudaStream_t st, const Geom& g, const unsigned long long* hitmask, CandBuf cb) {
    const int rows = g.F * g.rows;
    k_fill_candidates<<<(rows + 127) / 128, 128, 0, st>>>(g, hitmask, score_mask_words(g.W), cb);
}

// ---------------------------------------------------------------------------------------------------
// integral image of the difference frame, int32 (H+1) x (W+1) with a zero top row / left column
// ---------------------------------------------------------------------------------------------------
// Two passes over the image, one write of the integral (the classic row-scan + column-scan writes and re-reads it):
//   k_integral_left    per image row: left[y][s] = sum of the pixels left of 128-column strip s
//   k_integral_strips  one warp per (pair, strip) walks down the rows.  A lane owns 4 adjacent columns (one 32-bit read
//                      per row): byte prefix inside the lane, one warp scan of the lane totals, four running column sums
//                      in registers; the row's 128 results are transposed through shared memory so that every store
//                      instruction writes 128 contiguous bytes.  0.3 instructions per pixel (2.1 with one column per lane).
constexpr int IG_STRIP = 128;

__global__ void k_integral_left(const uint8_t* __restrict__ img, int32_t* __restrict__ left, int W, int H, int F, int strips) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= F * H) return;
    const uint8_t* row = img + (size_t)warp * W;                       // rows of all pairs are contiguous (layer 0 block)
    int32_t* out = left + (size_t)warp * strips;
    int carry = 0;
    for (int s = 0; s < strips; s++) {                                 // the warp's 32 words of this iteration ARE strip s
        const int x = s * IG_STRIP + 4 * lane;
        int sum = 0;
        if (x + 4 <= W) sum = (int)__dp4a(load4(row + x), 0x01010101u, 0u);
        else for (int j = 0; j < 4; j++) if (x + j < W) sum += row[x + j];
#pragma unroll
        for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (lane == 0) out[s] = carry;
        carry += sum;
    }
}

constexpr int IG_AHEAD = 8;       // rows fetched ahead of the dependent scan chain
constexpr int IG_WARPS = 4;

// one image row of one strip: the lane's four pixels v (packed), the sum of the row left of the strip -> running column sums,
// transposed through shared memory so that each store instruction covers 128 contiguous bytes
__device__ __forceinline__ void integral_row(uint32_t v, int left_of_strip, int lane, int& acc0, int& acc1, int& acc2, int& acc3,
                                             int32_t* xp, int32_t* o, const bool st0, const bool st1, const bool st2,
                                             const bool st3, const bool first_strip) {
    const int p0 = (int)(v & 0xffu), p1 = p0 + (int)((v >> 8) & 0xffu), p2 = p1 + (int)((v >> 16) & 0xffu), p3 = p2 + (int)(v >> 24);
    int incl = p3;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    const int t = incl - p3 + left_of_strip;                      // row sum left of this lane: inside the strip + left of the strip
    acc0 += t + p0; acc1 += t + p1; acc2 += t + p2; acc3 += t + p3;   // running column sums = the integral itself
    *reinterpret_cast<int4*>(&xp[4 * lane]) = make_int4(acc0, acc1, acc2, acc3);
    __syncwarp();
    if (st0) o[lane] = xp[lane];
    if (st1) o[32 + lane] = xp[32 + lane];
    if (st2) o[64 + lane] = xp[64 + lane];
    if (st3) o[96 + lane] = xp[96 + lane];
    if (first_strip && lane == 0) o[-1] = 0;                       // zero left column
    __syncwarp();
}

// ALIGNED: frame width and layer base are multiples of 4, so a lane's four pixels are one aligned word
template <bool ALIGNED>
__global__ void __launch_bounds__(IG_WARPS * 32)
k_integral_strips(const uint8_t* __restrict__ img, const int32_t* __restrict__ left, int32_t* __restrict__ integral,
                  int W, int H, int F, int strips) {
    __shared__ __align__(16) int32_t xpose[IG_WARPS][IG_STRIP];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (warp >= F * strips) return;
    const int f = warp / strips, s = warp - f * strips;
    const int x = s * IG_STRIP + 4 * lane;                               // first of this lane's four columns
    const int mine = min(max(W - x, 0), 4);                              // how many of them exist
    const uint32_t keep = mine == 4 ? 0xffffffffu : ((1u << (8 * mine)) - 1u);
    const uint8_t* col = img + (size_t)f * W * H + (mine ? x : 0);       // walks down one row per step
    const int32_t* lf = left + (size_t)f * H * strips + s;
    int32_t* o = integral + (size_t)f * (H + 1) * (W + 1) + 1 + s * IG_STRIP;        // column s * 128 of the image = integral column + 1
    const bool st0 = s * IG_STRIP + lane < W, st1 = s * IG_STRIP + 32 + lane < W, st2 = s * IG_STRIP + 64 + lane < W,
               st3 = s * IG_STRIP + 96 + lane < W;
    const bool first_strip = s == 0;
    if (st0) o[lane] = 0;                                                // zero top row
    if (st1) o[32 + lane] = 0;
    if (st2) o[64 + lane] = 0;
    if (st3) o[96 + lane] = 0;
    if (first_strip && lane == 0) o[-1] = 0;
    int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    int32_t* xp = xpose[wib];
    const int W1 = W + 1;
    int y = 0;
    for (; y + IG_AHEAD <= H; y += IG_AHEAD) {                           // whole groups: loads first, then the dependent chain
        uint32_t px[IG_AHEAD];
        int lv[IG_AHEAD];
#pragma unroll
        for (int k = 0; k < IG_AHEAD; k++) {
            px[k] = ALIGNED ? *reinterpret_cast<const uint32_t*>(col) : load4(col);
            lv[k] = *lf;
            col += W; lf += strips;
        }
#pragma unroll
        for (int k = 0; k < IG_AHEAD; k++) {
            o += W1;
            integral_row(px[k] & keep, lv[k], lane, acc0, acc1, acc2, acc3, xp, o, st0, st1, st2, st3, first_strip);
        }
    }
    for (; y < H; y++) {                                                 // the last H % 8 rows
        const uint32_t v = ALIGNED ? *reinterpret_cast<const uint32_t*>(col) : load4(col);
        const int l = *lf;
        col += W; lf += strips; o += W1;
        integral_row(v & keep, l, lane, acc0, acc1, acc2, acc3, xp, o, st0, st1, st2, st3, first_strip);
    }
}

void launch_integral(cudaStream_t st, const Geom& g, const uint8_t* img, int32_t* integral, int32_t* left, bool left_done) {
    const int strips = (g.W + IG_STRIP - 1) / IG_STRIP;
    const int warps1 = g.F * g.H;
    if (!left_done) k_integral_left<<<(warps1 * 32 + 255) / 256, 256, 0, st>>>(img, left, g.W, g.H, g.F, strips);
    const int warps2 = g.F * strips;
    const int blocks = (warps2 + IG_WARPS - 1) / IG_WARPS;
    if ((((uintptr_t)img | (uintptr_t)g.W) & 3) == 0) k_integral_strips<true><<<blocks, IG_WARPS * 32, 0, st>>>(img, left, integral, g.W, g.H, g.F, strips);
    else k_integral_strips<false><<<blocks, IG_WARPS * 32, 0, st>>>(img, left, integral, g.W, g.H, g.F, strips);
}

// ---------------------------------------------------------------------------------------------------
// exclusive scan (three small kernels; sizes here are at most a few million words)
// ---------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += n;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t before = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return before + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(uint32_t* data, int n_max, const uint32_t* d_n, uint32_t* partial) {
    const int n = d_n ? min((int)*d_n, n_max) : n_max;
    const int tiles = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int base = tile * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? data[base + k] : 0; sum += v[k]; }
        uint32_t total;
        uint32_t run = block_exclusive_scan(sum, &total);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) data[base + k] = run; run += v[k]; }
        if (threadIdx.x == 0) partial[tile] = total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partials(uint32_t* partial, int n_max, const uint32_t* d_n) {
    // single block, serial over chunks of SCAN_BLOCK partials; partial[tiles] receives the grand total
    const int n = d_n ? min((int)*d_n, n_max) : n_max;
    const int n_blocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int c = 0; c < n_blocks; c += SCAN_BLOCK) {
        const int base = c + threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n_blocks) ? partial[base + k] : 0; sum += v[k]; }
        uint32_t total;
        uint32_t run = block_exclusive_scan(sum, &total) + carry_s;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n_blocks) partial[base + k] = run; run += v[k]; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[n_blocks] = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* data, int n_max, const uint32_t* d_n, const uint32_t* partial) {
    const int n = d_n ? min((int)*d_n, n_max) : n_max;
    const int tiles = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t add = partial[tile];
        const int base = tile * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++)
            if (base + k < n) data[base + k] += add;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = partial[tiles];
}

void launch_exclusive_scan(cudaStream_t st, uint32_t* data, int n, const uint32_t* d_n, uint32_t* scratch) {
    const int tiles = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    const int grid = tiles < 148 * 8 ? (tiles > 0 ? tiles : 1) : 148 * 8;      // the live length may be far below the bound
    k_scan_blocks<<<grid, SCAN_THREADS, 0, st>>>(data, n, d_n, scratch);
    k_scan_partials<<<1, SCAN_THREADS, 0, st>>>(scratch, n, d_n);
    k_scan_add<<<grid, SCAN_THREADS, 0, st>>>(data, n, d_n, scratch);
}

}  // namespace mfk
