// Fused dense front of a launch sequence, one kernel instead of seven:
//   |I_t - I_{t-5}|              cv::absdiff                                  MoFREAKUtilities.cpp:413-414
//   pyramid layers 1..5          BriskLayer::twothirdsample / halfsample     brisk.cpp:1840-2065, driven as :572-588
//   per-row strip sums of layer 0 for the integral image (cv::integral inside cv::FREAK::compute)
// A block owns a BAND of 12 source rows of one frame pair: 12 = lcm of the samplers' row groups (3 for the two-third
// sampler, 2 for each of the three halvings), so that every output row of every layer depends on rows of ONE band:
//   L0 12 rows -> L1 8 -> L3 4 -> L5 2,   L0 12 -> L2 6 -> L4 3.
// The two frames' bands are staged in shared memory - when the frame width is a multiple of 16 by two bulk copies of the
// TMA engine (cp.async.bulk, 12 x W contiguous bytes each, completion on an mbarrier; SASS UBLKCP), else by aligned word
// reads - the difference is taken in place, and all six layers and the row sums are produced from shared memory: the
// frames are read once and every pyramid byte is written once.  (The separate kernels read layer 0 four times and
// layers 1..3 twice.)  The samplers' column rules (SSE blocks, tails, quirks) are those of dense_device.cuh.
#include "brisk_device.cuh"
#include "dense_device.cuh"
#include "kernels.h"
#include "tma_device.cuh"

namespace mfk {

namespace {

constexpr int FR_BAND = 12;
constexpr int FR_THREADS = 256;
constexpr int FR_MAX_W = 2048;            // wider frames take the separate kernels (shared memory per band)
constexpr int IG_STRIP_W = 128;           // strip width of k_integral_strips (kernels_dense.cu)
constexpr int FR_BAND_PX = 8640;          // pixels of layer 0 a block should own (12 rows of 720)
constexpr int FR_MULT_MAX = 8;
constexpr int FR_SUMS = 256;              // (rows, strip) sums of one band: rows * ceil(W / 128) <= 12 * mult * (W / 128 + 1)

struct FrontLayout {                      // shared-memory plan of one band
    int mult;                             // a band is 12 * mult source rows: narrow frames take taller bands, so that a block has ~8K pixels
    int pitch[N_LAYERS];                  // row pitch of each layer's band, a multiple of 16
    int off[N_LAYERS];                    // byte offsets: layer 0 = current frame (becomes the difference), then layers 1..5
    int off_prev;                         // previous frame's band
    int total;
    uint32_t mg_items[N_LAYERS];          // div_magic of the work items per destination row (group) of each layer's sampler
    uint32_t mg_words[N_LAYERS];          // div_magic of the words per row of each layer (copy_out)
};

// i / d for i < 65536, d <= 2048 by one multiply: m = floor(2^32 / d) + 1 is exact for i * d < 2^32 (0 stands for d = 1)
// (computed on the host, FrontLayout: the 64-bit division costs a thread more than all the divisions it replaces)
inline uint32_t div_magic(int d) { return d <= 1 ? 0u : (uint32_t)(0x100000000ull / (uint32_t)d) + 1u; }
__device__ __forceinline__ int fast_div(int i, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)i, magic) : i; }

// rows * w bytes from shared rows of pitch sp to global rows of pitch w (dense), widest stores the alignment allows
__device__ __forceinline__ void copy_out(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int sp, int w, int rows, int tid,
                                         const uint32_t mg) {
    if (rows <= 0) return;
    if (sp == w && (((uintptr_t)dst) & 15) == 0) {
        const uint4* s = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        const int n = rows * w / 16;                        // sp == w is a multiple of 16
        for (int i = tid; i < n; i += FR_THREADS) d[i] = s[i];
    } else if (((((uintptr_t)dst) | (uintptr_t)w) & 3) == 0) {
        const int words = w >> 2;
        for (int i = tid; i < rows * words; i += FR_THREADS) {
            const int r = fast_div(i, mg), k = i - r * words;
            reinterpret_cast<uint32_t*>(dst + (size_t)r * w)[k] = reinterpret_cast<const uint32_t*>(src + r * sp)[k];
        }
    } else {
        for (int i = tid; i < rows * w; i += FR_THREADS) {
            const int r = i / w, x = i - r * w;
            dst[(size_t)r * w + x] = src[r * sp + x];
        }
    }
}

// BriskLayer::halfsample on `rows` destination rows: source rows 2j, 2j+1 of a shared band (width sw, pitch sp) ->
// shared rows of pitch dp.  Inside the paired 16-byte blocks four outputs come from two 8-byte reads per source row
// (avg of avg, rounding up); the columns of an odd trailing block and of the leftover follow half_px's per-column rules.
__device__ __forceinline__ void half_band(const uint8_t* __restrict__ src, int sp, int sw, uint8_t* __restrict__ dst, int dp, int dw,
                                          int rows, int tid, const uint32_t mg) {
    if (rows <= 0) return;
    const int paired_out = ((sw >> 4) >> 1) * 16;
    const int quads = paired_out >> 2;
    const int items = quads + (dw - paired_out);
    for (int i = tid; i < rows * items; i += FR_THREADS) {
        const int j = fast_div(i, mg), t = i - j * items;
        if (t < quads) {
            const uint2 u = *reinterpret_cast<const uint2*>(src + (2 * j) * sp + 8 * t);
            const uint2 l = *reinterpret_cast<const uint2*>(src + (2 * j + 1) * sp + 8 * t);
            const uint32_t v0 = __vavgu4(u.x, l.x), v1 = __vavgu4(u.y, l.y);
            const uint32_t ev = __byte_perm(v0, v1, 0x6420), od = __byte_perm(v0, v1, 0x7531);
            *reinterpret_cast<uint32_t*>(dst + j * dp + 4 * t) = __vavgu4(ev, od);
        } else {
            const int ox = paired_out + (t - quads);
            dst[j * dp + ox] = (uint8_t)half_px(src, sp, sw, ox, j);
        }
    }
}

// one 3x3 source block -> its 2x2 outputs (the per-triple form of BriskLayer::twothirdsample)
__device__ __forceinline__ void twothird_triple(const uint8_t* __restrict__ r1, int sp, uint8_t* __restrict__ d1, int dp, bool in_block, bool fifth) {
    const uint8_t* r2 = r1 + sp;
    const uint8_t* r3 = r2 + sp;
    uint8_t* d2 = d1 + dp;
    const int a0 = r1[0], a1 = r1[1], a2 = r1[2], b0 = r2[0], b1 = r2[1], b2 = r2[2], c0 = r3[0], c1 = r3[1], c2 = r3[2];
    if (in_block) {
        // inside a 15-column block: avg(avg(a,b),a) chains, vertically then horizontally; the fifth triple takes its
        // middle sample from the block's byte 12, i.e. from its own first column (brisk.cpp:1982-1983)
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

// horizontal pass of four triples at once: L, M, R hold the left, middle and right column of each triple (one byte per
// triple); returns the eight outputs in order
__device__ __forceinline__ uint2 twothird_horizontal4(uint32_t L, uint32_t M, uint32_t R) {
    const uint32_t ol = __vavgu4(__vavgu4(L, M), L), orr = __vavgu4(__vavgu4(R, M), R);
    return make_uint2(__byte_perm(ol, orr, 0x5140), __byte_perm(ol, orr, 0x7362));
}

// BriskLayer::twothirdsample on `groups` row triples of a shared band.  A thread takes FOUR column triples of one row
// triple: 12 source bytes per row are three aligned words, the avg(avg(a,b),a) chains run on packed bytes (the SSE
// code's own formulation, _mm_avg_epu8), byte permutes gather each triple's left / middle / right column; triples of the
// trailing rule and a last incomplete quad take the per-triple form.
__device__ __forceinline__ void twothird_band(const uint8_t* __restrict__ src, int sp, int sw, uint8_t* __restrict__ dst, int dp, int dw,
                                              int groups, int tid, const uint32_t mg) {
    if (groups <= 0) return;
    const int triples = dw >> 1;
    const int in_blocks = 5 * (sw / 15);                   // triples that follow the block rule
    const int quads = (triples + 3) >> 2;
    for (int i = tid; i < groups * quads; i += FR_THREADS) {
        const int gt = fast_div(i, mg), q = i - gt * quads;
        const uint8_t* r1 = src + (3 * gt) * sp + 12 * q;
        uint8_t* d1 = dst + (2 * gt) * dp + 8 * q;
        if (4 * q + 4 <= in_blocks) {
            const uint32_t* a = reinterpret_cast<const uint32_t*>(r1);
            const uint32_t* b = reinterpret_cast<const uint32_t*>(r1 + sp);
            const uint32_t* c = reinterpret_cast<const uint32_t*>(r1 + 2 * sp);
            uint32_t U[3], D[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const uint32_t av = a[k], bv = b[k], cv = c[k];
                U[k] = __vavgu4(__vavgu4(av, bv), av);
                D[k] = __vavgu4(__vavgu4(cv, bv), cv);
            }
            // the triple whose index is 4 (mod 5) is the fifth of its 15-column block: its middle column is its left one
            const int kk = (4 + q) - 5 * (int)__umulhi((uint32_t)(4 + q), 0x33333334u);      // (4 + q) % 5
            const uint32_t fifth = kk < 4 ? 0xffu << (8 * kk) : 0u;
            uint32_t L = __byte_perm(__byte_perm(U[0], U[1], 0x0630), U[2], 0x5210);
            uint32_t M = __byte_perm(__byte_perm(U[0], U[1], 0x0741), U[2], 0x6210);
            uint32_t R = __byte_perm(__byte_perm(U[0], U[1], 0x0052), U[2], 0x7410);
            M = (M & ~fifth) | (L & fifth);
            *reinterpret_cast<uint2*>(d1) = twothird_horizontal4(L, M, R);
            L = __byte_perm(__byte_perm(D[0], D[1], 0x0630), D[2], 0x5210);
            M = __byte_perm(__byte_perm(D[0], D[1], 0x0741), D[2], 0x6210);
            R = __byte_perm(__byte_perm(D[0], D[1], 0x0052), D[2], 0x7410);
            M = (M & ~fifth) | (L & fifth);
            *reinterpret_cast<uint2*>(d1 + dp) = twothird_horizontal4(L, M, R);
        } else {
            for (int k = 0; k < 4 && 4 * q + k < triples; k++) {
                const int tc = 4 * q + k;
                twothird_triple(r1 + 3 * k, sp, d1 + 2 * k, dp, tc < in_blocks, (tc % 5) == 4);
            }
        }
    }
}

template <bool TMA>
__global__ void __launch_bounds__(FR_THREADS)
k_front(const __grid_constant__ Geom g, const __grid_constant__ FrontLayout fl, const uint8_t* __restrict__ store, int store_frames,
        PairList pl, uint8_t* __restrict__ img_all, int32_t* __restrict__ left, int strips) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int strip_sum[FR_SUMS];
    const int band = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int W = g.W, H = g.H;
    const int m = fl.mult;
    const int y0 = band * FR_BAND * m, rows = min(FR_BAND * m, H - y0);
    uint8_t* A = sm + fl.off[0];
    uint8_t* B = sm + fl.off_prev;
    const int p0 = fl.pitch[0];
    const uint8_t* cur = store + (size_t)pl.cur_slot[f] * g.L[0].px + (size_t)y0 * W;
    const uint8_t* prv = store + (size_t)pl.prev_slot[f] * g.L[0].px + (size_t)y0 * W;
    // ---- stage the two bands ----
    if (TMA) {                                   // W % 16 == 0: p0 == W, a band is rows * W contiguous, 16-byte aligned bytes
        if (tid == 0) {
            mbar_init(&bar, 1);
            const uint32_t bytes = (uint32_t)(rows * W);
            mbar_expect_tx(&bar, 2 * bytes);
            bulk_copy_g2s(A, cur, bytes, &bar);
            bulk_copy_g2s(B, prv, bytes, &bar);
        }
        __syncthreads();                         // the initialised barrier is visible to the waiters
        mbar_wait(&bar, 0);
    } else {
        const int words = (W + 3) >> 2;
        for (int i = tid; i < rows * words; i += FR_THREADS) {
            const int r = i / words, k = i - r * words;
            reinterpret_cast<uint32_t*>(A + r * p0)[k] = load4(cur + (size_t)r * W + 4 * k);
            reinterpret_cast<uint32_t*>(B + r * p0)[k] = load4(prv + (size_t)r * W + 4 * k);
        }
        __syncthreads();
    }
    // ---- |a - b| in place ----
    {
        uint4* a4 = reinterpret_cast<uint4*>(A);
        const uint4* b4 = reinterpret_cast<const uint4*>(B);
        const int n = rows * p0 / 16;
        for (int i = tid; i < n; i += FR_THREADS) {
            uint4 u = a4[i];
            const uint4 v = b4[i];
            u.x = __vabsdiffu4(u.x, v.x); u.y = __vabsdiffu4(u.y, v.y); u.z = __vabsdiffu4(u.z, v.z); u.w = __vabsdiffu4(u.w, v.w);
            a4[i] = u;
        }
    }
    __syncthreads();
    // ---- layer 0 out, its row strip sums, layers 1 and 2 into shared memory ----
    copy_out(img_all + g.L[0].base + (size_t)f * g.L[0].px + (size_t)y0 * W, A, p0, W, rows, tid, fl.mg_words[0]);
    {
        // sums of the 128-column strips of every row (a warp per (row, strip): one word per lane, one REDUX); the prefix over
        // the strips of a row - left[y][s], what k_integral_strips starts each row from - follows after the next barrier
        const int lane = tid & 31;
        for (int u = tid >> 5; u < rows * strips; u += FR_THREADS / 32) {
            const int r = u / strips, sidx = u - r * strips;
            const int x = sidx * IG_STRIP_W + 4 * lane;
            uint32_t v = 0;
            if (x < W) {
                v = *reinterpret_cast<const uint32_t*>(A + r * p0 + x);
                if (x + 4 > W) v &= (1u << (8 * (W - x))) - 1u;
            }
            const int sum = __reduce_add_sync(0xffffffffu, (int)__dp4a(v, 0x01010101u, 0u));
            if (lane == 0) strip_sum[u] = sum;
        }
    }
    uint8_t* S1 = sm + fl.off[1]; uint8_t* S2 = sm + fl.off[2]; uint8_t* S3 = sm + fl.off[3];
    uint8_t* S4 = sm + fl.off[4]; uint8_t* S5 = sm + fl.off[5];
    // rows of each layer this band produces (the layers' heights are not multiples of the band's share in general)
    const int n1 = min(8 * m, g.L[1].h - 8 * m * band), n2 = min(6 * m, g.L[2].h - 6 * m * band), n3 = min(4 * m, g.L[3].h - 4 * m * band);
    const int n4 = min(3 * m, g.L[4].h - 3 * m * band), n5 = min(2 * m, g.L[5].h - 2 * m * band);
    twothird_band(A, p0, W, S1, fl.pitch[1], g.L[1].w, n1 > 0 ? n1 / 2 : 0, tid, fl.mg_items[1]);
    half_band(A, p0, W, S2, fl.pitch[2], g.L[2].w, n2, tid, fl.mg_items[2]);
    __syncthreads();
    if (tid < rows) {
        int32_t* out = left + ((size_t)f * H + y0 + tid) * strips;
        int carry = 0;
        for (int sidx = 0; sidx < strips; sidx++) { out[sidx] = carry; carry += strip_sum[tid * strips + sidx]; }
    }
    copy_out(img_all + g.L[1].base + (size_t)f * g.L[1].px + (size_t)(8 * m * band) * g.L[1].w, S1, fl.pitch[1], g.L[1].w, n1, tid, fl.mg_words[1]);
    copy_out(img_all + g.L[2].base + (size_t)f * g.L[2].px + (size_t)(6 * m * band) * g.L[2].w, S2, fl.pitch[2], g.L[2].w, n2, tid, fl.mg_words[2]);
    half_band(S1, fl.pitch[1], g.L[1].w, S3, fl.pitch[3], g.L[3].w, n3, tid, fl.mg_items[3]);
    half_band(S2, fl.pitch[2], g.L[2].w, S4, fl.pitch[4], g.L[4].w, n4, tid, fl.mg_items[4]);
    __syncthreads();
    copy_out(img_all + g.L[3].base + (size_t)f * g.L[3].px + (size_t)(4 * m * band) * g.L[3].w, S3, fl.pitch[3], g.L[3].w, n3, tid, fl.mg_words[3]);
    copy_out(img_all + g.L[4].base + (size_t)f * g.L[4].px + (size_t)(3 * m * band) * g.L[4].w, S4, fl.pitch[4], g.L[4].w, n4, tid, fl.mg_words[4]);
    half_band(S3, fl.pitch[3], g.L[3].w, S5, fl.pitch[5], g.L[5].w, n5, tid, fl.mg_items[5]);
    __syncthreads();
    copy_out(img_all + g.L[5].base + (size_t)f * g.L[5].px + (size_t)(2 * m * band) * g.L[5].w, S5, fl.pitch[5], g.L[5].w, n5, tid, fl.mg_words[5]);
}

}  // namespace

bool front_supported(const Geom& g) { return g.W <= FR_MAX_W; }

cudaError_t launch_front(cudaStream_t st, const Geom& g, const uint8_t* store, int store_frames, PairList pl, uint8_t* img, int32_t* left) {
    FrontLayout fl;
    static const int band_rows[N_LAYERS] = {12, 8, 6, 4, 3, 2};
    const int strips = (g.W + IG_STRIP_W - 1) / IG_STRIP_W;
    // taller bands for narrow frames (a 12-row band of a 160-pixel frame is 1920 pixels: all overhead), bounded by the strip-sum
    // scratch and the frame height
    int mult = FR_BAND_PX / (FR_BAND * g.W);
    mult = mult < 1 ? 1 : (mult > FR_MULT_MAX ? FR_MULT_MAX : mult);
    while (mult > 1 && (FR_BAND * mult * strips > FR_SUMS || FR_BAND * (mult - 1) >= g.H)) mult--;
    fl.mult = mult;
    int at = 0;
    for (int l = 0; l < N_LAYERS; l++) {
        fl.pitch[l] = (g.L[l].w + 15) & ~15;
        fl.off[l] = at;
        at += band_rows[l] * mult * fl.pitch[l];
        if (l == 0) { fl.off_prev = at; at += band_rows[0] * mult * fl.pitch[0]; }
    }
    fl.total = at;
    static const int SRC[N_LAYERS] = {-1, 0, 0, 1, 2, 3};          // brisk.cpp:578-587
    for (int l = 0; l < N_LAYERS; l++) {
        fl.mg_words[l] = div_magic(g.L[l].w >> 2);
        if (l == 0) fl.mg_items[l] = 0;
        else if (l == 1) fl.mg_items[l] = div_magic(((g.L[1].w >> 1) + 3) >> 2);                      // twothird_band: quads of column triples
        else { const int paired_out = ((g.L[SRC[l]].w >> 4) >> 1) * 16; fl.mg_items[l] = div_magic((paired_out >> 2) + (g.L[l].w - paired_out)); }   // half_band: items
    }
    const dim3 grid((g.H + FR_BAND * mult - 1) / (FR_BAND * mult), g.F);
    const bool tma = (g.W & 15) == 0;            // slots, bands and rows then start 16-byte aligned and a band is a multiple of 16 bytes
    cudaError_t e;
    if (tma) {
        e = cudaFuncSetAttribute(k_front<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fl.total);
        if (e != cudaSuccess) return e;
        k_front<true><<<grid, FR_THREADS, fl.total, st>>>(g, fl, store, store_frames, pl, img, left, strips);
    } else {
        e = cudaFuncSetAttribute(k_front<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fl.total);
        if (e != cudaSuccess) return e;
        k_front<false><<<grid, FR_THREADS, fl.total, st>>>(g, fl, store, store_frames, pl, img, left, strips);
    }
    return cudaGetLastError();
}

}  // namespace mfk
