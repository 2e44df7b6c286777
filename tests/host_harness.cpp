// CPU harness for the product's DEVICE source.  mofreak_b200/csrc/{dense,brisk,sparse}_device.cuh are written
// so that they also compile as plain C++ once the CUDA qualifiers are defined away; this file does that and
// drives them with the same phase structure as the kernels (dense scores -> ordered candidates -> phase A ->
// phase B fixed point -> keypoint records), so the per-candidate logic that runs on the GPU is checked
// against the oracle in the CPU-only test suite.  Test infrastructure; not part of the product.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline
using std::max;
using std::min;

// scalar stand-ins for the packed 16x2 three-input min/max (VIMNMX3.U16x2 on sm_100a)
static inline uint32_t __vimin3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t lo = std::min(std::min(a & 0xffffu, b & 0xffffu), c & 0xffffu);
    const uint32_t hi = std::min(std::min(a >> 16, b >> 16), c >> 16);
    return lo | (hi << 16);
}
static inline uint32_t __vimax3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t lo = std::max(std::max(a & 0xffffu, b & 0xffffu), c & 0xffffu);
    const uint32_t hi = std::max(std::max(a >> 16, b >> 16), c >> 16);
    return lo | (hi << 16);
}

static inline uint32_t __vmaxu2(uint32_t a, uint32_t b) { return __vimax3_u16x2(a, b, b); }
static inline uint32_t __vminu2(uint32_t a, uint32_t b) { return __vimin3_u16x2(a, b, b); }

#include "../mofreak_b200/csrc/dense_device.cuh"
#include "../mofreak_b200/csrc/sparse_device.cuh"
#include "../mofreak_b200/csrc/desc_device.cuh"
#include "../mofreak_b200/csrc/text_device.cuh"

using namespace mfk;

namespace {

struct Harness {
    Geom g;
    std::vector<uint8_t> img, S, seen, status;
    std::vector<uint32_t> row_count, xy, meta, keep_scan, row_src, counters;
    std::vector<float> kx, ky, ksize, kresp;
    std::vector<uint8_t> flags, scale_idx;
    Maps maps;
    CandBuf cb;
    DeviceTables tb;

    void setup(int W, int H) {
        memset(&g, 0, sizeof g);
        int lw[6], lh[6];
        lw[0] = W; lh[0] = H; lw[1] = 2 * (W / 3); lh[1] = 2 * (H / 3);
        for (int l = 2; l < 6; l++) { lw[l] = lw[l - 2] / 2; lh[l] = lh[l - 2] / 2; }
        static const float scale[6] = {1.0f, 1.5f, 2.0f, 3.0f, 4.0f, 6.0f};
        size_t below = 0; int row0 = 0;
        for (int l = 0; l < 6; l++) {
            g.L[l].w = lw[l]; g.L[l].h = lh[l]; g.L[l].px = lw[l] * lh[l]; g.L[l].row0 = row0; g.L[l].base = below;
            g.L[l].scale = scale[l]; g.L[l].offset = (float)(0.5 * scale[l] - 0.5);
            below += g.L[l].px; row0 += lh[l];
        }
        g.W = W; g.H = H; g.F = 1; g.rows = row0; g.pyr_px = (int)below; g.cand_cap = (int)below; g.row_cap = (int)below;   // before slack
        const size_t alloc = below + 4 * (size_t)W + 64;      // same slack as the device allocations (windows read past the end)
        img.assign(alloc, 0); S.assign(alloc, 0); seen.assign(alloc, 0); status.assign(alloc, 0xAA);  // status is NOT zero-initialised on the device either
        maps.img = img.data(); maps.S = S.data(); maps.seen = seen.data(); maps.status = status.data();
        const Tables& T = tables();
        tb.points = T.points; tb.pattern_size = T.pattern_size; tb.scale_threshold = T.scale_threshold;
        tb.orient = T.orient; tb.used = T.used; tb.mip_needed = T.mip_needed; tb.n_mip_needed = T.n_mip_needed;
    }

    // prune_tile_losers: what k_scores does with MFK_EARLY_LOSERS (the default): a hit with a strictly larger score among its 8
    // neighbours that lies inside the 1-pixel frame of its 64 x 32 score tile gets status 0 and never becomes a candidate
    bool prune_tile_losers = false;

    void dense(const uint8_t* diff) {
        memcpy(img.data(), diff, g.L[0].px);
        static const int SRC[6] = {-1, 0, 0, 1, 2, 3};
        for (int l = 1; l < 6; l++) {
            const LayerGeom& D = g.L[l]; const LayerGeom& Sg = g.L[SRC[l]];
            for (int y = 0; y < D.h; y++)
                for (int x = 0; x < D.w; x++)
                    img[D.base + y * D.w + x] = (uint8_t)(l == 1 ? twothird_px(img.data() + Sg.base, Sg.w, x, y)
                                                                 : half_px(img.data() + Sg.base, Sg.w, Sg.w, x, y));
        }
        row_count.assign(g.rows + 1, 0);
        xy.clear(); meta.clear();
        for (int l = 0; l < 6; l++) {
            const LayerGeom& L = g.L[l];
            const uint8_t* im = img.data() + L.base;
            for (int y = 0; y < L.h; y++)
                for (int x = 0; x < L.w; x++) {
                    int s = 0;
                    if (x >= 3 && y >= 3 && x < L.w - 3 && y < L.h - 3) {
                        static const int RX[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};
                        static const int RY[16] = {0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1};
                        uint32_t r[16];
                        for (int i = 0; i < 16; i++) { const uint32_t v = im[(y + RY[i]) * L.w + x + RX[i]]; r[i] = v | ((255u - v) << 16); }
                        const uint32_t v = im[y * L.w + x];
                        s = oast_score_packed(r, v | ((255u - v) << 16));
                        const int bound = oast_score_bound_packed(r, v | ((255u - v) << 16));
                        if (bound <= 0 && s != 0) abort();                                                   // the bound must be sound
                        if (x + 1 < L.w - 3) {
                            // the two-centres-per-word form k_scores runs: each half must equal the single-centre bound (+ 256)
                            uint32_t rp[16];
                            for (int i = 0; i < 16; i++)
                                rp[i] = (uint32_t)im[(y + RY[i]) * L.w + x + RX[i]] | ((uint32_t)im[(y + RY[i]) * L.w + x + 1 + RX[i]] << 16);
                            const uint32_t both = oast_bound_pair(rp, v | ((uint32_t)im[y * L.w + x + 1] << 16));
                            uint32_t r2[16];
                            for (int i = 0; i < 16; i++) { const uint32_t u = im[(y + RY[i]) * L.w + x + 1 + RX[i]]; r2[i] = u | ((255u - u) << 16); }
                            const uint32_t v2 = im[y * L.w + x + 1];
                            const int bound2 = oast_score_bound_packed(r2, v2 | ((255u - v2) << 16));
                            if ((int)(both & 0xffffu) != bound + 256 || (int)(both >> 16) != bound2 + 256) abort();
                        }
                    }
                    S[L.base + y * L.w + x] = (uint8_t)s;
                }
            // the candidate list of the layer, in raster order (scores of the whole layer are known now, as in the kernel's tile)
            const uint8_t* Sl = S.data() + L.base;
            for (int y = 0; y < L.h; y++)
                for (int x = 0; x < L.w; x++) {
                    const int s = Sl[y * L.w + x];
                    if (s < BRISK_THRESHOLD) continue;
                    if (prune_tile_losers) {
                        const int lx = x % 64, ly = y % 32;                 // SC_TW x SC_TH tiles start at multiples of 64 / 32
                        if (lx >= 1 && lx <= 62 && ly >= 1 && ly <= 30) {
                            int mx = 0;
                            for (int dy = -1; dy <= 1; dy++)
                                for (int dx = -1; dx <= 1; dx++)
                                    if (dx || dy) mx = max(mx, (int)Sl[(y + dy) * L.w + x + dx]);
                            if (mx > s) { status[L.base + y * L.w + x] = 0; continue; }
                        }
                    }
                    xy.push_back((uint32_t)x | ((uint32_t)y << 16)); meta.push_back((uint32_t)l); row_count[L.row0 + y]++;
                }
        }
        uint32_t run = 0;
        for (int r = 0; r <= g.rows; r++) { const uint32_t c = row_count[r]; row_count[r] = run; run += c; }
        const size_t n = xy.size();
        kx.assign(n, 0); ky.assign(n, 0); ksize.assign(n, 0); kresp.assign(n, 0); flags.assign(n, 0); scale_idx.assign(n, 0);
        keep_scan.assign(n + 1, 0); row_src.assign(n, 0); counters.assign(8, 0);
        cb.row_count = row_count.data(); cb.xy = xy.data(); cb.meta = meta.data(); cb.kx = kx.data(); cb.ky = ky.data();
        cb.ksize = ksize.data(); cb.kresp = kresp.data(); cb.flags = flags.data(); cb.scale_idx = scale_idx.data();
        cb.keep_scan = keep_scan.data(); cb.row_src = row_src.data(); cb.counters = counters.data();
    }

    void sparse(bool reverse_rounds) {
        PairView pv; pv.g = &g; pv.m = maps; pv.f = 0;
        const uint32_t n = (uint32_t)xy.size();
        // phase A (k_phase_a body)
        for (uint32_t c = 0; c < n; c++) {
            const int x = xy[c] & 0xffff, y = xy[c] >> 16, li = meta[c] & 7;
            const LayerGeom& L = g.L[li];
            const uint8_t* Sp = maps.S + pv.off(li) + y * L.w + x;
            const int c0 = Sp[0];
            const int n0 = Sp[-L.w - 1], n1 = Sp[-L.w], n2 = Sp[-L.w + 1], n3 = Sp[-1], n4 = Sp[1], n5 = Sp[L.w - 1], n6 = Sp[L.w], n7 = Sp[L.w + 1];
            const int mx = max(max(max(n0, n1), max(n2, n3)), max(max(n4, n5), max(n6, n7)));
            uint8_t* st = maps.status + pv.off(li) + y * L.w + x;
            if (mx > c0) { *st = 0; flags[c] = 0; continue; }
            if (mx == c0) { *st = ST_PENDING; flags[c] = ST_PENDING; continue; }
            after_max2d(pv, cb, tb, li, x, y, c);
        }
        // phase B (k_phase_b's fixed point, one layer at a time, verdicts of a round applied after the round)
        for (int li = 0; li < 6; li++) {
            const LayerGeom& L = g.L[li];
            std::vector<uint32_t> list;
            for (uint32_t c = row_count[L.row0]; c < row_count[L.row0 + L.h]; c++) if (flags[c] & ST_PENDING) list.push_back(c);
            if (reverse_rounds) std::reverse(list.begin(), list.end());
            size_t left = list.size();
            while (left) {
                std::vector<uint8_t> verdict(list.size(), 0);
                for (size_t i = 0; i < list.size(); i++) {
                    if (list[i] == 0xffffffffu) continue;
                    const int x = xy[list[i]] & 0xffff, y = xy[list[i]] >> 16;
                    TieWindow tw;
                    tw.load(pv, li, x, y);
                    if (earlier_pending_nearby(tw)) continue;
                    verdict[i] = tie_break(tw) ? 2 : 1;
                }
                for (size_t i = 0; i < list.size(); i++) {
                    if (!verdict[i]) continue;
                    const uint32_t c = list[i];
                    const int x = xy[c] & 0xffff, y = xy[c] >> 16;
                    if (verdict[i] == 2) after_max2d(pv, cb, tb, li, x, y, c);
                    else { maps.status[pv.off(li) + y * L.w + x] = 0; flags[c] = 0; }
                    list[i] = 0xffffffffu; left--;
                }
            }
        }
    }
};

}  // namespace

extern "C" {

// out: {x, y, size, response, layer, keep} per BRISK keypoint in reference order; returns the count
// mode: bit 0 = phase B rounds in reverse order, bit 1 = drop 8-neighbour losers inside their score tile before the candidate list
int hh_detect(const uint8_t* diff, int w, int h, int mode, float* out, int cap) {
    Harness H;
    H.setup(w, h);
    H.prune_tile_losers = (mode & 2) != 0;
    H.dense(diff);
    H.sparse((mode & 1) != 0);
    int n = 0;
    for (size_t c = 0; c < H.xy.size(); c++) {
        if (!(H.flags[c] & ST_KEYPOINT)) continue;
        if (n < cap) {
            out[6 * n] = H.kx[c]; out[6 * n + 1] = H.ky[c]; out[6 * n + 2] = H.ksize[c]; out[6 * n + 3] = H.kresp[c];
            out[6 * n + 4] = (float)(H.meta[c] & 7); out[6 * n + 5] = (H.flags[c] & ST_KEEP) ? 1.f : 0.f;
        }
        n++;
    }
    return n;
}

// pyramid layer `layer` and its score map as the device functions compute them
void hh_layer(const uint8_t* diff, int w, int h, int layer, uint8_t* img_out, uint8_t* score_out, int* dims) {
    Harness H;
    H.setup(w, h);
    H.dense(diff);
    const LayerGeom& L = H.g.L[layer];
    memcpy(img_out, H.img.data() + L.base, L.px);
    memcpy(score_out, H.S.data() + L.base, L.px);
    dims[0] = L.w; dims[1] = L.h;
}

// The whole per-pair path with the device functions: rows (32 bytes each, mfk_row layout) in reference order.
int hh_extract_pair(const uint8_t* cur, const uint8_t* prev, int w, int h, int label, uint8_t* rows_out, int cap) {
    std::vector<uint8_t> diff((size_t)w * h);
    for (size_t i = 0; i < diff.size(); i++) diff[i] = (uint8_t)abs((int)cur[i] - (int)prev[i]);
    Harness H;
    H.setup(w, h);
    H.dense(diff.data());
    H.sparse(false);
    const int W1 = w + 1;
    std::vector<int32_t> I((size_t)W1 * (h + 1), 0);
    for (int y = 0; y < h; y++) {
        int run = 0;
        for (int x = 0; x < w; x++) { run += diff[(size_t)y * w + x]; I[(size_t)(y + 1) * W1 + x + 1] = I[(size_t)y * W1 + x + 1] + run; }
    }
    int n = 0;
    for (size_t c = 0; c < H.xy.size(); c++) {
        if (!(H.flags[c] & ST_KEEP)) continue;
        const float kx = H.kx[c], ky = H.ky[c], ksize = H.ksize[c];
        const int sidx = H.scale_idx[c];
        const FreakPoint* P0 = H.tb.points + (size_t)sidx * FREAK_ORIENTATIONS * FREAK_POINTS;
        uint8_t field[FREAK_POINTS];
        for (int p = 0; p < FREAK_POINTS; p++) field[p] = (uint8_t)mean_intensity(I.data(), W1, P0[p], kx, ky);
        int d0 = 0, d1 = 0;
        for (int q = 0; q < FREAK_ORIENT_PAIRS; q++) {
            const OrientPair op = H.tb.orient[q];
            const int delta = (int)field[op.i] - (int)field[op.j];
            d0 += delta * op.wdx / 2048;
            d1 += delta * op.wdy / 2048;
        }
        const int theta = freak_theta_index(d0, d1);
        const FreakPoint* PT = P0 + (size_t)theta * FREAK_POINTS;
        for (int p = 0; p < FREAK_POINTS; p++) field[p] = (uint8_t)mean_intensity(I.data(), W1, PT[p], kx, ky);
        uint8_t app[8] = {0}, mot[8] = {0};
        for (int i = 0; i < 64; i++) if (field[H.tb.used[i].i] >= field[H.tb.used[i].j]) app[i >> 3] |= (uint8_t)(1u << (i & 7));
        const int xi = (int)kx, yi = (int)ky, half = (int)ksize / 2, side = (int)ceilf(ksize);
        uint8_t a19[PATCH * PATCH], b19[PATCH * PATCH];
        RowTap rtap[PATCH]; ColTap ctap[PATCH];
        for (int d = 0; d < PATCH; d++) {
            const ResizeTap t = resize_tap(side, d);
            rtap[d] = row_tap(t, side, w); ctap[d] = col_tap(t, side, 0);
        }
        const uint8_t* c0 = cur + (size_t)(yi - half) * w + (xi - half);
        const uint8_t* p0 = prev + (size_t)(yi - half) * w + (xi - half);
        memset(a19, 0xEE, sizeof a19); memset(b19, 0xEE, sizeof b19);      // pixels the SSD tests never read stay unset
        for (int i = 0; i < H.tb.n_mip_needed; i++) {
            const int e = H.tb.mip_needed[i], fr = e >> 15, dx = e & 31, dy = (e >> 5) & 31;
            const int v = resize_px_flat(fr ? p0 : c0, rtap[dy], ctap[dx]);
            if (v != resize_px(fr ? p0 : c0, w, side, resize_tap(side, dy), resize_tap(side, dx))) abort();   // flat == reference form
            (fr ? b19 : a19)[dy * PATCH + dx] = (uint8_t)v;
        }
        for (int i = 0; i < 64; i++) if (mip_ssd(a19, b19, i >> 3, i & 7) > 288) mot[i >> 3] |= (uint8_t)(1u << (i & 7));
        if (n < cap) {
            uint8_t* r = rows_out + (size_t)32 * n;
            const int32_t frame = label;
            memcpy(r, &kx, 4); memcpy(r + 4, &ky, 4); memcpy(r + 8, &frame, 4); memcpy(r + 12, &ksize, 4);
            memcpy(r + 16, app, 8); memcpy(r + 24, mot, 8);
        }
        n++;
    }
    return n;
}

float hh_atan2f(float y, float x) { return atan2f_fdlibm(y, x); }
int hh_theta_index(int d0, int d1) { return freak_theta_index(d0, d1); }
int hh_scale_index_table(float size) { return scale_index_from_table(size); }
int hh_scale_index_direct(float size) { return scale_index_direct(size); }
// counts mismatches of the device atan2f against the host libm over a grid and random integer pairs
long hh_atan2f_sweep(int grid, long random_pairs, unsigned seed) {
    long bad = 0;
    for (int a = -grid; a <= grid; a++)
        for (int b = -grid; b <= grid; b++) {
            const float r = atan2f((float)a, (float)b), q = atan2f_fdlibm((float)a, (float)b);
            if (memcmp(&r, &q, 4) != 0) bad++;
        }
    unsigned s = seed ? seed : 1u;
    for (long i = 0; i < random_pairs; i++) {
        s = s * 1664525u + 1013904223u; const int a = (int)((s >> 8) % 400001u) - 200000;
        s = s * 1664525u + 1013904223u; const int b = (int)((s >> 8) % 400001u) - 200000;
        const float r = atan2f((float)a, (float)b), q = atan2f_fdlibm((float)a, (float)b);
        if (memcmp(&r, &q, 4) != 0) bad++;
    }
    return bad;
}

}  // extern "C"

// `.mofreak` text of n rows (text_device.cuh, the code k_text_lengths / k_text_write run): returns the length, or -(row + 1)
// for the first row that needs exponent notation.  `count_only` exercises the counting sink of the first pass.
extern "C" long hh_format_rows(const void* rows, long n, char* out, int count_only) {
    const TextRow* r = (const TextRow*)rows;
    long total = 0;
    for (long i = 0; i < n; i++) {
        CountSink cs;
        if (!put_row(cs, r[i])) return -(i + 1);
        if (cs.n > TEXT_ROW_MAX) return -(i + 1);
        if (!count_only) {
            BufSink bs(out + total);
            put_row(bs, r[i]);
            if (bs.n != cs.n) return -(i + 1);
        }
        total += cs.n;
    }
    return total;
}

// "%g" of one float through the device formatter; returns the length or -1
extern "C" int hh_format_float(float v, char* out) {
    BufSink bs(out);
    if (!put_float_g(bs, v)) return -1;
    return bs.n;
}
