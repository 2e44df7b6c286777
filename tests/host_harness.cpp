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

#include "../mofreak_b200/csrc/dense_device.cuh"
#include "../mofreak_b200/csrc/sparse_device.cuh"

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
        g.W = W; g.H = H; g.F = 1; g.rows = row0; g.pyr_px = (int)below; g.cand_cap = (int)below; g.row_cap = (int)below;
        img.assign(below, 0); S.assign(below, 0); seen.assign(below, 0); status.assign(below, 0xAA);  // status is NOT zero-initialised on the device either
        maps.img = img.data(); maps.S = S.data(); maps.seen = seen.data(); maps.status = status.data();
        const Tables& T = tables();
        tb.points = T.points; tb.pattern_size = T.pattern_size; tb.scale_threshold = T.scale_threshold;
        tb.orient = T.orient; tb.used = T.used;
    }

    void dense(const uint8_t* diff) {
        memcpy(img.data(), diff, g.L[0].px);
        static const int SRC[6] = {-1, 0, 0, 1, 2, 3};
        for (int l = 1; l < 6; l++) {
            const LayerGeom& D = g.L[l]; const LayerGeom& Sg = g.L[SRC[l]];
            for (int y = 0; y < D.h; y++)
                for (int x = 0; x < D.w; x++)
                    img[D.base + y * D.w + x] = (uint8_t)(l == 1 ? twothird_px(img.data() + Sg.base, Sg.w, x, y)
                                                                 : half_px(img.data() + Sg.base, Sg.w, x, y));
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
                    }
                    S[L.base + y * L.w + x] = (uint8_t)s;
                    if (s >= BRISK_THRESHOLD) { xy.push_back((uint32_t)x | ((uint32_t)y << 16)); meta.push_back((uint32_t)l); row_count[L.row0 + y]++; }
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
                    if (earlier_pending_nearby(pv, li, x, y)) continue;
                    verdict[i] = tie_break(pv, li, x, y) ? 2 : 1;
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
int hh_detect(const uint8_t* diff, int w, int h, int reverse_rounds, float* out, int cap) {
    Harness H;
    H.setup(w, h);
    H.dense(diff);
    H.sparse(reverse_rounds != 0);
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

}  // extern "C"
