// Per-candidate device functions of the sparse stage (see kernels_sparse.cu for the phase structure).  Kept in a
// header that also compiles as plain C++ (tests/host_harness.cpp defines the CUDA qualifiers away and runs
// exactly this source on the CPU against the oracle), so the candidate logic is covered without a GPU.
#pragma once
#include "brisk_device.cuh"
#include "kernels.h"

namespace mfk {

struct PairView {
    const Geom* g;
    Maps m;
    int f;
    __device__ __forceinline__ size_t off(int l) const { return g->L[l].base + (size_t)f * g->L[l].px; }
    __device__ __forceinline__ ScoreView scores(int l, bool mark) const {
        ScoreView v;
        v.S = m.S + off(l);
        v.seen = mark ? m.seen + off(l) : nullptr;
        v.w = g->L[l].w; v.h = g->L[l].h;
        return v;
    }
};

// FREAK scale index and border test of cv::FREAK::computeImpl (freak.cpp, call site MoFREAKUtilities.cpp:428)
__device__ __forceinline__ int freak_scale_index(const float* __restrict__ thr, float size) {
    int idx = 0;
#pragma unroll 9
    for (int k = 1; k < FREAK_SCALES; k++) idx += (size >= thr[k]);
    return idx;
}

// What follows a successful isMax2D for candidate c = (x, y) on layer li, in three stages so that the kernels can
// re-pack the surviving candidates onto full warps between them (about half drop out at each scan):
//   stage_above  getScoreMaxAbove incl. its write-back marks       (refine3D, brisk.cpp:945-950)
//   stage_below  getScoreMaxBelow                                   (:988-991, :1053-1056; top layer :654-658)
//   stage_emit   own 3x3 patch, layer-0 AGAST stand-in, scale/position refinement, keypoint record, FREAK border test
// Every stage fetches what it needs up front as register windows (Window4, 3x3 patch rows, 5x5 image window).
struct ScanPair { Scan up, dn; };

__device__ __forceinline__ Scan neutral_scan() { Scan s; s.ismax = true; s.max = 0.f; s.dx = 0.f; s.dy = 0.f; return s; }

__device__ __forceinline__ void finish_failed(const PairView& pv, const CandBuf& cb, int li, int x, int y, uint32_t c) {
    pv.m.status[pv.off(li) + y * pv.g->L[li].w + x] = ST_PASS2D;
    cb.flags[c] = ST_PASS2D;
}

// returns false when a larger score exists above (the candidate is then final)
// (center: the candidate's own score when the caller already holds it, else -1 = read it here)
__device__ inline bool stage_above(const PairView& pv, const CandBuf& cb, int li, int x, int y, uint32_t c, Scan& up, int center = -1) {
    up = neutral_scan();
    if (li == N_LAYERS - 1) return true;
    const Geom& g = *pv.g;
    if (center < 0) center = pv.m.S[pv.off(li) + y * g.L[li].w + x];
    const Footprint fa = scan_footprint(false, li, x, y);
    Window4 wa;
    wa.load(pv.m.S + pv.off(li + 1), g.L[li + 1].w, g.L[li + 1].h, window_origin(fa.x_1, fa.x1), window_origin(fa.y_1, fa.y1));
    up = scan_neighbour(wa, fa, false, li, x, y, center);
    wa.flush_marks(pv.m.seen + pv.off(li + 1));             // getScoreMaxAbove's write-backs, even on early exit
    if (!up.ismax) { finish_failed(pv, cb, li, x, y, c); return false; }
    return true;
}

// returns false when a larger score exists below
__device__ inline bool stage_below(const PairView& pv, const CandBuf& cb, int li, int x, int y, uint32_t c, Scan& dn, int center = -1) {
    dn = neutral_scan();
    if (li == 0) return true;
    const Geom& g = *pv.g;
    if (center < 0) center = pv.m.S[pv.off(li) + y * g.L[li].w + x];
    const Footprint fb = scan_footprint(true, li, x, y);
    Window4 wb;
    wb.load(pv.m.S + pv.off(li - 1), g.L[li - 1].w, g.L[li - 1].h, window_origin(fb.x_1, fb.x1), window_origin(fb.y_1, fb.y1));
    dn = scan_neighbour(wb, fb, true, li, x, y, center);
    if (!dn.ismax) { finish_failed(pv, cb, li, x, y, c); return false; }
    return true;
}

__device__ inline void stage_emit(const PairView& pv, const CandBuf& cb, const DeviceTables& tb, int li, int x, int y, uint32_t c,
                                  const Scan up, const Scan dn) {
    const Geom& g = *pv.g;
    const LayerGeom& L = g.L[li];
    const uint8_t* Sown = pv.m.S + pv.off(li);
    uint8_t* status = pv.m.status + pv.off(li) + y * L.w + x;
    // own-layer 3x3 patch (the dense map already holds 0 inside the 3-px border, as getAgastScore returns there)
    const uint32_t p0 = load4(Sown + (y - 1) * L.w + x - 1), p1 = load4(Sown + y * L.w + x - 1), p2 = load4(Sown + (y + 1) * L.w + x - 1);
    int a58[9];
    if (li == 0) agast58_patch9(pv.m.img + pv.off(0), L.w, x, y, a58);
    uint8_t st = ST_PASS2D | ST_SURVIVE;
    *status = st;       // from here on the candidate's own 3x3 patch counts as touched

    const int s_0_0 = p0 & 0xff, s_1_0 = (p0 >> 8) & 0xff, s_2_0 = (p0 >> 16) & 0xff;
    const int s_0_1 = p1 & 0xff, s_1_1 = (p1 >> 8) & 0xff, s_2_1 = (p1 >> 16) & 0xff;
    const int s_0_2 = p2 & 0xff, s_1_2 = (p2 >> 8) & 0xff, s_2_2 = (p2 >> 16) & 0xff;
    const int center = s_1_1;

    float delta_x_layer, delta_y_layer;
    const float max_layer = subpixel2D(s_0_0, s_0_1, s_0_2, s_1_0, s_1_1, s_1_2, s_2_0, s_2_1, s_2_2, delta_x_layer, delta_y_layer);

    float kx, ky, ksize, kresp;
    bool is_kp = true;
    if (li == N_LAYERS - 1) {
        // top layer (brisk.cpp:645-680): 2-D refinement only, no score test
        kx = (float(x) + delta_x_layer) * L.scale + L.offset;
        ky = (float(y) + delta_y_layer) * L.scale + L.offset;
        ksize = 12.0f * L.scale;
        kresp = max_layer;
    } else {
        float max_below = dn.max, delta_x_below = dn.dx, delta_y_below = dn.dy;
        if (li == 0) {
            // no layer below layer 0: 5/8 scores on the same image stand in (brisk.cpp:959-986)
            const int a_0_0 = a58[0], a_1_0 = a58[1], a_2_0 = a58[2], a_0_1 = a58[3], a_1_1 = a58[4], a_2_1 = a58[5];
            const int a_0_2 = a58[6], a_1_2 = a58[7], a_2_2 = a58[8];
            const int m = max(max(max(a_0_0, a_1_0), max(a_2_0, a_2_1)), max(max(a_1_1, a_0_1), max(max(a_0_2, a_1_2), a_2_2)));
            subpixel2D(a_0_0, a_0_1, a_0_2, a_1_0, a_1_1, a_1_2, a_2_0, a_2_1, a_2_2, delta_x_below, delta_y_below);
            max_below = (float)m;
        }
        const float mid = fmaxf(float(center), max_layer);
        float max, scale;
        if (li % 2 == 0) {
            scale = refine1D_fit(li == 0 ? 2 : 0, max_below, mid, up.max, max);
            if (scale > 1.0f) {
                const float r0 = (float)((1.5 - (double)scale) / .5);
                const float r1 = (float)(1.0 - (double)r0);
                kx = (r0 * delta_x_layer + r1 * up.dx + float(x)) * L.scale + L.offset;
                ky = (r0 * delta_y_layer + r1 * up.dy + float(y)) * L.scale + L.offset;
            } else if (li == 0) {
                const float r0 = (float)(((double)scale - 0.5) / 0.5);
                const float r_1 = (float)(1.0 - (double)r0);
                kx = r0 * delta_x_layer + r_1 * delta_x_below + float(x);
                ky = r0 * delta_y_layer + r_1 * delta_y_below + float(y);
            } else {
                const float r0 = (float)(((double)scale - 0.75) / 0.25);
                const float r_1 = (float)(1.0 - (double)r0);
                kx = (r0 * delta_x_layer + r_1 * delta_x_below + float(x)) * L.scale + L.offset;
                ky = (r0 * delta_y_layer + r_1 * delta_y_below + float(y)) * L.scale + L.offset;
            }
        } else {
            scale = refine1D_fit(1, max_below, mid, up.max, max);
            if (scale > 1.0f) {
                const float r0 = (float)(4.0 - (double)scale * 3.0);
                const float r1 = (float)(1.0 - (double)r0);
                kx = (r0 * delta_x_layer + r1 * up.dx + float(x)) * L.scale + L.offset;
                ky = (r0 * delta_y_layer + r1 * up.dy + float(y)) * L.scale + L.offset;
            } else {
                const float r0 = (float)((double)scale * 3.0 - 2.0);
                const float r_1 = (float)(1.0 - (double)r0);
                kx = (r0 * delta_x_layer + r_1 * delta_x_below + float(x)) * L.scale + L.offset;
                ky = (r0 * delta_y_layer + r_1 * delta_y_below + float(y)) * L.scale + L.offset;
            }
        }
        scale *= L.scale;
        ksize = 12.0f * scale;
        kresp = max;
        is_kp = max > float(BRISK_THRESHOLD);       // brisk.cpp:698
    }
    if (is_kp) {
        cb.kx[c] = kx; cb.ky[c] = ky; cb.ksize[c] = ksize; cb.kresp[c] = kresp;
        // cv::FREAK drops keypoints whose pattern does not fit into the image
        const int sidx = freak_scale_index(tb.scale_threshold, ksize);
        const float ps = (float)tb.pattern_size[sidx];
        const bool drop = kx <= ps || ky <= ps || kx >= (float)(g.W - tb.pattern_size[sidx]) ||
                          ky >= (float)(g.H - tb.pattern_size[sidx]);
        cb.scale_idx[c] = (uint8_t)sidx;
        st |= ST_KEYPOINT;
        if (!drop) st |= ST_KEEP;
    }
    cb.flags[c] = st;
}

// the three stages back to back (phase B and the CPU harness; phase A re-packs candidates between the stages).  The two
// scans run through ONE copy of the scan code with a run-time side (above first, then below, as refine3D does): phase B's
// warps each follow their own pair, so the kernel is bound by instruction fetch, and two inlined copies of the scan are a
// third of its code.
__device__ inline void after_max2d(const PairView& pv, const CandBuf& cb, const DeviceTables& tb, int li, int x, int y, uint32_t c,
                                   int center = -1) {
    const Geom& g = *pv.g;
    if (center < 0) center = pv.m.S[pv.off(li) + y * g.L[li].w + x];
    Scan up = neutral_scan(), dn = neutral_scan();
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        const bool below = side != 0;
        const int ln = below ? li - 1 : li + 1;                   // the neighbouring layer
        if (ln < 0 || ln >= N_LAYERS) continue;
        const Footprint fp = scan_footprint(below, li, x, y);
        Window4 wn;
        wn.load(pv.m.S + pv.off(ln), g.L[ln].w, g.L[ln].h, window_origin(fp.x_1, fp.x1), window_origin(fp.y_1, fp.y1));
        const Scan sc = scan_neighbour(wn, fp, below, li, x, y, center);
        if (!below) wn.flush_marks(pv.m.seen + pv.off(ln));       // getScoreMaxAbove's write-backs, even on early exit
        if (!sc.ismax) { finish_failed(pv, cb, li, x, y, c); return; }
        if (below) dn = sc; else up = sc;
    }
    stage_emit(pv, cb, tb, li, x, y, c, up, dn);
}

// ---- tie candidates (phase B) ----------------------------------------------------------------------------------
// Everything a tie candidate at (x, y) needs lies within 3 px: the scores and candidate statuses of the 7x7
// neighbourhood and the "seen" marks of the 5x5.  They are fetched as packed rows with independent aligned loads
// (x in [3, w-4], y in [3, h-4], so the 7x7 is inside the layer) and all decisions are taken from registers.
struct TieWindow {
    uint32_t S[7][2], st[7][2], seen[5][2];

    __device__ __forceinline__ void load(const PairView& pv, int li, int x, int y) {
        const int w = pv.g->L[li].w;
        const size_t off = pv.off(li);
#pragma unroll
        for (int r = 0; r < 7; r++) {
            load8(pv.m.S + off + (y - 3 + r) * w + x - 3, S[r][0], S[r][1]);
            load8(pv.m.status + off + (y - 3 + r) * w + x - 3, st[r][0], st[r][1]);
        }
#pragma unroll
        for (int r = 0; r < 5; r++) load8(pv.m.seen + off + (y - 2 + r) * w + x - 2, seen[r][0], seen[r][1]);
    }
};

__device__ __forceinline__ int packed_byte(const uint32_t r[2], int b) { return (int)((r[b >> 2] >> (8 * (b & 3))) & 0xffu); }

// true while a raster-earlier tie candidate within 3 px is still undecided (status bytes are only defined on
// candidates, i.e. where S >= threshold)
__device__ inline bool earlier_pending_nearby(const TieWindow& W) {
    bool pending = false;
#pragma unroll
    for (int r = 0; r <= 3; r++)
#pragma unroll
        for (int c = 0; c < 7; c++) {
            if (r == 3 && c >= 3) continue;
            pending |= packed_byte(W.S[r], c) >= BRISK_THRESHOLD && (packed_byte(W.st[r], c) & ST_PENDING);
        }
    return pending;
}

// what the reference's raw score map holds at window cell (r, c) (7x7 coordinates, candidate at (3, 3)) when the
// candidate runs isMax2D: the score if it was written by getAgastPoints (>= threshold), or read by a
// getScoreMaxAbove scan from the layer below, or lies in the 3x3 patch of a raster-earlier survivor; else 0
__device__ __forceinline__ int raw_cell(const TieWindow& W, const int r, const int c) {
    const int s = packed_byte(W.S[r], c);
    if (s == 0 || s >= BRISK_THRESHOLD) return s;
    if (packed_byte(W.seen[r - 1], c - 1)) return s;
    bool touched = false;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++) {
            const int a = r + dr, b = c + dc;
            if (a > 3 || (a == 3 && b >= 3)) continue;                       // not raster-earlier than the candidate
            touched |= packed_byte(W.S[a], b) >= BRISK_THRESHOLD && (packed_byte(W.st[a], b) & ST_SURVIVE);
        }
    return touched ? s : 0;
}

// the smoothed tie-break of isMax2D, brisk.cpp:870-933
__device__ inline bool tie_break(const TieWindow& W) {
    int r[5][5];
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) r[i][j] = raw_cell(W, i + 1, j + 1);
    const int center = r[2][2];
    bool ok = true;
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        // k = 0 is the candidate itself, then its eight neighbours
        const int cy = (k == 0) ? 2 : 1 + ((k - 1 + (k > 4)) / 3), cx = (k == 0) ? 2 : 1 + ((k - 1 + (k > 4)) % 3);
        const int sm = 4 * r[cy][cx] + 2 * (r[cy][cx - 1] + r[cy][cx + 1] + r[cy - 1][cx] + r[cy + 1][cx]) + r[cy - 1][cx - 1] +
                       r[cy - 1][cx + 1] + r[cy + 1][cx - 1] + r[cy + 1][cx + 1];
        if (k == 0) mine = sm;
        else if (r[cy][cx] == center && sm > mine) ok = false;
    }
    return ok;
}

}  // namespace mfk
