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

// Everything after a successful isMax2D for candidate c = (x, y) on layer li: the scans, and for survivors the
// refined keypoint.  Writes the candidate's final status byte and, for keypoints, its record.
__device__ inline void after_max2d(const PairView& pv, const CandBuf& cb, const DeviceTables& tb, int li, int x, int y, uint32_t c) {
    const Geom& g = *pv.g;
    const LayerGeom& L = g.L[li];
    const ScoreView own = pv.scores(li, false);
    uint8_t* status = pv.m.status + pv.off(li) + y * L.w + x;
    const int center = own.S[y * L.w + x];
    uint8_t st = ST_PASS2D;

    Scan up, dn;
    up.ismax = true; up.max = 0.f; up.dx = 0.f; up.dy = 0.f;
    dn = up;
    if (li < N_LAYERS - 1) {
        up = scan_neighbour(pv.scores(li + 1, true), false, li, x, y, center);      // marks layer li+1
        if (!up.ismax) { *status = st; cb.flags[c] = st; return; }
        if (li > 0) {
            dn = scan_neighbour(pv.scores(li - 1, false), true, li, x, y, center);
            if (!dn.ismax) { *status = st; cb.flags[c] = st; return; }
        }
    } else {
        dn = scan_neighbour(pv.scores(li - 1, false), true, li, x, y, center);
        if (!dn.ismax) { *status = st; cb.flags[c] = st; return; }
    }
    st |= ST_SURVIVE;
    *status = st;       // from here on the candidate's own 3x3 patch counts as touched

    // own-layer patch
    const int s_0_0 = own.at(x - 1, y - 1), s_1_0 = own.at(x, y - 1), s_2_0 = own.at(x + 1, y - 1);
    const int s_2_1 = own.at(x + 1, y), s_1_1 = own.at(x, y), s_0_1 = own.at(x - 1, y);
    const int s_0_2 = own.at(x - 1, y + 1), s_1_2 = own.at(x, y + 1), s_2_2 = own.at(x + 1, y + 1);
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
            const uint8_t* img0 = pv.m.img + pv.off(0);
            const int a_0_0 = agast58_score(img0, L.w, L.h, x - 1, y - 1), a_1_0 = agast58_score(img0, L.w, L.h, x, y - 1);
            const int a_2_0 = agast58_score(img0, L.w, L.h, x + 1, y - 1), a_2_1 = agast58_score(img0, L.w, L.h, x + 1, y);
            const int a_1_1 = agast58_score(img0, L.w, L.h, x, y), a_0_1 = agast58_score(img0, L.w, L.h, x - 1, y);
            const int a_0_2 = agast58_score(img0, L.w, L.h, x - 1, y + 1), a_1_2 = agast58_score(img0, L.w, L.h, x, y + 1);
            const int a_2_2 = agast58_score(img0, L.w, L.h, x + 1, y + 1);
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

// what the reference's raw score map holds at q = (qx, qy) when candidate index cidx of this layer runs isMax2D
__device__ __forceinline__ int raw_at(const ScoreView& own, const uint8_t* __restrict__ seen, const uint8_t* __restrict__ status,
                                      int qx, int qy, int cidx) {
    const int s = own.at(qx, qy);
    if (s == 0 || s >= BRISK_THRESHOLD) return s;
    if (seen[qy * own.w + qx]) return s;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            const int nx = qx + dx, ny = qy + dy;
            const int nidx = ny * own.w + nx;
            if (nidx >= cidx) continue;
            if (own.at(nx, ny) < BRISK_THRESHOLD) continue;
            if (status[nidx] & ST_SURVIVE) return s;
        }
    return 0;
}

// the smoothed tie-break of isMax2D, brisk.cpp:870-933
__device__ inline bool tie_break(const PairView& pv, int li, int x, int y) {
    const LayerGeom& L = pv.g->L[li];
    const ScoreView own = pv.scores(li, false);
    const uint8_t* seen = pv.m.seen + pv.off(li);
    const uint8_t* status = pv.m.status + pv.off(li);
    const int cidx = y * L.w + x;
    int r[5][5];
#pragma unroll
    for (int dy = -2; dy <= 2; dy++)
#pragma unroll
        for (int dx = -2; dx <= 2; dx++) r[dy + 2][dx + 2] = raw_at(own, seen, status, x + dx, y + dy, cidx);
    const int center = r[2][2];
    auto smooth = [&](int cx, int cy) {
        return 4 * r[cy][cx] + 2 * (r[cy][cx - 1] + r[cy][cx + 1] + r[cy - 1][cx] + r[cy + 1][cx]) + r[cy - 1][cx - 1] +
               r[cy - 1][cx + 1] + r[cy + 1][cx - 1] + r[cy + 1][cx + 1];
    };
    const int mine = smooth(2, 2);
    bool ok = true;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            if (!dx && !dy) continue;
            if (r[dy + 2][dx + 2] == center && smooth(dx + 2, dy + 2) > mine) ok = false;
        }
    return ok;
}

__device__ inline bool earlier_pending_nearby(const PairView& pv, int li, int x, int y) {
    const LayerGeom& L = pv.g->L[li];
    const ScoreView own = pv.scores(li, false);
    const uint8_t* status = pv.m.status + pv.off(li);
    const int cidx = y * L.w + x;
    for (int dy = -3; dy <= 0; dy++)
        for (int dx = -3; dx <= 3; dx++) {
            const int nx = x + dx, ny = y + dy;
            const int nidx = ny * L.w + nx;
            if (nidx >= cidx) continue;
            if (own.at(nx, ny) < BRISK_THRESHOLD) continue;      // status bytes are only defined on candidates
            if (status[nidx] & ST_PENDING) return true;
        }
    return false;
}


}  // namespace mfk
