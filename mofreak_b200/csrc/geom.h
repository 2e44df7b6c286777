// Shared host/device description of one extraction batch: F frame pairs, each with the 6-layer BRISK
// pyramid of reference brisk.cpp:572-588 (layer sizes :1660, :1666; scale/offset :1662-1669).
//
// HBM layout.  Every per-pixel map (pyramid image, score map S, "seen" marks, candidate status) is one
// allocation of F * pyr_px bytes, layer-major then pair-major:
//     address(layer l, pair f, x, y) = base + L[l].base + f * L[l].px + y * L[l].w + x
// with L[l].base = F * sum_{k<l} L[k].px, so that one layer of the whole batch is a dense
// [F][h][w] array (coalesced row reads, one launch covers all pairs).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace mfk {

enum { N_LAYERS = 6, BRISK_THRESHOLD = 30 };

struct LayerGeom {
    int w, h;
    int px;            // w * h
    int row0;          // index of this layer's first row in the per-pair row table (sum of lower h)
    size_t base;       // F * (pixels of all lower layers)
    float scale;       // 1, 1.5, 2, 3, 4, 6
    float offset;      // 0.5 * scale - 0.5
};

struct Geom {
    LayerGeom L[N_LAYERS];
    int W, H;          // frame size
    int F;             // pairs in this batch (<= max_pairs)
    int rows;          // rows per pair over all layers
    int pyr_px;        // pixels per pair over all layers
    int cand_cap;      // candidate budget of the batch
    int row_cap;       // output-row budget of the whole accumulation buffer
};

// candidate status bits (same meaning as in oracle/brisk_port.cpp)
enum {
    ST_PASS2D = 1,      // isMax2D succeeded (its getScoreMaxAbove scan has marked the layer above)
    ST_SURVIVE = 2,     // also passed the scale-space checks: its own 3x3 patch counts as touched
    ST_PENDING = 4,     // has an equal-score neighbour, waits for phase B
    ST_KEEP = 8,        // BRISK keypoint that also fits FREAK's border: becomes an output row
    ST_KEYPOINT = 16,   // BRISK keypoint (before FREAK's border filter)
};

constexpr int TEXT_ROW_MAX = 128;     // upper bound of one `.mofreak` text line (text_device.cuh)

}  // namespace mfk
