// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restatement of the per-frame loop of MoFREAKUtilities::computeMoFREAKFromFile (reference
// src/MoFREAK/MoFREAKUtilities.cpp:374-498) on gray frames already in memory (video decode and
// BGR->gray, :380-410, are outside the hot path) and of writeMoFREAKFeaturesToFile (:691-719).
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

namespace orc {

static const int GAP_FOR_FRAME_DIFFERENCE = 5;   // MoFREAKUtilities.cpp:378

// sufficientMotion(Mat&, Mat, float, float, float), :329-351.  Does the crop + two resizes + one MIP
// and then returns true unconditionally; only run in "as written" mode, for timing fidelity.
static bool sufficient_motion_as_written(const uint8_t* cur, const uint8_t* prev, int w, float x, float y, float scale) {
    const int tl_x = x - (int)scale / 2;
    const int tl_y = y - (int)scale / 2;
    const int side = (int)ceil(scale);
    uint8_t a[361], b[361];
    resize_linear_u8_19(cur + (size_t)tl_y * w + tl_x, w, side, a);
    resize_linear_u8_19(prev + (size_t)tl_y * w + tl_x, w, side, b);
    volatile unsigned sink = mip_byte(a, b, 9, 9);
    (void)sink;
    return true;
}

void extract_clip(const uint8_t* frames, int w, int h, int n_frames, const ExtractOptions& opt,
                  std::vector<Row>& rows) {
    const size_t P = (size_t)w * h;
    std::vector<uint8_t> diff(P);
    std::vector<Keypoint> kps;
    std::vector<uint8_t> desc;
    // frames 0..4 prime the queue; frame c >= 5 is compared with frame c-5 and labelled c-1
    unsigned frame_num = GAP_FOR_FRAME_DIFFERENCE - 1;
    for (int c = GAP_FOR_FRAME_DIFFERENCE; c < n_frames; c++, frame_num++) {
        const uint8_t* cur = frames + (size_t)c * P;
        const uint8_t* prev = frames + (size_t)(c - GAP_FOR_FRAME_DIFFERENCE) * P;
        for (size_t i = 0; i < P; i++) diff[i] = (uint8_t)abs((int)cur[i] - (int)prev[i]);   // cv::absdiff

        opt.detect(diff.data(), w, h, kps);           // BriskFeatureDetector(30).detect(diff_img)

        // cv::FREAK extractor; extractor.compute(diff_img, keypoints, descriptors)
        if (opt.as_written) {
            FreakPattern* fresh = freak_build_pattern();   // the extractor is constructed inside the loop
            freak_compute(*fresh, diff.data(), w, h, kps, desc, 0);
            freak_free_pattern(fresh);
        } else {
            freak_compute(freak_pattern(), diff.data(), w, h, kps, desc, 0);
        }

        for (size_t k = 0; k < kps.size(); k++) {
            if (opt.as_written && !sufficient_motion_as_written(cur, prev, w, kps[k].x, kps[k].y, kps[k].size)) continue;
            Row r;
            r.frame = (int32_t)frame_num;
            r.scale = kps[k].size;
            r.x = kps[k].x;
            r.y = kps[k].y;
            memcpy(r.appearance, desc.data() + k * 64, 8);
            // float coordinates are truncated by the int parameters of the callee (:460 vs :290)
            mip_extract(cur, prev, w, h, kps[k].size, (int)kps[k].x, (int)kps[k].y, r.motion);
            rows.push_back(r);
        }
    }
}

// operator<<(float) with the default precision of 6 significant digits is printf("%g").
std::string format_mofreak(const Row* rows, size_t n) {
    std::string out;
    out.reserve(n * 96);
    char buf[64];
    for (size_t i = 0; i < n; i++) {
        const Row& r = rows[i];
        snprintf(buf, sizeof buf, "%g ", (double)r.x); out += buf;
        snprintf(buf, sizeof buf, "%g ", (double)r.y); out += buf;
        snprintf(buf, sizeof buf, "%d ", r.frame); out += buf;
        snprintf(buf, sizeof buf, "%g ", (double)r.scale); out += buf;
        out += "0 0 ";                                   // motion_x, motion_y (:476-477)
        for (int k = 0; k < 8; k++) { snprintf(buf, sizeof buf, "%u ", (unsigned)r.appearance[k]); out += buf; }
        for (int k = 0; k < 8; k++) { snprintf(buf, sizeof buf, "%d ", (int)r.motion[k]); out += buf; }
        out += "\n";
    }
    return out;
}

}  // namespace orc
