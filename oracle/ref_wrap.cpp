// TEST INFRASTRUCTURE ONLY — never linked into, imported by or shipped with the product.
//
// C entry points around the reference's OWN detector object code.  The reference translation unit
// src/MoFREAK/brisk.cpp is compiled unchanged, from where it lies under $MOFREAK_REF (default
// /root/reference), by including it here (its `static inline` pyramid samplers are only reachable
// from inside that translation unit); oast9_16*.cc, agast5_8*.cc and AstDetector.cc are compiled
// as separate objects by oracle/Makefile.  OpenCV is replaced by oracle/ref_shim/.  Output goes to
// oracle/_ref/ only.  No reference source text is copied into this repository.
//
// Exposed (all used by tests/ to pin oracle/brisk_port.cpp and as the detector inside the
// reference-anchored extraction oracle):
//   ref_brisk_detect       BriskFeatureDetector(thresh, octaves).detect       brisk.cpp:544-559
//   ref_halfsample         BriskLayer::halfsample                             brisk.cpp:1840-1972
//   ref_twothirdsample     BriskLayer::twothirdsample                         brisk.cpp:1974-2065
//   ref_oast_detect/score  OastDetector9_16::detect / cornerScore             oast9_16.cc:46, oast9_16_nms.cc:42
//   ref_agast58_score      AgastDetector5_8::cornerScore                      agast5_8_nms.cc:42
//   ref_brisk_dump         pyramid images + final raw score maps after getKeypoints (brisk.cpp:572-704)

#ifndef MFK_REF_BRISK_CPP
#error "compile with -DMFK_REF_BRISK_CPP=\"<reference>/src/MoFREAK/brisk.cpp\""
#endif
#include MFK_REF_BRISK_CPP

#include <stdint.h>
#include <string.h>

namespace {

cv::Mat wrap_u8(const uint8_t* img, int w, int h) {
    cv::Mat m(h, w, CV_8U);
    memcpy(m.data, img, (size_t)w * h);
    return m;
}

// Grants access to the protected pyramid for stage dumps.
class OpenScaleSpace : public cv::BriskScaleSpace {
public:
    explicit OpenScaleSpace(uint8_t octaves) : cv::BriskScaleSpace(octaves) {}
    int layers() const { return layers_; }
    const cv::BriskLayer& layer(int i) const { return pyramid_[i]; }
};

}  // namespace

extern "C" {

// out: n rows of {x, y, size, response, octave}; returns the number of keypoints (may exceed cap;
// only the first cap are written).
int ref_brisk_detect(const uint8_t* img, int w, int h, int thresh, int octaves, float* out, int cap) {
    cv::Mat m = wrap_u8(img, w, h);
    std::vector<cv::KeyPoint> kps;
    cv::BriskFeatureDetector det(thresh, octaves);
    det.detect(m, kps);
    int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        out[5 * i + 0] = kps[i].pt.x;
        out[5 * i + 1] = kps[i].pt.y;
        out[5 * i + 2] = kps[i].size;
        out[5 * i + 3] = kps[i].response;
        out[5 * i + 4] = (float)kps[i].octave;
    }
    return n;
}

void ref_halfsample(const uint8_t* src, int w, int h, uint8_t* dst) {
    cv::Mat s = wrap_u8(src, w, h);
    cv::Mat d(h / 2, w / 2, CV_8U);
    cv::BriskLayer::halfsample(s, d);
    memcpy(dst, d.data, (size_t)(w / 2) * (h / 2));
}

void ref_twothirdsample(const uint8_t* src, int w, int h, uint8_t* dst) {
    cv::Mat s = wrap_u8(src, w, h);
    cv::Mat d(2 * (h / 3), 2 * (w / 3), CV_8U);
    cv::BriskLayer::twothirdsample(s, d);
    memcpy(dst, d.data, (size_t)(2 * (w / 3)) * (2 * (h / 3)));
}

// xy: up to cap (x, y) pairs in the reference's push order; returns the number of hits.
int ref_oast_detect(const uint8_t* img, int w, int h, int b, int* xy, int cap) {
    agast::OastDetector9_16 det(w, h, b);
    std::vector<CvPoint> pts;
    det.detect(img, pts);
    int n = (int)pts.size();
    for (int i = 0; i < n && i < cap; i++) { xy[2 * i] = pts[i].x; xy[2 * i + 1] = pts[i].y; }
    return n;
}

int ref_oast_score(const uint8_t* img, int w, int h, int x, int y, int b) {
    agast::OastDetector9_16 det(w, h, b);
    return det.cornerScore(img + (size_t)y * w + x);
}

int ref_agast58_score(const uint8_t* img, int w, int h, int x, int y, int b) {
    agast::AgastDetector5_8 det(w, h, b);
    return det.cornerScore(img + (size_t)y * w + x);
}

// Runs constructPyramid + getKeypoints and copies out layer `layer`'s image and its FINAL raw score
// map (what scores_ holds after every lazy write-back).  dims[0..1] = layer width, height.
// Returns the number of layers.
int ref_brisk_dump(const uint8_t* img, int w, int h, int thresh, int octaves, int layer,
                   uint8_t* layer_img, uint8_t* layer_scores, int* dims) {
    cv::Mat m = wrap_u8(img, w, h);
    OpenScaleSpace ss((uint8_t)octaves);
    ss.constructPyramid(m);
    std::vector<cv::KeyPoint> kps;
    ss.getKeypoints((uint8_t)thresh, kps);
    if (layer >= 0 && layer < ss.layers()) {
        const cv::BriskLayer& l = ss.layer(layer);
        dims[0] = l.img().cols; dims[1] = l.img().rows;
        if (layer_img) memcpy(layer_img, l.img().data, (size_t)dims[0] * dims[1]);
        if (layer_scores) memcpy(layer_scores, l.scores().data, (size_t)dims[0] * dims[1]);
    }
    return ss.layers();
}

}  // extern "C"
