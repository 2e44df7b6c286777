// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restatement of cv::FREAK with default parameters (orientationNormalized = scaleNormalized = true,
// patternScale = 22, nOctaves = 4, default pairs), OpenCV 2.4.2 modules/features2d/src/freak.cpp.
// The reference does not vendor OpenCV; its call site is MoFREAKUtilities.cpp:427-428 and it keeps the
// first 8 of the 64 descriptor bytes (:453-456).
//
// PARITY UNPINNED: no FREAK implementation is available in this container (cv2 4.13 headless has no
// xfeatures2d) and the reference ships no golden vectors, so this follows the published algorithm
// (Alahi et al., CVPR 2012, as implemented in OpenCV 2.4.x) from SURVEY.md Appendix A.3 only.
#include "oracle.h"

#include <math.h>
#include <string.h>
#include <algorithm>

namespace orc {

static const int NB_SCALES = 64, NB_ORIENTATION = 256, NB_POINTS = 43, NB_PAIRS = 512, NB_ORIENPAIRS = 45;
static const int SMALLEST_KP_SIZE = 7;
static const double FREAK_LOG2 = 0.693147180559945;
static const double PI_CV = 3.1415926535897932384626433832795;   // CV_PI

// default description pairs (indices into the i>j enumeration of all 903 point pairs)
static const int DEF_PAIRS[NB_PAIRS] = {
    404,431,818,511,181,52,311,874,774,543,719,230,417,205,11,560,149,265,39,306,165,857,250,8,61,15,55,717,44,412,592,134,
    761,695,660,782,625,487,549,516,271,665,762,392,178,796,773,31,672,845,548,794,677,654,241,831,225,238,849,83,691,484,826,707,
    122,517,583,731,328,339,571,475,394,472,580,381,137,93,380,327,619,729,808,218,213,459,141,806,341,95,382,568,124,750,193,749,
    706,843,79,199,317,329,768,198,100,466,613,78,562,783,689,136,838,94,142,164,679,219,419,366,418,423,77,89,523,259,683,312,
    555,20,470,684,123,458,453,833,72,113,253,108,313,25,153,648,411,607,618,128,305,232,301,84,56,264,371,46,407,360,38,99,
    176,710,114,578,66,372,653,129,359,424,159,821,10,323,393,5,340,891,9,790,47,0,175,346,236,26,172,147,574,561,32,294,
    429,724,755,398,787,288,299,769,565,767,722,757,224,465,723,498,467,235,127,802,446,233,544,482,800,318,16,532,801,441,554,173,
    60,530,713,469,30,212,630,899,170,266,799,88,49,512,399,23,500,107,524,90,194,143,135,192,206,345,148,71,119,101,563,870,
    158,254,214,276,464,332,725,188,385,24,476,40,231,620,171,258,67,109,844,244,187,388,701,690,50,7,850,479,48,522,22,154,
    12,659,736,655,577,737,830,811,174,21,237,335,353,234,53,270,62,182,45,177,245,812,673,355,556,612,166,204,54,248,365,226,
    242,452,700,685,573,14,842,481,468,781,564,416,179,405,35,819,608,624,367,98,643,448,2,460,676,440,240,130,146,184,185,430,
    65,807,377,82,121,708,239,310,138,596,730,575,477,851,797,247,27,85,586,307,779,326,494,856,324,827,96,748,13,397,125,688,
    702,92,293,716,277,140,112,4,80,855,839,1,413,347,584,493,289,696,19,751,379,76,73,115,6,590,183,734,197,483,217,344,
    330,400,186,243,587,220,780,200,793,246,824,41,735,579,81,703,322,760,720,139,480,490,91,814,813,163,152,488,763,263,425,410,
    576,120,319,668,150,160,302,491,515,260,145,428,97,251,395,272,252,18,106,358,854,485,144,550,131,133,378,68,102,104,58,361,
    275,209,697,582,338,742,589,325,408,229,28,304,191,189,110,126,486,211,547,533,70,215,670,249,36,581,389,605,331,518,442,822};

struct PatternPoint { float x, y, sigma; };
struct OrientationPair { int i, j, weight_dx, weight_dy; };
struct DescriptionPair { uint8_t i, j; };

struct FreakPattern {
    std::vector<PatternPoint> lookup;       // [scale][orientation][point]
    int sizes[NB_SCALES];
    OrientationPair orient[NB_ORIENPAIRS];
    DescriptionPair pairs[NB_PAIRS];
};

static void build(FreakPattern& P) {
    const double patternScale = 22.0f;
    const int nOctaves = 4;
    P.lookup.resize((size_t)NB_SCALES * NB_ORIENTATION * NB_POINTS);
    const double scaleStep = pow(2.0, (double)(nOctaves) / NB_SCALES);
    // rings from outer to inner, radius normalised to 1
    const int n[8] = {6, 6, 6, 6, 6, 6, 6, 1};
    const double bigR(2.0 / 3.0);
    const double smallR(2.0 / 24.0);
    const double unitSpace((bigR - smallR) / 21.0);
    const double radius[8] = {bigR, bigR - 6 * unitSpace, bigR - 11 * unitSpace, bigR - 15 * unitSpace,
                              bigR - 18 * unitSpace, bigR - 20 * unitSpace, smallR, 0.0};
    const double sigma[8] = {radius[0] / 2.0, radius[1] / 2.0, radius[2] / 2.0, radius[3] / 2.0,
                             radius[4] / 2.0, radius[5] / 2.0, radius[6] / 2.0, radius[6] / 2.0};
    for (int scaleIdx = 0; scaleIdx < NB_SCALES; ++scaleIdx) {
        P.sizes[scaleIdx] = 0;
        const double scalingFactor = pow(scaleStep, scaleIdx);
        for (int orientationIdx = 0; orientationIdx < NB_ORIENTATION; ++orientationIdx) {
            const double theta = double(orientationIdx) * 2 * PI_CV / double(NB_ORIENTATION);
            int pointIdx = 0;
            for (size_t i = 0; i < 8; ++i) {
                for (int k = 0; k < n[i]; ++k) {
                    const double beta = PI_CV / n[i] * (i % 2);   // alternate rings are staggered
                    const double alpha = double(k) * 2 * PI_CV / double(n[i]) + beta + theta;
                    PatternPoint& point =
                        P.lookup[(size_t)scaleIdx * NB_ORIENTATION * NB_POINTS + orientationIdx * NB_POINTS + pointIdx];
                    point.x = static_cast<float>(radius[i] * cos(alpha) * scalingFactor * patternScale);
                    point.y = static_cast<float>(radius[i] * sin(alpha) * scalingFactor * patternScale);
                    point.sigma = static_cast<float>(sigma[i] * scalingFactor * patternScale);
                    const int sizeMax =
                        static_cast<int>(ceil((radius[i] + sigma[i]) * scalingFactor * patternScale)) + 1;
                    if (P.sizes[scaleIdx] < sizeMax) P.sizes[scaleIdx] = sizeMax;
                    ++pointIdx;
                }
            }
        }
    }
    // orientation pairs: four outer rings, 9 pairs each, then one diameter triple per inner ring pair
    int m = 0;
    for (int b = 0; b < 24; b += 6) {
        const int ij[9][2] = {{0, 3}, {1, 4}, {2, 5}, {0, 2}, {1, 3}, {2, 4}, {3, 5}, {4, 0}, {5, 1}};
        for (int q = 0; q < 9; q++, m++) { P.orient[m].i = b + ij[q][0]; P.orient[m].j = b + ij[q][1]; }
    }
    {
        const int ij[9][2] = {{24, 27}, {25, 28}, {26, 29}, {30, 33}, {31, 34}, {32, 35}, {36, 39}, {37, 40}, {38, 41}};
        for (int q = 0; q < 9; q++, m++) { P.orient[m].i = ij[q][0]; P.orient[m].j = ij[q][1]; }
    }
    for (unsigned q = NB_ORIENPAIRS; q--;) {
        const float dx = P.lookup[P.orient[q].i].x - P.lookup[P.orient[q].j].x;
        const float dy = P.lookup[P.orient[q].i].y - P.lookup[P.orient[q].j].y;
        const float norm_sq = (dx * dx + dy * dy);
        P.orient[q].weight_dx = int((dx / (norm_sq)) * 4096.0 + 0.5);
        P.orient[q].weight_dy = int((dy / (norm_sq)) * 4096.0 + 0.5);
    }
    std::vector<DescriptionPair> allPairs;
    for (unsigned i = 1; i < (unsigned)NB_POINTS; ++i)
        for (unsigned j = 0; j < i; ++j) { DescriptionPair pr = {(uint8_t)i, (uint8_t)j}; allPairs.push_back(pr); }
    for (int i = 0; i < NB_PAIRS; ++i) P.pairs[i] = allPairs[DEF_PAIRS[i]];
}

const FreakPattern& freak_pattern() {
    static FreakPattern* P = 0;
    if (!P) { P = new FreakPattern(); build(*P); }
    return *P;
}
FreakPattern* freak_build_pattern() { FreakPattern* P = new FreakPattern(); build(*P); return P; }
void freak_free_pattern(FreakPattern* P) { delete P; }
int freak_pattern_size(const FreakPattern& P, int s) { return P.sizes[s]; }

int freak_scale_index(float kp_size) {
    const float sizeCst = static_cast<float>(NB_SCALES / (FREAK_LOG2 * 4));
    // log() on a float argument resolves to the float overload, i.e. logf
    int idx = std::max((int)(logf(kp_size / SMALLEST_KP_SIZE) * sizeCst + 0.5), 0);
    if (idx >= NB_SCALES) idx = NB_SCALES - 1;
    return idx;
}

void integral_u8(const uint8_t* img, int w, int h, int32_t* sum) {
    const int W = w + 1;
    memset(sum, 0, sizeof(int32_t) * W);
    for (int y = 0; y < h; y++) {
        int run = 0;
        sum[(size_t)(y + 1) * W] = 0;
        for (int x = 0; x < w; x++) {
            run += img[(size_t)y * w + x];
            sum[(size_t)(y + 1) * W + x + 1] = sum[(size_t)y * W + x + 1] + run;
        }
    }
}

// FREAK::meanIntensity: box mean of side ~2*sigma around the pattern point, from the integral image
static inline uint8_t mean_intensity(const FreakPattern& P, const uint8_t* img, const int32_t* integral, int w,
                                     const float kp_x, const float kp_y, const unsigned scale, const unsigned rot,
                                     const unsigned point) {
    const PatternPoint& FreakPoint = P.lookup[(size_t)scale * NB_ORIENTATION * NB_POINTS + rot * NB_POINTS + point];
    const float xf = FreakPoint.x + kp_x;
    const float yf = FreakPoint.y + kp_y;
    const int x = int(xf);
    const int y = int(yf);
    const int imagecols = w;
    const float radius = FreakPoint.sigma;
    int ret_val;
    if (radius < 0.5) {
        // 4-tap interpolation; unreachable with patternScale 22 (smallest sigma is 0.9167)
        const int r_x = static_cast<int>((xf - x) * 1024);
        const int r_y = static_cast<int>((yf - y) * 1024);
        const int r_x_1 = (1024 - r_x);
        const int r_y_1 = (1024 - r_y);
        const uint8_t* ptr = img + x + y * imagecols;
        ret_val = (r_x_1 * r_y_1 * int(*ptr));
        ptr++;
        ret_val += (r_x * r_y_1 * int(*ptr));
        ptr += imagecols;
        ret_val += (r_x * r_y * int(*ptr));
        ptr--;
        ret_val += (r_x_1 * r_y * int(*ptr));
        return static_cast<uint8_t>((ret_val + 512) / 1024);
    }
    const int x_left = int(xf - radius + 0.5);
    const int y_top = int(yf - radius + 0.5);
    const int x_right = int(xf + radius + 1.5);    // integral image is 1 px wider
    const int y_bottom = int(yf + radius + 1.5);   // and 1 px higher
    const int W = w + 1;
    ret_val = integral[(size_t)y_bottom * W + x_right];
    ret_val -= integral[(size_t)y_bottom * W + x_left];
    ret_val += integral[(size_t)y_top * W + x_left];
    ret_val -= integral[(size_t)y_top * W + x_right];
    ret_val = ret_val / ((x_right - x_left) * (y_bottom - y_top));
    return static_cast<uint8_t>(ret_val);
}

void freak_compute(const FreakPattern& P, const uint8_t* img, int w, int h, std::vector<Keypoint>& kps,
                   std::vector<uint8_t>& desc, std::vector<int>* theta_out) {
    desc.clear();
    if (theta_out) theta_out->clear();
    if (w == 0 || h == 0 || kps.empty()) return;
    std::vector<int32_t> integral((size_t)(w + 1) * (h + 1));
    integral_u8(img, w, h, integral.data());

    // scale index per keypoint; drop keypoints whose pattern does not fit inside the image
    std::vector<int> kpScaleIdx(kps.size());
    for (size_t k = kps.size(); k--;) {
        kpScaleIdx[k] = freak_scale_index(kps[k].size);
        const int ps = P.sizes[kpScaleIdx[k]];
        if (kps[k].x <= ps || kps[k].y <= ps || kps[k].x >= w - ps || kps[k].y >= h - ps) {
            kps.erase(kps.begin() + k);
            kpScaleIdx.erase(kpScaleIdx.begin() + k);
        }
    }
    desc.assign(kps.size() * 64, 0);
    if (theta_out) theta_out->assign(kps.size(), 0);
    uint8_t pointsValue[NB_POINTS];
    for (size_t k = kps.size(); k--;) {
        for (int i = NB_POINTS; i--;)
            pointsValue[i] = mean_intensity(P, img, integral.data(), w, kps[k].x, kps[k].y, kpScaleIdx[k], 0, i);
        int direction0 = 0, direction1 = 0;
        for (int m = 45; m--;) {
            const int delta = (pointsValue[P.orient[m].i] - pointsValue[P.orient[m].j]);
            direction0 += delta * (P.orient[m].weight_dx) / 2048;
            direction1 += delta * (P.orient[m].weight_dy) / 2048;
        }
        // atan2() on float arguments resolves to the float overload, i.e. atan2f
        const float angle = static_cast<float>(atan2f((float)direction1, (float)direction0) * (180.0 / PI_CV));
        int thetaIdx = int(NB_ORIENTATION * angle * (1 / 360.0) + 0.5);
        if (thetaIdx < 0) thetaIdx += NB_ORIENTATION;
        if (thetaIdx >= NB_ORIENTATION) thetaIdx -= NB_ORIENTATION;
        if (theta_out) (*theta_out)[k] = thetaIdx;
        for (int i = NB_POINTS; i--;)
            pointsValue[i] = mean_intensity(P, img, integral.data(), w, kps[k].x, kps[k].y, kpScaleIdx[k], thetaIdx, i);
        // bit layout of the SSE2 build (the scalar build reproduces it): inside each block of 128
        // pairs, pair p lands in byte 15 - (p % 16), bit p / 16
        uint8_t* row = desc.data() + k * 64;
        for (int p = 0; p < NB_PAIRS; p++) {
            if (pointsValue[P.pairs[p].i] >= pointsValue[P.pairs[p].j]) {
                const int blk = p / 128, q = p % 128;
                row[blk * 16 + 15 - (q % 16)] |= (uint8_t)(1u << (q / 16));
            }
        }
    }
}

}  // namespace orc
