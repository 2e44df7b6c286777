#include "MotionInterchangePatterns.h"

#include <stdlib.h>

#include <iostream>

static void mipDie(const char* what) { std::cout << "mofreak_b200: " << what << std::endl; exit(1); }

MotionInterchangePatterns::MotionInterchangePatterns(int width, int height, int device) : ctx_(nullptr), w_(width), h_(height) {
    if (mfk_create(&ctx_, device, width, height, MFK_FRAME_GAP + 1, 1) != MFK_OK) mipDie(mfk_last_error(nullptr));
}

MotionInterchangePatterns::~MotionInterchangePatterns() { mfk_destroy(ctx_); }

void MotionInterchangePatterns::setFrames(const uint8_t* current_frame, const uint8_t* prev_frame) {
    if (mfk_upload_frames(ctx_, current_frame, (size_t)w_, 1, 1) != MFK_OK || mfk_upload_frames(ctx_, prev_frame, (size_t)w_, 1, 0) != MFK_OK)
        mipDie(mfk_last_error(ctx_));
}

void MotionInterchangePatterns::extractMany(const float* kp_xys, int n, uint8_t* out) {
    // an out-of-frame region of interest is an OpenCV assertion (abort) in the reference; here it is reported and fatal
    if (mfk_mip_describe(ctx_, 1, 0, kp_xys, n, out) != MFK_OK) mipDie(mfk_last_error(ctx_));
}

void MotionInterchangePatterns::extractMotionByMotionInterchangePatterns(std::vector<unsigned int>& motion_descriptor, float scale, int x, int y) {
    const float kp[3] = {(float)x, (float)y, scale};
    uint8_t out[8];
    extractMany(kp, 1, out);
    for (int i = 0; i < 8; i++) motion_descriptor.push_back(out[i]);
}
