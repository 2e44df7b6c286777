// The MIP facade BASELINE.json's north_star asks for: the reference keeps these functions as private members of
// MoFREAKUtilities (MoFREAKUtilities.h:83-90, MoFREAKUtilities.cpp:46-99, 288-325); here they are one class with the
// reference's parameter order, running on the GPU through mfk_mip_describe.  The per-patch SSD test
// (motionInterchangePattern) is the device function mip_ssd() of csrc/desc_device.cuh and is not exposed on the host.
#ifndef MOFREAK_B200_MOTIONINTERCHANGEPATTERNS_H
#define MOFREAK_B200_MOTIONINTERCHANGEPATTERNS_H

#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/mofreak_b200.h"

class MotionInterchangePatterns {
public:
    MotionInterchangePatterns(int width, int height, int device = 0);
    ~MotionInterchangePatterns();
    MotionInterchangePatterns(const MotionInterchangePatterns&) = delete;
    MotionInterchangePatterns& operator=(const MotionInterchangePatterns&) = delete;

    // the two gray frames (width*height bytes each) all following calls refer to
    void setFrames(const uint8_t* current_frame, const uint8_t* prev_frame);
    // reference signature minus the cv::Mat frames: appends the 8 motion bytes to motion_descriptor
    void extractMotionByMotionInterchangePatterns(std::vector<unsigned int>& motion_descriptor, float scale, int x, int y);
    // many keypoints at once: kp = n x {x, y, size}; out = n x 8 bytes
    void extractMany(const float* kp_xys, int n, uint8_t* out);

private:
    mfk_ctx* ctx_;
    int w_, h_;
};

#endif
