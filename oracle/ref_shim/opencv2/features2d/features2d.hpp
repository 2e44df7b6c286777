// TEST INFRASTRUCTURE ONLY (oracle/): see core/core.hpp.
#pragma once
#include <opencv2/core/core.hpp>

namespace cv {

class FeatureDetector {
public:
    virtual ~FeatureDetector() {}
    // OpenCV 2.4 semantics: clear the output, return on an empty image, else dispatch.
    void detect(const Mat& image, std::vector<KeyPoint>& keypoints, const Mat& mask = Mat()) const {
        keypoints.clear();
        if (image.empty()) return;
        detectImpl(image, keypoints, mask);
    }
protected:
    virtual void detectImpl(const Mat& image, std::vector<KeyPoint>& keypoints,
                            const Mat& mask = Mat()) const = 0;
    // The reference only ever passes the empty mask, for which this is a no-op.
    static void removeInvalidPoints(const Mat& mask, std::vector<KeyPoint>&) { assert(mask.empty()); (void)mask; }
};

class DescriptorExtractor {
public:
    virtual ~DescriptorExtractor() {}
    virtual int descriptorSize() const = 0;
    virtual int descriptorType() const = 0;
protected:
    virtual void computeImpl(const Mat& image, std::vector<KeyPoint>& keypoints, Mat& descriptors) const = 0;
};

}  // namespace cv
