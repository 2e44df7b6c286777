// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for the handful of OpenCV 2.4 types the
// reference's vendored BRISK/AGAST sources touch, so that those sources can be compiled UNCHANGED,
// in place under $MOFREAK_REF (default /root/reference), into oracle/_ref/.  Nothing here is
// product code and nothing in the product links it.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <assert.h>
#include <vector>
#include <memory>
#include <algorithm>

typedef unsigned char uchar;
#define CV_EXPORTS
#define CV_8U 0
#define CV_32S 4

#ifndef MFK_SHIM_CVPOINT
#define MFK_SHIM_CVPOINT
typedef struct CvPoint { int x; int y; } CvPoint;
#endif

namespace cv {

struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

// Dense, single-channel matrix.  Copies share the pixel buffer (the reference relies on that:
// BriskLayer keeps `img_ = img`), rows are contiguous (step == cols * elemsize), the base pointer is
// 64-byte aligned (the reference uses aligned SSE loads when cols % 16 == 0) and the allocation is
// padded because the reference's 16-byte loads/stores run past the end of the last row.
class Mat {
public:
    int rows, cols;
    uchar* data;

    Mat() : rows(0), cols(0), data(0), type_(CV_8U) {}
    Mat(int r, int c, int t) : rows(0), cols(0), data(0), type_(CV_8U) { create(r, c, t); }

    void create(int r, int c, int t) {
        rows = r; cols = c; type_ = t;
        size_t bytes = (size_t)r * c * elem(t);
        store_.reset(new std::vector<uchar>(bytes + 192, 0));
        data = (uchar*)(((uintptr_t)store_->data() + 63) & ~(uintptr_t)63);
    }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }   // create() zero-fills
    Mat clone() const {
        Mat m(rows, cols, type_);
        if (data) memcpy(m.data, data, (size_t)rows * cols * elem(type_));
        return m;
    }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    void release() { store_.reset(); data = 0; rows = cols = 0; }
    int type() const { return type_; }

private:
    static size_t elem(int t) { return t == CV_8U ? 1 : 4; }
    int type_;
    std::shared_ptr<std::vector<uchar> > store_;
};

template <class T> class Ptr {
public:
    Ptr() {}
    Ptr(T* raw) : p_(raw) {}
    Ptr& operator=(T* raw) { p_.reset(raw); return *this; }
    T* operator->() const { return p_.get(); }
    T& operator*() const { return *p_; }
private:
    std::shared_ptr<T> p_;
};

// 32-bit summed-area table with a leading zero row/column, (rows+1) x (cols+1).
inline void integral(const Mat& src, Mat& sum) {
    sum.create(src.rows + 1, src.cols + 1, CV_32S);
    int* S = (int*)sum.data;
    const int W = src.cols + 1;
    for (int y = 0; y < src.rows; y++) {
        int run = 0;
        for (int x = 0; x < src.cols; x++) {
            run += src.data[(size_t)y * src.cols + x];
            S[(size_t)(y + 1) * W + x + 1] = S[(size_t)y * W + x + 1] + run;
        }
    }
}

}  // namespace cv
