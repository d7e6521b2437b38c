// Minimal stand-in for <opencv.hpp>: this image has no OpenCV C++ headers.
// The LSD-SLAM sources use cv::Mat purely as a ref-counted 2-D array
// (Mat::zeros, ptr<T>(row), rows, cols, size[i], clone, release — and, in the catkin snapshot's
// FeatureAssociation.cpp, ptr<T>(row, col), colRange / row views and copyTo), so that is all this header provides.  It is used (a) to compile the unmodified
// reference into oracle/_ref and (b) to compile this repo's host wrappers.
//
// A test-only allocation registry lets a harness recover Mats that are local
// to a callee (usedMap, degMap, ... in myLineSegmentDetector): when armed with
// a budget K, the first K allocations are retained and can be read back.
#ifndef LSDB_SHIM_OPENCV_HPP
#define LSDB_SHIM_OPENCV_HPP

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>

#define CV_8U 0
#define CV_16U 2
#define CV_64F 6
#define CV_8UC1 0
#define CV_16UC1 2
#define CV_64FC1 6

namespace cv {

struct MatBuf {
    unsigned char* data;
    int refcount;
};

class Mat;
struct MatRegistry {
    int budget;                 // how many more allocations to retain
    std::vector<Mat>* kept;     // retained handles (owned by the harness)
    MatRegistry() : budget(0), kept(0) {}
};
inline MatRegistry& mat_registry() {
    static thread_local MatRegistry r;
    return r;
}

struct MatSize {
    int dims[2];
    int operator[](int i) const { return dims[i]; }
};

class Mat {
public:
    int rows, cols, type_;
    MatSize size;
    size_t step;
    unsigned char* data;

    Mat() : rows(0), cols(0), type_(0), step(0), data(0), buf_(0) { size.dims[0] = size.dims[1] = 0; }
    Mat(int r, int c, int type) : buf_(0) { create(r, c, type); }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), type_(o.type_), size(o.size), step(o.step), data(o.data), buf_(o.buf_) {
        if (buf_) __sync_fetch_and_add(&buf_->refcount, 1);
    }
    Mat& operator=(const Mat& o) {
        if (this == &o) return *this;
        if (o.buf_) __sync_fetch_and_add(&o.buf_->refcount, 1);
        release();
        rows = o.rows; cols = o.cols; type_ = o.type_; size = o.size; step = o.step; data = o.data; buf_ = o.buf_;
        return *this;
    }
    ~Mat() { release(); }

    static size_t elemSizeOf(int type) { return type == CV_64F ? 8 : (type == CV_16U ? 2 : 1); }
    size_t elemSize() const { return elemSizeOf(type_); }
    int type() const { return type_; }
    bool empty() const { return data == 0; }

    void create(int r, int c, int type) {
        release();
        rows = r; cols = c; type_ = type;
        size.dims[0] = r; size.dims[1] = c;
        step = (size_t)c * elemSizeOf(type);
        buf_ = (MatBuf*)malloc(sizeof(MatBuf));
        // +16 bytes of slack: the reference's driver fscanf("%d")s into uint8 slots
        buf_->data = (unsigned char*)calloc((size_t)r * step + 16, 1);
        buf_->refcount = 1;
        data = buf_->data;
        MatRegistry& reg = mat_registry();
        if (reg.budget > 0 && reg.kept) { reg.budget--; reg.kept->push_back(*this); }
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }

    void release() {
        if (buf_ && __sync_sub_and_fetch(&buf_->refcount, 1) == 0) {
            free(buf_->data);
            free(buf_);
        }
        buf_ = 0; data = 0; rows = cols = 0; size.dims[0] = size.dims[1] = 0;
    }
    Mat clone() const {
        Mat m;
        if (!data) return m;
        MatRegistry& reg = mat_registry();
        int saved = reg.budget; reg.budget = 0;       // clones are never "internal" Mats
        m.create(rows, cols, type_);
        reg.budget = saved;
        memcpy(m.data, data, (size_t)rows * step);
        return m;
    }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> T* ptr(int r, int c) { return (T*)(data + (size_t)r * step) + c; }
    template <typename T> const T* ptr(int r, int c) const { return (const T*)(data + (size_t)r * step) + c; }

    // views share the buffer (and its reference count) with the matrix they were cut from
    Mat colRange(int c0, int c1) const {
        Mat v(*this);
        v.data = data + (size_t)c0 * elemSize();
        v.cols = c1 - c0; v.size.dims[1] = v.cols;
        return v;
    }
    Mat row(int r) const {
        Mat v(*this);
        v.data = data + (size_t)r * step;
        v.rows = 1; v.size.dims[0] = 1;
        return v;
    }
    // dst of the same shape (a view, typically) is written in place; anything else is re-created, like cv::Mat::copyTo
    void copyTo(Mat dst) const {
        if (!dst.data || dst.rows != rows || dst.cols != cols || dst.type_ != type_) return;
        for (int r = 0; r < rows; r++) memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * elemSize());
    }

private:
    MatBuf* buf_;
};

}  // namespace cv
#endif
