// Stand-in for the OpenCV 2.4 C++ API (not installed here).  TEST INFRASTRUCTURE, see Eigen/Dense.
// Only what HFTest.cpp touches.  cv::blur is stated here the way OpenCV 2.4's boxFilter evaluates CV_32F input:
// row sums and column sums in double (sumType CV_64F), one multiplication by the double scale 1/(kw*kh), narrowed to
// float, BORDER_REFLECT_101 (imgproc/src/smooth.cpp, createBoxFilter).  tests/test_golden.py compares the same statement
// with the real cv2.blur.
#ifndef HF6D_SHIM_CV_HPP
#define HF6D_SHIM_CV_HPP
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_LOAD_IMAGE_ANYDEPTH 2
#define CV_LOAD_IMAGE_ANYCOLOR 4

typedef unsigned char uchar;
typedef unsigned short ushort;

namespace cv {

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
template <typename T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
typedef Size Size2i;

class Mat {
    struct Buf {
        uchar* p;
        int refs;
        bool owned;
    };
    Buf* b_;
    void release() {
        if (b_ && --b_->refs == 0) {
            if (b_->owned) std::free(b_->p);
            delete b_;
        }
        b_ = 0;
    }
    static int elem_size(int type) {
        const int depth = type & 7, cn = (type >> 3) + 1;
        const int ds = depth == CV_8U ? 1 : depth == CV_16U ? 2 : 4;
        return ds * cn;
    }

  public:
    int rows, cols, type_;
    uchar* data;
    uchar* datastart;
    uchar* dataend;
    Mat() : b_(0), rows(0), cols(0), type_(0), data(0), datastart(0), dataend(0) {}
    Mat(int r, int c, int type, const Scalar& s = Scalar()) : b_(0) { create(r, c, type); fill(s); }
    // wraps caller memory (not owned)
    Mat(int r, int c, int type, void* ext) : b_(0), rows(r), cols(c), type_(type) {
        b_ = new Buf;
        b_->p = (uchar*)ext; b_->refs = 1; b_->owned = false;
        data = datastart = b_->p;
        dataend = data + (size_t)r * c * elem_size(type);
    }
    Mat(const Mat& o) : b_(o.b_), rows(o.rows), cols(o.cols), type_(o.type_), data(o.data), datastart(o.datastart), dataend(o.dataend) {
        if (b_) ++b_->refs;
    }
    Mat& operator=(const Mat& o) {
        if (this != &o) {
            if (o.b_) ++o.b_->refs;
            release();
            b_ = o.b_; rows = o.rows; cols = o.cols; type_ = o.type_; data = o.data; datastart = o.datastart; dataend = o.dataend;
        }
        return *this;
    }
    ~Mat() { release(); }
    void create(int r, int c, int type) {
        release();
        rows = r; cols = c; type_ = type;
        const size_t bytes = (size_t)r * c * elem_size(type);
        b_ = new Buf;
        b_->p = (uchar*)std::calloc(bytes ? bytes : 1, 1);
        b_->refs = 1; b_->owned = true;
        data = datastart = b_->p;
        dataend = data + bytes;
    }
    void fill(const Scalar& s) {
        const int depth = type_ & 7, cn = (type_ >> 3) + 1;
        for (size_t i = 0; i < (size_t)rows * cols; ++i)
            for (int k = 0; k < cn; ++k) {
                if (depth == CV_8U) data[i * cn + k] = (uchar)s.val[k];
                else if (depth == CV_16U) ((ushort*)data)[i * cn + k] = (ushort)s.val[k];
                else ((float*)data)[i * cn + k] = (float)s.val[k];
            }
    }
    int type() const { return type_; }
    bool empty() const { return data == 0 || rows * cols == 0; }
    size_t elemSize() const { return (size_t)elem_size(type_); }
    template <typename T> T& at(int r, int c) { return *(T*)(data + ((size_t)r * cols + c) * elem_size(type_)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + ((size_t)r * cols + c) * elem_size(type_)); }
    template <typename T> T& at(int i) { return *(T*)(data + (size_t)i * elem_size(type_)); }
    template <typename T> const T& at(int i) const { return *(const T*)(data + (size_t)i * elem_size(type_)); }
    void copyTo(Mat& dst) const {
        Mat t;
        t.create(rows, cols, type_);
        std::memcpy(t.data, data, (size_t)(dataend - datastart));
        dst = t;
    }
    Mat& operator+=(const Mat& o) {  // CV_32FC1 only (the vote-map merge, HFTest.cpp:649)
        float* a = (float*)data;
        const float* b = (const float*)o.data;
        for (size_t i = 0; i < (size_t)rows * cols; ++i) a[i] = a[i] + b[i];
        return *this;
    }
    Mat operator*(double) const { return *this; }  // display only
    Mat& operator/=(double) { return *this; }      // display only
};

enum { MORPH_ELLIPSE = 2 };

void blur(const Mat& src, Mat& dst, Size ksize);
Mat imread(const std::string& filename, int flags = 1);
bool imwrite(const std::string& filename, const Mat& img);
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return 0; }
inline void circle(Mat&, Point, int, const Scalar&, int = 1) {}
inline void minMaxLoc(const Mat&, double*, double*, Point* = 0, Point* = 0) {}
inline Mat getStructuringElement(int, Size, Point = Point(-1, -1)) { return Mat(); }
inline void erode(const Mat&, Mat&, const Mat&) {}
inline void dilate(const Mat&, Mat&, const Mat&) {}

}  // namespace cv
#endif
