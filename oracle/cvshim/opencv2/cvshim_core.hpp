// Test infrastructure only (oracle build): a minimal header-only stand-in for the parts of
// OpenCV's cv::Mat that the reference's hot-path translation units touch.  OpenCV C++ is not
// installed in this image; the reference only uses cv::Mat as a typed, ref-counted 2-D array
// (SURVEY.md §8c).  Type codes follow OpenCV's numeric values so `depth()==CV_32FC1` etc.
// behave identically.  Copies are shallow (shared buffer), clone() is deep -- the reference
// relies on header aliasing (RGBDFrame.gray vs pyramid level 0).
#ifndef OPB_CVSHIM_CORE_HPP
#define OPB_CVSHIM_CORE_HPP
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <set>
#include <map>
#include <string>
#include <iostream>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC2 CV_MAKETYPE(CV_32S, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv
{
class Mat;
template <typename T, int N> struct Vec
{
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(const Mat &) { for (int i = 0; i < N; ++i) val[i] = T(0); } // only GlobalRegistration.cpp's unused Eigen2OpenCV needs it to parse
    Vec(T a, T b) { static_assert(N >= 2, ""); val[0] = a; val[1] = b; for (int i = 2; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N >= 3, ""); val[0] = a; val[1] = b; val[2] = c; for (int i = 3; i < N; ++i) val[i] = T(0); }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
    T &operator()(int i) { return val[i]; }
    const T &operator()(int i) const { return val[i]; }
};
typedef Vec<unsigned char, 3> Vec3b;
typedef Vec<int, 2> Vec2i;
typedef Vec<float, 3> Vec3f;

struct Scalar
{
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};
struct Size
{
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};
struct Point2f { float x, y; Point2f(float _x = 0, float _y = 0) : x(_x), y(_y) {} };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };

class Mat
{
  public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); fill(s); }
    // wrap external memory without owning it (like cv::Mat(rows, cols, type, void*))
    Mat(int r, int c, int type, void *ext) : rows(r), cols(c), data((unsigned char *)ext), type_(type) {}
    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type;
        buf_ = std::shared_ptr<unsigned char>(new unsigned char[(size_t)r * c * elemSize() + 64], std::default_delete<unsigned char[]>());
        data = buf_.get();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; }
    Mat clone() const
    {
        Mat m;
        if (!data) return m;
        m.create(rows, cols, type_);
        std::memcpy(m.data, data, (size_t)rows * cols * elemSize());
        return m;
    }
    void copyTo(Mat &o) const { o = clone(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int depth() const { return type_ & ((1 << CV_CN_SHIFT) - 1); }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize1() const
    {
        switch (depth()) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2;
                           case CV_32S: case CV_32F: return 4; default: return 8; }
    }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t total() const { return (size_t)rows * cols; }
    template <typename T> T &at(int r, int c) { return ((T *)data)[(size_t)r * cols + c]; }
    template <typename T> const T &at(int r, int c) const { return ((const T *)data)[(size_t)r * cols + c]; }
    template <typename T> T &at(int i) { return ((T *)data)[i]; }
    template <typename T> const T &at(int i) const { return ((const T *)data)[i]; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * cols * elemSize()); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * cols * elemSize()); }

  private:
    void fill(const Scalar &s)
    {
        const int cn = channels();
        const size_t n = total();
        for (size_t i = 0; i < n; ++i)
            for (int c = 0; c < cn; ++c)
            {
                const double v = s.val[c];
                unsigned char *p = data + (i * cn + c) * elemSize1();
                switch (depth())
                {
                case CV_8U: *(unsigned char *)p = (unsigned char)v; break;
                case CV_8S: *(signed char *)p = (signed char)v; break;
                case CV_16U: *(unsigned short *)p = (unsigned short)v; break;
                case CV_16S: *(short *)p = (short)v; break;
                case CV_32S: *(int *)p = (int)v; break;
                case CV_32F: *(float *)p = (float)v; break;
                default: *(double *)p = v; break;
                }
            }
    }
    int type_ = 0;
    std::shared_ptr<unsigned char> buf_;
};
} // namespace cv
#endif
