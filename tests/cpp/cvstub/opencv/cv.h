// TEST BUILD ONLY.  The reference's src/Odometry/Odometry.h pulls in ORB / BFMatcher / MILD declarations (its sparse
// tracking path) through <opencv/cv.h>.  OpenCV C++ is not installed in this image, so the caller-style test binary
// (tests/cpp/dropin_main.cpp) parses that header against these empty stand-ins.  Nothing here computes anything and
// none of it is part of the product: a maintainer builds against real OpenCV.
#ifndef OPB_TEST_CV_STUB_H
#define OPB_TEST_CV_STUB_H
#include <opencv2/opencv.hpp>
#include <vector>
typedef unsigned char uchar;
#define CV_FM_RANSAC 8
namespace cv
{
enum { NORM_HAMMING = 6 };
template <typename T> struct Ptr
{
    std::shared_ptr<T> p;
    T *operator->() const { return p.get(); }
};
struct ORB
{
    static Ptr<ORB> create() { Ptr<ORB> p; p.p = std::make_shared<ORB>(); return p; }
    void setMaxFeatures(int) {}
    template <typename... A> void detectAndCompute(A &&...) const {}
    template <typename... A> void detect(A &&...) const {}
    template <typename... A> void compute(A &&...) const {}
};
struct BFMatcher
{
    BFMatcher(int = 0, bool = false) {}
    template <typename... A> void knnMatch(A &&...) const {}
    template <typename... A> void match(A &&...) const {}
};
template <typename... A> Mat findHomography(A &&...) { return Mat(); }
// window / drawing calls of the sparse path's debug output (Odometry.cpp:125-131,154-157, ...): never reached by the dense path
struct NoArray {};
inline NoArray noArray() { return NoArray(); }
struct DrawMatchesFlags { enum { DEFAULT = 0 }; };
template <typename... A> void drawKeypoints(A &&...) {}
template <typename... A> void drawMatches(A &&...) {}
template <typename... A> void imshow(A &&...) {}
template <typename... A> void destroyWindow(A &&...) {}
inline int waitKey(int = 0) { return -1; }
} // namespace cv
#endif
