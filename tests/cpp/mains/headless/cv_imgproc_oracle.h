// TEST BUILD ONLY (force-included into the REFERENCE-side build of the example mains): declarations of the five OpenCV imgproc
// functions src/Tool/ImageProcessing.cpp calls, so that the reference's own translation unit compiles unmodified without
// OpenCV.  Their bodies (cv_imgproc_oracle.cpp) run the oracle's restatements of those filters (oracle/opb_oracle.c, pinned to
// cv2 outputs: tests/test_oracle_filters.py).  Never part of the product and never linked into the drop-in build.
#ifndef OPB_CV_IMGPROC_ORACLE_H
#define OPB_CV_IMGPROC_ORACLE_H
#include <opencv2/opencv.hpp>
#ifndef CV_RGB2GRAY
#define CV_RGB2GRAY 7
#endif
namespace cv
{
void pyrDown(const Mat &src, Mat &dst, const Size &size);
void cvtColor(const Mat &src, Mat &dst, int code);
void Sobel(const Mat &src, Mat &dst, int ddepth, int dx, int dy);
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigma);
void bilateralFilter(const Mat &src, Mat &dst, int d, double sigma_color, double sigma_space);
} // namespace cv
#endif
