// TEST BUILD ONLY: see cv_imgproc_oracle.h.  Arguments outside what the oracle restates abort loudly.
#include <cstdio>
#include <cstdlib>

extern "C" {
#include "opb_oracle.h"
}
namespace cv
{
static void need(bool ok, const char *what)
{
    if (!ok) { std::fprintf(stderr, "cv_imgproc_oracle: %s is outside the restated filters\n", what); std::abort(); }
}
void pyrDown(const Mat &src, Mat &dst, const Size &size)
{
    need(src.type() == CV_32FC1 && size.width == src.cols / 2 && size.height == src.rows / 2, "pyrDown");
    Mat out(size.height, size.width, CV_32FC1);
    orc_pyr_down((const float *)src.data, src.cols, src.rows, (float *)out.data);
    dst = out;
}
void cvtColor(const Mat &src, Mat &dst, int code)
{
    need(src.type() == CV_8UC3 && code == CV_RGB2GRAY, "cvtColor");
    Mat out(src.rows, src.cols, CV_8UC1);
    orc_gray_u8(src.data, src.rows * src.cols, out.data);
    dst = out;
}
void Sobel(const Mat &src, Mat &dst, int ddepth, int dx, int dy)
{
    need(src.type() == CV_32FC1 && ddepth == CV_32F && dx + dy == 1, "Sobel");
    Mat out(src.rows, src.cols, CV_32FC1);
    orc_sobel3((const float *)src.data, src.cols, src.rows, dx, (float *)out.data);
    dst = out;
}
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigma)
{
    need(src.type() == CV_32FC1 && ksize.width == 3 && ksize.height == 3 && sigma == 0, "GaussianBlur");
    Mat out(src.rows, src.cols, CV_32FC1);
    orc_blur3((const float *)src.data, src.cols, src.rows, (float *)out.data);
    dst = out;
}
void bilateralFilter(const Mat &src, Mat &dst, int d, double sigma_color, double sigma_space)
{
    need(src.type() == CV_32FC1, "bilateralFilter");
    Mat out(src.rows, src.cols, CV_32FC1);
    orc_bilateral_filter((const float *)src.data, src.cols, src.rows, d, sigma_color, sigma_space, (float *)out.data);
    dst = out;
}
} // namespace cv
