// TEST BUILD ONLY (force-included): the one OpenCV highgui function the reference's example mains call.  The image files of
// the synthetic datasets are raw containers written by tests/test_reference_mains.py: "OPBIMG\0\0", int32 rows, cols, type, data.
#ifndef OPB_HEADLESS_CV_H
#define OPB_HEADLESS_CV_H
#include <opencv2/opencv.hpp>
#include <string>
namespace cv
{
Mat imread(const std::string &path, int flags = 1);
}
#endif
