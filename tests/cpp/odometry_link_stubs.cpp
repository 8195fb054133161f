// TEST BUILD ONLY: link stand-ins for the sparse-tracking members that one_piece::odometry::Odometry's inline
// constructors reference (MILD's SparseMatcher lives in 3rdparty/MILD and needs real OpenCV).  Never called.
#include "Odometry/Odometry.h"
namespace MILD
{
SparseMatcher::SparseMatcher(int, int, int, float) {}
SparseMatcher::~SparseMatcher() {}
// referenced by the reference's src/Odometry/SparseMatcher.cpp (linked into the reference-side build of the example mains)
void SparseMatcher::train(cv::Mat) {}
void SparseMatcher::search_8(cv::Mat, std::vector<cv::DMatch> &, int) {}
void SparseMatcher::search_8_with_range(cv::Mat, std::vector<cv::DMatch> &, const std::vector<cv::KeyPoint> &, const std::vector<cv::KeyPoint> &, float,
                                        int)
{
}
} // namespace MILD
