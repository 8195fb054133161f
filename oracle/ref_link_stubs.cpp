// TEST INFRASTRUCTURE ONLY.  Link-time stand-ins for the two tool:: functions that
// src/Odometry/DenseOdometryFunction.cpp references but whose home translation unit
// (src/Tool/ImageProcessing.cpp) needs real OpenCV imgproc and therefore cannot be compiled here.
#include <cstdlib>
#include <iostream>
#include <opencv2/opencv.hpp>
#ifdef USING_FLOAT64
#include "Geometry/Geometry.h"
#include "Geometry/PointCloud.h"
#endif
namespace one_piece
{
namespace tool
{
// ImageProcessing.cpp:21-24 is cv::cvtColor(CV_RGB2GRAY); the oracle's odometry driver supplies gray images
// through its own explicit restatement of that filter, so this symbol must never be reached.
void Convert2Gray(const cv::Mat &, cv::Mat &)
{
    std::cerr << "oracle: tool::Convert2Gray stub reached" << std::endl;
    std::abort();
}
// ImageProcessing.cpp:56-63 semantics: in-place a*x+b on a CV_32FC1 image
void LinearTransform(cv::Mat &source, float scale, float offset)
{
    for (int i = 0; i != source.rows; ++i)
        for (int j = 0; j != source.cols; ++j) source.at<float>(i, j) = source.at<float>(i, j) * scale + offset;
}
} // namespace tool
#ifdef USING_FLOAT64
// PointCloud.cpp does not compile with -DUSING_FLOAT64 (KDTree.h:237 passes float* for double*); the only
// symbols of it the float64 oracle needs are these two (PointCloud.cpp:72-100, 238-243 semantics).
namespace geometry
{
void PointCloud::Transform(const TransformationMatrix &T)
{
    geometry::TransformPoints(T, points);
    if (HasNormals()) geometry::TransformNormals(T, normals);
}
void PointCloud::LoadFromDepth(const cv::Mat &depth, const camera::PinholeCamera &camera)
{
    Reset();
    float fx = camera.GetFx(), fy = camera.GetFy(), cx = camera.GetCx(), cy = camera.GetCy();
    float depth_scale = camera.GetDepthScale();
    points.resize(depth.rows * depth.cols);
    int cnt = 0;
    for (int i = 0; i < depth.rows; i++)
        for (int j = 0; j < depth.cols; j++)
        {
            float z = depth.depth() == CV_32FC1 ? depth.at<float>(i, j) : depth.at<unsigned short>(i, j) / depth_scale;
            if (z > 0) points[cnt++] = geometry::Point3((j - cx) * z / fx, (i - cy) * z / fy, z);
        }
    points.resize(cnt);
}
} // namespace geometry
#endif
} // namespace one_piece
