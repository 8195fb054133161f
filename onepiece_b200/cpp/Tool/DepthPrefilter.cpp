// Drop-in definitions of the two one_piece::tool functions that sit directly in front of CubeHandler::IntegrateImage in
// every fusion main of the reference (example/ImageSequenceIntegration.cpp:36-38, example/DenseFusion/DenseFusion.cpp:92-95):
//   tool::BilateralFilter     reference src/Tool/ImageProcessing.cpp:64-67   (cv::bilateralFilter(source, target, range, 0.03, 4.5))
//   tool::ConvertDepthTo32F   reference src/Tool/ImageProcessing.cpp:68-91
// Signatures are the reference's (src/Tool/ImageProcessing.h:19-20, included unchanged); the work is done on the GPU through
// the C-ABI (opb_prefilter_*).  To integrate: delete those two function bodies from the reference's ImageProcessing.cpp (the
// other functions of that file stay) and add this file to the library sources.
#include <cstdlib>
#include <iostream>
#include <map>

#include "Tool/ImageProcessing.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace tool
{
namespace
{
// one device workspace per image size, like the reference's callers reuse their cv::Mat buffers
opb_prefilter *Workspace(int width, int height)
{
    static std::map<std::pair<int, int>, opb_prefilter *> cache;
    auto it = cache.find(std::make_pair(width, height));
    if (it != cache.end()) return it->second;
    opb_prefilter *f = nullptr;
    if (opb_prefilter_create(0, nullptr, width, height, &f) != OPB_OK)
    {
        std::cout << RED << "[ImageProcessing]::[ERROR]::" << opb_last_error() << RESET << std::endl;
        std::exit(1); // no device, no result: there is no CPU path to fall back to
    }
    cache[std::make_pair(width, height)] = f;
    return f;
}
int DepthType(const cv::Mat &depth)
{
    if (depth.depth() == CV_32FC1) return OPB_DEPTH_F32;
    if (depth.depth() == CV_16UC1) return OPB_DEPTH_U16;
    return -1;
}
} // namespace

void BilateralFilter(const cv::Mat &source, cv::Mat &target, int range)
{
    cv::Mat result(source.rows, source.cols, CV_32FC1); // cv::bilateralFilter never works in place
    int rc = opb_prefilter_run(Workspace(source.cols, source.rows), source.data, DepthType(source), 1.0f, range, 0.03, 4.5, nullptr,
                               (float *)result.data);
    if (rc != OPB_OK) std::cout << RED << "[ImageProcessing]::[ERROR]::" << opb_last_error() << RESET << std::endl;
    target = result;
}

void ConvertDepthTo32F(const cv::Mat &depth, cv::Mat &refined_depth, float depth_scale)
{
    refined_depth.create(depth.rows, depth.cols, CV_32FC1);
    int rc = opb_prefilter_run(Workspace(depth.cols, depth.rows), depth.data, DepthType(depth), depth_scale, 7, 0.03, 4.5,
                               (float *)refined_depth.data, nullptr);
    if (rc == OPB_ERR_UNSUPPORTED)
    {
        // ImageProcessing.cpp:86-90
        std::cout << RED << "[ImageProcessing]::[ERROR]::Unknown depth image type: " << depth.depth() << RESET << std::endl;
        std::exit(1);
    }
    if (rc != OPB_OK) std::cout << RED << "[ImageProcessing]::[ERROR]::" << opb_last_error() << RESET << std::endl;
}
} // namespace tool
} // namespace one_piece
