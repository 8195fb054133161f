// Drop-in replacement for the DENSE half of the reference's src/Odometry/Odometry.cpp (:436-685): the two
// Odometry::DenseTracking overloads declared in the reference's own src/Odometry/Odometry.h (:78-84), unchanged
// signatures, computed by libonepiece_b200 on the GPU.  A maintainer deletes lines 436-685 of Odometry.cpp (DenseTracking,
// CreateImagePyramid, CreateImageXYZPyramid, InitializeRGBDDenseTracking, MultiScaleComputing -- they become device
// kernels) and compiles this file next to what is left (the sparse ORB path, untouched).  Odometry.h, RGBDFrame.h and
// DenseOdometryFunction.h stay as they are.
//
// geometry::RGBDFrame has no slot for a device handle, so the frame's dense cache lives in a side table keyed by the
// frame's colour buffer (cv::Mat copies of a frame share it, like they share the reference's cached pyramids).  As in
// the reference the cache is filled on first use, `frame.image_xyz` is made non-empty so IsPreprocessedDense() holds,
// and the level-0 intensity is re-normalised in place on every call.  The table keeps the 64 most recently used frames.
#include <cstdint>
#include <iostream>
#include <map>
#include <vector>

#include "Odometry/Odometry.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace odometry
{
namespace
{
struct DeviceOdometry
{
    opb_odometry_desc desc;
    opb_odometry *handle;
};
std::vector<DeviceOdometry> g_odometries; // callers are single-threaded

struct DeviceFrame
{
    opb_frame *handle;
    opb_odometry *owner;
    const void *depth_data;
    uint64_t stamp;
};
std::map<const void *, DeviceFrame> g_frames;
uint64_t g_stamp = 0;
const size_t kMaxCachedFrames = 64;

void Fail(const char *where)
{
    std::cout << RED << "[ERROR]::[" << where << "]::" << opb_last_error() << RESET << std::endl;
}

opb_odometry *Acquire(const camera::PinholeCamera &camera, int levels, const std::vector<int> &iterations)
{
    opb_odometry_desc d;
    opb_odometry_desc_default(&d);
    d.fx = camera.GetFx(); d.fy = camera.GetFy(); d.cx = camera.GetCx(); d.cy = camera.GetCy();
    d.width = camera.GetWidth(); d.height = camera.GetHeight(); d.depth_scale = camera.GetDepthScale();
    d.levels = levels;
    for (int i = 0; i < OPB_ODO_MAX_LEVELS; ++i) d.iterations[i] = i < (int)iterations.size() && i < levels ? iterations[i] : 0;
    for (auto &e : g_odometries)
        if (e.desc.fx == d.fx && e.desc.fy == d.fy && e.desc.cx == d.cx && e.desc.cy == d.cy && e.desc.width == d.width &&
            e.desc.height == d.height && e.desc.depth_scale == d.depth_scale && e.desc.levels == d.levels &&
            std::equal(d.iterations, d.iterations + OPB_ODO_MAX_LEVELS, e.desc.iterations))
            return e.handle;
    opb_odometry *h = nullptr;
    if (opb_odometry_create(&d, &h) != OPB_OK)
    {
        Fail("DenseTracking");
        std::exit(1); // no device / bad camera: there is no CPU path
    }
    g_odometries.push_back({d, h});
    return h;
}

int DepthType(const cv::Mat &depth)
{
    if (depth.depth() == CV_32FC1) return OPB_DEPTH_F32;
    if (depth.depth() == CV_16UC1) return OPB_DEPTH_U16;
    return -1;
}

opb_frame *DeviceFrameOf(opb_odometry *o, geometry::RGBDFrame &frame, int levels)
{
    auto it = g_frames.find(frame.rgb.data);
    if (it != g_frames.end() && (it->second.owner != o || it->second.depth_data != frame.depth.data || !frame.IsPreprocessedDense()))
    {   // same buffer, other images / other camera / frame was reset: start over
        opb_frame_destroy(it->second.handle);
        g_frames.erase(it);
        it = g_frames.end();
    }
    if (it == g_frames.end())
    {
        if (g_frames.size() >= kMaxCachedFrames)
        {
            auto oldest = g_frames.begin();
            for (auto k = g_frames.begin(); k != g_frames.end(); ++k)
                if (k->second.stamp < oldest->second.stamp) oldest = k;
            opb_frame_destroy(oldest->second.handle);
            g_frames.erase(oldest);
        }
        opb_frame *h = nullptr;
        if (opb_frame_create(o, frame.rgb.data, frame.depth.data, DepthType(frame.depth), &h) != OPB_OK)
        {
            // ConvertDepthTo32FNaN: message + exit(1) on an unknown depth type (DenseOdometryFunction.cpp:51-55)
            Fail("ImageProcessing");
            std::exit(1);
        }
        it = g_frames.insert(std::make_pair((const void *)frame.rgb.data, DeviceFrame{h, o, frame.depth.data, 0})).first;
    }
    it->second.stamp = ++g_stamp;
    if (!frame.IsPreprocessedDense()) frame.image_xyz.resize(levels); // the dense cache now exists (on the device)
    return it->second.handle;
}

std::shared_ptr<DenseTrackingResult> Assemble(const opb_tracking_result &r, const std::vector<uint32_t> &pairs, const std::vector<float> &xyz)
{
    DenseTrackingResult out;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row) out.T(row, c) = (geometry::scalar)r.T[c * 4 + row];
    const size_t n = r.n_correspondences;
    out.pixel_correspondence_set.resize(n);
    out.correspondence_set.resize(n);
    for (size_t k = 0; k < n; ++k)
    {
        out.pixel_correspondence_set[k] = std::make_pair(geometry::Point2ui(pairs[4 * k], pairs[4 * k + 1]), geometry::Point2ui(pairs[4 * k + 2], pairs[4 * k + 3]));
        out.correspondence_set[k] = std::make_pair(geometry::Point3(xyz[6 * k], xyz[6 * k + 1], xyz[6 * k + 2]),
                                                   geometry::Point3(xyz[6 * k + 3], xyz[6 * k + 4], xyz[6 * k + 5]));
    }
    out.rmse = r.rmse;
    out.tracking_success = r.tracking_success != 0;
    return std::make_shared<DenseTrackingResult>(out);
}

void DebugTerm(int term_type)
{
#if DEBUG_MODE
    if (term_type == 0) std::cout << BLUE << "[DEBUG]::Using hybrid term" << RESET << std::endl;
    else if (term_type == 1) std::cout << BLUE << "[DEBUG]::Using photo term" << RESET << std::endl;
    else if (term_type == 2) std::cout << BLUE << "[DEBUG]::Using geometry term" << RESET << std::endl;
#endif
}
} // namespace

std::shared_ptr<DenseTrackingResult> Odometry::DenseTracking(const cv::Mat &source_color, const cv::Mat &target_color, const cv::Mat &source_depth,
                                                             const cv::Mat &target_depth, const geometry::TransformationMatrix &initial_T,
                                                             int term_type)
{
    DebugTerm(term_type);
    opb_odometry *o = Acquire(camera, multi_scale_level, iter_count_per_level);
    const size_t npx = (size_t)camera.GetWidth() * camera.GetHeight();
    std::vector<uint32_t> pairs(npx * 4);
    std::vector<float> xyz(npx * 6);
    Eigen::Matrix4f T0 = initial_T.cast<float>();
    opb_tracking_result r;
    if (opb_odometry_dense_tracking(o, source_color.data, target_color.data, source_depth.data, target_depth.data, DepthType(source_depth),
                                    T0.data(), term_type, &r, pairs.data(), npx, xyz.data()) != OPB_OK)
    {
        Fail("DenseTracking");
        if (DepthType(source_depth) < 0) std::exit(1); // the reference exits on an unknown depth type
        return std::make_shared<DenseTrackingResult>(DenseTrackingResult());
    }
    return Assemble(r, pairs, xyz);
}

std::shared_ptr<DenseTrackingResult> Odometry::DenseTracking(geometry::RGBDFrame &source_frame, geometry::RGBDFrame &target_frame,
                                                             const geometry::TransformationMatrix &initial_T, int term_type)
{
    DebugTerm(term_type);
    opb_odometry *o = Acquire(camera, multi_scale_level, iter_count_per_level);
    opb_frame *s = DeviceFrameOf(o, source_frame, multi_scale_level);
    opb_frame *t = DeviceFrameOf(o, target_frame, multi_scale_level);
    const size_t npx = (size_t)camera.GetWidth() * camera.GetHeight();
    std::vector<uint32_t> pairs(npx * 4);
    std::vector<float> xyz(npx * 6);
    Eigen::Matrix4f T0 = initial_T.cast<float>();
    opb_tracking_result r;
    if (opb_odometry_dense_tracking_frames(o, s, t, T0.data(), term_type, &r, pairs.data(), npx, xyz.data()) != OPB_OK)
    {
        Fail("DenseTracking");
        return std::make_shared<DenseTrackingResult>(DenseTrackingResult());
    }
    return Assemble(r, pairs, xyz);
}
} // namespace odometry
} // namespace one_piece
