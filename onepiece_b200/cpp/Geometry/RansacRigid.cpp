// Drop-in definition of geometry::EstimateRigidTransformationRANSAC (reference src/Geometry/Ransac.cpp:7-41, declared in
// src/Geometry/Ransac.h:11-12) over the C-ABI: the max_iteration eight-point hypotheses of the reference's GRANSAC estimator are
// scored in parallel on the GPU (opb_ransac_rigid_transformation) and the first strictly best one wins, as in GRANSAC.hpp:113-122.
// Like the reference, every call draws fresh samples (the reference seeds its engines from std::random_device); setting
// OPB_RANSAC_SEED pins them.  `inliers` / `inlier_ids` come back in ascending pair order (the reference's are in its shuffled order).
// FitPlaneRANSAC, the other function of Ransac.cpp, is not on the path and stays with the reference: delete only this body there,
// or link this object first with -Wl,--allow-multiple-definition as tests/cpp/Makefile does.
#include <cstdlib>
#include <iostream>
#include <random>
#include <vector>

#include "Geometry/Ransac.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace geometry
{
TransformationMatrix EstimateRigidTransformationRANSAC(const PointCorrespondenceSet &correspondence_set, PointCorrespondenceSet &inliers,
                                                       std::vector<int> &inlier_ids, int max_iteration, float threshold)
{
    if (correspondence_set.size() < MIN_INLIER_SIZE_RANSAC_TRANSFORMATION)
    {
        std::cout << YELLOW << "[Warning]::[FitPlaneRANSAC]::Too few canidate point pair." << RESET << std::endl;
        return TransformationMatrix::Zero();
    }
    static opb_kdtree *ws = nullptr; // callers are single-threaded; lends its stream and buffers
    if (!ws && opb_kdtree_create(0, &ws) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[RANSAC]::" << opb_last_error() << RESET << std::endl;
        std::exit(1); // no device: there is no CPU path
    }
    const size_t n = correspondence_set.size();
    std::vector<float> a(3 * n), b(3 * n);
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k)
        {
            a[3 * i + k] = (float)correspondence_set[i].first(k);
            b[3 * i + k] = (float)correspondence_set[i].second(k);
        }
    const char *pinned = std::getenv("OPB_RANSAC_SEED");
    const uint64_t seed = pinned ? std::strtoull(pinned, nullptr, 10) : ((uint64_t)std::random_device()() << 32) ^ std::random_device()();
    float T[16];
    std::vector<int32_t> ids(n);
    size_t m = 0;
    if (opb_ransac_rigid_transformation(ws, a.data(), b.data(), n, max_iteration, (double)threshold, seed, nullptr, T, ids.data(), &m, nullptr, nullptr) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[RANSAC]::" << opb_last_error() << RESET << std::endl;
        return TransformationMatrix::Zero();
    }
    for (size_t i = 0; i < m; ++i)
    {
        inliers.push_back(correspondence_set[ids[i]]);
        inlier_ids.push_back(ids[i]);
    }
    TransformationMatrix out;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out(r, c) = T[c * 4 + r];
    return out;
}
} // namespace geometry
} // namespace one_piece
