// Drop-in definition of registration::ComputeFPFHFeature (reference src/Registration/3DFeature.cpp:83-131, declared in
// src/Registration/3DFeature.h:24) computed on the GPU through the C-ABI (opb_kdtree_build + opb_kdtree_fpfh).  The device builds
// the reference's own k-d tree and visits it in the reference's order, so the 33-bin descriptors are the reference's bit for
// bit -- including which 2.5 * knn points the early-stopping radius search gets to see.  Replaces 3DFeature.cpp in the build
// (ComputePairDescriptor / ComputeSPFH, its two helpers, have no other caller).
#include <cstdlib>
#include <iostream>
#include <vector>

#include "Registration/3DFeature.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace registration
{
void ComputeFPFHFeature(const geometry::PointCloud &pcd, FeatureSet &fpfh_features, int knn, float radius)
{
    static opb_kdtree *ws = nullptr; // callers are single-threaded; one search workspace per process
    if (!ws && opb_kdtree_create(0, &ws) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[FPFHFeature]::" << opb_last_error() << RESET << std::endl;
        std::exit(1); // no device: there is no CPU path
    }
    const size_t n = pcd.points.size();
    std::vector<float> p(n * 3), nr(n * 3), f(n * 33);
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k)
        {
            p[3 * i + k] = (float)pcd.points[i](k);
            nr[3 * i + k] = (float)pcd.normals[i](k);
        }
    fpfh_features.clear();
    if (opb_kdtree_build(ws, p.data(), n) != OPB_OK || opb_kdtree_fpfh(ws, nr.data(), knn, radius, f.data()) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[FPFHFeature]::" << opb_last_error() << RESET << std::endl;
        return;
    }
    Feature zero;
    zero.resize(33);
    fpfh_features.resize(n, zero);
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 33; ++k) fpfh_features[i](k) = f[33 * i + k];
}
} // namespace registration
} // namespace one_piece
