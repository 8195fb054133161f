// Drop-in definition of optimization::SimpleBA (reference src/Optimization/SimpleBA.cpp:80-157, declared in SimpleBA.h:18),
// which Optimizer::FastBA forwards to (Optimizer.h:22-25): the per-pair normal-equation sums are reduced on the GPU and the small
// pose system is solved on the host, all behind opb_simple_ba.  Replaces SimpleBA.cpp in the build (ComputeJTJAndJTr, its helper,
// has no other caller).  Poses agree with the reference's to ~1e-4 (double accumulation here, float there), not bit for bit.
// NOTE: compiled and linked by tests/cpp/Makefile but not called by the caller test yet -- opb_simple_ba has not run on a GPU.
#include <cstdint>
#include <iostream>
#include <vector>

#include "Optimization/SimpleBA.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace optimization
{
void SimpleBA(const std::vector<Correspondence> &correspondences, geometry::SE3List &camera_poses, int max_iteration)
{
    if (camera_poses.size() < 3)
    {
        std::cout << BLUE << "[INFO]::[SimpleBA]::Too few optimization variables, No need to optimize." << RESET << std::endl;
        return;
    }
    if (correspondences.size() < camera_poses.size() - 1)
    {
        std::cout << RED << "[ERROR]::[SimpleBA]::There are unconnected components." << RESET << std::endl;
        return;
    }
    const int n_poses = (int)camera_poses.size(), n_corr = (int)correspondences.size();
    std::vector<float> poses(16 * (size_t)n_poses), a, b;
    std::vector<int32_t> src(n_corr), tgt(n_corr);
    std::vector<int64_t> offset(n_corr + 1, 0);
    for (int i = 0; i < n_poses; ++i)
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) poses[16 * i + 4 * c + r] = (float)camera_poses[i](r, c);
    for (int k = 0; k < n_corr; ++k)
    {
        src[k] = correspondences[k].source_id;
        tgt[k] = correspondences[k].target_id;
        const auto &set = correspondences[k].correspondence_set;
        offset[k + 1] = offset[k] + (int64_t)set.size();
        for (size_t j = 0; j < set.size(); ++j)
            for (int e = 0; e < 3; ++e)
            {
                a.push_back((float)set[j].first(e));
                b.push_back((float)set[j].second(e));
            }
    }
    if (opb_simple_ba(0, n_poses, poses.data(), n_corr, src.data(), tgt.data(), offset.data(), a.data(), b.data(), max_iteration) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[SimpleBA]::" << opb_last_error() << RESET << std::endl;
        return;
    }
    for (int i = 1; i < n_poses; ++i)
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) camera_poses[i](r, c) = poses[16 * i + 4 * c + r];
}
} // namespace optimization
} // namespace one_piece
