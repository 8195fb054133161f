// Drop-in replacement for the reference's src/Registration/ICP.cpp: the two free functions declared in the
// reference's own src/Registration/ICP.h (:23-26) with unchanged signatures, computed by libonepiece_b200 on
// the GPU.  Compile this file instead of src/Registration/ICP.cpp; ICP.h and RegistrationResult.h stay as they are.
#include <iostream>

#include "Registration/ICP.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace registration
{
namespace
{
opb_icp *Workspace()
{
    static opb_icp *ws = nullptr; // callers are single-threaded; one workspace per process
    if (!ws && opb_icp_create(0, nullptr, &ws) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[ICP]::" << opb_last_error() << RESET << std::endl;
        std::exit(1); // no device: there is no CPU path
    }
    return ws;
}
void Flatten(const geometry::Point3List &pts, std::vector<float> &out)
{
    out.resize(pts.size() * 3);
    for (size_t i = 0; i < pts.size(); ++i)
        for (int k = 0; k < 3; ++k) out[3 * i + k] = (float)pts[i](k);
}
std::shared_ptr<RegistrationResult> Run(const geometry::PointCloud &source, const geometry::PointCloud &target,
                                        const geometry::TransformationMatrix &init_T, const ICPParameter &icp_para, bool plane)
{
    RegistrationResult result;
    if (plane && (!target.HasNormals() || icp_para.scaling != 1))
    {   // ICP.cpp:159-163
        std::cout << RED << "[ERROR]::[ICPPointToPlane]::target point cloud need to have normals." << RESET << std::endl;
        return std::make_shared<RegistrationResult>(RegistrationResult());
    }
    std::vector<float> s, t, n;
    Flatten(source.points, s);
    Flatten(target.points, t);
    if (plane) Flatten(target.normals, n);
    opb_icp_params par;
    par.max_iteration = icp_para.max_iteration;
    par.threshold = icp_para.threshold;
    par.scaling = icp_para.scaling;
    Eigen::Matrix4f T0 = init_T.cast<float>();
    opb_icp_result r;
    std::vector<int32_t> pairs(source.points.size() * 2 + 2);
    int rc = plane ? opb_icp_point_to_plane(Workspace(), s.data(), source.points.size(), t.data(), n.data(), target.points.size(),
                                            T0.data(), &par, &r, pairs.data(), source.points.size())
                   : opb_icp_point_to_point(Workspace(), s.data(), source.points.size(), t.data(), target.points.size(), T0.data(),
                                            &par, &r, pairs.data(), source.points.size());
    if (rc != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[ICP]::" << opb_last_error() << RESET << std::endl;
        return std::make_shared<RegistrationResult>(RegistrationResult());
    }
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) result.T(rr, c) = r.T[c * 4 + rr];
    result.rmse = r.rmse;
    for (size_t i = 0; i < r.n_inliers; ++i)
    {
        result.correspondence_set_index.push_back(std::make_pair(pairs[2 * i], pairs[2 * i + 1]));
        result.correspondence_set.push_back(std::make_pair(source.points[pairs[2 * i]], target.points[pairs[2 * i + 1]]));
    }
    return std::make_shared<RegistrationResult>(result);
}
} // namespace

std::shared_ptr<RegistrationResult> PointToPoint(const geometry::PointCloud &source, const geometry::PointCloud &target,
                                                 const geometry::TransformationMatrix &init_T, const ICPParameter &icp_para)
{
    return Run(source, target, init_T, icp_para, false);
}
std::shared_ptr<RegistrationResult> PointToPlane(const geometry::PointCloud &source, const geometry::PointCloud &target,
                                                 const geometry::TransformationMatrix &init_T, const ICPParameter &icp_para)
{
    return Run(source, target, init_T, icp_para, true);
}
} // namespace registration
} // namespace one_piece
