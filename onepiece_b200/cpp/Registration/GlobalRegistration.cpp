// Drop-in for the whole-call entry points of the reference's src/Registration/GlobalRegistration.cpp:
//   registration::RansacRegistration(source, target, r_para)                                  (:120-206)
//   registration::DownSampleAndExtractFeature(pcd, r_para)                                    (:207-218)
//   registration::RansacRegistration(source_feature_pcd, target_feature_pcd, features..., r_para)  (:219-268)
// DenseSlam::RegisterSubmap (example/DenseFusion/DenseSlam.cpp:69-125) calls the last two.  Every step they are made of runs on
// the GPU through libonepiece_b200.so: PointCloud::DownSample / EstimateNormals (cpp/Geometry/MeshPost.cpp), ComputeFPFHFeature
// (cpp/Registration/Feature3D.cpp), FeatureMatching3D / RejectMatchesRanSaPC (cpp/Registration/FeatureMatching.cpp),
// geometry::EstimateRigidTransformationRANSAC (cpp/Geometry/RansacRigid.cpp).  This file only sequences them the way the
// reference does and assembles RegistrationResult; it replaces the reference's translation unit on the link line (the
// reference's FeatureMatching3D and RejectMatchesRanSaPC live in that same unit, so it cannot be linked next to this one).
#include "Registration/GlobalRegistration.h"

#include <cmath>
#include <random>

namespace one_piece
{
namespace registration
{
namespace
{
// sqrt(sum |R a + t - b|^2 / n) over the inlier pairs, accumulated in float like the reference's ComputeRMSE (:6-14)
float InlierRmse(const geometry::PointCorrespondenceSet &pairs, const geometry::TransformationMatrix &T)
{
    float sum = 0.0;
    for (const auto &p : pairs) sum += (T.block<3, 3>(0, 0) * p.first + T.block<3, 1>(0, 3) - p.second).squaredNorm();
    return sqrt(sum / pairs.size());
}

// descriptor matching, three rejection passes with one engine, RANSAC over the surviving pairs (common tail of both overloads);
// pair_scale divides the matched points before the estimation (the first overload un-scales them), threshold is already scaled
std::shared_ptr<RegistrationResult> MatchAndEstimate(const geometry::Point3List &source_points, const geometry::Point3List &target_points,
                                                     const FeatureSet &source_features, const FeatureSet &target_features, double pair_scale,
                                                     int max_iteration, double threshold)
{
    geometry::FMatchSet matches;
    FeatureMatching3D(source_features, target_features, matches);
    std::default_random_engine engine; // default seed; the three passes continue one sequence
    for (int pass = 0; pass < 3; ++pass) RejectMatchesRanSaPC(source_points, target_points, engine, matches);
    geometry::PointCorrespondenceSet pairs;
    pairs.reserve(matches.size());
    for (const auto &m : matches) pairs.push_back(std::make_pair(source_points[m.first], target_points[m.second]));
    if (pair_scale != 1)
        for (auto &p : pairs) { p.first = p.first / pair_scale; p.second = p.second / pair_scale; }
    auto result = std::make_shared<RegistrationResult>();
    std::vector<int> inlier_ids;
    result->T = geometry::EstimateRigidTransformationRANSAC(pairs, result->correspondence_set, inlier_ids, max_iteration, threshold);
    result->rmse = InlierRmse(result->correspondence_set, result->T);
    for (int id : inlier_ids) result->correspondence_set_index.push_back(matches[id]);
    return result;
}
} // namespace

std::tuple<geometry::PointCloud, FeatureSet> DownSampleAndExtractFeature(const geometry::PointCloud &pcd, const RANSACParameter &r_para)
{
    auto down = pcd.DownSample(r_para.voxel_len);
    if (!down->HasNormals()) down->EstimateNormals(r_para.search_radius_normal, r_para.max_nn_normal);
    FeatureSet features;
    ComputeFPFHFeature(*down, features, r_para.max_nn, r_para.search_radius);
    return std::make_tuple(*down, features);
}

std::shared_ptr<RegistrationResult> RansacRegistration(const geometry::PointCloud &source_feature_pcd, const geometry::PointCloud &target_feature_pcd,
                                                       const FeatureSet &source_features, const FeatureSet &target_features,
                                                       const RANSACParameter &r_para)
{
    return MatchAndEstimate(source_feature_pcd.points, target_feature_pcd.points, source_features, target_features, 1.0, r_para.max_iteration,
                            r_para.threshold);
}

std::shared_ptr<RegistrationResult> RansacRegistration(const geometry::PointCloud &source_pcd, const geometry::PointCloud &target_pcd,
                                                       const RANSACParameter &r_para)
{
    geometry::PointCloud source = source_pcd, target = target_pcd;
    if (r_para.scaling != 1)
    {
        for (auto &p : source.points) p = p * r_para.scaling;
        for (auto &p : target.points) p = p * r_para.scaling;
    }
    geometry::PointCloud s_down, t_down;
    FeatureSet s_feat, t_feat;
    std::tie(s_down, s_feat) = DownSampleAndExtractFeature(source, r_para);
    std::tie(t_down, t_feat) = DownSampleAndExtractFeature(target, r_para);
    return MatchAndEstimate(s_down.points, t_down.points, s_feat, t_feat, r_para.scaling, r_para.max_iteration, r_para.threshold / r_para.scaling);
}
} // namespace registration
} // namespace one_piece
