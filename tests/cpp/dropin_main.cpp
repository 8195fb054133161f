// Reads like example/ImageSequenceIntegration.cpp:20-53, example/ICPTest.cpp:14-34 and example/DenseOdometry.cpp:14-31 of the
// reference, minus
// file IO and the viewer: the reference's caller code compiled UNCHANGED against the drop-in classes of
// onepiece_b200/cpp.  Inputs come from raw binary files written by tests/test_dropin_cpp.py; results are written
// back as raw binary for comparison with the oracle.
#include <fstream>
#include <iostream>
#include <vector>

#include "Geometry/PointCloud.h"
#include "Geometry/TriangleMesh.h"
#include "Geometry/RGBDFrame.h"
#include "Integration/CubeHandler.h"
#include "Odometry/Odometry.h"
#include "Registration/3DFeature.h"
#include "Registration/GlobalRegistration.h"
#include "Registration/ICP.h"
#include "Tool/ImageProcessing.h"

using namespace one_piece;

template <typename T> std::vector<T> ReadAll(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    f.seekg(0, f.end);
    size_t n = f.tellg();
    f.seekg(0, f.beg);
    std::vector<T> v(n / sizeof(T));
    f.read((char *)v.data(), n);
    return v;
}
template <typename T> void WriteAll(const std::string &path, const std::vector<T> &v)
{
    std::ofstream f(path, std::ios::binary);
    f.write((const char *)v.data(), v.size() * sizeof(T));
}

int main(int argc, char **argv)
{
    if (argc < 3) { std::cout << "usage: dropin_main <dir> <n_frames>" << std::endl; return 2; }
    const std::string dir = argv[1];
    const int n_frames = atoi(argv[2]);
    const int W = 160, H = 120;
    camera::PinholeCamera camera(514.817f / 4, 515.375f / 4, 318.771f / 4, 238.447f / 4, W, H, 1000.0f);

    // --- example/ImageSequenceIntegration.cpp:20-53 ---------------------------------------------------------
    integration::CubeHandler cube_handler(camera);
    cube_handler.SetVoxelResolution(0.02);
    std::vector<float> poses = ReadAll<float>(dir + "/poses.bin"); // row-major 4x4 per frame, like trajectory.txt
    for (int i = 0; i < n_frames; ++i)
    {
        std::vector<float> d = ReadAll<float>(dir + "/depth" + std::to_string(i) + ".bin");
        std::vector<unsigned char> c = ReadAll<unsigned char>(dir + "/bgr" + std::to_string(i) + ".bin");
        cv::Mat depth(H, W, CV_32FC1, d.data()), rgb(H, W, CV_8UC3, c.data());
        geometry::TransformationMatrix pose;
        for (int r = 0; r < 4; ++r)
            for (int col = 0; col < 4; ++col) pose(r, col) = poses[16 * i + 4 * r + col];
        cube_handler.IntegrateImage(depth, rgb, pose);
    }
    geometry::TriangleMesh mesh;
    cube_handler.ExtractTriangleMesh(mesh);
    std::vector<float> mesh_out;
    for (size_t i = 0; i < mesh.points.size(); ++i)
        for (int k = 0; k < 3; ++k) mesh_out.push_back(mesh.points[i](k));
    WriteAll(dir + "/mesh_points.bin", mesh_out);
    std::vector<float> mesh_colors;
    for (size_t i = 0; i < mesh.colors.size(); ++i)
        for (int k = 0; k < 3; ++k) mesh_colors.push_back(mesh.colors[i](k));
    WriteAll(dir + "/mesh_colors.bin", mesh_colors);
    integration::CubeMap m = cube_handler.GetCubeMap();
    std::vector<float> vox;
    std::vector<int> ids;
    for (auto it = m.begin(); it != m.end(); ++it)
    {
        for (int k = 0; k < 3; ++k) ids.push_back(it->first(k));
        for (int j = 0; j < 512; ++j)
        {
            vox.push_back(it->second.voxels[j].sdf);
            vox.push_back(it->second.voxels[j].weight);
            for (int k = 0; k < 3; ++k) vox.push_back(it->second.voxels[j].color(k));
        }
    }
    WriteAll(dir + "/ids.bin", ids);
    WriteAll(dir + "/voxels.bin", vox);
    cube_handler.WriteToFile(dir + "/volume.cubes");
    // example/ImageSequenceIntegration.cpp:48-53: resample the volume into the frame of the middle pose, mesh the result
    {
        geometry::TransformationMatrix mid;
        for (int r = 0; r < 4; ++r)
            for (int col = 0; col < 4; ++col) mid(r, col) = poses[16 * (n_frames / 2) + 4 * r + col];
        auto transformed_cube_handler = cube_handler.TransformNearest(mid);
        geometry::TriangleMesh tmesh;
        transformed_cube_handler->ExtractTriangleMesh(tmesh);
        // example/MergeMultipleSubmaps.cpp:40-41
        integration::CubeHandler merged(camera);
        merged.SetVoxelResolution(0.02);
        auto transformed_handler = cube_handler.Transform(mid);
        merged.Merge(*transformed_handler);
        merged.Merge(cube_handler);
        std::vector<double> out;
        out.push_back((double)transformed_cube_handler->GetCubeMap().size());
        out.push_back((double)tmesh.points.size());
        out.push_back((double)transformed_handler->GetCubeMap().size());
        out.push_back((double)merged.GetCubeMap().size());
        double wsum = 0;
        integration::CubeMap mm = merged.GetCubeMap();
        for (auto it = mm.begin(); it != mm.end(); ++it)
            for (int j = 0; j < 512; ++j) wsum += it->second.voxels[j].weight;
        out.push_back(wsum);
        WriteAll(dir + "/resample.bin", out);
    }
    // example/DenseFusion/DenseFusion.cpp:99-105 and example/MergeMultipleSubmaps.cpp:45-46: simplify the extracted mesh, normals,
    // then a down-sampled point cloud of its vertices (example/ReadPLYPointCloud.cpp:24)
    {
        auto c_mesh = mesh.ClusteringSimplify(0.02);
        if (!c_mesh->HasNormals()) c_mesh->ComputeNormals();
        auto pcd = c_mesh->GetPointCloud()->DownSample(0.05);
        std::vector<float> post;
        post.push_back((float)c_mesh->points.size());
        post.push_back((float)c_mesh->triangles.size());
        post.push_back((float)pcd->points.size());
        for (size_t i = 0; i < c_mesh->points.size(); ++i)
            for (int k = 0; k < 3; ++k) { post.push_back(c_mesh->points[i](k)); post.push_back(c_mesh->normals[i](k)); }
        for (size_t i = 0; i < pcd->points.size(); ++i)
            for (int k = 0; k < 3; ++k) { post.push_back(pcd->points[i](k)); post.push_back(pcd->normals[i](k)); post.push_back(pcd->colors[i](k)); }
        WriteAll(dir + "/meshpost.bin", post);
    }

    // --- example/ICPTest.cpp:14-34 ---------------------------------------------------------------------------
    geometry::PointCloud s_pcd, t_pcd;
    std::vector<float> s = ReadAll<float>(dir + "/src.bin"), t = ReadAll<float>(dir + "/tgt.bin"), n = ReadAll<float>(dir + "/nrm.bin");
    for (size_t i = 0; i < s.size() / 3; ++i) s_pcd.points.push_back(geometry::Point3(s[3 * i], s[3 * i + 1], s[3 * i + 2]));
    for (size_t i = 0; i < t.size() / 3; ++i)
    {
        t_pcd.points.push_back(geometry::Point3(t[3 * i], t[3 * i + 1], t[3 * i + 2]));
        t_pcd.normals.push_back(geometry::Point3(n[3 * i], n[3 * i + 1], n[3 * i + 2]));
    }
    // example/ICPTest.cpp:27-29: normals from the cloud itself when none are given (checked separately; the registration below
    // keeps the analytic ones)
    {
        geometry::PointCloud e_pcd;
        e_pcd.points = t_pcd.points;
        e_pcd.EstimateNormals();
        std::vector<float> en;
        for (size_t i = 0; i < e_pcd.normals.size(); ++i)
            for (int k = 0; k < 3; ++k) en.push_back(e_pcd.normals[i](k));
        WriteAll(dir + "/estimated_normals.bin", en);
        // the submap back end's descriptor step (example/DenseFusion/DenseSlam.cpp:76 -> DownSampleAndExtractFeature):
        // down-sample, normals, FPFH with DenseSlam's parameters (DenseSlam.h:49-56)
        auto down = e_pcd.DownSample(0.05);
        down->EstimateNormals(0.1, 30);
        registration::FeatureSet features;
        registration::ComputeFPFHFeature(*down, features, 100, 0.25);
        std::vector<float> ff, fp;
        for (size_t i = 0; i < features.size(); ++i)
        {
            for (int k = 0; k < 33; ++k) ff.push_back(features[i](k));
            for (int k = 0; k < 3; ++k) fp.push_back(down->points[i](k));
        }
        WriteAll(dir + "/fpfh.bin", ff);
        WriteAll(dir + "/fpfh_points.bin", fp);
        // the rest of registration::RansacRegistration on precomputed features, statement for statement
        // (GlobalRegistration.cpp:225-266): matching, the three rejection passes and the RANSAC estimator all through the drop-ins
        // (the estimator draws fresh samples per call like the reference's: its T is checked against the true motion)
        geometry::PointCloud s_full;
        s_full.points = s_pcd.points;
        auto s_down = s_full.DownSample(0.05);
        s_down->EstimateNormals(0.1, 30);
        registration::FeatureSet s_features;
        registration::ComputeFPFHFeature(*s_down, s_features, 100, 0.25);
        geometry::FMatchSet matches;
        registration::FeatureMatching3D(s_features, features, matches);
        std::vector<int32_t> m0, m3;
        for (size_t i = 0; i < matches.size(); ++i) { m0.push_back(matches[i].first); m0.push_back(matches[i].second); }
        std::default_random_engine engine;
        registration::RejectMatchesRanSaPC(s_down->points, down->points, engine, matches);
        registration::RejectMatchesRanSaPC(s_down->points, down->points, engine, matches);
        registration::RejectMatchesRanSaPC(s_down->points, down->points, engine, matches);
        for (size_t i = 0; i < matches.size(); ++i) { m3.push_back(matches[i].first); m3.push_back(matches[i].second); }
        geometry::PointCorrespondenceSet correspondence_set, inliers;
        for (size_t i = 0; i != matches.size(); ++i)
            correspondence_set.push_back(std::make_pair(s_down->points[matches[i].first], down->points[matches[i].second]));
        std::vector<int> inlier_ids;
        auto T = geometry::EstimateRigidTransformationRANSAC(correspondence_set, inliers, inlier_ids, 2000, 0.1);
        std::vector<double> rr;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) rr.push_back(T(r, c));
        rr.push_back((double)inliers.size());
        rr.push_back((double)correspondence_set.size());
        for (size_t i = 0; i < inlier_ids.size(); ++i) rr.push_back((double)inlier_ids[i]);
        std::vector<float> sp;
        for (size_t i = 0; i < s_down->points.size(); ++i)
            for (int k = 0; k < 3; ++k) sp.push_back(s_down->points[i](k));
        WriteAll(dir + "/ransac.bin", rr);
        WriteAll(dir + "/matches_initial.bin", m0);
        WriteAll(dir + "/matches_kept.bin", m3);
        WriteAll(dir + "/fpfh_source_points.bin", sp);
    }
    registration::ICPParameter icp_para;
    icp_para.threshold = 0.05;
    icp_para.max_iteration = 10;
    auto result = registration::PointToPlane(s_pcd, t_pcd, geometry::TransformationMatrix::Identity(), icp_para);
    std::vector<double> icp_out;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) icp_out.push_back(result->T(r, c));
    icp_out.push_back(result->rmse);
    icp_out.push_back((double)result->correspondence_set_index.size());
    WriteAll(dir + "/icp.bin", icp_out);
    // --- example/DenseOdometry.cpp:14-31 (+ the chained use of example/DenseFusion/DenseSlam.cpp:22) ---------
    odometry::Odometry rgbd_odometry(camera);
    std::vector<double> odo_out;
    {
        std::vector<unsigned short> sd = ReadAll<unsigned short>(dir + "/odo_depth1.bin"), td = ReadAll<unsigned short>(dir + "/odo_depth0.bin");
        std::vector<unsigned char> sc = ReadAll<unsigned char>(dir + "/odo_bgr1.bin"), tc = ReadAll<unsigned char>(dir + "/odo_bgr0.bin");
        cv::Mat source_rgb(H, W, CV_8UC3, sc.data()), source_depth(H, W, CV_16UC1, sd.data());
        cv::Mat target_rgb(H, W, CV_8UC3, tc.data()), target_depth(H, W, CV_16UC1, td.data());
        geometry::RGBDFrame source_frame(source_rgb, source_depth);
        geometry::RGBDFrame target_frame(target_rgb, target_depth);
        geometry::Matrix4 T = geometry::Matrix4::Identity();
        for (int call = 0; call < 2; ++call) // the second call re-normalises the cached frames, like the reference
        {
            auto tracking_result = rgbd_odometry.DenseTracking(source_frame, target_frame, T, 0);
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) odo_out.push_back(tracking_result->T(r, c));
            odo_out.push_back(tracking_result->rmse);
            odo_out.push_back((double)tracking_result->pixel_correspondence_set.size());
            odo_out.push_back(tracking_result->tracking_success ? 1.0 : 0.0);
            odo_out.push_back((double)tracking_result->correspondence_set.size());
            odo_out.push_back(source_frame.IsPreprocessedDense() && target_frame.IsPreprocessedDense() ? 1.0 : 0.0);
        }
        auto mat_result = rgbd_odometry.DenseTracking(source_rgb, target_rgb, source_depth, target_depth, T, 0);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) odo_out.push_back(mat_result->T(r, c));
        odo_out.push_back(mat_result->rmse);
        odo_out.push_back((double)mat_result->pixel_correspondence_set.size());
    }
    WriteAll(dir + "/odometry.bin", odo_out);
    // --- example/ImageSequenceIntegration.cpp:33-40: the depth pre-filter in front of IntegrateImage ------------
    {
        std::vector<unsigned short> raw = ReadAll<unsigned short>(dir + "/odo_depth1.bin");
        std::vector<unsigned char> c = ReadAll<unsigned char>(dir + "/odo_bgr1.bin");
        cv::Mat depth(H, W, CV_16UC1, raw.data()), rgb(H, W, CV_8UC3, c.data());
        cv::Mat refined_depth, filtered_depth;
        tool::ConvertDepthTo32F(depth, refined_depth, camera.GetDepthScale());
        tool::BilateralFilter(refined_depth, filtered_depth);
        integration::CubeHandler filtered_handler(camera);
        filtered_handler.SetVoxelResolution(0.02);
        filtered_handler.IntegrateImage(filtered_depth, rgb, geometry::TransformationMatrix::Identity());
        WriteAll(dir + "/refined_depth.bin", std::vector<float>((float *)refined_depth.data, (float *)refined_depth.data + W * H));
        WriteAll(dir + "/filtered_depth.bin", std::vector<float>((float *)filtered_depth.data, (float *)filtered_depth.data + W * H));
        std::vector<double> n_cubes(1, (double)filtered_handler.GetCubeMap().size());
        WriteAll(dir + "/filtered_cubes.bin", n_cubes);
    }
    std::cout << "dropin ok: " << m.size() << " cubes, " << mesh.triangles.size() << " triangles, "
              << result->correspondence_set_index.size() << " ICP inliers, " << (size_t)odo_out[17] << " odometry correspondences" << std::endl;
    return 0;
}
