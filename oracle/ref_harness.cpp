// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C entry points around the UNMODIFIED reference translation units (compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libopref_{f32,f64}.so).  Used by tests/ and by
// bench.py's cpu_baseline / --impl reference legs to (1) pin the C restatement in oracle/opb_oracle.c,
// (2) generate tests/golden/*.npz and (3) time the reference's own CPU path.
//
// Reference entry points exercised:
//   integration::CubeHandler::{IntegrateImage,PrepareCubes,ComputeBounding,ExtractTriangleMesh,GetCubeMap}
//       src/Integration/CubeHandler.cpp:9-44,116-214, CubeHandler.h:36,141,339,349-356
//   registration::{PointToPlane,PointToPoint,EstimateRigidTransformationPointToPlane,CountInliers}
//       src/Registration/ICP.cpp:9-224
//   geometry::{TransformPoints,EstimateRigidTransformation,Se3ToSE3}  src/Geometry/Geometry.cpp:9-27,107-151
//   geometry::PointCloud::{LoadFromDepth,EstimateNormals}             src/Geometry/PointCloud.cpp:72-144
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "Camera/Camera.h"
#include "Geometry/Geometry.h"
#include "Geometry/KDTree.h"
#include "Geometry/PointCloud.h"
#include "Geometry/TriangleMesh.h"
#include "Integration/CubeHandler.h"
#include "Integration/Frustum.h"
#include "Integration/MarchingCube.h"
#include "Registration/ICP.h"
#ifndef USING_FLOAT64
#include "Registration/3DFeature.h"
#include "Registration/GlobalRegistration.h"
#include "Geometry/Ransac.h"
#include "Optimization/SimpleBA.h"
#endif

using namespace one_piece;
typedef geometry::scalar scalar;

namespace one_piece
{
namespace registration
{
// external linkage in ICP.cpp:9 but not declared in ICP.h
double CountInliers(const geometry::Point3List &source, const geometry::Point3List target,
                    const std::vector<int> &correspondence_index, const geometry::TransformationMatrix &T,
                    double threshold, geometry::FMatchSet &inliers);
} // namespace registration
} // namespace one_piece

namespace
{
struct RefVolume
{
    integration::CubeHandler handler;
    camera::PinholeCamera cam;
    geometry::TriangleMesh mesh;
    // CubeHandler keeps camera/c_para protected; expose what the harness needs
};
// access to protected members for read-only inspection (bounding box, cube list)
struct HandlerPeek : public integration::CubeHandler
{
    const integration::CubeMap &map() const { return cube_map; }
    const integration::CubePara &para() const { return c_para; }
};

geometry::TransformationMatrix PoseFromColMajor(const float *p)
{
    geometry::TransformationMatrix T;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) T(r, c) = (scalar)p[c * 4 + r];
    return T;
}
void PoseToColMajor(const geometry::TransformationMatrix &T, double *p)
{
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) p[c * 4 + r] = (double)T(r, c);
}
cv::Mat WrapDepth(const void *depth, int is_u16, int w, int h)
{
    return cv::Mat(h, w, is_u16 ? CV_16UC1 : CV_32FC1, const_cast<void *>(depth));
}
cv::Mat WrapBgr(const uint8_t *bgr, int w, int h) { return cv::Mat(h, w, CV_8UC3, const_cast<uint8_t *>(bgr)); }
void ToList(const float *xyz, size_t n, geometry::Point3List &out)
{
    out.resize(n);
    for (size_t i = 0; i < n; ++i) out[i] = geometry::Point3((scalar)xyz[3 * i], (scalar)xyz[3 * i + 1], (scalar)xyz[3 * i + 2]);
}
double Now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace

extern "C"
{
int ref_scalar_bytes() { return (int)sizeof(scalar); }
// the reference prints a line per call (DEBUG_MODE 1); keep test logs readable
void ref_set_quiet(int quiet)
{
    if (quiet) std::cout.setstate(std::ios_base::failbit);
    else std::cout.clear();
}

void *ref_volume_create(float fx, float fy, float cx, float cy, int w, int h, float depth_scale, float voxel_res,
                        float truncation, float near_plane, float far_plane)
{
    RefVolume *v = new RefVolume();
    v->cam = camera::PinholeCamera(fx, fy, cx, cy, w, h, depth_scale);
    v->handler.SetCamera(v->cam);
    v->handler.SetVoxelResolution(voxel_res);
    v->handler.SetTruncation(truncation);
    v->handler.SetNearPlane(near_plane);
    v->handler.SetFarPlane(far_plane);
    return v;
}
void ref_volume_destroy(void *h) { delete (RefVolume *)h; }
void ref_volume_clear(void *h) { ((RefVolume *)h)->handler.Clear(); }

// CubeHandler::IntegrateImage (CubeHandler.cpp:197-210). Returns seconds spent inside the call.
double ref_volume_integrate(void *h, const void *depth, int is_u16, const uint8_t *bgr, const float *pose_cm)
{
    RefVolume *v = (RefVolume *)h;
    int w = v->cam.GetWidth(), hh = v->cam.GetHeight();
    cv::Mat d = WrapDepth(depth, is_u16, w, hh), c = WrapBgr(bgr, w, hh);
    geometry::TransformationMatrix pose = PoseFromColMajor(pose_cm);
    double t0 = Now();
    v->handler.IntegrateImage(d, c, pose);
    return Now() - t0;
}
// CubeHandler::ComputeBounding (CubeHandler.cpp:116-145)
void ref_volume_bounding(void *h, const void *depth, int is_u16, const float *pose_cm, double *max_pos, double *min_pos)
{
    RefVolume *v = (RefVolume *)h;
    int w = v->cam.GetWidth(), hh = v->cam.GetHeight();
    cv::Mat d = WrapDepth(depth, is_u16, w, hh);
    geometry::Point3 mx, mn;
    v->handler.ComputeBounding(d, PoseFromColMajor(pose_cm), mx, mn);
    for (int i = 0; i < 3; ++i) { max_pos[i] = mx(i); min_pos[i] = mn(i); }
}
// CubeHandler::PrepareCubes (CubeHandler.cpp:147-196): allocates absent cubes and returns the frame's list
long ref_volume_prepare_cubes(void *h, const void *depth, int is_u16, const float *pose_cm, int32_t *ids, long cap)
{
    RefVolume *v = (RefVolume *)h;
    int w = v->cam.GetWidth(), hh = v->cam.GetHeight();
    cv::Mat d = WrapDepth(depth, is_u16, w, hh);
    std::vector<integration::CubeID> list;
    v->handler.PrepareCubes(d, PoseFromColMajor(pose_cm), list);
    long n = (long)list.size();
    for (long i = 0; i < n && i < cap; ++i)
        for (int k = 0; k < 3; ++k) ids[3 * i + k] = list[i](k);
    return n;
}
long ref_volume_num_cubes(void *h)
{
    return (long)static_cast<HandlerPeek *>(&((RefVolume *)h)->handler)->map().size();
}
// ids: n*3 int32; voxels: n*512*5 float (sdf, weight, c0, c1, c2) in the reference's voxel order x+8y+64z
void ref_volume_download(void *h, int32_t *ids, float *voxels)
{
    const integration::CubeMap &m = static_cast<HandlerPeek *>(&((RefVolume *)h)->handler)->map();
    size_t i = 0;
    for (auto it = m.begin(); it != m.end(); ++it, ++i)
    {
        for (int k = 0; k < 3; ++k) ids[3 * i + k] = it->first(k);
        float *dst = voxels + i * 512 * 5;
        for (int j = 0; j < 512; ++j)
        {
            const integration::TSDFVoxel &vx = it->second.voxels[j];
            dst[5 * j + 0] = (float)vx.sdf;
            dst[5 * j + 1] = (float)vx.weight;
            dst[5 * j + 2] = (float)vx.color(0);
            dst[5 * j + 3] = (float)vx.color(1);
            dst[5 * j + 4] = (float)vx.color(2);
        }
    }
}
// replaces the whole map (CubeHandler::SetCubeMap, CubeHandler.h:344)
void ref_volume_upload(void *h, const int32_t *ids, const float *voxels, long n)
{
    integration::CubeMap m;
    for (long i = 0; i < n; ++i)
    {
        integration::CubeID id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]);
        integration::VoxelCube c(id);
        const float *src = voxels + (size_t)i * 512 * 5;
        for (int j = 0; j < 512; ++j)
        {
            c.voxels[j].sdf = src[5 * j];
            c.voxels[j].weight = src[5 * j + 1];
            c.voxels[j].color = geometry::Point3(src[5 * j + 2], src[5 * j + 3], src[5 * j + 4]);
        }
        m[id] = c;
    }
    bool was_quiet = std::cout.fail();
    std::cout.setstate(std::ios_base::failbit);
    ((RefVolume *)h)->handler.SetCubeMap(m);
    if (!was_quiet) std::cout.clear();
}
// CubeHandler::Transform (CubeHandler.h:242-298) / TransformNearest (:299-338): returns a new handle holding the result
void *ref_volume_transform(void *h, const float *trans_cm, int nearest)
{
    RefVolume *v = (RefVolume *)h;
    geometry::TransformationMatrix T = PoseFromColMajor(trans_cm);
    std::shared_ptr<integration::CubeHandler> r = nearest ? v->handler.TransformNearest(T) : v->handler.Transform(T);
    RefVolume *out = new RefVolume();
    out->cam = v->cam;
    out->handler = *r;
    return out;
}
// CubeHandler::Merge(another) (CubeHandler.h:145-167)
void ref_volume_merge(void *h, void *other) { ((RefVolume *)h)->handler.Merge(((RefVolume *)other)->handler); }
// CubeHandler::Merge(another, trans) (CubeHandler.h:168-177)
void ref_volume_merge_transformed(void *h, void *other, const float *trans_cm)
{
    ((RefVolume *)h)->handler.Merge(((RefVolume *)other)->handler, PoseFromColMajor(trans_cm));
}
float ref_volume_resolution(void *h) { return (float)static_cast<HandlerPeek *>(&((RefVolume *)h)->handler)->para().VoxelResolution; }
// TriangleMesh::ClusteringSimplify (TriangleMesh.cpp:53-58) on caller-supplied arrays; results copied back (buffers must hold
// the input sizes).  with_normals != 0: the mesh gets ComputeNormals() first, so CompactMesh recomputes them (normals out).
// Returns seconds spent in ClusteringSimplify.
double ref_clustering_simplify(float *points, float *colors, long *n_points, uint32_t *tri, long *n_tris, float grid_len, int with_normals,
                               float *normals)
{
    geometry::TriangleMesh mesh;
    mesh.points.resize(*n_points);
    if (colors) mesh.colors.resize(*n_points);
    for (long i = 0; i < *n_points; ++i)
    {
        mesh.points[i] = geometry::Point3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
        if (colors) mesh.colors[i] = geometry::Point3(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]);
    }
    mesh.triangles.resize(*n_tris);
    for (long i = 0; i < *n_tris; ++i) mesh.triangles[i] = geometry::Point3ui(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
    if (with_normals) mesh.ComputeNormals();
    double t0 = Now();
    auto out = mesh.ClusteringSimplify(grid_len);
    double dt = Now() - t0;
    *n_points = (long)out->points.size();
    *n_tris = (long)out->triangles.size();
    for (long i = 0; i < *n_points; ++i)
        for (int k = 0; k < 3; ++k)
        {
            points[3 * i + k] = (float)out->points[i](k);
            if (colors) colors[3 * i + k] = (float)out->colors[i](k);
            if (with_normals && normals) normals[3 * i + k] = (float)out->normals[i](k);
        }
    for (long i = 0; i < *n_tris; ++i)
        for (int k = 0; k < 3; ++k) tri[3 * i + k] = out->triangles[i](k);
    return dt;
}
#ifndef USING_FLOAT64
// PointCloud::DownSample (PointCloud.cpp:145-189); outputs hold n entries, returns the number written
long ref_downsample(const float *points, const float *colors, const float *normals, long n, float grid_len, float *out_points,
                    float *out_colors, float *out_normals)
{
    geometry::PointCloud pcd;
    pcd.points.resize(n);
    if (colors) pcd.colors.resize(n);
    if (normals) pcd.normals.resize(n);
    for (long i = 0; i < n; ++i)
    {
        pcd.points[i] = geometry::Point3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
        if (colors) pcd.colors[i] = geometry::Point3(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]);
        if (normals) pcd.normals[i] = geometry::Point3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
    }
    auto out = pcd.DownSample(grid_len);
    const long m = (long)out->points.size();
    for (long i = 0; i < m; ++i)
        for (int k = 0; k < 3; ++k)
        {
            out_points[3 * i + k] = (float)out->points[i](k);
            if (colors) out_colors[3 * i + k] = (float)out->colors[i](k);
            if (normals) out_normals[3 * i + k] = (float)out->normals[i](k);
        }
    return m;
}
#endif
// TriangleMesh::WriteToPLY (TriangleMesh.cpp:128-131 -> tool::WritePLY, PLYManager.cpp:188-276)
bool ref_write_ply(const char *path, const float *points, const float *normals, const float *colors, long n_points, const uint32_t *tri, long n_tris)
{
    geometry::TriangleMesh mesh;
    mesh.points.resize(n_points);
    if (normals) mesh.normals.resize(n_points);
    if (colors) mesh.colors.resize(n_points);
    for (long i = 0; i < n_points; ++i)
    {
        mesh.points[i] = geometry::Point3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
        if (normals) mesh.normals[i] = geometry::Point3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        if (colors) mesh.colors[i] = geometry::Point3(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]);
    }
    mesh.triangles.resize(n_tris);
    for (long i = 0; i < n_tris; ++i) mesh.triangles[i] = geometry::Point3ui(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
    return mesh.WriteToPLY(path);
}
// TriangleMesh::ComputeNormals (TriangleMesh.cpp:95-127)
void ref_compute_normals(const float *points, long n_points, const uint32_t *tri, long n_tris, float *normals)
{
    geometry::TriangleMesh mesh;
    mesh.points.resize(n_points);
    for (long i = 0; i < n_points; ++i) mesh.points[i] = geometry::Point3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    mesh.triangles.resize(n_tris);
    for (long i = 0; i < n_tris; ++i) mesh.triangles[i] = geometry::Point3ui(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
    mesh.ComputeNormals();
    for (long i = 0; i < n_points; ++i)
        for (int k = 0; k < 3; ++k) normals[3 * i + k] = (float)mesh.normals[i](k);
}
// CubeHandler::ExtractTriangleMesh (CubeHandler.cpp:9-44). Returns seconds; sizes through out params.
double ref_volume_extract_mesh(void *h, long *n_points, long *n_triangles)
{
    RefVolume *v = (RefVolume *)h;
    v->mesh.Reset();
    double t0 = Now();
    v->handler.ExtractTriangleMesh(v->mesh);
    double dt = Now() - t0;
    *n_points = (long)v->mesh.points.size();
    *n_triangles = (long)v->mesh.triangles.size();
    return dt;
}
void ref_volume_mesh_copy(void *h, float *points, float *colors, uint32_t *triangles)
{
    RefVolume *v = (RefVolume *)h;
    for (size_t i = 0; i < v->mesh.points.size(); ++i)
        for (int k = 0; k < 3; ++k)
        {
            points[3 * i + k] = (float)v->mesh.points[i](k);
            colors[3 * i + k] = (float)v->mesh.colors[i](k);
        }
    for (size_t i = 0; i < v->mesh.triangles.size(); ++i)
        for (int k = 0; k < 3; ++k) triangles[3 * i + k] = v->mesh.triangles[i](k);
}
bool ref_volume_write(void *h, const char *path) { return ((RefVolume *)h)->handler.WriteToFile(path); }
bool ref_volume_read(void *h, const char *path) { return ((RefVolume *)h)->handler.ReadFromFile(path); }


// integration::MarchingCube on one cell (MarchingCube.cpp:31-74): corners 8x3, sdf 8, colors 8x3.
// Writes up to 15 vertices (xyz, rgb); returns the vertex count (3 per triangle).
int ref_marching_cube_cell(const float *corners, const float *sdf, const float *colors, float *out_xyz, float *out_rgb)
{
    geometry::Point3List c(8);
    std::vector<integration::TSDFVoxel> v(8);
    for (int i = 0; i < 8; ++i)
    {
        c[i] = geometry::Point3(corners[3 * i], corners[3 * i + 1], corners[3 * i + 2]);
        v[i] = integration::TSDFVoxel(sdf[i], 1.0f, geometry::Point3(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]));
    }
    geometry::TriangleMesh m;
    integration::MarchingCube(c, v, m);
    for (size_t i = 0; i < m.points.size(); ++i)
        for (int k = 0; k < 3; ++k)
        {
            out_xyz[3 * i + k] = (float)m.points[i](k);
            out_rgb[3 * i + k] = (float)m.colors[i](k);
        }
    return (int)m.points.size();
}
// integration::Frustum::{ComputeFromCamera,ContainPoint} (Frustum.cpp:7-25, Frustum.h:74-103).
// planes: 6x4 in the order top,left,right,bottom,near,far; mask[i] = ContainPoint(points[i]).
void ref_frustum(float fx, float fy, float cx, float cy, int w, int h, const float *pose_cm, float far_d, float near_d,
                 double *planes, const float *points, long n, uint8_t *mask)
{
    camera::PinholeCamera cam(fx, fy, cx, cy, w, h, 1000.0f);
    integration::Frustum f;
    f.ComputeFromCamera(cam, PoseFromColMajor(pose_cm), far_d, near_d);
    const geometry::Plane *pl[6] = {&f.top_plane, &f.left_plane, &f.right_plane, &f.bottom_plane, &f.near_plane, &f.far_plane};
    for (int i = 0; i < 6; ++i)
        for (int k = 0; k < 4; ++k) planes[4 * i + k] = (*pl[i])(k);
    for (long i = 0; i < n; ++i)
        mask[i] = f.ContainPoint(geometry::Point3(points[3 * i], points[3 * i + 1], points[3 * i + 2])) ? 1 : 0;
}
// Eigen's Matrix4 inverse as used at Integrator.cpp:18,48
void ref_pose_inverse(const float *pose_cm, double *inv_cm)
{
    geometry::TransformationMatrix inv = PoseFromColMajor(pose_cm).inverse();
    PoseToColMajor(inv, inv_cm);
}
// Integrator::GetSDF (Integrator.cpp:8-35) at explicit points
void ref_get_sdf(void *h, const void *depth, int is_u16, const float *pose_cm, const float *points, long n, float *sdf)
{
    RefVolume *v = (RefVolume *)h;
    int w = v->cam.GetWidth(), hh = v->cam.GetHeight();
    cv::Mat d = WrapDepth(depth, is_u16, w, hh);
    integration::Integrator integ;
    geometry::TransformationMatrix pose = PoseFromColMajor(pose_cm);
    for (long i = 0; i < n; ++i)
        sdf[i] = integ.GetSDF(geometry::Point3(points[3 * i], points[3 * i + 1], points[3 * i + 2]), v->cam, pose, d);
}

// ---------------------------------------------------------------------------------------------
// Registration
// ---------------------------------------------------------------------------------------------
// registration::PointToPlane / PointToPoint called exactly as example/ICPTest.cpp:33 does.
// tgt_nrm == NULL -> PointToPoint.  out_T column-major double[16]; pairs: 2*n int32 (source, target).
// Returns seconds spent in the call; *n_pairs = -1 if the reference returned its default (error) result.
double ref_icp(const float *src, long ns, const float *tgt, const float *tgt_nrm, long nt, const float *init_T_cm,
               int max_iter, double threshold, double *out_T_cm, int32_t *pairs, long pairs_cap, long *n_pairs,
               double *rmse)
{
    geometry::PointCloud s, t;
    ToList(src, ns, s.points);
    ToList(tgt, nt, t.points);
    if (tgt_nrm) ToList(tgt_nrm, nt, t.normals);
    registration::ICPParameter para;
    para.max_iteration = max_iter;
    para.threshold = threshold;
    geometry::TransformationMatrix T0 = PoseFromColMajor(init_T_cm);
    double t0 = Now();
    std::shared_ptr<registration::RegistrationResult> r =
        tgt_nrm ? registration::PointToPlane(s, t, T0, para) : registration::PointToPoint(s, t, T0, para);
    double dt = Now() - t0;
    PoseToColMajor(r->T, out_T_cm);
    *rmse = r->rmse;
    long n = (long)r->correspondence_set_index.size();
    *n_pairs = n;
    for (long i = 0; i < n && i < pairs_cap; ++i)
    {
        pairs[2 * i] = r->correspondence_set_index[i].first;
        pairs[2 * i + 1] = r->correspondence_set_index[i].second;
    }
    return dt;
}

// One loop body of PointToPlane (ICP.cpp:175-205) for teacher-forced comparison: given T_in, return the
// nearest-neighbour indices, the inlier count, the 6x6 system the reference accumulates (recomputed with the
// same statement order as ICP.cpp:121-136) and T_out = Se3ToSE3(x) * T_in.
struct RefIcpState
{
    geometry::PointCloud source, target;
    geometry::KDTree<> kdtree;
};
void *ref_icp_state_create(const float *src, long ns, const float *tgt, const float *tgt_nrm, long nt)
{
    RefIcpState *st = new RefIcpState();
    ToList(src, ns, st->source.points);
    ToList(tgt, nt, st->target.points);
    if (tgt_nrm) ToList(tgt_nrm, nt, st->target.normals);
    st->kdtree.BuildTree(st->target.points);
    return st;
}
void ref_icp_state_destroy(void *p) { delete (RefIcpState *)p; }
long ref_icp_iteration(void *p, const double *T_in_cm, double threshold, int32_t *nn_index, double *JTJ36,
                       double *JTr6, double *x6, double *T_out_cm, double *rmse)
{
    RefIcpState *st = (RefIcpState *)p;
    geometry::TransformationMatrix T;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) T(r, c) = (scalar)T_in_cm[c * 4 + r];
    std::vector<int> corr(st->source.points.size(), -1);
    geometry::Point3List transformed = st->source.points;
    geometry::TransformPoints(T, transformed);
#pragma omp parallel for
    for (size_t i = 0; i < transformed.size(); ++i)
    {
        std::vector<int> indices(1);
        std::vector<float> dists(1);
        st->kdtree.KnnSearch(transformed[i], indices, dists, 1, geometry::SearchParameter(128));
        if (indices.size() > 0) corr[i] = indices[0];
    }
    geometry::FMatchSet inliers;
    *rmse = registration::CountInliers(st->source.points, st->target.points, corr, T, threshold, inliers);
    for (size_t i = 0; i < corr.size(); ++i) nn_index[i] = corr[i];
    if (st->target.HasNormals())
    {
        geometry::Matrix6 JTJ = geometry::Matrix6::Zero();
        geometry::Se3 JTr = geometry::Se3::Zero();
        for (size_t i = 0; i != inliers.size(); ++i)
        {
            int sid = inliers[i].first, tid = inliers[i].second;
            geometry::Se3 row;
            double r = (st->target.normals[tid].transpose() * transformed[sid] -
                        st->target.normals[tid].transpose() * st->target.points[tid])(0);
            row.block<3, 1>(0, 0) = st->target.normals[tid];
            row.block<3, 1>(3, 0) = transformed[sid].cross(st->target.normals[tid]);
            JTJ.noalias() += row * row.transpose();
            JTr.noalias() += r * row;
        }
        Eigen::JacobiSVD<geometry::MatrixX> svd(JTJ, Eigen::ComputeThinU | Eigen::ComputeThinV);
        geometry::Se3 x = svd.solve(-JTr);
        for (int i = 0; i < 36; ++i) JTJ36[i] = JTJ(i / 6, i % 6);
        for (int i = 0; i < 6; ++i) { JTr6[i] = JTr(i); x6[i] = x(i); }
        geometry::TransformationMatrix dT = registration::EstimateRigidTransformationPointToPlane(
            transformed, st->target.points, st->target.normals, inliers);
        PoseToColMajor(dT * T, T_out_cm);
    }
    else
    {
        geometry::PointCorrespondenceSet cs;
        for (size_t i = 0; i != inliers.size(); ++i)
            cs.push_back(std::make_pair(transformed[inliers[i].first], st->target.points[inliers[i].second]));
        geometry::TransformationMatrix dT = geometry::EstimateRigidTransformation(cs);
        PoseToColMajor(dT * T, T_out_cm);
    }
    return (long)inliers.size();
}

// geometry::Se3ToSE3 (Geometry.cpp:9-13)
void ref_se3_exp(const double *x6, double *T_cm)
{
    geometry::Se3 x;
    for (int i = 0; i < 6; ++i) x(i) = (scalar)x6[i];
    PoseToColMajor(geometry::Se3ToSE3(x), T_cm);
}
// geometry::EstimateRigidTransformation (Geometry.cpp:107-151) on explicit pairs
void ref_kabsch(const float *a, const float *b, long n, double *T_cm)
{
    geometry::PointCorrespondenceSet cs(n);
    for (long i = 0; i < n; ++i)
    {
        cs[i].first = geometry::Point3((scalar)a[3 * i], (scalar)a[3 * i + 1], (scalar)a[3 * i + 2]);
        cs[i].second = geometry::Point3((scalar)b[3 * i], (scalar)b[3 * i + 1], (scalar)b[3 * i + 2]);
    }
    PoseToColMajor(geometry::EstimateRigidTransformation(cs), T_cm);
}
#ifndef USING_FLOAT64
// PointCloud::LoadFromDepth (PointCloud.cpp:72-100); returns the number of points written
long ref_load_from_depth(const void *depth, int is_u16, int w, int h, float fx, float fy, float cx, float cy,
                         float depth_scale, float *xyz)
{
    camera::PinholeCamera cam(fx, fy, cx, cy, w, h, depth_scale);
    geometry::PointCloud pcd;
    pcd.LoadFromDepth(WrapDepth(depth, is_u16, w, h), cam);
    for (size_t i = 0; i < pcd.points.size(); ++i)
        for (int k = 0; k < 3; ++k) xyz[3 * i + k] = pcd.points[i](k);
    return (long)pcd.points.size();
}
// PointCloud::EstimateNormals (PointCloud.cpp:102-144)
double ref_estimate_normals(const float *xyz, long n, float radius, int knn, float *normals)
{
    geometry::PointCloud pcd;
    ToList(xyz, n, pcd.points);
    double t0 = Now();
    pcd.EstimateNormals(radius, knn);
    double dt = Now() - t0;
    for (long i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) normals[3 * i + k] = pcd.normals[i](k);
    return dt;
}
// geometry::KDTree<3>::KnnSearch / RadiusSearch (KDTree.h:93-196) on a tree built over `xyz`, queried with
// `queries`: out_index/out_dist are nq rows of `cap` entries (-1 padded), out_count the entries returned
void ref_kdtree_search(const float *xyz, long n, const float *queries, long nq, int mode, int k, float radius, long cap,
                       int32_t *out_index, float *out_dist, int32_t *out_count)
{
    geometry::Point3List pts;
    ToList(xyz, n, pts);
    geometry::KDTree<> tree;
    tree.BuildTree(pts);
    for (long q = 0; q < nq; ++q)
    {
        std::vector<int> indices;
        std::vector<float> dists;
        geometry::Point3 query(queries[3 * q], queries[3 * q + 1], queries[3 * q + 2]);
        if (mode == 0)
            tree.KnnSearch(query, indices, dists, k, geometry::SearchParameter(1024));
        else if (mode == 1)
            tree.RadiusSearch(query, indices, dists, radius, (size_t)k, geometry::SearchParameter(1024));
        else
            tree.KnnRadiusSearch(query, indices, dists, k, radius, geometry::SearchParameter(1024));
        out_count[q] = (int32_t)indices.size();
        for (long j = 0; j < cap; ++j)
        {
            out_index[q * cap + j] = j < (long)indices.size() ? indices[j] : -1;
            out_dist[q * cap + j] = j < (long)dists.size() ? dists[j] : -1.0f;
        }
    }
}
// registration::ComputeFPFHFeature (3DFeature.cpp:83-131): n rows of 33 floats
double ref_fpfh(const float *xyz, const float *normals, long n, int knn, float radius, float *features)
{
    geometry::PointCloud pcd;
    ToList(xyz, n, pcd.points);
    ToList(normals, n, pcd.normals);
    registration::FeatureSet fs;
    double t0 = Now();
    registration::ComputeFPFHFeature(pcd, fs, knn, radius);
    double dt = Now() - t0;
    for (long i = 0; i < n; ++i)
        for (int k = 0; k < 33; ++k) features[33 * i + k] = fs[i](k);
    return dt;
}
// registration::FeatureMatching3D (GlobalRegistration.cpp:29-73): nearest target feature (33-D) of every source feature
long ref_feature_matching(const float *src_feat, long ns, const float *tgt_feat, long nt, int32_t *pairs)
{
    registration::FeatureSet sf(ns), tf(nt);
    for (long i = 0; i < ns; ++i) { sf[i].resize(33); for (int k = 0; k < 33; ++k) sf[i](k) = src_feat[33 * i + k]; }
    for (long i = 0; i < nt; ++i) { tf[i].resize(33); for (int k = 0; k < 33; ++k) tf[i](k) = tgt_feat[33 * i + k]; }
    geometry::FMatchSet m;
    registration::FeatureMatching3D(sf, tf, m);
    for (size_t i = 0; i < m.size(); ++i) { pairs[2 * i] = m[i].first; pairs[2 * i + 1] = m[i].second; }
    return (long)m.size();
}
// registration::RejectMatchesRanSaPC (GlobalRegistration.cpp:75-108), `rounds` calls on one default-seeded engine as in
// RansacRegistration (:168-172,236-240); matches in/out as (source, target) index pairs
long ref_reject_matches(const float *src, long ns, const float *tgt, long nt, int32_t *pairs, long n, int rounds, int candidate_num,
                        float difference)
{
    geometry::Point3List s, t;
    ToList(src, ns, s);
    ToList(tgt, nt, t);
    geometry::FMatchSet m(n);
    for (long i = 0; i < n; ++i) m[i] = std::make_pair(pairs[2 * i], pairs[2 * i + 1]);
    std::default_random_engine engine;
    for (int r = 0; r < rounds; ++r) registration::RejectMatchesRanSaPC(s, t, engine, m, candidate_num, difference);
    for (size_t i = 0; i < m.size(); ++i) { pairs[2 * i] = m[i].first; pairs[2 * i + 1] = m[i].second; }
    return (long)m.size();
}
// one RANSAC hypothesis of geometry::EstimateRigidTransformationRANSAC (Ransac.cpp:7-41): TransformationModel over the eight
// sampled pairs (TransformationModel.hpp:54-75), Evaluate over all pairs in the given order (:77-94) -> inlier fraction, flags
double ref_ransac_hypothesis(const float *a, const float *b, long n, const int32_t *sample8, double threshold, uint8_t *inlier)
{
    std::vector<std::shared_ptr<GRANSAC::AbstractParameter>> all, sample;
    for (long i = 0; i < n; ++i)
        all.push_back(std::make_shared<geometry::Point3fPair>(geometry::Point3(a[3 * i], a[3 * i + 1], a[3 * i + 2]),
                                                              geometry::Point3(b[3 * i], b[3 * i + 1], b[3 * i + 2]), (int)i));
    for (int k = 0; k < 8; ++k) sample.push_back(all[sample8[k]]);
    geometry::TransformationModel model(sample);
    auto eval = model.Evaluate(all, threshold);
    for (long i = 0; i < n; ++i) inlier[i] = 0;
    for (auto &p : eval.second) inlier[std::dynamic_pointer_cast<geometry::Point3fPair>(p)->id] = 1;
    return eval.first;
}
// registration::RansacRegistration on precomputed features (GlobalRegistration.cpp:219-267); seeded from std::random_device
// inside GRANSAC, so two runs differ: for statistics and timing only
double ref_ransac_registration(const float *src, long ns, const float *tgt, long nt, const float *src_feat, const float *tgt_feat,
                               int max_iteration, double threshold, double *T_cm, long *n_inliers, double *rmse)
{
    geometry::PointCloud sp, tp;
    ToList(src, ns, sp.points);
    ToList(tgt, nt, tp.points);
    registration::FeatureSet sf(ns), tf(nt);
    for (long i = 0; i < ns; ++i) { sf[i].resize(33); for (int k = 0; k < 33; ++k) sf[i](k) = src_feat[33 * i + k]; }
    for (long i = 0; i < nt; ++i) { tf[i].resize(33); for (int k = 0; k < 33; ++k) tf[i](k) = tgt_feat[33 * i + k]; }
    registration::RANSACParameter para;
    para.max_iteration = max_iteration;
    para.threshold = threshold;
    double t0 = Now();
    auto r = registration::RansacRegistration(sp, tp, sf, tf, para);
    double dt = Now() - t0;
    PoseToColMajor(r->T, T_cm);
    *n_inliers = (long)r->correspondence_set.size();
    *rmse = r->rmse;
    return dt;
}
// optimization::SimpleBA (= Optimizer::FastBA, src/Optimization/SimpleBA.cpp:80-157): correspondence k links frames
// src_id[k] -> tgt_id[k] through the point pairs [offset[k], offset[k+1]) of a / b; poses column-major 4x4 floats, in place
double ref_simple_ba(int n_poses, float *poses_cm, int n_corr, const int32_t *src_id, const int32_t *tgt_id, const int64_t *offset,
                     const float *a, const float *b, int max_iteration)
{
    geometry::SE3List poses(n_poses);
    for (int i = 0; i < n_poses; ++i)
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) poses[i](r, c) = poses_cm[16 * i + 4 * c + r];
    std::vector<optimization::Correspondence> cs;
    for (int k = 0; k < n_corr; ++k)
    {
        geometry::PointCorrespondenceSet set;
        for (int64_t j = offset[k]; j < offset[k + 1]; ++j)
            set.push_back(std::make_pair(geometry::Point3(a[3 * j], a[3 * j + 1], a[3 * j + 2]), geometry::Point3(b[3 * j], b[3 * j + 1], b[3 * j + 2])));
        cs.push_back(optimization::Correspondence(src_id[k], tgt_id[k], set));
    }
    double t0 = Now();
    optimization::SimpleBA(cs, poses, max_iteration);
    double dt = Now() - t0;
    for (int i = 0; i < n_poses; ++i)
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) poses_cm[16 * i + 4 * c + r] = (float)poses[i](r, c);
    return dt;
}
// optimization::ComputeJTJAndJTr (SimpleBA.cpp:18-78) for one correspondence: out = JTJ_ss, JTJ_tt, JTJ_st, JTJ_ts (row-major 6x6
// each), JTr_s, JTr_t (6 each) = 156 floats
void ref_ba_blocks(const float *pose_s_cm, const float *pose_t_cm, const float *a, const float *b, long n, float *out)
{
    geometry::SE3List poses(2);
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) { poses[0](r, c) = pose_s_cm[4 * c + r]; poses[1](r, c) = pose_t_cm[4 * c + r]; }
    geometry::PointCorrespondenceSet set;
    for (long j = 0; j < n; ++j)
        set.push_back(std::make_pair(geometry::Point3(a[3 * j], a[3 * j + 1], a[3 * j + 2]), geometry::Point3(b[3 * j], b[3 * j + 1], b[3 * j + 2])));
    optimization::Correspondence corr(0, 1, set);
    geometry::Matrix6 ss, tt, st, ts;
    geometry::Se3 rs, rt;
    std::tie(ss, tt, st, ts, rs, rt) = optimization::ComputeJTJAndJTr(corr, poses);
    const geometry::Matrix6 *m[4] = {&ss, &tt, &st, &ts};
    for (int k = 0; k < 4; ++k)
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) out[36 * k + 6 * r + c] = (float)(*m[k])(r, c);
    for (int r = 0; r < 6; ++r) { out[144 + r] = (float)rs(r); out[150 + r] = (float)rt(r); }
}
#endif
} // extern "C"
