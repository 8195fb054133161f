// Drop-in definitions of three reference functions that post-process what the hot path produces, computed on the GPU through
// the C-ABI (opb_mesh_clustering_simplify, opb_mesh_compute_normals, opb_pointcloud_downsample); results are bit-identical to
// the reference's, vertex / triangle / point order included, so nothing downstream changes:
//   geometry::ClusteringSimplification(TriangleMesh &, float)   reference src/Geometry/MeshSimplification.cpp:579-657
//       (the body of TriangleMesh::ClusteringSimplify, src/Geometry/TriangleMesh.cpp:53-58)
//   geometry::TriangleMesh::ComputeNormals()                     reference src/Geometry/TriangleMesh.cpp:95-127
//   geometry::PointCloud::DownSample(float) const                reference src/Geometry/PointCloud.cpp:145-189
//   geometry::PointCloud::EstimateNormals(float, int)            reference src/Geometry/PointCloud.cpp:102-144 (bit-identical unless
//       two neighbours of a point are at exactly the same distance: nanoflann orders such ties by tree traversal)
// Signatures are the reference's own (its headers are included unchanged).  To integrate, delete those three bodies from the
// reference's sources and add this file -- or, without touching the reference, put this object in front of the reference's
// on the link line with -Wl,--allow-multiple-definition, which is what tests/cpp/Makefile does.
#include <cstdlib>
#include <iostream>
#include <vector>

#include "Geometry/MeshSimplification.h"
#include "Geometry/PointCloud.h"
#include "Geometry/TriangleMesh.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace geometry
{
namespace
{
void Flatten(const Point3List &pts, std::vector<float> &out)
{
    out.resize(pts.size() * 3);
    for (size_t i = 0; i < pts.size(); ++i)
        for (int k = 0; k < 3; ++k) out[3 * i + k] = (float)pts[i](k);
}
void Unflatten(const float *in, size_t n, Point3List &out)
{
    out.resize(n);
    for (size_t i = 0; i < n; ++i) out[i] = Point3(in[3 * i], in[3 * i + 1], in[3 * i + 2]);
}
void Report(const char *what)
{
    std::cout << RED << "[ERROR]::[" << what << "]::" << opb_last_error() << RESET << std::endl;
}
} // namespace

void ClusteringSimplification(TriangleMesh &wait_to_simplify, float grid_len)
{
    if (grid_len <= 0)
    {
        std::cout << RED << "[ClusteringMeshSimplification]::[ERROR]::Grid length cannot be less than 0." << RESET << std::endl;
        return;
    }
    std::vector<float> p, c;
    std::vector<uint32_t> t(wait_to_simplify.triangles.size() * 3);
    Flatten(wait_to_simplify.points, p);
    const bool has_colors = wait_to_simplify.HasColors();
    if (has_colors) Flatten(wait_to_simplify.colors, c);
    for (size_t i = 0; i < wait_to_simplify.triangles.size(); ++i)
        for (int k = 0; k < 3; ++k) t[3 * i + k] = wait_to_simplify.triangles[i](k);
    float *op = nullptr, *oc = nullptr;
    uint32_t *ot = nullptr;
    size_t nv = 0, nt = 0;
    int rc = opb_mesh_clustering_simplify(0, p.data(), has_colors ? c.data() : nullptr, wait_to_simplify.points.size(), t.data(),
                                          wait_to_simplify.triangles.size(), grid_len, &op, has_colors ? &oc : nullptr, &ot, &nv, &nt);
    if (rc != OPB_OK) { Report("ClusteringMeshSimplification"); return; }
    const bool had_normals = wait_to_simplify.HasNormals();
    Unflatten(op, nv, wait_to_simplify.points);
    if (has_colors) Unflatten(oc, nv, wait_to_simplify.colors);
    wait_to_simplify.triangles.resize(nt);
    for (size_t i = 0; i < nt; ++i) wait_to_simplify.triangles[i] = Point3ui(ot[3 * i], ot[3 * i + 1], ot[3 * i + 2]);
    opb_free(op); opb_free(oc); opb_free(ot);
    if (had_normals) wait_to_simplify.ComputeNormals(); // CompactMesh recomputes them (MeshSimplification.cpp:338-342)
    std::cout << GREEN << "[ClusteringMeshSimplification]::[INFO]::Simplify done." << RESET << std::endl;
}

void TriangleMesh::ComputeNormals()
{
    std::vector<float> p, n(points.size() * 3);
    std::vector<uint32_t> t(triangles.size() * 3);
    Flatten(points, p);
    for (size_t i = 0; i < triangles.size(); ++i)
        for (int k = 0; k < 3; ++k) t[3 * i + k] = triangles[i](k);
    if (opb_mesh_compute_normals(0, p.data(), points.size(), t.data(), triangles.size(), n.data()) != OPB_OK) { Report("ComputeNormals"); return; }
    Unflatten(n.data(), points.size(), normals);
}

void PointCloud::EstimateNormals(float radius, int knn)
{
    static opb_kdtree *ws = nullptr; // callers are single-threaded; one search workspace per process
    if (!ws && opb_kdtree_create(0, &ws) != OPB_OK)
    {
        Report("EstimateNormals");
        std::exit(1); // no device: there is no CPU path
    }
    std::cout << BLUE << "[EstimateNormals]::[INFO]::RadiusSearch " << knn << " nearest points, radius: " << radius << RESET << std::endl;
    std::vector<float> p, n(points.size() * 3);
    Flatten(points, p);
    // the reference's own k-d tree and visiting order on the device: identical neighbours, identical order
    if (opb_kdtree_build(ws, p.data(), points.size()) != OPB_OK || opb_kdtree_estimate_normals(ws, radius, knn, n.data()) != OPB_OK)
    {
        Report("EstimateNormals");
        return;
    }
    Unflatten(n.data(), points.size(), normals);
}

std::shared_ptr<PointCloud> PointCloud::DownSample(float grid_len) const
{
    PointCloud pcd;
    std::vector<float> p, c, n;
    Flatten(points, p);
    const bool has_color = HasColors(), has_normal = HasNormals();
    if (has_color) Flatten(colors, c);
    if (has_normal) Flatten(normals, n);
    float *op = nullptr, *oc = nullptr, *on = nullptr;
    size_t m = 0;
    int rc = opb_pointcloud_downsample(0, p.data(), has_color ? c.data() : nullptr, has_normal ? n.data() : nullptr, points.size(), grid_len, &op,
                                       has_color ? &oc : nullptr, has_normal ? &on : nullptr, &m);
    if (rc != OPB_OK) { Report("DownSample"); return std::make_shared<PointCloud>(*this); }
    Unflatten(op, m, pcd.points);
    if (has_color) Unflatten(oc, m, pcd.colors);
    if (has_normal) Unflatten(on, m, pcd.normals);
    opb_free(op); opb_free(oc); opb_free(on);
    return std::make_shared<PointCloud>(pcd);
}
} // namespace geometry
} // namespace one_piece
