// Drop-in replacement for the reference's src/Integration/CubeHandler.h: the same class, namespace, public
// member functions and argument meaning (one_piece::integration::CubeHandler, reference CubeHandler.h:24-366),
// with the volume held in GPU memory by libonepiece_b200 instead of a host std::unordered_map.
//
// It compiles against the reference's own headers (Geometry/, Camera/, Integration/{TSDFVoxel,VoxelCube}.h) --
// put onepiece_b200/cpp in front of src/ on the include path (this file then shadows
// src/Integration/CubeHandler.h), or copy it over the original -- so example/ImageSequenceIntegration.cpp and example/DenseFusion/*.cpp
// keep compiling unchanged.  See INTEGRATION.md.
//
// Everything computational goes through the C-ABI in include/onepiece_b200.h; there is no CPU fallback.
#ifndef VOXEL_HASHING_H
#define VOXEL_HASHING_H
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "Camera/Camera.h"
#include "Geometry/Geometry.h"
#include "Geometry/RGBDFrame.h"
#include "Geometry/TriangleMesh.h"
#include "Integration/VoxelCube.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace integration
{
typedef std::unordered_map<CubeID, VoxelCube, CubeHasher> CubeMap;

class CubeHandler
{
  public:
    CubeHandler();
    CubeHandler(const camera::PinholeCamera &_camera);
    CubeHandler(const CubeHandler &) = delete;
    CubeHandler &operator=(const CubeHandler &) = delete;
    ~CubeHandler();

    // capacity of the device block pool (cubes of 8^3 voxels); call before the first IntegrateImage.
    // 262,144 cubes = a 512^3-voxel volume (2.7 GB at 20 B/voxel).
    void SetMaxCubes(int max_cubes);
    void SetDevice(int device);

    void SetVoxelResolution(float resolution);
    void SetTruncation(float trunc);
    void SetCamera(const camera::PinholeCamera &_camera);
    void SetFarPlane(float _far);
    void SetNearPlane(float _near);
    void Clear();

    // A volumetric method for building complex models from range images, Curless & Levoy 1996
    void IntegrateImage(const cv::Mat &depth, const cv::Mat &rgb, const geometry::TransformationMatrix &pose);
    void IntegrateImage(const geometry::RGBDFrame &rgbd, const geometry::TransformationMatrix &pose);
    void PrepareCubes(const cv::Mat &depth, const geometry::TransformationMatrix &pose, std::vector<CubeID> &cube_id_list);
    void ComputeBounding(const cv::Mat &depth, const geometry::TransformationMatrix &pose, geometry::Point3 &max_pos,
                         geometry::Point3 &min_pos);
    void ExtractTriangleMesh(geometry::TriangleMesh &mesh);
    std::shared_ptr<geometry::PointCloud> GetPointCloud() const;

    CubeID GetCubeID(const geometry::Point3 &point) const { return c_para.GetCubeID(point); }
    bool HasCube(const CubeID &cube_id) const;
    CubeMap GetCubeMap();
    void SetCubeMap(const CubeMap &_cube_map);
    bool WriteToFile(const std::string &filename) const;
    bool ReadFromFile(const std::string &filename);
    bool ReadFromFileFloat(const std::string &filename); // the older stream with a size word and separate colour records (:73-111)

    // CubeHandler.h:145-177,242-338: volume resampling / merging, on the device
    void Merge(const CubeHandler &another);
    void Merge(const CubeHandler &another, const geometry::TransformationMatrix &trans);
    std::shared_ptr<CubeHandler> Transform(const geometry::TransformationMatrix &trans) const;
    std::shared_ptr<CubeHandler> TransformNearest(const geometry::TransformationMatrix &trans);

    size_t CubeCount() const;

  protected:
    void EnsureVolume() const;
    void PushParams();
    void Download(std::vector<int32_t> &ids, std::vector<float> &voxels) const;

    // The reference's ORDER.  Its cube_map is a std::unordered_map whose iteration order -- a function of the insertion history --
    // decides the sequence of the extracted mesh, of the cubes a resampled volume creates, of the .cubes stream.  The block pool
    // has no such order, so the class keeps a mirror of the KEYS with the same key type, hasher and insertion history (new cubes
    // of a frame in PrepareCubes' i / j / k nesting, CubeHandler.cpp:163-191; resampled cubes at their first touch; merged cubes in
    // the other map's order) and hands its iteration order to the library wherever the reference iterates its map.
    typedef std::unordered_map<CubeID, int, CubeHasher> OrderMap;
    void NoteFrame();                                   // the cubes created since the last note came from one PrepareCubes pass
    void SyncOrder() const;                             // bring order_ up to date with the block pool
    std::vector<int32_t> IterationIds() const;          // cube ids (3 ints each) in the reference's iteration order
    void AdoptOrder(const int32_t *ids, size_t n);      // a fresh handler whose cubes were inserted in this sequence
    mutable OrderMap order_;
    mutable size_t ordered_cubes_ = 0;                  // pool cubes [0, ordered_cubes_) are in order_
    mutable std::vector<size_t> frame_ends_;            // pool cube count after every frame since then

    camera::PinholeCamera camera;
    CubePara c_para;
    float truncation = 0.1f;
    float far = 5.0f;
    float near = 0.5f;
    int max_cubes = 1 << 18;
    int device = 0;
    mutable opb_volume *volume = nullptr;
};
} // namespace integration
} // namespace one_piece
#endif
