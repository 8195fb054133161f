// See CubeHandler.h.  Each member cites the reference member it stands in for (reference
// src/Integration/CubeHandler.{h,cpp}).
#include "Integration/CubeHandler.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <set>

namespace one_piece
{
namespace integration
{
namespace
{
bool Check(int rc, const char *what)
{
    if (rc != OPB_OK)
    {
        // the reference has no error channel (no exceptions, no status codes): report like its other messages
        std::cout << RED << "[ERROR]::[" << what << "]::" << opb_last_error() << RESET << std::endl;
        if (rc == OPB_ERR_CUDA) std::exit(1); // no device, no result: there is no CPU path to fall back to
    }
    return rc == OPB_OK;
}
int DepthType(const cv::Mat &depth)
{
    if (depth.depth() == CV_32FC1) return OPB_DEPTH_F32;
    if (depth.depth() == CV_16UC1) return OPB_DEPTH_U16;
    return -1;
}
void PoseToArray(const geometry::TransformationMatrix &pose, float *out)
{
    Eigen::Matrix4f p = pose.cast<float>(); // column-major
    for (int i = 0; i < 16; ++i) out[i] = p.data()[i];
}
} // namespace

CubeHandler::CubeHandler() { c_para.InitializeVoxelCube(); }
CubeHandler::CubeHandler(const camera::PinholeCamera &_camera) : camera(_camera) { c_para.InitializeVoxelCube(); }
CubeHandler::~CubeHandler() { opb_volume_destroy(volume); }

static void FillDesc(opb_volume_desc &d, const camera::PinholeCamera &camera, const CubePara &c_para, float truncation,
                     float near, float far, int max_cubes, int device)
{
    opb_volume_desc_default(&d);
    d.fx = camera.GetFx(); d.fy = camera.GetFy(); d.cx = camera.GetCx(); d.cy = camera.GetCy();
    d.width = (int)camera.GetWidth(); d.height = (int)camera.GetHeight(); d.depth_scale = camera.GetDepthScale();
    d.voxel_resolution = c_para.VoxelResolution;
    d.truncation = truncation;
    d.near_plane = near; d.far_plane = far;
    d.max_cubes = max_cubes; d.device = device;
}
void CubeHandler::EnsureVolume() const
{
    if (volume) return;
    opb_volume_desc d;
    FillDesc(d, camera, c_para, truncation, near, far, max_cubes, device);
    Check(opb_volume_create(&d, &volume), "CubeHandler");
}
void CubeHandler::PushParams()
{
    if (!volume) return;
    opb_volume_desc d;
    FillDesc(d, camera, c_para, truncation, near, far, max_cubes, device);
    Check(opb_volume_set_params(volume, &d), "CubeHandler");
}
// ---- the mirror of the reference's cube map (see CubeHandler.h) ----
void CubeHandler::NoteFrame()
{
    size_t n = 0;
    if (volume && opb_volume_num_cubes(volume, &n) == OPB_OK && (frame_ends_.empty() ? n > ordered_cubes_ : n > frame_ends_.back())) frame_ends_.push_back(n);
}
void CubeHandler::SyncOrder() const
{
    EnsureVolume();
    size_t n = 0;
    if (opb_volume_num_cubes(volume, &n) != OPB_OK || n <= ordered_cubes_) { frame_ends_.clear(); return; }
    std::vector<int32_t> ids(n * 3);
    size_t cap = n;
    if (!Check(opb_volume_download(volume, ids.data(), nullptr, &cap), "CubeHandler")) return;
    if (frame_ends_.empty() || frame_ends_.back() < n) frame_ends_.push_back(n); // cubes of unknown history: treated as one frame
    size_t from = ordered_cubes_;
    for (size_t end : frame_ends_)
    {
        if (end <= from) continue;
        // the cubes a frame creates enter the reference's map in PrepareCubes' loop nesting: i outermost, k innermost
        std::vector<size_t> idx(end - from);
        for (size_t q = 0; q < idx.size(); ++q) idx[q] = from + q;
        std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
            for (int c = 0; c < 3; ++c)
                if (ids[3 * a + c] != ids[3 * b + c]) return ids[3 * a + c] < ids[3 * b + c];
            return false;
        });
        for (size_t q : idx) order_[CubeID(ids[3 * q], ids[3 * q + 1], ids[3 * q + 2])] = 0;
        from = end;
    }
    ordered_cubes_ = n;
    frame_ends_.clear();
}
std::vector<int32_t> CubeHandler::IterationIds() const
{
    SyncOrder();
    std::vector<int32_t> out;
    out.reserve(order_.size() * 3);
    for (auto it = order_.begin(); it != order_.end(); ++it)
        for (int c = 0; c < 3; ++c) out.push_back(it->first(c));
    return out;
}
void CubeHandler::AdoptOrder(const int32_t *ids, size_t n)
{
    order_.clear();
    for (size_t q = 0; q < n; ++q) order_[CubeID(ids[3 * q], ids[3 * q + 1], ids[3 * q + 2])] = 0;
    ordered_cubes_ = n;
    frame_ends_.clear();
}
void CubeHandler::SetMaxCubes(int n) { max_cubes = n; }
void CubeHandler::SetDevice(int dev) { device = dev; }
// CubeHandler.h:36 (prints through CubePara::SetVoxelResolution like the reference)
void CubeHandler::SetVoxelResolution(float resolution) { c_para.SetVoxelResolution(resolution); PushParams(); }
void CubeHandler::SetTruncation(float trunc) { truncation = trunc; PushParams(); }                // :141
void CubeHandler::SetCamera(const camera::PinholeCamera &_camera) { camera = _camera; PushParams(); } // :137
void CubeHandler::SetFarPlane(float _far) { far = _far; PushParams(); }                            // :349
void CubeHandler::SetNearPlane(float _near) { near = _near; PushParams(); }                        // :353
void CubeHandler::Clear()                                                                        // :133
{
    if (volume) Check(opb_volume_clear(volume), "Clear");
    order_.clear(); // (keeps its buckets, like cube_map.clear())
    ordered_cubes_ = 0;
    frame_ends_.clear();
}

// CubeHandler.cpp:197-210
void CubeHandler::IntegrateImage(const cv::Mat &depth, const cv::Mat &rgb, const geometry::TransformationMatrix &pose)
{
    EnsureVolume();
    float p[16];
    PoseToArray(pose, p);
    int rc = opb_volume_integrate(volume, depth.data, DepthType(depth), rgb.data, p);
    Check(rc, "IntegrateImage");
    NoteFrame();
#if DEBUG_MODE
    opb_frame_stats st;
    if (opb_volume_frame_stats(volume, &st) == OPB_OK)
        std::cout << BLUE << "[PrepareCubes]::[DEBUG]::Number of Candidate Cubes: " << st.frame_cubes << RESET << std::endl;
#endif
    if (rc == OPB_OK) std::cout << GREEN << "[IntegrateImage]::[Info]::Finish image integration." << RESET << std::endl;
}
// CubeHandler.cpp:211-214
void CubeHandler::IntegrateImage(const geometry::RGBDFrame &rgbd, const geometry::TransformationMatrix &pose)
{
    IntegrateImage(rgbd.depth, rgbd.rgb, pose);
}
// CubeHandler.cpp:147-196
void CubeHandler::PrepareCubes(const cv::Mat &depth, const geometry::TransformationMatrix &pose, std::vector<CubeID> &cube_id_list)
{
    EnsureVolume();
    float p[16];
    PoseToArray(pose, p);
    size_t n = 0;
    if (!Check(opb_volume_prepare_cubes(volume, depth.data, DepthType(depth), p, nullptr, &n), "PrepareCubes")) { cube_id_list.clear(); return; }
    NoteFrame();
    std::vector<int32_t> ids(n * 3 + 3);
    Check(opb_volume_last_frame_cubes(volume, ids.data(), &n), "PrepareCubes");
    cube_id_list.clear();
    for (size_t i = 0; i < n; ++i) cube_id_list.push_back(CubeID(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]));
}
// CubeHandler.cpp:116-145 (read-only, like the reference's: nothing is selected or allocated)
void CubeHandler::ComputeBounding(const cv::Mat &depth, const geometry::TransformationMatrix &pose, geometry::Point3 &max_pos,
                                  geometry::Point3 &min_pos)
{
    EnsureVolume();
    float p[16], mn[3], mx[3];
    PoseToArray(pose, p);
    if (!Check(opb_volume_compute_bounding(volume, depth.data, DepthType(depth), p, mn, mx), "ComputeBounding")) return;
    max_pos = geometry::Point3(mx[0], mx[1], mx[2]);
    min_pos = geometry::Point3(mn[0], mn[1], mn[2]);
}
// CubeHandler.cpp:9-44 + TriangleMesh::LoadFromMeshes (TriangleMesh.cpp:73-94)
void CubeHandler::ExtractTriangleMesh(geometry::TriangleMesh &mesh)
{
    EnsureVolume();
    float *xyz = nullptr, *rgb = nullptr;
    uint32_t *tri = nullptr;
    size_t nv = 0, nt = 0;
    // cube by cube in the iteration order of the reference's map (CubeHandler.cpp:27-40), cells in GenerateMeshByCube's nesting
    const std::vector<int32_t> order = IterationIds();
    Check(opb_volume_extract_mesh_ordered(volume, order.data(), order.size() / 3, &xyz, &rgb, &tri, &nv, &nt), "ExtractTriangleMesh");
    mesh.Reset();
    mesh.points.resize(nv);
    mesh.colors.resize(nv);
    mesh.triangles.resize(nt);
    for (size_t i = 0; i < nv; ++i)
    {
        mesh.points[i] = geometry::Point3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        mesh.colors[i] = geometry::Point3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    }
    for (size_t i = 0; i < nt; ++i) mesh.triangles[i] = geometry::Point3ui(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
    opb_free(xyz); opb_free(rgb); opb_free(tri);
    std::cout << GREEN << "[ExtractTriangleMesh]::[Info]::Finish mesh extraction( sm_100a kernels)." << RESET << std::endl;
}
void CubeHandler::Download(std::vector<int32_t> &ids, std::vector<float> &voxels) const
{
    EnsureVolume();
    size_t n = 0;
    Check(opb_volume_num_cubes(volume, &n), "GetCubeMap");
    ids.resize(n * 3);
    voxels.resize(n * 512 * 5);
    size_t cap = n;
    if (n) Check(opb_volume_download(volume, ids.data(), voxels.data(), &cap), "GetCubeMap");
}
size_t CubeHandler::CubeCount() const
{
    EnsureVolume();
    size_t n = 0;
    Check(opb_volume_num_cubes(volume, &n), "CubeCount");
    return n;
}
// CubeHandler.h:339-343 (returns a deep copy, like the reference)
CubeMap CubeHandler::GetCubeMap()
{
    std::vector<int32_t> ids;
    std::vector<float> vox;
    Download(ids, vox);
    // The copy iterates like the reference's map: same bucket count, keys inserted in REVERSE iteration order (a node goes to the
    // front of its bucket, or of the whole list when the bucket is empty, so this rebuilds the list back to front).
    SyncOrder();
    std::unordered_map<CubeID, size_t, CubeHasher> slot_of;
    for (size_t c = 0; c < ids.size() / 3; ++c) slot_of[CubeID(ids[3 * c], ids[3 * c + 1], ids[3 * c + 2])] = c;
    std::vector<CubeID> seq;
    for (auto it = order_.begin(); it != order_.end(); ++it) seq.push_back(it->first);
    CubeMap m;
    m.rehash(order_.bucket_count());
    for (size_t q = seq.size(); q-- > 0;)
    {
        const CubeID &id = seq[q];
        VoxelCube cube(id);
        const float *src = &vox[slot_of[id] * 512 * 5];
        for (int j = 0; j < 512; ++j)
        {
            cube.voxels[j].sdf = src[5 * j];
            cube.voxels[j].weight = src[5 * j + 1];
            cube.voxels[j].color = geometry::Point3(src[5 * j + 2], src[5 * j + 3], src[5 * j + 4]);
        }
        m[id] = cube;
    }
    return m;
}
// CubeHandler.h:344-348
void CubeHandler::SetCubeMap(const CubeMap &_cube_map)
{
    std::cout << YELLOW << "[WARNING]::[SetCubeMap]::Note that you are changing the hashing map directly." << RESET << std::endl;
    EnsureVolume();
    std::vector<int32_t> ids;
    std::vector<float> vox;
    ids.reserve(_cube_map.size() * 3);
    vox.reserve(_cube_map.size() * 512 * 5);
    for (auto it = _cube_map.begin(); it != _cube_map.end(); ++it)
    {
        for (int k = 0; k < 3; ++k) ids.push_back(it->first(k));
        for (int j = 0; j < 512; ++j)
        {
            const TSDFVoxel &v = it->second.voxels[j];
            vox.push_back(v.sdf); vox.push_back(v.weight);
            vox.push_back(v.color(0)); vox.push_back(v.color(1)); vox.push_back(v.color(2));
        }
    }
    Check(opb_volume_upload(volume, ids.data(), vox.data(), _cube_map.size()), "SetCubeMap");
    // cube_map = _cube_map: the copy iterates like the original (same buckets, same list)
    order_.clear();
    order_.rehash(_cube_map.bucket_count());
    for (size_t q = ids.size() / 3; q-- > 0;) order_[CubeID(ids[3 * q], ids[3 * q + 1], ids[3 * q + 2])] = 0;
    ordered_cubes_ = _cube_map.size();
    frame_ends_.clear();
}
// CubeHandler.h:129-132
bool CubeHandler::HasCube(const CubeID &cube_id) const
{
    std::vector<int32_t> ids;
    size_t n = 0;
    EnsureVolume();
    Check(opb_volume_num_cubes(volume, &n), "HasCube");
    ids.resize(n * 3);
    size_t cap = n;
    if (n) Check(opb_volume_download(volume, ids.data(), nullptr, &cap), "HasCube");
    for (size_t c = 0; c < n; ++c)
        if (ids[3 * c] == cube_id(0) && ids[3 * c + 1] == cube_id(1) && ids[3 * c + 2] == cube_id(2)) return true;
    return false;
}
// CubeHandler.cpp:45-69
std::shared_ptr<geometry::PointCloud> CubeHandler::GetPointCloud() const
{
    std::vector<int32_t> ids;
    std::vector<float> vox;
    Download(ids, vox);
    geometry::PointCloud pcd;
    const float cube_resolution = CUBE_SIZE * c_para.VoxelResolution;
    SyncOrder();
    std::unordered_map<CubeID, size_t, CubeHasher> slot_of;
    for (size_t c = 0; c < ids.size() / 3; ++c) slot_of[CubeID(ids[3 * c], ids[3 * c + 1], ids[3 * c + 2])] = c;
    for (auto it = order_.begin(); it != order_.end(); ++it) // the reference walks its map (CubeHandler.cpp:48)
    {
        const size_t c = slot_of[it->first];
        for (size_t x = 0; x != CUBE_SIZE; ++x)
            for (size_t y = 0; y != CUBE_SIZE; ++y)
                for (size_t z = 0; z != CUBE_SIZE; ++z)
                {
                    const size_t voxel_id = x + y * CUBE_SIZE + z * CUBE_SIZE * CUBE_SIZE;
                    const float sdf = vox[(c * 512 + voxel_id) * 5], w = vox[(c * 512 + voxel_id) * 5 + 1];
                    if (w != 0 && std::fabs(sdf) < truncation)
                    {
                        const float f = std::fabs(sdf) / truncation;
                        geometry::Point3 origin(ids[3 * c] * cube_resolution, ids[3 * c + 1] * cube_resolution, ids[3 * c + 2] * cube_resolution);
                        pcd.points.push_back(origin + c_para.VoxelCentroidOffSet[voxel_id]);
                        pcd.colors.push_back(geometry::Point3(f, f, f));
                    }
                }
    }
    return std::make_shared<geometry::PointCloud>(pcd);
}
// CubeHandler.h:242-298 (trilinear) and :299-338 (nearest).  Like the reference, Transform copies c_para into the result and
// TransformNearest does not: its result keeps CubePara's default VoxelResolution (0.01).
static std::shared_ptr<CubeHandler> MakeResult(const camera::PinholeCamera &camera) { return std::make_shared<CubeHandler>(camera); }
std::shared_ptr<CubeHandler> CubeHandler::Transform(const geometry::TransformationMatrix &trans) const
{
    EnsureVolume();
    std::shared_ptr<CubeHandler> after_trans = MakeResult(camera);
    after_trans->far = far; after_trans->near = near; after_trans->truncation = truncation; after_trans->device = device;
    after_trans->c_para = c_para;
    float t[16];
    PoseToArray(trans, t);
    const std::vector<int32_t> order = IterationIds();
    int32_t *created = nullptr;
    size_t n_created = 0;
    if (Check(opb_volume_transform_ordered(volume, t, 0, after_trans->c_para.VoxelResolution, 0, order.data(), order.size() / 3, &after_trans->volume,
                                           &created, &n_created), "Transform"))
        after_trans->AdoptOrder(created, n_created);
    opb_free(created);
    opb_volume_desc d;
    if (after_trans->volume && opb_volume_get_desc(after_trans->volume, &d) == OPB_OK) after_trans->max_cubes = d.max_cubes;
    return after_trans;
}
std::shared_ptr<CubeHandler> CubeHandler::TransformNearest(const geometry::TransformationMatrix &trans)
{
    EnsureVolume();
    std::shared_ptr<CubeHandler> after_trans = MakeResult(camera);
    after_trans->far = far; after_trans->near = near; after_trans->truncation = truncation; after_trans->device = device;
    float t[16];
    PoseToArray(trans, t);
    const std::vector<int32_t> order = IterationIds();
    int32_t *created = nullptr;
    size_t n_created = 0;
    if (Check(opb_volume_transform_ordered(volume, t, 1, after_trans->c_para.VoxelResolution, 0, order.data(), order.size() / 3, &after_trans->volume,
                                           &created, &n_created), "TransformNearest"))
        after_trans->AdoptOrder(created, n_created);
    opb_free(created);
    opb_volume_desc d;
    if (after_trans->volume && opb_volume_get_desc(after_trans->volume, &d) == OPB_OK) after_trans->max_cubes = d.max_cubes;
    return after_trans;
}
// CubeHandler.h:145-167
void CubeHandler::Merge(const CubeHandler &another)
{
    if (c_para.VoxelResolution != another.c_para.VoxelResolution)
    {
        std::cout << YELLOW << "[Warning]::[MergeVoxelHash]::Voxel resolution is not identical." << RESET << std::endl;
        return;
    }
    EnsureVolume();
    another.EnsureVolume();
    SyncOrder();
    const std::vector<int32_t> theirs = another.IterationIds();
    if (!Check(opb_volume_merge(volume, another.volume), "Merge")) return;
    // cubes this map did not have are inserted while walking the other map (CubeHandler.h:152-158)
    for (size_t q = 0; q < theirs.size() / 3; ++q)
    {
        const CubeID id(theirs[3 * q], theirs[3 * q + 1], theirs[3 * q + 2]);
        if (order_.find(id) == order_.end()) order_[id] = 0;
    }
    size_t n = 0;
    if (opb_volume_num_cubes(volume, &n) == OPB_OK) ordered_cubes_ = n;
    frame_ends_.clear();
}
// CubeHandler.h:168-177
void CubeHandler::Merge(const CubeHandler &another, const geometry::TransformationMatrix &trans)
{
    if (c_para.VoxelResolution != another.c_para.VoxelResolution)
    {
        std::cout << YELLOW << "[Warning]::[MergeVoxelHash]::Voxel resolution is not identical." << RESET << std::endl;
        return;
    }
    auto another_ptr = another.Transform(trans);
    Merge(*another_ptr);
}
// CubeHandler.h:113-128: the same float stream ([u32 n] then per cube VoxelCube::WriteToBuffer, VoxelCube.h:128-142)
bool CubeHandler::WriteToFile(const std::string &filename) const
{
    CubeMap m = const_cast<CubeHandler *>(this)->GetCubeMap();
    std::ofstream ofs(filename, std::ios::binary);
    std::vector<float> buffer;
    unsigned int size = m.size();
    buffer.push_back(0);
    *((unsigned int *)(&buffer[0])) = size;
    for (auto it = m.begin(); it != m.end(); ++it) it->second.WriteToBuffer(buffer);
    ofs.write((char *)&buffer[0], sizeof(float) * buffer.size());
    std::cout << GREEN << "[CubeHandler]::[INFO]::Write TSDF field done!(To BinaryFile) " << RESET << std::endl;
    return true;
}
// CubeHandler.h:40-69
bool CubeHandler::ReadFromFile(const std::string &filename)
{
    std::ifstream ifs(filename, std::ifstream::binary);
    if (!ifs) return false;
    std::vector<float> buffer;
    ifs.seekg(0, ifs.end);
    size_t length = ifs.tellg();
    ifs.seekg(0, ifs.beg);
    buffer.resize(length / sizeof(float));
    ifs.read((char *)&buffer[0], length);
    unsigned int cube_size = *((unsigned int *)(&buffer[0]));
    size_t ptr = 1;
    CubeMap m;
    for (size_t i = 0; i < cube_size; ++i)
    {
        float x = buffer[ptr++], y = buffer[ptr++], z = buffer[ptr++];
        CubeID cube_id = CubeID(x, y, z);
        m[cube_id] = VoxelCube(cube_id);
        m[cube_id].ReadFromBuffer(buffer, ptr);
    }
    bool quiet = std::cout.fail();
    std::cout.setstate(std::ios_base::failbit);
    SetCubeMap(m);
    if (!quiet) std::cout.clear();
    std::cout << GREEN << "[CubeHandler]::[INFO]::Load TSDF field done!(From BinaryFile) " << RESET << std::endl;
    return true;
}
// CubeHandler.h:73-111: [size word][cube count] then per cube 3 id floats and VoxelCube::ReadFromBufferFloat (VoxelCube.h:168-194)
bool CubeHandler::ReadFromFileFloat(const std::string &filename)
{
    std::ifstream ifs(filename, std::ifstream::binary);
    if (!ifs) return false;
    std::vector<float> buffer;
    ifs.seekg(0, ifs.end);
    size_t length = ifs.tellg();
    ifs.seekg(0, ifs.beg);
    buffer.resize(length / sizeof(float));
    ifs.read((char *)&buffer[0], length);
    if (buffer.size() < 2) return false;
    unsigned int cube_size = buffer[1];
    size_t ptr = 2;
    CubeMap m;
    for (size_t i = 0; i < cube_size; ++i)
    {
        float x = buffer[ptr++], y = buffer[ptr++], z = buffer[ptr++];
        CubeID cube_id = CubeID(x, y, z);
        m[cube_id] = VoxelCube(cube_id);
        m[cube_id].ReadFromBufferFloat(buffer, ptr);
    }
    bool quiet = std::cout.fail();
    std::cout.setstate(std::ios_base::failbit);
    SetCubeMap(m);
    if (!quiet) std::cout.clear();
    std::cout << GREEN << "[CubeHandler]::[INFO]::Load TSDF field done!(From BinaryFile) " << RESET << std::endl;
    return true;
}
} // namespace integration
} // namespace one_piece
