/*
 * onepiece_b200 -- C-ABI of the B200-native (sm_100a) dense-reconstruction hot path.
 *
 * The reference (wlsdzyzl/OnePiece) has no FFI: its hot path sits behind C++ symbols of libone_piece.so.
 * This header is the one device boundary the rebuild introduces, directly underneath those symbols
 * (SURVEY.md §8b).  Every entry point names the reference interface it replaces (file:line relative to the
 * reference tree).  Plain pointers and sizes only; no C++/torch types.  The reference-side C++ classes that
 * call these functions (drop-in for CubeHandler / PointToPlane / Odometry::DenseTracking) live in
 * onepiece_b200/cpp/ and are described in INTEGRATION.md.
 *
 * Conventions
 *   - all functions return OPB_OK (0) or a negative opb_status; opb_last_error() gives the message of the
 *     last failure on the calling thread.  Nothing ever falls back to a CPU path: without a CUDA device the
 *     create calls fail with OPB_ERR_CUDA.
 *   - poses are camera-to-world 4x4 float, COLUMN-major (Eigen's default, geometry::TransformationMatrix).
 *   - depth is either float32 metres (OPB_DEPTH_F32 == CV_32FC1) or uint16 raw units divided by
 *     depth_scale (OPB_DEPTH_U16 == CV_16UC1), row-major W x H; colour is uint8 x 3 in file order (BGR).
 *   - an object may be driven from one host thread at a time (same as the reference's classes).
 */
#ifndef ONEPIECE_B200_H
#define ONEPIECE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPB_IPC_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t): handles of device buffers that peer ranks map (opb_ipc_open) */

typedef enum
{
    OPB_OK = 0,
    OPB_ERR_INVALID = -1,  /* bad argument */
    OPB_ERR_CUDA = -2,     /* CUDA runtime failure or no device */
    OPB_ERR_CAPACITY = -3, /* block pool / output buffer too small */
    OPB_ERR_UNSUPPORTED = -4
} opb_status;

enum { OPB_DEPTH_F32 = 5 /* CV_32FC1 */, OPB_DEPTH_U16 = 2 /* CV_16UC1 */ };
enum
{
    OPB_STORAGE_F32 = 0,     /* 20 B/voxel: f32 sdf, weight, 3 x colour -- bit-exact with TSDFVoxel.h:8-82 */
    OPB_STORAGE_PACKED16 = 1 /* 8 B/voxel: f16 sdf, f16 weight, rgb8 (throughput mode, error reported) */
};

const char *opb_last_error(void);
int opb_device_count(void);
/* pinned host staging buffers for the *_async entry points */
int opb_host_alloc(void **ptr, size_t bytes);
void opb_host_free(void *ptr);
/* frees memory returned through out-parameters of this library (opb_volume_extract_mesh) */
void opb_free(void *ptr);

/* Host-side helpers, exported because the per-frame constants they produce feed discrete decisions on the
 * device and must match the reference bit for bit (no GPU needed to call them):
 *   opb_pose_inverse   = Eigen::Matrix4f::inverse() as used at Integrator.cpp:18,48 (column-major in/out)
 *   opb_frustum_planes = Frustum::ComputeFromCamera (Frustum.cpp:7-52); 6 planes x 4 floats in the order
 *                        Frustum::ContainPoint tests them: top, left, right, bottom, near, far */
void opb_pose_inverse(const float pose_colmajor[16], float inverse_colmajor[16]);
void opb_frustum_planes(float fx, float fy, float cy, int width, int height, float near_plane, float far_plane,
                        const float pose_colmajor[16], float planes[24]);

/* Device self-test: compares the shared-reciprocal quotient used by the voxel-update kernel with div.rn on n
 * pseudo-random operand pairs (mode 0: integer divisors 1..2^24 as in TSDFVoxel::operator+; mode 1: arbitrary
 * operands with exponents within +-40 of 1.0 as in the projection) and returns the number of bitwise mismatches. */
int opb_selftest_quotient(int device, uint64_t n, uint64_t seed, int mode, uint64_t *mismatches);

/* ------------------------------------------------------------------------------------------------------
 * TSDF volume  (replaces one_piece::integration::CubeHandler, src/Integration/CubeHandler.h:24-366)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct opb_volume opb_volume;

typedef struct
{
    /* camera::PinholeCamera (src/Camera/Camera.h:13-130) */
    float fx, fy, cx, cy;
    int32_t width, height;
    float depth_scale;
    /* CubeHandler::SetVoxelResolution (CubeHandler.h:36), SetTruncation (:141), SetNearPlane/SetFarPlane (:349-356) */
    float voxel_resolution; /* reference default 0.01 */
    float truncation;       /* reference default 0.1  */
    float near_plane;       /* reference default 0.5  */
    float far_plane;        /* reference default 5.0  */
    /* device-side block pool: the unordered_map<CubeID,VoxelCube> (CubeHandler.h:22) becomes a bounded pool of
     * 8^3-voxel cubes plus an open-addressing hash table cube-id -> slot */
    int32_t max_cubes;
    int32_t storage; /* OPB_STORAGE_* */
    int32_t device;  /* CUDA ordinal */
    /* sub-volume ownership for multi-GPU fusion: this volume allocates/integrates cube (i,j,k)
     * only if floor_mod(floor_div(id[shard_axis], shard_slab_cubes), shard_world) == shard_rank.
     * shard_world <= 1 disables sharding. */
    int32_t shard_rank, shard_world, shard_axis, shard_slab_cubes;
    /* optional cudaStream_t to run on (NULL: the volume creates its own non-blocking stream) */
    void *stream;
} opb_volume_desc;

/* per-frame counters of the last integrate call (device counters, read back on request) */
typedef struct
{
    int32_t candidate_cubes; /* cubes tested in [min-1,max+1]^3   (CubeHandler.cpp:165-171) */
    int32_t frame_cubes;     /* cubes listed for this frame        (:181-191)               */
    int32_t total_cubes;     /* cubes allocated in the volume                                */
    int32_t overflow;        /* how many times the block pool has grown so far (the synchronous calls grow it instead
                                of dropping cubes; asynchronous ones return OPB_ERR_CAPACITY at the next synchronisation) */
    int64_t updated_voxels;  /* voxels whose |sdf| < truncation this frame (Integrator.cpp:74) */
    float bbox_min[3], bbox_max[3]; /* CubeHandler::ComputeBounding result (CubeHandler.cpp:116-145) */
    float select_ms;         /* device time of cube selection, when profiling is on */
    float integrate_ms;      /* device time of the voxel-update kernel, when profiling is on */
} opb_frame_stats;

void opb_volume_desc_default(opb_volume_desc *desc);
int opb_volume_create(const opb_volume_desc *desc, opb_volume **out);
void opb_volume_destroy(opb_volume *v);
/* CubeHandler::Clear (CubeHandler.h:133-136) */
int opb_volume_clear(opb_volume *v);
/* setters of CubeHandler callable at any time (SetCamera :137, SetVoxelResolution :36, SetTruncation :141,
 * SetFarPlane/SetNearPlane :349-356); pool/storage/device/shard fields of desc are ignored */
int opb_volume_set_params(opb_volume *v, const opb_volume_desc *desc);
/* record CUDA events around the selection and update kernels of every frame (for bench.py) */
int opb_volume_set_profiling(opb_volume *v, int on);
/* accumulated device milliseconds of cube selection (K1) and voxel update (K2) over *frames profiled frames */
int opb_volume_profile_read(opb_volume *v, double *select_ms, double *integrate_ms, int64_t *frames, int reset);

/* CubeHandler::IntegrateImage(depth, rgb, pose) (CubeHandler.cpp:197-210; Integrator.cpp:36-94):
 * host buffers, synchronous -- returns after the volume has been updated. */
int opb_volume_integrate(opb_volume *v, const void *depth, int depth_type, const uint8_t *bgr,
                         const float pose_colmajor[16]);
/* same, but only enqueues (host buffers should be pinned: opb_host_alloc); up to two frames in flight, the
 * call blocks only while the staging buffer it needs is still busy.  opb_volume_synchronize() drains. */
int opb_volume_integrate_async(opb_volume *v, const void *depth, int depth_type, const uint8_t *bgr,
                               const float pose_colmajor[16]);
/* same with inputs already resident in device memory (enqueue only) */
int opb_volume_integrate_device(opb_volume *v, const void *d_depth, int depth_type, const uint8_t *d_bgr,
                                const float pose_colmajor[16]);
int opb_volume_synchronize(opb_volume *v);
int opb_volume_frame_stats(opb_volume *v, opb_frame_stats *out); /* synchronizes */

/* CubeHandler::PrepareCubes alone (CubeHandler.cpp:147-196): selects + allocates, returns the frame's cube
 * list (ids: 3 x int32 per cube, unordered).  *n_cubes in: capacity, out: count. */
int opb_volume_prepare_cubes(opb_volume *v, const void *depth, int depth_type, const float pose_colmajor[16],
                             int32_t *cube_ids, size_t *n_cubes);
/* CubeHandler::ComputeBounding(depth, pose, max_pos, min_pos) (CubeHandler.cpp:116-145) by itself: read-only like the
 * reference's -- nothing is selected or allocated.  Host depth image, synchronous. */
int opb_volume_compute_bounding(opb_volume *v, const void *depth, int depth_type, const float pose_colmajor[16],
                                float bbox_min[3], float bbox_max[3]);
/* the cube ids the last frame listed (PrepareCubes' cube_id_list); *n_cubes in: capacity of cube_ids (3 int32 each), out: count;
 * cube_ids may be NULL to ask for the count */
int opb_volume_last_frame_cubes(opb_volume *v, int32_t *cube_ids, size_t *n_cubes);

int opb_volume_num_cubes(opb_volume *v, size_t *n);
/* CubeHandler::GetCubeMap (CubeHandler.h:339): cube_ids 3 x int32 per cube; voxels 512 x 5 float per cube in
 * the reference's AoS order (sdf, weight, c0, c1, c2), voxel index x + 8y + 64z (VoxelCube.h:56).
 * *n_cubes in: capacity of both arrays (in cubes), out: cubes written.  Either array may be NULL. */
int opb_volume_download(opb_volume *v, int32_t *cube_ids, float *voxels_aos, size_t *n_cubes);
/* CubeHandler::SetCubeMap (CubeHandler.h:344) / ReadFromFile (:40-69): replaces the volume content */
int opb_volume_upload(opb_volume *v, const int32_t *cube_ids, const float *voxels_aos, size_t n_cubes);

/* CubeHandler::ExtractTriangleMesh (CubeHandler.cpp:9-44,70-114; MarchingCube.cpp:9-74):
 * 3 fresh vertices per triangle, xyz/rgb float[3*nv], tri uint32[3*nt] (tri[i] = 3i,3i+1,3i+2 like
 * TriangleMesh::LoadFromMeshes produces).  Buffers are malloc'ed by the library: release with opb_free. */
int opb_volume_extract_mesh(opb_volume *v, float **xyz, float **rgb, uint32_t **tri, size_t *nv, size_t *nt);
/* counts only (no download) */
int opb_volume_count_mesh(opb_volume *v, size_t *nv, size_t *nt);

/* TriangleMesh::ClusteringSimplify(grid_len) (src/Geometry/TriangleMesh.cpp:53-58 -> ClusteringSimplification,
 * src/Geometry/MeshSimplification.cpp:579-657, UpdateMesh :114-139, CompactMesh :314-343): vertex clustering on a hashed grid
 * -- the first vertex (in triangle order) of every grid cell becomes its representative and moves to the mean of the vertex
 * references that fell into the cell, triangles with two corners in one cell are dropped, unreferenced vertices removed.
 * Bit-identical to the reference (the float sums are taken in its order), same vertex and triangle order.  For meshes
 * without normals, which is what ExtractTriangleMesh produces (with normals the reference recomputes them afterwards:
 * call opb_mesh_compute_normals on the result).  Inputs may be host or device pointers; colors may be NULL; outputs are
 * malloc'ed by the library (opb_free).  grid_len <= 0 -> OPB_ERR_INVALID (the reference prints an error and returns the mesh
 * unchanged). */
int opb_mesh_clustering_simplify(int device, const float *points, const float *colors, size_t nv, const uint32_t *triangles, size_t nt,
                                 float grid_len, float **out_points, float **out_colors, uint32_t **out_triangles, size_t *out_nv,
                                 size_t *out_nt);
/* PointCloud::DownSample(grid_len) (src/Geometry/PointCloud.cpp:145-189; SURVEY.md §8f rank 5, first item): voxel-grid
 * down-sampling -- one output point per occupied grid cell, in the order in which the cells are first met, each the mean of
 * the cell's points summed in input order (bit-identical to the reference); colors / normals (optional, 3 floats per point)
 * are averaged alike.  Inputs host or device; outputs malloc'ed by the library (opb_free). */
int opb_pointcloud_downsample(int device, const float *points, const float *colors, const float *normals, size_t n, float grid_len,
                              float **out_points, float **out_colors, float **out_normals, size_t *out_n);
/* TriangleMesh::ComputeNormals (src/Geometry/TriangleMesh.cpp:95-127): unit face normals, vertex normal = normalised sum of
 * the normals of the faces referencing the vertex, summed in the reference's order.  normals: 3 x nv floats, host or device. */
int opb_mesh_compute_normals(int device, const float *points, size_t nv, const uint32_t *triangles, size_t nt, float *normals);
/* ExtractTriangleMesh + ClusteringSimplify(grid_len) as the fusion mains chain them (example/DenseFusion/DenseFusion.cpp:99-105)
 * without the raw mesh leaving the device.  Buffers are malloc'ed by the library: release with opb_free. */
int opb_volume_extract_mesh_clustered(opb_volume *v, float grid_len, float **xyz, float **rgb, uint32_t **tri, size_t *nv, size_t *nt);

/* CubeHandler::Transform (CubeHandler.h:242-298: trilinear, ReadVoxelInterpolate VoxelCube.cpp:6-50) and
 * CubeHandler::TransformNearest (:299-338): resamples the volume under the rigid transform trans into a NEW volume on the
 * same device (*out; release with opb_volume_destroy).  Bit-identical to the reference, including its quirk:
 * TransformNearest does not copy the voxel resolution into its result, whose cubes are therefore allocated -- and later
 * meshed -- with CubePara's default 0.01 while the values are sampled at the source resolution.
 *   result_voxel_resolution  VoxelResolution of the result; <= 0: the source's.  (What the reference does: the source's
 *                            for Transform, 0.01 for TransformNearest.)
 *   result_max_cubes         block-pool capacity of the result; <= 0: sized automatically (grown and retried if needed)
 * The result is never sharded (shard_world = 1). */
int opb_volume_transform(opb_volume *src, const float trans_colmajor[16], int nearest, float result_voxel_resolution,
                         int32_t result_max_cubes, opb_volume **out);
/* CubeHandler::Merge(another) (CubeHandler.h:145-167): cubes missing from dst are copied, voxels of common cubes are
 * combined with TSDFVoxel::operator+ (weighted mean).  Different voxel resolutions -> OPB_ERR_INVALID and no change (the
 * reference prints "[Warning]::[MergeVoxelHash]::Voxel resolution is not identical." and returns).  Merge(another, trans)
 * (:168-177) is opb_volume_transform(another, trans, 0, ...) followed by this call.  Both volumes on one device. */
int opb_volume_merge(opb_volume *dst, opb_volume *another);
/* The reference's ORDER, for callers whose output depends on it.  The reference keeps its cubes in a std::unordered_map, emits the
 * mesh cube by cube in that map's iteration order (CubeHandler.cpp:27-40) and creates the cubes of a resampled volume at their first
 * touch while walking the source in that order (CubeHandler.h:198-241); TriangleMesh::ClusteringSimplify then averages vertices in
 * arrival order, so the bytes of a written PLY depend on the sequence.  The block pool has no such order, so the caller supplies it:
 *   opb_volume_extract_mesh_ordered  the mesh with the cubes taken in the given sequence (every cube of the volume exactly once; ids
 *                                    as 3 x int32), cells inside a cube in GenerateMeshByCube's x / y / z nesting;
 *   opb_volume_transform_ordered     opb_volume_transform plus, in *result_ids_in_order (malloc'ed, opb_free), the result's cubes in
 *                                    the order the reference would have inserted them, given the source's cubes in the source map's
 *                                    iteration order.
 * onepiece_b200/cpp/Integration/CubeHandler.cpp keeps a mirror map of ids with the reference's insertion history to produce these
 * sequences. */
int opb_volume_extract_mesh_ordered(opb_volume *v, const int32_t *cube_ids, size_t n_ids, float **xyz, float **rgb, uint32_t **tri, size_t *nv,
                                    size_t *nt);
int opb_volume_transform_ordered(opb_volume *src, const float trans_colmajor[16], int nearest, float result_voxel_resolution,
                                 int32_t result_max_cubes, const int32_t *src_ids_in_order, size_t n_src_ids, opb_volume **out,
                                 int32_t **result_ids_in_order, size_t *n_result);
/* the descriptor a volume currently runs with (e.g. of a transform result) */
int opb_volume_get_desc(opb_volume *v, opb_volume_desc *out);

/* Marching Cubes over a volume partitioned across GPUs (no counterpart in the single-process
 * reference, whose GenerateMeshByCube reads the +x/+y/+z neighbour cubes from the same map, CubeHandler.cpp:83-99).
 * With sub-volume ownership (shard_* above) the +1 neighbours of the cubes in the last layer of a slab live on the owner
 * of the next slab.  Before extracting its part of the mesh every rank
 *   1. exports the cubes in the FIRST layer of its slabs: ids (3 x int32 per cube) and the one voxel layer Marching Cubes
 *      reads from them (axis coordinate 0; 5 planes sdf, weight, c0, c1, c2 x 64 voxels = 320 floats per cube),
 *   2. sends both buffers to rank-1 and receives those of rank+1 (mod world; the host program's transport -- NCCL
 *      send/recv on the device buffers in onepiece_b200/fusion.py),
 *   3. imports what it received as ghost cubes: visible to the mesh kernels as neighbours only; never integrated,
 *      downloaded or meshed themselves.  Ghosts are dropped by opb_volume_halo_clear and by the next integrate / upload /
 *      clear.
 * ids / layers may be host or device pointers.  Export with ids == layers == NULL only counts (*n). */
int opb_volume_halo_export(opb_volume *v, int32_t *ids, float *layers, size_t cap_cubes, size_t *n);
int opb_volume_halo_import(opb_volume *v, const int32_t *ids, const float *layers, size_t n);
int opb_volume_halo_clear(opb_volume *v);
int opb_volume_num_ghost_cubes(opb_volume *v, size_t *n);
/* The same exchange with the transport inside the kernels (steps 1-3 as two launches, no host program in between):
 *   opb_volume_halo_peer_buffer    this volume's receive box for up to cap_cubes boundary cubes (device memory; created once)
 *                                  and its cudaIpc handle, which the host program hands to the neighbouring ranks
 *                                  (opb_ipc_open maps it there);
 *   opb_volume_halo_peer_attach    dst_buffer = the mapped box of rank-1 (this rank exports into it; dst_cap_cubes = its
 *                                  capacity), src_buffer = the mapped box of rank+1 (this rank acknowledges there).  With two
 *                                  ranks both are the same box.  NULL, 0, NULL detaches;
 *   opb_volume_halo_exchange_peer  collective over the ranks: the export kernel stores ids and layers straight into the
 *                                  destination's box over NVLink and raises a flag there; the import kernel, queued behind it,
 *                                  waits for the flag of rank+1 in its own box, registers the ghosts and acknowledges.  Returns
 *                                  the cubes sent and the ghosts imported.  A peer that never makes the call ends the wait
 *                                  after 4 s with OPB_ERR_CUDA; OPB_ERR_CAPACITY when a box or the block pool is too small.
 *   opb_volume_halo_exchange_begin / _end   the same call in two halves (enqueue / synchronise and read the counters) for a
 *                                  host thread that drives several ranks' volumes itself.
 * No counterpart in the single-process reference (CubeHandler.cpp:83-99 reads the neighbour cubes from its own map). */
int opb_volume_halo_peer_buffer(opb_volume *v, size_t cap_cubes, void **d_buffer, unsigned char ipc_handle[OPB_IPC_HANDLE_BYTES]);
int opb_volume_halo_peer_attach(opb_volume *v, void *dst_buffer, size_t dst_cap_cubes, void *src_buffer);
int opb_volume_halo_exchange_peer(opb_volume *v, size_t *n_sent, size_t *n_imported);
int opb_volume_halo_exchange_begin(opb_volume *v);
int opb_volume_halo_exchange_end(opb_volume *v, size_t *n_sent, size_t *n_imported);
/* One frame uploaded in row bands by the ranks that fuse a stream into a partitioned volume (every rank needs the whole frame
 * to update its own cubes; handing each rank the whole frame from host memory makes N ranks pull N copies through the host):
 *   opb_volume_frame_ring_buffer     this volume's frame ring (two frame slots + flags, device memory) and its cudaIpc handle;
 *   opb_volume_frame_ring_attach     buffers[r] = rank r's ring as mapped here (buffers[rank] = the own ring); NULL detaches;
 *   opb_volume_integrate_rows_async  collective, asynchronous: this rank uploads rows [row0, row0 + n_rows) of the frame (depth_rows /
 *                                    bgr_rows point at the first of those rows in host memory, pinned for overlap), a kernel stores the
 *                                    band into every peer's ring over NVLink and raises a flag there; the frame is integrated (as
 *                                    opb_volume_integrate_async does) once the bands of all ranks have landed.  The bands of the
 *                                    ranks must tile the image.  opb_volume_synchronize waits for the frame;
 *   opb_volume_frame_ring_status     synchronises and reports OPB_ERR_CUDA if a wait on a peer ran into the 4 s limit.
 * No counterpart in the single-process reference (IntegrateImage is handed the frame, CubeHandler.cpp:198-227). */
int opb_volume_frame_ring_buffer(opb_volume *v, void **d_buffer, unsigned char ipc_handle[OPB_IPC_HANDLE_BYTES]);
int opb_volume_frame_ring_attach(opb_volume *v, int rank, int world, void *const *buffers);
int opb_volume_integrate_rows_async(opb_volume *v, const void *depth_rows, int depth_type, const uint8_t *bgr_rows, int row0, int n_rows,
                                    const float pose_cm[16]);
int opb_volume_frame_ring_status(opb_volume *v);

/* ------------------------------------------------------------------------------------------------------
 * Depth pre-filter  (replaces one_piece::tool::ConvertDepthTo32F and tool::BilateralFilter,
 *                    src/Tool/ImageProcessing.cpp:64-91 -- the two calls every fusion main makes right before
 *                    IntegrateImage: example/ImageSequenceIntegration.cpp:36-38, example/DenseFusion/DenseFusion.cpp:92-95)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct opb_prefilter opb_prefilter; /* device workspace for one image size; one host thread at a time */
int opb_prefilter_create(int device, void *stream /* cudaStream_t or NULL */, int width, int height, opb_prefilter **out);
void opb_prefilter_destroy(opb_prefilter *f);
/* converted = ConvertDepthTo32F(depth, depth_scale): float copy, or u16 / depth_scale clamped at 0 (ImageProcessing.cpp:68-91);
 * filtered  = cv::bilateralFilter(converted, d, sigma_color, sigma_space), BORDER_REFLECT_101 -- tool::BilateralFilter calls
 *             it with (range = 7, 0.03, 4.5) (ImageProcessing.cpp:64-67).
 * depth / converted / filtered may be host or device pointers; converted and filtered are optional (NULL: the result only
 * stays in the workspace, see opb_prefilter_device_result).  Unknown depth_type -> OPB_ERR_UNSUPPORTED (the reference exits). */
int opb_prefilter_run(opb_prefilter *f, const void *depth, int depth_type, float depth_scale, int d, double sigma_color,
                      double sigma_space, float *converted, float *filtered);
/* device pointer of the last filtered image (valid until the next run; ordered on the workspace's stream) */
const float *opb_prefilter_device_result(opb_prefilter *f);
int opb_prefilter_synchronize(opb_prefilter *f);
/* ConvertDepthTo32F + BilateralFilter + CubeHandler::IntegrateImage(filtered_depth, rgb, pose) as the fusion mains chain them,
 * without the filtered image leaving the device.  Host buffers, synchronous. */
int opb_volume_integrate_prefiltered(opb_volume *v, opb_prefilter *f, const void *depth, int depth_type, const uint8_t *bgr,
                                     const float pose_colmajor[16], int d, double sigma_color, double sigma_space);

/* ------------------------------------------------------------------------------------------------------
 * ICP  (replaces one_piece::registration::PointToPlane / PointToPoint, src/Registration/ICP.h:23-26,
 *       src/Registration/ICP.cpp:31-224)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct opb_icp opb_icp; /* reusable device workspace; one host thread at a time */

/* registration::ICPParameter (ICP.h:13-19) */
typedef struct
{
    int32_t max_iteration; /* 30  */
    double threshold;      /* 0.2 : max distance of a correspondence */
    double scaling;        /* 1.0 : PointToPlane refuses anything else, PointToPoint scales both clouds */
} opb_icp_params;

/* registration::RegistrationResult (RegistrationResult.h:8-16) */
typedef struct
{
    float T[16];          /* result.T: Kabsch fit over the final inlier pairs of the original clouds (ICP.cpp:221),
                             column-major; NOT the iterated transform */
    float T_iterated[16]; /* start_T after the last iteration (what the loop converged to), column-major */
    double rmse;          /* CountInliers' sqrt(sum_error / inliers) with the final transform */
    size_t n_inliers;     /* correspondence_set_index.size() */
    int32_t iterations;
    int32_t status;       /* OPB_OK, or the error the reference reports by returning a default result */
    size_t n_local_pairs; /* pairs written by this call: n_inliers, or this rank's share of them when the source
                             cloud is split across ranks (opb_icp_comm_attach) */
} opb_icp_result;

void opb_icp_params_default(opb_icp_params *p);
int opb_icp_create(int device, void *stream /* cudaStream_t or NULL */, opb_icp **out);
void opb_icp_destroy(opb_icp *c);
/* PointToPlane(source, target, init_T, params).  xyz arrays are 3 floats per point; they may be host or
 * device pointers.  pairs (optional): up to pairs_cap (source index, target index) int32 pairs, ascending
 * source index like correspondence_set_index.  Returns OPB_ERR_INVALID (and result->status) when the target has
 * no normals or scaling != 1, the case in which the reference prints an error and returns a default result. */
int opb_icp_point_to_plane(opb_icp *c, const float *src_xyz, size_t ns, const float *tgt_xyz, const float *tgt_normals,
                           size_t nt, const float init_T_colmajor[16], const opb_icp_params *params,
                           opb_icp_result *result, int32_t *pairs, size_t pairs_cap);
int opb_icp_point_to_point(opb_icp *c, const float *src_xyz, size_t ns, const float *tgt_xyz, size_t nt,
                           const float init_T_colmajor[16], const opb_icp_params *params, opb_icp_result *result,
                           int32_t *pairs, size_t pairs_cap);
/* ---- geometry::PointCloud resident on the device (src/Geometry/PointCloud.h:52-54) ----
 * The reference's callers back-project every frame on the host (PointCloud::LoadFromDepth, src/Geometry/PointCloud.cpp:72-100)
 * and hand the cloud to ICP, which copies it again (ICP.cpp:150-151); a frame is the source of one registration and the target
 * of the next, and is then integrated.  An opb_cloud uploads the depth (and colour) image once, builds the cloud on the device
 * with the reference's arithmetic (bit-identical points, raster order, z > 0 only) and keeps images, points and normals in
 * HBM for all three uses.  Calls that take images or arrays only ENQUEUE on the cloud's own stream: host buffers must stay
 * valid until opb_cloud_size (or any call that consumes the cloud) has returned; pinned host memory makes the copies truly
 * asynchronous, so the next frame can be loaded while the current one is being registered. */
typedef struct opb_cloud opb_cloud;
int opb_cloud_create(int device, void *stream /* cudaStream_t or NULL: own stream */, opb_cloud **out);
void opb_cloud_destroy(opb_cloud *c);
/* PointCloud::LoadFromDepth(depth, camera); bgr (optional, W*H*3) is kept with the depth image for opb_volume_integrate_cloud.
 * depth / bgr: host or device pointers. */
int opb_cloud_load_from_depth(opb_cloud *c, const void *depth, int depth_type, const uint8_t *bgr, float fx, float fy, float cx,
                              float cy, int width, int height, float depth_scale);
/* an arbitrary cloud / its normals from arrays (3 floats per point, host or device) */
int opb_cloud_set_points(opb_cloud *c, const float *xyz, size_t n);
int opb_cloud_set_normals(opb_cloud *c, const float *normals, size_t n);
/* waits for everything enqueued on the cloud; number of points */
int opb_cloud_size(opb_cloud *c, size_t *n);
/* points / normals back to the host (either may be NULL) */
int opb_cloud_download(opb_cloud *c, float *xyz, float *normals);
/* PointToPlane / PointToPoint on device-resident clouds: the buffers are used where they lie, nothing is copied (PointToPoint
 * with scaling != 1 works on scaled copies like the reference).  Everything else as opb_icp_point_to_plane / _point. */
int opb_icp_point_to_plane_clouds(opb_icp *c, opb_cloud *source, opb_cloud *target, const float init_T_colmajor[16],
                                  const opb_icp_params *params, opb_icp_result *result, int32_t *pairs, size_t pairs_cap);
int opb_icp_point_to_point_clouds(opb_icp *c, opb_cloud *source, opb_cloud *target, const float init_T_colmajor[16],
                                  const opb_icp_params *params, opb_icp_result *result, int32_t *pairs, size_t pairs_cap);
/* CubeHandler::IntegrateImage(depth, rgb, pose) with the images the cloud was loaded from (already in HBM); synchronous, grows
 * the pool like opb_volume_integrate */
int opb_volume_integrate_cloud(opb_volume *v, opb_cloud *frame, const float pose_colmajor[16]);

/* PointCloud::EstimateNormals(radius = 0.1, knn = 30) (src/Geometry/PointCloud.cpp:102-144), the step in front of PointToPlane
 * when the clouds come without normals (example/ICPTest.cpp:27-33): per point the knn nearest points (the point itself
 * included) whose SQUARED distance does not exceed `radius` (the reference's KnnRadiusSearch compares dist^2 with radius,
 * KDTree.h:248-254), geometry::FitPlane on them (Geometry.cpp:172-199: float mean and covariance, Eigen JacobiSVD, third
 * column of U, normalised); fewer than 3 points -> the zero vector.  Bit-identical to the reference when no two neighbours of
 * a point are at exactly the same distance; exact ties are ordered by index here and by k-d tree traversal there
 * (opb_kdtree_estimate_normals below follows the reference's order and is what the drop-ins call).
 * xyz / normals: 3 floats per point, host or device pointers; knn <= 64. */
int opb_icp_estimate_normals(opb_icp *c, const float *xyz, size_t n, float radius, int knn, float *normals);
/* ---- geometry::KDTree<3> with the reference's visiting order, and what the reference builds on it ----
 * The reference's searches go through its vendored, modified nanoflann (src/Geometry/KDTree.h:60-262).  Which neighbours come
 * back and in which order depends on the tree, not only on the distances: equal distances are ordered by the traversal, the
 * radius search stops after (size_t)(2.5 k) hits in traversal order and sorts them with std::sort (unstable).  The float sums
 * downstream depend on that order, so the device builds the same tree (nanoflann.hpp:843-1010, leaf size 10) and walks it the
 * same way (:1354-1417); results are identical to the reference's, ties and early stops included. */
typedef struct opb_kdtree opb_kdtree;
int opb_kdtree_create(int device, opb_kdtree **out);
void opb_kdtree_destroy(opb_kdtree *t);
/* KDTree::BuildTree (KDTree.h:84-92): xyz = 3 floats per point (finite), host or device pointer; copied. */
int opb_kdtree_build(opb_kdtree *t, const float *xyz, size_t n);
/* mode 0: KnnSearch (KDTree.h:176-195), k <= 64.  mode 1: RadiusSearch(radius, max_result = k) (:125-143; `radius` is compared
 * with SQUARED distances), k <= 409.  mode 2: KnnRadiusSearch(k, radius) (:230-256), k <= 64.  queries: 3 floats each; outputs
 * are nq rows of k entries, padded with -1 behind out_count[q] results; host or device pointers. */
int opb_kdtree_search(opb_kdtree *t, const float *queries, size_t nq, int mode, int k, float radius, int32_t *out_index,
                      float *out_dist, int32_t *out_count);
/* PointCloud::EstimateNormals(radius, knn) (PointCloud.cpp:102-144) over the tree's points; knn <= 64.  Bit-identical to the
 * reference on every cloud, tied distances included. */
int opb_kdtree_estimate_normals(opb_kdtree *t, float radius, int knn, float *normals);
/* registration::ComputeFPFHFeature(pcd, features, knn = 100, radius = 0.1) (src/Registration/3DFeature.cpp:83-131) over the
 * tree's points with the given normals: 33 floats per point.  Points without neighbours get NaN rows like the reference's
 * (0 * inf).  The descriptor angle goes through a double atan2 rounded to float as in the reference; the device's atan2 may
 * differ from glibc's in the last place of the DOUBLE, which survives the float rounding about once in 1e8 pairs. knn <= 256. */
int opb_kdtree_fpfh(opb_kdtree *t, const float *normals, int knn, float radius, float *features33);
/* registration::FeatureMatching3D (src/Registration/GlobalRegistration.cpp:29-73): for every source feature (33 floats) the nearest
 * target feature; pairs = (source index, target index), 2 * ns ints of room; a NaN source feature matches nothing and is left
 * out, like the reference's empty KnnSearch result.  Answered the reference's way: the device builds the KDTree<33> nanoflann
 * would build over the targets and walks it per source, so exactly equidistant targets (duplicate descriptors on flat
 * surfaces) come back in the reference's order -- identical index for index.  One exception: if a TARGET row is NaN or
 * infinite (descriptor of an isolated point), nanoflann's bounding boxes are poisoned and the reference prunes real
 * neighbours (a few percent of its answers change); such sets are scanned exhaustively here and every source gets its true
 * nearest finite target (lowest index on exact ties).  `t` lends its stream and buffers; its own tree is untouched. */
int opb_kdtree_feature_matching(opb_kdtree *t, const float *src_feat33, size_t ns, const float *tgt_feat33, size_t nt, int32_t *pairs,
                                size_t *n_pairs);
/* registration::RejectMatchesRanSaPC(source_points, target_points, engine, init_matches, candidate_num = 4, difference = 0.1)
 * (GlobalRegistration.cpp:75-108), one call: keeps a match if one of candidate_num randomly drawn other matches preserves the
 * distance to it within `difference`.  *engine_state is the std::default_random_engine (libstdc++: minstd_rand0) the reference
 * threads through its three calls: 1 for a default-constructed engine, updated on return.  pairs / *n_pairs in place.  HOST
 * code and host pointers: the draws form one sequential data-dependent chain; identical to the reference index for index. */
int opb_reject_matches(const float *src_xyz, size_t ns, const float *tgt_xyz, size_t nt, int32_t *pairs, size_t *n_pairs,
                       uint32_t *engine_state, int candidate_num, float difference);
/* geometry::EstimateRigidTransformationRANSAC(correspondence_set, inliers, inlier_ids, max_iteration, threshold)
 * (src/Geometry/Ransac.cpp:7-41 over 3rdparty/GRANSAC/GRANSAC.hpp:71-131 and TransformationModel.hpp:28-96), the last step of
 * RansacRegistration (GlobalRegistration.cpp:252): max_iteration hypotheses, each the rigid motion of eight distinct pairs
 * (EstimateRigidTransformation in float, operation for operation), scored by the number of pairs with |R a + t - b| < threshold;
 * the first hypothesis with the strictly largest score wins; T = its eight-point motion (column-major), inlier_ids = the pairs
 * it explains, ascending (the reference lists them in its shuffled order), room for n.  The reference draws its samples from
 * engines seeded by std::random_device, so its own result differs from run to run; here hypothesis h takes its sample from a
 * counter-based generator keyed by (seed, h), or from forced_samples (max_iteration x 8 pair indices) when that is not NULL --
 * with the same samples the winner, its motion and its inliers are the reference's bit for bit.  n < 8: T = 0, no inliers
 * (the reference's warning path); n == 8, or no hypothesis with an inlier: OPB_ERR_INVALID (the reference crashes).
 * best_iteration / best_sample (8 ints) may be NULL.  `t` lends its stream and buffers. */
int opb_ransac_rigid_transformation(opb_kdtree *t, const float *src_xyz, const float *tgt_xyz, size_t n, int max_iteration, double threshold,
                                    uint64_t seed, const int32_t *forced_samples, float T_colmajor[16], int32_t *inlier_ids, size_t *n_inliers,
                                    int32_t *best_iteration, int32_t best_sample[8]);
/* optimization::SimpleBA(correspondences, poses, max_iteration = 5) = Optimizer::FastBA (src/Optimization/SimpleBA.cpp:80-157,
 * Optimizer.h:22-25), the pose-graph refinement over submaps: frame pair k links poses src_id[k] -> tgt_id[k] (tgt_id >= 1: the
 * reference indexes block tgt_id - 1) through the point pairs [offset[k], offset[k+1]) of a_xyz / b_xyz (points in the two
 * frames' own coordinates); poses are camera-to-world 4x4 column-major floats, refined in place, pose 0 fixed.  The per-pair
 * normal-equation blocks are reduced on the device (28 sums per pair, float products, double accumulation), the 6 (n_poses - 1)
 * unknowns are solved on the host (dense LDL^T in double; the reference: float sums, SimplicialLDLT<float>), so poses agree
 * with the reference to ~1e-4, not bit for bit.  n_poses < 3: nothing to do (OPB_OK); n_corr < n_poses - 1: OPB_ERR_INVALID
 * (the reference prints "unconnected components" and returns).  NOTE: not yet exercised on a GPU (see csrc/opb_ba.cu). */
int opb_simple_ba(int device, int n_poses, float *poses_colmajor, int n_corr, const int32_t *src_id, const int32_t *tgt_id,
                  const int64_t *offset, const float *a_xyz, const float *b_xyz, int max_iteration);
/* test hook, host only: one iteration's assembly + solve + pose update from per-pair sums (28 doubles each: N, sum q1, sum q2,
 * upper triangles of sum q1 q1^T and sum q2 q2^T, sum q1 q2^T row-major; q = pose applied to the point) */
int opb_simple_ba_from_sums(int n_poses, float *poses_colmajor, int n_corr, const int32_t *src_id, const int32_t *tgt_id,
                            const double *sums28);
/* test hook: the built tree (vind: n ints; nodes in allocation order, root = 0: left, right, child1, child2 (-1 = leaf), divfeat;
 * divlow, divhigh); all output pointers NULL -> only n_nodes */
int opb_kdtree_dump(opb_kdtree *t, int32_t *vind, int32_t *node_ints5, float *node_floats2, float root_box[6], size_t *n_nodes);

/* One registration over several GPUs (no counterpart in the reference).  The source points are split
 * across the ranks, every rank holds the whole target; each iteration every rank reduces its share to the 30-scalar packet
 * (21 J^T J + 6 J^T r + 2 error + count) and the packets are summed across ranks inside the reduction kernel's last CTA through
 * peer memory: plain stores into every peer's mailbox over NVLink, a flag, a bounded wait, a fixed-order sum.  All ranks then
 * solve the identical 6x6 system, so T, rmse and n_inliers of the result are global and identical on every rank; pairs holds
 * the rank's own inliers with LOCAL source indices (n_local_pairs of them).  Setup, once per workspace:
 *   1. opb_icp_comm_buffer  -> this rank's mailbox: device pointer and cudaIpc handle (64 bytes to send to the peers)
 *   2. peers in other processes open the handle with opb_ipc_open; workspaces of the same process use the pointer itself
 *   3. opb_icp_comm_attach(rank, world, buffers) on every rank, buffers[r] = mailbox of rank r as visible from this process
 * Afterwards opb_icp_point_to_plane / opb_icp_point_to_point are COLLECTIVE calls: every rank must make them in the same
 * order with its share of the source and identical target, init_T and params (a missing peer makes the call fail with
 * OPB_ERR_CUDA after 4 s instead of hanging).  A rank's share may be empty. */
/* Pair list off the critical path.  A fusion loop needs the pose of a registration before it can integrate the frame, the inlier
 * pairs (8 bytes per source point) only afterwards.  With async pairs on, a registration whose `pairs` buffer is page-locked host
 * memory returns as soon as pose, rmse and counters are on the host; the list follows on the workspace's copy stream, under
 * whatever the caller enqueues next, and is complete after opb_icp_wait_pairs (also implied by the next registration on the
 * workspace).  The first n_local_pairs entries are the result, as in the synchronous call.  Off by default; ignored for pageable
 * buffers and for workspaces that exchange packets with peer ranks. */
int opb_icp_set_async_pairs(opb_icp *c, int on);
int opb_icp_wait_pairs(opb_icp *c);
/* Sizes the workspace for clouds of up to n_source / n_target points now instead of inside the first registration.  Device
 * allocation can wait for running kernels; a workspace that takes part in a split registration with a peer on the SAME device (a
 * peer's kernel spins until this workspace has launched its own) must not allocate in the middle of a collective call. */
int opb_icp_reserve(opb_icp *c, size_t n_source, size_t n_target);
int opb_icp_comm_buffer(opb_icp *c, void **d_buffer, unsigned char ipc_handle[OPB_IPC_HANDLE_BYTES]);
int opb_ipc_open(int device, const unsigned char ipc_handle[OPB_IPC_HANDLE_BYTES], void **d_ptr);
int opb_ipc_close(int device, void *d_ptr);
int opb_icp_comm_attach(opb_icp *c, int rank, int world, void *const *buffers);
int opb_icp_comm_detach(opb_icp *c);
/* nearest-neighbour index per source point from the last search of the previous call -- the search of the LAST ITERATION, under
 * the pose before its update; -1: none that could pass the closing inlier test -- for parity tests against KDTree::KnnSearch */
int opb_icp_last_nn(opb_icp *c, int32_t *nn, size_t n);
/* the pose those neighbours were searched under: start_T BEFORE the last iteration's update (column-major).  The reference's
 * closing CountInliers re-tests exactly these neighbours with the final pose and runs no search of its own (ICP.cpp:90,206) */
int opb_icp_last_prev_pose(opb_icp *c, float T_colmajor[16]);
/* how many exact grid searches the last call performed, of n_source * (max_iteration + 1) queries: the others kept the
 * certified nearest neighbour of an earlier pass (see csrc/opb_icp.cu, icp_certify_kernel) */
int opb_icp_last_search_count(opb_icp *c, uint64_t *full_searches);
/* kernels the last call launched (grid construction 8, the pass loop 1 as a persistent launch or 3 per pass, final sums 2, ...) */
int opb_icp_last_launch_count(opb_icp *c, int *launches);
/* the same per pass (pass 0 searches everything), for the first min(cap, 64) passes */
int opb_icp_last_search_trace(opb_icp *c, uint32_t *per_pass, int cap);
/* CUDA-event timing of the last call when enabled: grid construction and the iteration loop */
int opb_icp_set_profiling(opb_icp *c, int on);
int opb_icp_last_timing(opb_icp *c, float *grid_build_ms, float *iterations_ms);
/* per-pass time stamps of the persistent loop kernel of the last call (globaltimer ns), 8 per pass for the first min(passes, 48)
 * passes.  CTA 0: [0] pass start, [1] own points certified / searched, [2] accumulated + partial published, [3] released into
 * the next pass; the CTA that finished last: [4] knows it is last, [5] partials summed, [6] solved + pose updated; [7] unused */
int opb_icp_last_stamps(opb_icp *c, uint64_t *stamps, int passes);


/* ------------------------------------------------------------------------------------------------------
 * Dense RGB-D odometry  (replaces one_piece::odometry::Odometry::DenseTracking, src/Odometry/Odometry.h:78-84,
 *                        src/Odometry/Odometry.cpp:436-685, src/Odometry/DenseOdometryFunction.cpp)
 * ---------------------------------------------------------------------------------------------------- */
#define OPB_ODO_MAX_LEVELS 6
#define OPB_ODO_MAX_TRACE 64

typedef struct opb_odometry opb_odometry; /* Odometry object: camera, schedule, device workspace; one host thread at a time */
typedef struct opb_frame opb_frame;       /* geometry::RGBDFrame's dense-tracking cache on the device (RGBDFrame.h:33-44) */

typedef struct
{
    /* camera::PinholeCamera of level 0 (Odometry::SetCamera, Odometry.h:86-89); levels halve it (Camera.h:38-42) */
    float fx, fy, cx, cy;
    int32_t width, height; /* both divisible by 2^(levels-1) */
    float depth_scale;
    /* Odometry::SetMultiScale (Odometry.h:101-105): multi_scale_level and iter_count_per_level, indexed by LEVEL
     * (level 0 = full resolution runs last); reference default 3 levels, {4, 8, 16} */
    int32_t levels;
    int32_t iterations[OPB_ODO_MAX_LEVELS];
    int32_t device;
    void *stream; /* optional cudaStream_t */
} opb_odometry_desc;

/* odometry::DenseTrackingResult (Odometry.h:29-37) plus a per-iteration trace for parity tests */
typedef struct
{
    float T[16];               /* source -> target, column-major */
    double rmse;               /* ComputeReprojectionError3D over correspondence_set (Odometry.cpp:606) */
    int32_t tracking_success;  /* correspondences / (W*H) >= MIN_INLIER_RATIO_DENSE (Odometry.cpp:684) */
    int32_t status;
    size_t n_correspondences;  /* pixel_correspondence_set.size() */
    int32_t iterations;        /* executed solver iterations (a level ends early above MAX_INLIER_RATIO_DENSE) */
    int32_t corr_per_iteration[OPB_ODO_MAX_TRACE];
    float T_per_iteration[OPB_ODO_MAX_TRACE][16];
} opb_tracking_result;

void opb_odometry_desc_default(opb_odometry_desc *desc);
int opb_odometry_create(const opb_odometry_desc *desc, opb_odometry **out);
void opb_odometry_destroy(opb_odometry *o);
int opb_odometry_set_profiling(opb_odometry *o, int on);
/* How MultiScaleComputing (Odometry.cpp:621-685) is launched: 1 (default) the second persistent form -- one cooperative launch for all
 * levels and iterations, the 29 sums as 8x8 outer-product accumulations on the FP64 tensor-core op; 2 the first persistent form; 0 one
 * launch pair per iteration.  All forms sum exact (double) products of the float Jacobian rows in double; they differ only in the
 * order of the additions (~1e-16 relative).  -1 restores the default.  For A/B runs and tests. */
int opb_odometry_set_loop_form(opb_odometry *o, int form);
/* CUDA-event time of the last tracking call (pre-processing of new frames included) and the mean time per solver
 * iteration that the last CTA spent on the fixed-order partial sum + 6x6 solve + pose update (%globaltimer) */
int opb_odometry_last_timing(opb_odometry *o, float *tracking_ms, float *solve_tail_us);
/* persistent solver loop of the last tracking call, seen from CTA 0 (%globaltimer ns summed over the iterations): candidate
 * pass, mid-iteration grid barrier, reduction (+ solve when CTA 0 finished last), wait for the release into the next iteration */
int opb_odometry_last_phases(opb_odometry *o, uint64_t phase_ns[4]);

/* geometry::RGBDFrame(rgb, depth) (RGBDFrame.h:14-19): uploads the raw images (host or device pointers; returns when
 * the copies are done).  depth_type other than OPB_DEPTH_F32 / OPB_DEPTH_U16 -> OPB_ERR_UNSUPPORTED (the reference
 * prints "Unknown depth image type" and exits, DenseOdometryFunction.cpp:51-55). */
int opb_frame_create(opb_odometry *o, const uint8_t *bgr, const void *depth, int depth_type, opb_frame **out);
void opb_frame_destroy(opb_frame *f);
/* InitializeRGBDDenseTracking + CreateImagePyramid (Odometry.cpp:571-587,609-620,436-449); idempotent like
 * RGBDFrame::IsPreprocessedDense */
int opb_frame_preprocess(opb_odometry *o, opb_frame *f);
int opb_frame_is_preprocessed(const opb_frame *f);
/* download one cached image: what = 0 gray, 1 depth32f, 2 gray dx, 3 gray dy, 4 depth dx, 5 depth dy */
int opb_frame_image(opb_odometry *o, opb_frame *f, int what, int level, float *out);

/* DenseTracking(RGBDFrame &source, RGBDFrame &target, initial_T, term_type) (Odometry.cpp:526-608).  Both frames are
 * MUTATED like the reference's: caches are filled on first use and the level-0 gray image is re-normalised in place on
 * every call.  term_type 0 hybrid / 1 photometric / 2 geometric.  Optional outputs, up to pairs_cap entries each:
 *   pixel_pairs  4 x uint32 per correspondence (v_s, u_s, v_t, u_t), raster order of the source
 *   corr_xyz     6 x float  per correspondence: correspondence_set[i].first / .second (Odometry.cpp:672-682) */
int opb_odometry_dense_tracking_frames(opb_odometry *o, opb_frame *source, opb_frame *target, const float init_T_colmajor[16],
                                       int term_type, opb_tracking_result *result, uint32_t *pixel_pairs, size_t pairs_cap,
                                       float *corr_xyz);
/* DenseTracking(source_color, target_color, source_depth, target_depth, initial_T, term_type) (Odometry.cpp:463-523):
 * intensities are normalised BEFORE the pyramids are built (unlike the RGBDFrame overload). */
int opb_odometry_dense_tracking(opb_odometry *o, const uint8_t *src_bgr, const uint8_t *tgt_bgr, const void *src_depth,
                                const void *tgt_depth, int depth_type, const float init_T_colmajor[16], int term_type,
                                opb_tracking_result *result, uint32_t *pixel_pairs, size_t pairs_cap, float *corr_xyz);
/* One DoSingleIteration{,PhotoTerm,DepthTerm} (DenseOdometryFunction.cpp:382-475) at a pyramid level from a given pose
 * (teacher forcing, for parity tests): T in/out; sums43 = J^T J (36, row-major), J^T r (6), sum r^2; the
 * correspondence list of that iteration. */
int opb_odometry_single_iteration(opb_odometry *o, opb_frame *source, opb_frame *target, int level, float T_inout_colmajor[16],
                                  int term_type, double sums43[43], uint32_t *pixel_pairs, size_t pairs_cap, size_t *n_pairs);

#ifdef __cplusplus
}
#endif
#endif /* ONEPIECE_B200_H */
