/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle.  Never linked into, imported by or called from the product
 * (onepiece_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and there only as the checker or as the CPU baseline.
 *
 * Plain-C restatement of the reference's algorithm for the hot path (file:line relative to the reference
 * tree are given at every function in opb_oracle.c).  It is pinned, bit for bit on the integration and
 * Marching Cubes paths, against the reference's own translation units compiled unmodified
 * (oracle/_ref, see oracle/Makefile) by tests/test_oracle_vs_ref.py and against the committed golden
 * fixtures in tests/golden/ (generated from oracle/_ref by tests/golden/gen_golden.py).
 */
#ifndef OPB_ORACLE_H
#define OPB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_volume orc_volume;

orc_volume *orc_volume_create(float fx, float fy, float cx, float cy, int width, int height, float depth_scale,
                              float voxel_resolution, float truncation, float near_plane, float far_plane);
void orc_volume_destroy(orc_volume *v);
void orc_volume_clear(orc_volume *v);
/* Eigen Matrix4f inverse / frustum planes as the reference evaluates them (column-major in and out) */
void orc_pose_inverse(const float *pose_cm, float *inv_cm);
void orc_frustum_planes(const orc_volume *v, const float *pose_cm, float *planes24);
int orc_frustum_contains(const float *planes24, float x, float y, float z);
/* CubeHandler::ComputeBounding */
void orc_volume_bounding(const orc_volume *v, const void *depth, int is_u16, const float *pose_cm, float *max3,
                         float *min3);
/* Integrator::GetSDF at n points */
void orc_volume_get_sdf(const orc_volume *v, const void *depth, int is_u16, const float *pose_cm, const float *points,
                        long n, float *sdf);
/* CubeHandler::PrepareCubes: allocates, returns the list (3 ints per cube) in the reference's loop order */
long orc_volume_prepare_cubes(orc_volume *v, const void *depth, int is_u16, const float *pose_cm, int32_t *ids, long cap);
/* CubeHandler::IntegrateImage; returns the number of cubes listed for the frame */
long orc_volume_integrate(orc_volume *v, const void *depth, int is_u16, const uint8_t *bgr, const float *pose_cm);
long orc_volume_num_cubes(const orc_volume *v);
/* ids n*3, voxels n*512*5 (sdf, weight, c0, c1, c2), insertion order */
void orc_volume_download(const orc_volume *v, int32_t *ids, float *voxels);
void orc_volume_upload(orc_volume *v, const int32_t *ids, const float *voxels, long n);
/* CubeHandler::ExtractTriangleMesh: returns the vertex count (3 per triangle); buffers are malloc'ed */
long orc_volume_extract_mesh(const orc_volume *v, float **xyz, float **rgb);
/* integration::MarchingCube on one cell: returns the vertex count, writes up to 15 xyz / rgb triples */
int orc_marching_cube_cell(const float *corners24, const float *sdf8, const float *colors24, float *xyz, float *rgb);
void orc_free(void *p);

/* ---- registration (src/Registration/ICP.cpp) ------------------------------------------------------------
 * Point clouds are 3 floats per point.  Per-point arithmetic (transform, inlier test, Jacobian rows, nearest
 * neighbour distances) is float32 in the reference's operation order; the 6x6 / Kabsch sums are accumulated in
 * double and solved in double (the reference accumulates sequentially in float32 -- its own deviation from its
 * -DUSING_FLOAT64 build is the noise floor), so poses are pinned to oracle/_ref within a
 * tolerance, index sets exactly. */
/* exact nearest neighbour (nanoflann L2_Simple distance, ties -> lowest index) of every query; brute force */
void orc_nearest(const float *query, long nq, const float *target, long nt, int32_t *nn);
/* registration::PointToPlane (tgt_nrm != NULL) / PointToPoint (NULL).  T in/out column-major.  out_T = the
 * Kabsch fit the reference returns as result.T; out_T_iter = the iterated transform.  pairs: 2 ints per inlier.
 * Returns the inlier count, or -1 for the reference's "default result" error path. */
long orc_icp(const float *src, long ns, const float *tgt, const float *tgt_nrm, long nt, const float *init_T_cm,
             int max_iteration, double threshold, double scaling, double *out_T_cm, double *out_T_iter_cm,
             int32_t *pairs, double *rmse);
/* geometry::Se3ToSE3 (Sophus SE3 exp), double, column-major out */
void orc_se3_exp(const double *x6, double *T_cm);

/* ---- dense RGB-D odometry (src/Odometry/Odometry.cpp:436-685, DenseOdometryFunction.cpp) -----------------
 * The five OpenCV filters the path calls (cvtColor RGB2GRAY, GaussianBlur 3x3, pyrDown, Sobel 3x3; OpenCV is an
 * un-vendored dependency, README pins 3.4) are DEFINED here by explicit formulas (opb_oracle.c) and cross-checked
 * against cv2 4.13 in tests/test_odometry_filters.py: parity at that boundary is unpinned by the reference. */
void orc_gray_u8(const uint8_t *bgr, int n_pixels, uint8_t *gray);         /* cvtColor(CV_RGB2GRAY) on the stored bytes */
void orc_blur3(const float *src, int w, int h, float *dst);                /* GaussianBlur(3x3, sigma 0), REFLECT_101 */
void orc_pyr_down(const float *src, int w, int h, float *dst);             /* pyrDown to (w/2, h/2) */
void orc_sobel3(const float *src, int w, int h, int dx, float *dst);       /* Sobel(CV_32F, dx, 1-dx, ksize 3) */

typedef struct orc_frame orc_frame; /* geometry::RGBDFrame's dense cache: gray, depth32f, 6 pyramids */
orc_frame *orc_frame_create(const uint8_t *bgr, const void *depth, int is_u16, int w, int h);
void orc_frame_destroy(orc_frame *f);
/* what: 0 gray pyramid, 1 depth pyramid, 2 gray dx, 3 gray dy, 4 depth dx, 5 depth dy; returns w*h of the level */
long orc_frame_image(const orc_frame *f, int what, int level, float *out);

typedef struct
{
    double T[16];            /* column-major */
    double rmse;
    int tracking_success;
    long n_correspondences;
    int iterations;
    long corr_per_iteration[64];
    double T_per_iteration[64][16];
} orc_tracking_result;

/* Odometry::DenseTracking(RGBDFrame &source, RGBDFrame &target, initial_T, term_type) (Odometry.cpp:526-608): frames are
 * pre-processed on first use and their level-0 gray is re-normalised IN PLACE on every call, like the reference.
 * pixel_pairs (optional): 4 uint32 per correspondence (v_s, u_s, v_t, u_t), raster order of the source. */
void orc_dense_tracking_frames(orc_frame *source, orc_frame *target, float fx, float fy, float cx, float cy, float depth_scale,
                               const float *init_T_cm, int term_type, orc_tracking_result *out, uint32_t *pixel_pairs, long cap);
/* Odometry::DenseTracking(cv::Mat overload) (Odometry.cpp:463-523): normalises before building the pyramids */
void orc_dense_tracking(const uint8_t *src_bgr, const uint8_t *tgt_bgr, const void *src_depth, const void *tgt_depth, int is_u16,
                        int w, int h, float fx, float fy, float cx, float cy, float depth_scale, const float *init_T_cm,
                        int term_type, orc_tracking_result *out, uint32_t *pixel_pairs, long cap);
/* ComputeCorrespondencePixelWise (DenseOdometryFunction.cpp:72-128) on two NaN-masked depth maps; returns the count */
long orc_correspondences(const float *src_depth, const float *tgt_depth, int w, int h, float fx, float fy, float cx, float cy,
                         const float *T_cm, uint32_t *pixel_pairs, long cap);

/* teacher-forced single iteration at a pyramid level (intrinsics of level 0); sums43 = 36 JTJ, 6 JTr, sum r^2 */
long orc_single_iteration(orc_frame *source, orc_frame *target, int level, float fx, float fy, float cx, float cy, float *T_cm,
                          int term_type, double *sums43, uint32_t *pixel_pairs, long cap);
void orc_frame_preprocess(orc_frame *f, float depth_scale);


/* CubeHandler::Transform / TransformNearest (CubeHandler.h:242-338) -> new volume; CubeHandler::Merge (:145-167) */
orc_volume *orc_volume_transform(const orc_volume *v, const float *trans_cm, int nearest, float alloc_res);
int orc_volume_merge(orc_volume *v, const orc_volume *other);
float orc_volume_resolution(const orc_volume *v);

/* TriangleMesh::ClusteringSimplify (TriangleMesh.cpp:53-58, MeshSimplification.cpp:579-657,114-139,314-343) in place; 0 or -1 */
int orc_clustering_simplify(float *points, float *colors, long *n_points, uint32_t *tri, long *n_tris, float grid_len);
/* TriangleMesh::ComputeNormals (TriangleMesh.cpp:95-127) */
void orc_compute_normals(const float *points, long n_points, const uint32_t *tri, long n_tris, float *normals);

/* PointCloud::DownSample (PointCloud.cpp:145-189); returns the number of output points */
long orc_downsample(const float *points, const float *colors, const float *normals, long n, float grid_len, float *out_points,
                    float *out_colors, float *out_normals);

/* PointCloud::EstimateNormals (PointCloud.cpp:102-144): knn nearest points within sqrt(radius), FitPlane, Eigen JacobiSVD */
void orc_estimate_normals(const float *pts, long n, float radius, int knn, float *normals);

/* geometry::KDTree<3> searches (KDTree.h:93-256 over the vendored nanoflann): mode 0 KnnSearch, 1 RadiusSearch, 2 KnnRadiusSearch;
 * nq rows of `cap` entries, -1 padded */
void orc_kdtree_search(const float *pts, long n, const float *queries, long nq, int mode, int k, float radius, long cap,
                       int32_t *out_index, float *out_dist, int32_t *out_count);
long orc_kdtree_dump(const float *pts, long n, int32_t *vind, int32_t *node_ints, float *node_floats, float *root_box);
/* registration::ComputeFPFHFeature (3DFeature.cpp:83-131): n x 33 floats; -1 if std::sort's heap fallback would be needed */
int orc_fpfh(const float *pts, const float *normals, long n, int knn, float radius, float *features);

/* registration::FeatureMatching3D (GlobalRegistration.cpp:29-73): (source, target) index pairs, returns their number */
long orc_feature_matching(const float *src_feat, long ns, const float *tgt_feat, long nt, int32_t *pairs);
/* registration::RejectMatchesRanSaPC (GlobalRegistration.cpp:75-108) applied `rounds` times with one default-seeded engine, in place */
long orc_reject_matches(const float *src, const float *tgt, int32_t *pairs, long n, int rounds, int candidate_num, float difference);

/* geometry::EstimateRigidTransformation in the float build, bit for bit (Geometry.cpp:107-151); T row-major 4x4 */
void orc_kabsch_f32(const float *a, const float *b, long n, float *T_rowmajor);
/* one hypothesis / the whole selection of geometry::EstimateRigidTransformationRANSAC with given samples (Ransac.cpp:7-41) */
long orc_ransac_hypothesis(const float *a, const float *b, long n, const int32_t *sample8, double threshold, float *T_rowmajor, uint8_t *inlier);
long orc_ransac_select(const float *a, const float *b, long n, const int32_t *samples, long iterations, double threshold, float *T_rowmajor,
                       uint8_t *inlier, long *best_count);

/* optimization::SimpleBA = Optimizer::FastBA (SimpleBA.cpp:18-157): per-pair blocks (156 floats) and the whole refinement; poses
 * column-major 4x4 in place; pinned by tolerance (the solve is dense double here, SimplicialLDLT<float> there) */
void orc_ba_blocks(const float *pose_s_cm, const float *pose_t_cm, const float *a, const float *b, long n, float *out156);
int orc_simple_ba(int n_poses, float *poses_cm, int n_corr, const int32_t *src_id, const int32_t *tgt_id, const int64_t *offset,
                  const float *a, const float *b, int max_iteration);

/* caller-side depth pre-filter: tool::ConvertDepthTo32F (ImageProcessing.cpp:68-91), tool::BilateralFilter (:64-67) */
void orc_convert_depth_32f(const void *depth, int is_u16, long n, float depth_scale, float *out);
void orc_bilateral_filter(const float *src, int w, int h, int d, double sigma_color, double sigma_space, float *dst);

#ifdef __cplusplus
}
#endif
#endif
