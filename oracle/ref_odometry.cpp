// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C entry points around the reference's dense-odometry translation unit, compiled unmodified
// (src/Odometry/DenseOdometryFunction.cpp): ComputeCorrespondencePixelWise (:72-128), NormalizeIntensity (:129-144),
// DoSingleIteration{,PhotoTerm,DepthTerm} (:382-475), plus geometry::TransformToMatXYZ (Geometry.cpp:72-106) and
// ComputeReprojectionError3D (:45-59).  src/Odometry/Odometry.cpp itself cannot be compiled here (ORB / BFMatcher /
// MILD need real OpenCV), so the ~40 lines of Odometry::MultiScaleComputing (Odometry.cpp:621-685) that sequence
// those calls are restated below, statement for statement.  Image pyramids are supplied by the caller (the
// OpenCV filters that build them are outside the reference tree; see oracle/opb_oracle.c).
#include <cstdint>
#include <cstring>
#include <tuple>
#include <vector>

#include "Camera/Camera.h"
#include "Geometry/Geometry.h"
#include "Odometry/DenseOdometryFunction.h"

using namespace one_piece;
typedef geometry::scalar scalar;

namespace
{
cv::Mat Wrap(const float *p, int w, int h) { return cv::Mat(h, w, CV_32FC1, const_cast<float *>(p)); }
geometry::TransformationMatrix FromCm(const double *p)
{
    geometry::TransformationMatrix T;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) T(r, c) = (scalar)p[c * 4 + r];
    return T;
}
void ToCm(const geometry::TransformationMatrix &T, double *p)
{
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) p[c * 4 + r] = (double)T(r, c);
}
long CopyPairs(const geometry::PixelCorrespondenceSet &c, uint32_t *pairs, long cap)
{
    long n = (long)c.size();
    for (long k = 0; k < n && k < cap; ++k)
    {
        pairs[4 * k] = c[k].first(0); pairs[4 * k + 1] = c[k].first(1);
        pairs[4 * k + 2] = c[k].second(0); pairs[4 * k + 3] = c[k].second(1);
    }
    return n;
}
} // namespace

extern "C"
{
long ref_correspondences(const float *sd, const float *td, int w, int h, float fx, float fy, float cx, float cy, const double *T_cm,
                         uint32_t *pairs, long cap)
{
    camera::PinholeCamera cam(fx, fy, cx, cy, w, h, 1000.0f);
    geometry::PixelCorrespondenceSet c;
    odometry::ComputeCorrespondencePixelWise(Wrap(sd, w, h), Wrap(td, w, h), cam, FromCm(T_cm), c);
    return CopyPairs(c, pairs, cap);
}
void ref_normalize_intensity(float *sg, float *tg, int w, int h, const uint32_t *pairs, long n)
{
    geometry::PixelCorrespondenceSet c(n);
    for (long k = 0; k < n; ++k)
    {
        c[k].first = geometry::Point2ui(pairs[4 * k], pairs[4 * k + 1]);
        c[k].second = geometry::Point2ui(pairs[4 * k + 2], pairs[4 * k + 3]);
    }
    cv::Mat s = Wrap(sg, w, h), t = Wrap(tg, w, h);
    odometry::NormalizeIntensity(s, t, c);
}
// img[what][level]: what = 0 gray, 1 depth, 2 gray dx, 3 gray dy, 4 depth dx, 5 depth dy (source needs 0,1; target all)
// Restates Odometry::MultiScaleComputing (Odometry.cpp:621-685) + the result assembly of DenseTracking (:596-607).
long ref_multiscale(const float *const *src_img, const float *const *tgt_img, int w, int h, float fx, float fy, float cx, float cy,
                    const double *init_T_cm, int term, double *out_T_cm, double *rmse, int *success, uint32_t *pairs, long cap,
                    long *corr_per_iteration, double *T_per_iteration, int *n_iterations)
{
    const int levels = 3;
    const int iter_count_per_level[3] = {4, 8, 16}; // Odometry.h:170
    std::vector<camera::PinholeCamera> cams;
    cams.push_back(camera::PinholeCamera(fx, fy, cx, cy, w, h, 1000.0f));
    for (int i = 1; i < levels; ++i) cams.push_back(cams[i - 1].GenerateNextPyramid());
    std::vector<geometry::ImageXYZ> sxyz(levels), txyz(levels);
    for (int i = 0; i < levels; ++i)
    {
        geometry::TransformToMatXYZ(Wrap(src_img[1 * 3 + i], w >> i, h >> i), cams[i], sxyz[i]);
        geometry::TransformToMatXYZ(Wrap(tgt_img[1 * 3 + i], w >> i, h >> i), cams[i], txyz[i]);
    }
    geometry::TransformationMatrix T = FromCm(init_T_cm);
    geometry::PixelCorrespondenceSet correspondences;
    int it = 0;
    for (int i = levels - 1; i >= 0; --i)
    {
        const int lw = w >> i, lh = h >> i;
        cv::Mat sc = Wrap(src_img[0 * 3 + i], lw, lh), sd = Wrap(src_img[1 * 3 + i], lw, lh);
        cv::Mat tc = Wrap(tgt_img[0 * 3 + i], lw, lh), td = Wrap(tgt_img[1 * 3 + i], lw, lh);
        cv::Mat tcdx = Wrap(tgt_img[2 * 3 + i], lw, lh), tcdy = Wrap(tgt_img[3 * 3 + i], lw, lh);
        cv::Mat tddx = Wrap(tgt_img[4 * 3 + i], lw, lh), tddy = Wrap(tgt_img[5 * 3 + i], lw, lh);
        for (int j = 0; j != iter_count_per_level[i]; ++j)
        {
            correspondences.clear();
            if (term == 0) odometry::DoSingleIteration(sc, sd, tc, td, tcdx, tddx, tcdy, tddy, sxyz[i], cams[i], T, correspondences);
            else if (term == 1) odometry::DoSingleIterationPhotoTerm(sc, sd, tc, td, tcdx, tcdy, sxyz[i], cams[i], T, correspondences);
            else odometry::DoSingleIterationDepthTerm(sc, sd, tc, td, tddx, tddy, sxyz[i], cams[i], T, correspondences);
            if (it < 64)
            {
                corr_per_iteration[it] = (long)correspondences.size();
                ToCm(T, T_per_iteration + 16 * it);
            }
            ++it;
            if ((float)correspondences.size() / (h * w) > MAX_INLIER_RATIO_DENSE) break;
        }
    }
    *n_iterations = it;
    geometry::PointCorrespondenceSet correspondence_set;
    for (size_t i = 0; i < correspondences.size(); ++i)
    {
        int v_s = correspondences[i].first(0), u_s = correspondences[i].first(1);
        correspondence_set.push_back(std::make_pair(sxyz[0][v_s][u_s], txyz[0][v_s][u_s]));
    }
    *success = (float)correspondences.size() / (h * w) >= MIN_INLIER_RATIO_DENSE;
    *rmse = geometry::ComputeReprojectionError3D(correspondence_set, T);
    ToCm(T, out_T_cm);
    return CopyPairs(correspondences, pairs, cap);
}
// One solver iteration at pyramid level `level` from a caller-supplied pose ("teacher forcing"): the reference's
// correspondences, its J^T J / J^T r / sum r^2 (ComputeJTJandJTr*Term, DenseOdometryFunction.cpp:297-381; hybrid and
// photo terms are declared in the header, the depth term is reached through DoSingleIterationDepthTerm only) and
// the pose after DoSingleIteration*.  w, h and the intrinsics are those of level 0.
long ref_single_iteration(const float *const *src_img, const float *const *tgt_img, int level, int w, int h, float fx, float fy,
                          float cx, float cy, const double *T_in_cm, int term, double *T_out_cm, double *JTJ36, double *JTr6,
                          double *r2, uint32_t *pairs, long cap)
{
    camera::PinholeCamera cam(fx, fy, cx, cy, w, h, 1000.0f);
    for (int i = 0; i < level; ++i) cam = cam.GenerateNextPyramid();
    const int lw = w >> level, lh = h >> level, i = level;
    geometry::ImageXYZ sxyz;
    geometry::TransformToMatXYZ(Wrap(src_img[1 * 3 + i], lw, lh), cam, sxyz);
    cv::Mat sc = Wrap(src_img[0 * 3 + i], lw, lh), sd = Wrap(src_img[1 * 3 + i], lw, lh);
    cv::Mat tc = Wrap(tgt_img[0 * 3 + i], lw, lh), td = Wrap(tgt_img[1 * 3 + i], lw, lh);
    cv::Mat tcdx = Wrap(tgt_img[2 * 3 + i], lw, lh), tcdy = Wrap(tgt_img[3 * 3 + i], lw, lh);
    cv::Mat tddx = Wrap(tgt_img[4 * 3 + i], lw, lh), tddy = Wrap(tgt_img[5 * 3 + i], lw, lh);
    geometry::TransformationMatrix T = FromCm(T_in_cm);
    geometry::PixelCorrespondenceSet c;
    if (term != 2)
    {
        odometry::ComputeCorrespondencePixelWise(sd, td, cam, T, c);
        geometry::Matrix6 JTJ;
        geometry::Se3 JTr;
        float r;
        if (term == 0) std::tie(JTJ, JTr, r) = odometry::ComputeJTJandJTrHybridTerm(sc, sd, tc, td, tcdx, tddx, tcdy, tddy, sxyz, cam, T, c);
        else std::tie(JTJ, JTr, r) = odometry::ComputeJTJandJTrPhotoTerm(sc, sd, tc, td, tcdx, tcdy, sxyz, cam, T, c);
        for (int a = 0; a < 6; ++a)
        {
            for (int b = 0; b < 6; ++b) JTJ36[a * 6 + b] = (double)JTJ(a, b);
            JTr6[a] = (double)JTr(a);
        }
        *r2 = r;
        c.clear();
    }
    if (term == 0) odometry::DoSingleIteration(sc, sd, tc, td, tcdx, tddx, tcdy, tddy, sxyz, cam, T, c);
    else if (term == 1) odometry::DoSingleIterationPhotoTerm(sc, sd, tc, td, tcdx, tcdy, sxyz, cam, T, c);
    else odometry::DoSingleIterationDepthTerm(sc, sd, tc, td, tddx, tddy, sxyz, cam, T, c);
    ToCm(T, T_out_cm);
    return CopyPairs(c, pairs, cap);
}
} // extern "C"
