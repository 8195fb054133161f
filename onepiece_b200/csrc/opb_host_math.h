// Host-side float32 restatements of the two small Eigen computations whose results are consumed by
// discrete decisions on the device: the camera-pose inverse (Integrator.cpp:18,48 call
// Eigen::Matrix4f::inverse()) and the view-frustum planes (Frustum.cpp:7-52, Geometry.cpp:170-176).
// Both are evaluated once per frame on the host, exactly as the reference does, and handed to the kernels.
//
// Third-party arithmetic restated here: Eigen 3.3.7 (vendored by the reference under 3rdparty/Eigen,
// Eigen/src/Core/util/Macros.h:14-16).  Its Matrix4f inverse on SSE targets is the 2x2-block cofactor scheme
// (Eigen/src/LU/arch/Inverse_SSE.h); fixed-size-3 reductions associate as a0 + (a1 + a2).  Parity of these
// restatements with the compiled reference is pinned by tests/test_host_math.py.
//
// Must be compiled without FMA contraction (-ffp-contract=off) so that every product and sum rounds
// separately like the reference's -msse4.2 build.
#pragma once
#include <cmath>

namespace opb
{
namespace hostmath
{
struct Quad // one 2x2 block held like a 4-lane register
{
    float l[4];
};
static inline Quad q_mul(const Quad &a, const Quad &b) { return Quad{{a.l[0] * b.l[0], a.l[1] * b.l[1], a.l[2] * b.l[2], a.l[3] * b.l[3]}}; }
static inline Quad q_add(const Quad &a, const Quad &b) { return Quad{{a.l[0] + b.l[0], a.l[1] + b.l[1], a.l[2] + b.l[2], a.l[3] + b.l[3]}}; }
static inline Quad q_sub(const Quad &a, const Quad &b) { return Quad{{a.l[0] - b.l[0], a.l[1] - b.l[1], a.l[2] - b.l[2], a.l[3] - b.l[3]}}; }
// lanes (i0,i1) from a, (i2,i3) from b
static inline Quad q_pick(const Quad &a, const Quad &b, int i0, int i1, int i2, int i3) { return Quad{{a.l[i0], a.l[i1], b.l[i2], b.l[i3]}}; }
static inline Quad q_perm(const Quad &a, int i0, int i1, int i2, int i3) { return q_pick(a, a, i0, i1, i2, i3); }
static inline Quad q_splat(float s) { return Quad{{s, s, s, s}}; }
// 2x2 determinant of a block stored (p, q, r, s): p*s - q*r  with the products formed lane-wise first
static inline float q_det(const Quad &a)
{
    Quad m = q_mul(q_perm(a, 3, 3, 1, 1), a); // (s*p, s*q, q*r, q*s)
    return m.l[0] - m.l[2];
}

// inverse of a column-major 4x4 float matrix, Eigen 3.3.7 SSE operation order
static inline void mat4_inverse_colmajor(const float *m, float *out)
{
    // 2x2 blocks of the column-major storage: A = cols 0-1 rows 0-1, B = cols 0-1 rows 2-3,
    // C = cols 2-3 rows 0-1, D = cols 2-3 rows 2-3, each as (c0r0, c0r1, c1r0, c1r1)
    Quad A{{m[0], m[1], m[4], m[5]}}, B{{m[2], m[3], m[6], m[7]}};
    Quad C{{m[8], m[9], m[12], m[13]}}, D{{m[10], m[11], m[14], m[15]}};

    // AB = adj(A) * B ; DC = adj(D) * C
    Quad AB = q_sub(q_mul(q_perm(A, 3, 3, 0, 0), B), q_mul(q_perm(A, 1, 1, 2, 2), q_perm(B, 2, 3, 0, 1)));
    Quad DC = q_sub(q_mul(q_perm(D, 3, 3, 0, 0), C), q_mul(q_perm(D, 1, 1, 2, 2), q_perm(C, 2, 3, 0, 1)));
    float dA = q_det(A), dB = q_det(B), dC = q_det(C), dD = q_det(D);

    // tr = trace(AB * DC), pairwise: (t0 + t2) + (t1 + t3)
    Quad t = q_mul(q_perm(DC, 0, 2, 1, 3), AB);
    float tr = (t.l[0] + t.l[2]) + (t.l[1] + t.l[3]);

    // iD = D*|A| - C*AB ; iA = A*|D| - B*DC
    Quad iD = q_add(q_mul(q_perm(C, 0, 0, 2, 2), q_perm(AB, 0, 1, 0, 1)), q_mul(q_perm(C, 1, 1, 3, 3), q_perm(AB, 2, 3, 2, 3)));
    Quad iA = q_add(q_mul(q_perm(B, 0, 0, 2, 2), q_perm(DC, 0, 1, 0, 1)), q_mul(q_perm(B, 1, 1, 3, 3), q_perm(DC, 2, 3, 2, 3)));
    iD = q_sub(q_mul(D, q_splat(dA)), iD);
    iA = q_sub(q_mul(A, q_splat(dD)), iA);

    float det = (dA * dD + dB * dC) - tr;
    float rd = 1.0f / det;

    // iB = C*|B| - D*adj(AB) ; iC = B*|C| - A*adj(DC)
    Quad iB = q_sub(q_mul(D, q_perm(AB, 3, 0, 3, 0)), q_mul(q_perm(D, 1, 0, 3, 2), q_perm(AB, 2, 1, 2, 1)));
    Quad iC = q_sub(q_mul(A, q_perm(DC, 3, 0, 3, 0)), q_mul(q_perm(A, 1, 0, 3, 2), q_perm(DC, 2, 1, 2, 1)));
    iB = q_sub(q_mul(C, q_splat(dB)), iB);
    iC = q_sub(q_mul(B, q_splat(dC)), iC);

    const Quad sgn{{rd, -rd, -rd, rd}};
    iA = q_mul(sgn, iA);
    iB = q_mul(sgn, iB);
    iC = q_mul(sgn, iC);
    iD = q_mul(sgn, iD);

    Quad c0 = q_pick(iA, iB, 3, 1, 3, 1), c1 = q_pick(iA, iB, 2, 0, 2, 0);
    Quad c2 = q_pick(iC, iD, 3, 1, 3, 1), c3 = q_pick(iC, iD, 2, 0, 2, 0);
    for (int i = 0; i < 4; ++i)
    {
        out[i] = c0.l[i];
        out[4 + i] = c1.l[i];
        out[8 + i] = c2.l[i];
        out[12 + i] = c3.l[i];
    }
}

struct V3
{
    float x, y, z;
};
static inline V3 v_add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 v_sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 v_scale(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline float v_dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }

// geometry::GetPlane (Geometry.cpp:170-176): unit normal of (p2-p1)x(p3-p1) and d = -p1.n
static inline void plane_from_points(V3 p1, V3 p2, V3 p3, float *pl)
{
    V3 a = v_sub(p2, p1), b = v_sub(p3, p1);
    V3 n{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    float z = v_dot(n, n);
    if (z > 0.0f)
    {
        float len = std::sqrt(z);
        n = V3{n.x / len, n.y / len, n.z / len};
    }
    double d = -v_dot(p1, n);
    pl[0] = n.x;
    pl[1] = n.y;
    pl[2] = n.z;
    pl[3] = (float)d;
}

// Frustum::ComputeFromCamera + ComputeFromVectors (Frustum.cpp:7-52).  planes: 6 x 4 floats in the order
// ContainPoint tests them (Frustum.h:74-103): top, left, right, bottom, near, far.
static inline void frustum_planes(const float *pose_cm, float fx, float fy, float cy, float width, float height,
                                  float far_dist, float near_dist, float *planes)
{
    V3 right{pose_cm[0], pose_cm[1], pose_cm[2]};
    V3 up{-pose_cm[4], -pose_cm[5], -pose_cm[6]};
    V3 fwd{pose_cm[8], pose_cm[9], pose_cm[10]};
    V3 pos{pose_cm[12], pose_cm[13], pose_cm[14]};
    float aspect = (fy * width) / (fx * height);
    // Frustum.cpp:23,29 call unqualified atan2()/tan() on floats from a plain C++ translation unit: they bind to
    // the C library's double versions, and the result is rounded once on assignment
    float fov = (float)(std::atan2((double)cy, (double)fy) + std::atan2((double)(height - cy), (double)fy));
    float angle_tangent = (float)std::tan((double)(fov / 2));
    float height_far = angle_tangent * far_dist;
    float width_far = height_far * aspect;
    float height_near = angle_tangent * near_dist;
    float width_near = height_near * aspect;
    V3 fc = v_add(pos, v_scale(fwd, far_dist));
    V3 ftl = v_sub(v_add(fc, v_scale(up, height_far)), v_scale(right, width_far));
    V3 ftr = v_add(v_add(fc, v_scale(up, height_far)), v_scale(right, width_far));
    V3 fbl = v_sub(v_sub(fc, v_scale(up, height_far)), v_scale(right, width_far));
    V3 fbr = v_add(v_sub(fc, v_scale(up, height_far)), v_scale(right, width_far));
    V3 nc = v_add(pos, v_scale(fwd, near_dist));
    V3 ntl = v_sub(v_add(nc, v_scale(up, height_near)), v_scale(right, width_near));
    V3 ntr = v_add(v_add(nc, v_scale(up, height_near)), v_scale(right, width_near));
    V3 nbl = v_sub(v_sub(nc, v_scale(up, height_near)), v_scale(right, width_near));
    V3 nbr = v_add(v_sub(nc, v_scale(up, height_near)), v_scale(right, width_near));
    plane_from_points(ntl, ftl, ntr, planes + 0);   // top
    plane_from_points(ftl, ntl, fbl, planes + 4);   // left
    plane_from_points(ntr, ftr, nbr, planes + 8);   // right
    plane_from_points(nbr, fbl, nbl, planes + 12);  // bottom
    plane_from_points(nbl, ntl, nbr, planes + 16);  // near
    plane_from_points(ftr, ftl, fbr, planes + 20);  // far
}

} // namespace hostmath
} // namespace opb
