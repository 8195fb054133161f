/*
 * TEST INFRASTRUCTURE ONLY -- see opb_oracle.h.  CPU restatement of the reference algorithm, written as
 * straight-line C that mirrors the reference's statement order so the float32 results are bit-identical to
 * the reference built with its own flags (-O3 -msse4.2, no FMA).  Build: `make -C oracle oracle`
 * (gcc -O2 -msse4.2 -ffp-contract=off).
 *
 * Parity status: PINNED against oracle/_ref (the reference's translation units compiled unmodified) by
 * tests/test_oracle_vs_ref.py and against tests/golden/ fixtures by tests/test_oracle_golden.py.
 */
#include "opb_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../onepiece_b200/csrc/opb_mc_table.h" /* generated data: Marching Cubes case table */

#define CUBE 8 /* CUBE_SIZE, src/Integration/VoxelCube.h:4 */
#define NVOX 512

/* TSDFVoxel, src/Integration/TSDFVoxel.h:8-82 (20 bytes) */
typedef struct
{
    float sdf, weight, c[3];
} voxel_t;
typedef struct
{
    int id[3];
    voxel_t vox[NVOX];
} cube_t;

struct orc_volume
{
    float fx, fy, cx, cy, depth_scale;
    int width, height;
    float res, trunc, near_plane, far_plane;
    float centroid[NVOX][3]; /* CubePara::VoxelCentroidOffSet, VoxelCube.h:48-61 */
    cube_t **cubes;          /* insertion order */
    long n_cubes, cap_cubes;
    long *table;             /* open addressing: index into cubes or -1 */
    long table_cap;
};

/* x86 float/double -> int conversion as g++ emits it (cvttss2si / cvttsd2si): out-of-range and NaN give INT_MIN */
static int cvtt_d(double d) { return (d > -2147483649.0 && d < 2147483648.0) ? (int)d : INT_MIN; }
static int cvtt_f(float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : INT_MIN; }

/* ------------------------------------------------------------------------------------------------------- */
/* cube container (std::unordered_map<CubeID, VoxelCube>, CubeHandler.h:22)                               */
/* ------------------------------------------------------------------------------------------------------- */
static unsigned long hash_id(const int *id)
{
    /* VoxelGridHasher, src/Geometry/Geometry.h:101-112 (the oracle's container order is its own) */
    return ((unsigned long)(long)id[0] * 73856093ul) ^ ((unsigned long)(long)id[1] * 19349663ul) ^
           ((unsigned long)(long)id[2] * 83492791ul);
}
static void table_rebuild(orc_volume *v, long cap)
{
    free(v->table);
    v->table_cap = cap;
    v->table = (long *)malloc(sizeof(long) * cap);
    for (long i = 0; i < cap; ++i) v->table[i] = -1;
    for (long c = 0; c < v->n_cubes; ++c)
    {
        unsigned long h = hash_id(v->cubes[c]->id) & (cap - 1);
        while (v->table[h] >= 0) h = (h + 1) & (cap - 1);
        v->table[h] = c;
    }
}
static cube_t *find_cube(const orc_volume *v, int i, int j, int k)
{
    const int id[3] = {i, j, k};
    unsigned long h = hash_id(id) & (v->table_cap - 1);
    while (v->table[h] >= 0)
    {
        cube_t *c = v->cubes[v->table[h]];
        if (c->id[0] == i && c->id[1] == j && c->id[2] == k) return c;
        h = (h + 1) & (v->table_cap - 1);
    }
    return NULL;
}
/* VoxelCube(const CubeID&), VoxelCube.h:102-106 with TSDFVoxel defaults TSDFVoxel.h:79-81 */
static cube_t *add_cube(orc_volume *v, int i, int j, int k)
{
    if (v->n_cubes == v->cap_cubes)
    {
        v->cap_cubes = v->cap_cubes ? v->cap_cubes * 2 : 1024;
        v->cubes = (cube_t **)realloc(v->cubes, sizeof(cube_t *) * v->cap_cubes);
    }
    cube_t *c = (cube_t *)malloc(sizeof(cube_t));
    c->id[0] = i; c->id[1] = j; c->id[2] = k;
    for (int n = 0; n < NVOX; ++n)
    {
        c->vox[n].sdf = 999; c->vox[n].weight = 0;
        c->vox[n].c[0] = c->vox[n].c[1] = c->vox[n].c[2] = -1;
    }
    v->cubes[v->n_cubes++] = c;
    if (v->n_cubes * 2 > v->table_cap) table_rebuild(v, v->table_cap * 2);
    else
    {
        unsigned long h = hash_id(c->id) & (v->table_cap - 1);
        while (v->table[h] >= 0) h = (h + 1) & (v->table_cap - 1);
        v->table[h] = v->n_cubes - 1;
    }
    return c;
}

orc_volume *orc_volume_create(float fx, float fy, float cx, float cy, int width, int height, float depth_scale,
                              float res, float trunc, float near_plane, float far_plane)
{
    orc_volume *v = (orc_volume *)calloc(1, sizeof(orc_volume));
    v->fx = fx; v->fy = fy; v->cx = cx; v->cy = cy; v->width = width; v->height = height; v->depth_scale = depth_scale;
    v->res = res; v->trunc = trunc; v->near_plane = near_plane; v->far_plane = far_plane;
    /* CubePara::InitializeVoxelCube, VoxelCube.h:48-61 */
    float half = res / 2;
    for (int x = 0; x < CUBE; ++x)
        for (int y = 0; y < CUBE; ++y)
            for (int z = 0; z < CUBE; ++z)
            {
                float *o = v->centroid[x + y * CUBE + z * CUBE * CUBE];
                o[0] = x * res + half; o[1] = y * res + half; o[2] = z * res + half;
            }
    table_rebuild(v, 4096);
    return v;
}
void orc_volume_clear(orc_volume *v)
{
    for (long c = 0; c < v->n_cubes; ++c) free(v->cubes[c]);
    v->n_cubes = 0;
    table_rebuild(v, 4096);
}
void orc_volume_destroy(orc_volume *v)
{
    if (!v) return;
    orc_volume_clear(v);
    free(v->cubes); free(v->table); free(v);
}
void orc_free(void *p) { free(p); }
long orc_volume_num_cubes(const orc_volume *v) { return v->n_cubes; }

/* ------------------------------------------------------------------------------------------------------- */
/* small Eigen computations                                                                                */
/* ------------------------------------------------------------------------------------------------------- */
/* Eigen 3.3.7 Matrix4f::inverse() on SSE targets (3rdparty/Eigen/Eigen/src/LU/arch/Inverse_SSE.h): 2x2-block
 * cofactor scheme; called at Integrator.cpp:18,48.  Lanes written out explicitly.                          */
void orc_pose_inverse(const float *m, float *out)
{
    const float A[4] = {m[0], m[1], m[4], m[5]}, B[4] = {m[2], m[3], m[6], m[7]};
    const float C[4] = {m[8], m[9], m[12], m[13]}, D[4] = {m[10], m[11], m[14], m[15]};
    float AB[4], DC[4], iA[4], iB[4], iC[4], iD[4];
    /* AB = adj(A)*B, DC = adj(D)*C */
    AB[0] = A[3] * B[0] - A[1] * B[2]; AB[1] = A[3] * B[1] - A[1] * B[3];
    AB[2] = A[0] * B[2] - A[2] * B[0]; AB[3] = A[0] * B[3] - A[2] * B[1];
    DC[0] = D[3] * C[0] - D[1] * C[2]; DC[1] = D[3] * C[1] - D[1] * C[3];
    DC[2] = D[0] * C[2] - D[2] * C[0]; DC[3] = D[0] * C[3] - D[2] * C[1];
    const float dA = A[3] * A[0] - A[1] * A[2], dB = B[3] * B[0] - B[1] * B[2];
    const float dC = C[3] * C[0] - C[1] * C[2], dD = D[3] * D[0] - D[1] * D[2];
    const float t0 = DC[0] * AB[0], t1 = DC[2] * AB[1], t2 = DC[1] * AB[2], t3 = DC[3] * AB[3];
    const float tr = (t0 + t2) + (t1 + t3);
    iD[0] = C[0] * AB[0] + C[1] * AB[2]; iD[1] = C[0] * AB[1] + C[1] * AB[3];
    iD[2] = C[2] * AB[0] + C[3] * AB[2]; iD[3] = C[2] * AB[1] + C[3] * AB[3];
    iA[0] = B[0] * DC[0] + B[1] * DC[2]; iA[1] = B[0] * DC[1] + B[1] * DC[3];
    iA[2] = B[2] * DC[0] + B[3] * DC[2]; iA[3] = B[2] * DC[1] + B[3] * DC[3];
    for (int i = 0; i < 4; ++i) { iD[i] = D[i] * dA - iD[i]; iA[i] = A[i] * dD - iA[i]; }
    const float det = (dA * dD + dB * dC) - tr;
    const float rd = 1.0f / det;
    iB[0] = D[0] * AB[3] - D[1] * AB[2]; iB[1] = D[1] * AB[0] - D[0] * AB[1];
    iB[2] = D[2] * AB[3] - D[3] * AB[2]; iB[3] = D[3] * AB[0] - D[2] * AB[1];
    iC[0] = A[0] * DC[3] - A[1] * DC[2]; iC[1] = A[1] * DC[0] - A[0] * DC[1];
    iC[2] = A[2] * DC[3] - A[3] * DC[2]; iC[3] = A[3] * DC[0] - A[2] * DC[1];
    for (int i = 0; i < 4; ++i) { iB[i] = C[i] * dB - iB[i]; iC[i] = B[i] * dC - iC[i]; }
    const float s[4] = {rd, -rd, -rd, rd};
    for (int i = 0; i < 4; ++i) { iA[i] = s[i] * iA[i]; iB[i] = s[i] * iB[i]; iC[i] = s[i] * iC[i]; iD[i] = s[i] * iD[i]; }
    out[0] = iA[3]; out[1] = iA[1]; out[2] = iB[3]; out[3] = iB[1];
    out[4] = iA[2]; out[5] = iA[0]; out[6] = iB[2]; out[7] = iB[0];
    out[8] = iC[3]; out[9] = iC[1]; out[10] = iD[3]; out[11] = iD[1];
    out[12] = iC[2]; out[13] = iC[0]; out[14] = iD[2]; out[15] = iD[0];
}

/* Eigen fixed-size-3 reduction order: a0*b0 + (a1*b1 + a2*b2) */
static float dot3(const float *a, const float *b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }

/* geometry::GetPlane, src/Geometry/Geometry.cpp:170-176 */
static void get_plane(const float *p1, const float *p2, const float *p3, float *pl)
{
    float a[3], b[3], n[3];
    for (int i = 0; i < 3; ++i) { a[i] = p2[i] - p1[i]; b[i] = p3[i] - p1[i]; }
    n[0] = a[1] * b[2] - a[2] * b[1]; n[1] = a[2] * b[0] - a[0] * b[2]; n[2] = a[0] * b[1] - a[1] * b[0];
    float z = dot3(n, n);
    if (z > 0) { float len = sqrtf(z); n[0] = n[0] / len; n[1] = n[1] / len; n[2] = n[2] / len; }
    double d = -dot3(p1, n);
    pl[0] = n[0]; pl[1] = n[1]; pl[2] = n[2]; pl[3] = (float)d;
}
/* Frustum::ComputeFromCamera / ComputeFromVectors, src/Integration/Frustum.cpp:7-52; plane order = the order
 * ContainPoint tests them (Frustum.h:74-103): top, left, right, bottom, near, far */
void orc_frustum_planes(const orc_volume *v, const float *T, float *planes)
{
    const float right[3] = {T[0], T[1], T[2]}, up[3] = {-T[4], -T[5], -T[6]}, fwd[3] = {T[8], T[9], T[10]};
    const float pos[3] = {T[12], T[13], T[14]};
    const float width = (float)v->width, height = (float)v->height;
    float aspect = (v->fy * width) / (v->fx * height);
    /* unqualified atan2()/tan() on floats in a plain C++ translation unit bind to the C library's double
     * versions (Frustum.cpp:23,29): evaluate in double, round once on assignment */
    float fov = (float)(atan2((double)v->cy, (double)v->fy) + atan2((double)(height - v->cy), (double)v->fy));
    float tang = (float)tan((double)(fov / 2));
    float hf = tang * v->far_plane, wf = hf * aspect, hn = tang * v->near_plane, wn = hn * aspect;
    float ftl[3], ftr[3], fbl[3], fbr[3], ntl[3], ntr[3], nbl[3], nbr[3];
    for (int i = 0; i < 3; ++i)
    {
        float fc = pos[i] + fwd[i] * v->far_plane, nc = pos[i] + fwd[i] * v->near_plane;
        ftl[i] = fc + (up[i] * hf) - (right[i] * wf); ftr[i] = fc + (up[i] * hf) + (right[i] * wf);
        fbl[i] = fc - (up[i] * hf) - (right[i] * wf); fbr[i] = fc - (up[i] * hf) + (right[i] * wf);
        ntl[i] = nc + (up[i] * hn) - (right[i] * wn); ntr[i] = nc + (up[i] * hn) + (right[i] * wn);
        nbl[i] = nc - (up[i] * hn) - (right[i] * wn); nbr[i] = nc - (up[i] * hn) + (right[i] * wn);
    }
    get_plane(ntl, ftl, ntr, planes + 0);
    get_plane(ftl, ntl, fbl, planes + 4);
    get_plane(ntr, ftr, nbr, planes + 8);
    get_plane(nbr, fbl, nbl, planes + 12);
    get_plane(nbl, ntl, nbr, planes + 16);
    get_plane(ftr, ftl, fbr, planes + 20);
}
/* Frustum::ContainPoint, src/Integration/Frustum.h:74-103 */
int orc_frustum_contains(const float *pl, float x, float y, float z)
{
    const float p[3] = {x, y, z};
    for (int i = 0; i < 6; ++i)
    {
        float d = dot3(pl + 4 * i, p) + pl[4 * i + 3];
        if (d < 0) return 0;
        if (d == 0) return 1;
    }
    return 1;
}
/* Eigen Matrix4f * Vector4f(x,y,z,1) row r (vectorised gemv): ((m0*x + m1*y) + m2*z) + m3*1 */
static float row_xyz1(const float *m, int r, float x, float y, float z)
{
    return ((m[r] * x + m[4 + r] * y) + m[8 + r] * z) + m[12 + r];
}
static float depth_at(const orc_volume *v, const void *depth, int is_u16, int row, int col)
{
    if (is_u16) return ((const unsigned short *)depth)[row * v->width + col] / v->depth_scale;
    return ((const float *)depth)[row * v->width + col];
}

/* ------------------------------------------------------------------------------------------------------- */
/* cube selection                                                                                          */
/* ------------------------------------------------------------------------------------------------------- */
/* CubeHandler::ComputeBounding, src/Integration/CubeHandler.cpp:116-145
 * (PointCloud::LoadFromDepth PointCloud.cpp:72-100, geometry::TransformPoints Geometry.cpp:19-27) */
void orc_volume_bounding(const orc_volume *v, const void *depth, int is_u16, const float *T, float *mx, float *mn)
{
    float planes[24];
    orc_frustum_planes(v, T, planes);
    for (int a = 0; a < 3; ++a) { mx[a] = -FLT_MAX; mn[a] = FLT_MAX; }
    for (int i = 0; i < v->height; ++i)
        for (int j = 0; j < v->width; ++j)
        {
            float z = depth_at(v, depth, is_u16, i, j);
            if (!(z > 0)) continue;
            float x = (j - v->cx) * z / v->fx;
            float y = (i - v->cy) * z / v->fy;
            float w = row_xyz1(T, 3, x, y, z);
            float p[3] = {row_xyz1(T, 0, x, y, z) / w, row_xyz1(T, 1, x, y, z) / w, row_xyz1(T, 2, x, y, z) / w};
            if (!orc_frustum_contains(planes, p[0], p[1], p[2])) continue;
            for (int a = 0; a < 3; ++a)
            {
                mx[a] = (p[a] < mx[a]) ? mx[a] : p[a]; /* std::max(p, max) */
                mn[a] = (mn[a] < p[a]) ? mn[a] : p[a]; /* std::min(p, min) */
            }
        }
}
/* Integrator::GetSDF, src/Integration/Integrator.cpp:8-35 (pose_inv passed in: it is loop-invariant) */
static float get_sdf(const orc_volume *v, const void *depth, int is_u16, const float *pinv, float x, float y, float z)
{
    float X = row_xyz1(pinv, 0, x, y, z), Y = row_xyz1(pinv, 1, x, y, z), Z = row_xyz1(pinv, 2, x, y, z);
    int u = cvtt_d(v->fx * X / Z + 0.5 + v->cx);
    int w = cvtt_d(v->fy * Y / Z + 0.5 + v->cy);
    if (w < 0 || w >= v->height || u < 0 || u >= v->width) return 999;
    float d = depth_at(v, depth, is_u16, w, u);
    if (d <= 0) return 999;
    return d - Z;
}
void orc_volume_get_sdf(const orc_volume *v, const void *depth, int is_u16, const float *pose_cm, const float *pts, long n,
                        float *sdf)
{
    float pinv[16];
    orc_pose_inverse(pose_cm, pinv);
    for (long i = 0; i < n; ++i) sdf[i] = get_sdf(v, depth, is_u16, pinv, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
}
/* CubePara::GetCubeID(Point3), VoxelCube.h:63-74 */
static int cube_id_of(float p, float res)
{
    int voxel = cvtt_f(floorf(p / res));
    return cvtt_d(floor((voxel + 0.0) / CUBE));
}
/* CubeHandler::PrepareCubes, src/Integration/CubeHandler.cpp:147-196 */
static long prepare_cubes(orc_volume *v, const void *depth, int is_u16, const float *pose_cm, const float *pinv,
                          cube_t ***list_out)
{
    float mx[3], mn[3];
    orc_volume_bounding(v, depth, is_u16, pose_cm, mx, mn);
    int hi[3], lo[3];
    for (int a = 0; a < 3; ++a) { hi[a] = cube_id_of(mx[a], v->res); lo[a] = cube_id_of(mn[a], v->res); }
    static const int corner_voxel[8] = {0, CUBE - 1, (CUBE - 1) * CUBE, (CUBE - 1) * CUBE + CUBE - 1,
                                        CUBE * CUBE * (CUBE - 1), CUBE * CUBE * (CUBE - 1) + CUBE - 1,
                                        CUBE * CUBE * (CUBE - 1) + CUBE * (CUBE - 1),
                                        CUBE * CUBE * (CUBE - 1) + CUBE * (CUBE - 1) + CUBE - 1};
    float cube_res = v->res * CUBE;
    long n = 0, cap = 1024;
    cube_t **list = (cube_t **)malloc(sizeof(cube_t *) * cap);
    for (long i = (long)lo[0] - 1; i <= (long)hi[0] + 1; ++i)
        for (long j = (long)lo[1] - 1; j <= (long)hi[1] + 1; ++j)
            for (long k = (long)lo[2] - 1; k <= (long)hi[2] + 1; ++k)
            {
                float min_sdf = FLT_MAX;
                for (int c = 0; c < 8; ++c)
                {
                    const float *o = v->centroid[corner_voxel[c]];
                    float sdf = get_sdf(v, depth, is_u16, pinv, (int)i * cube_res + o[0], (int)j * cube_res + o[1],
                                        (int)k * cube_res + o[2]);
                    if (min_sdf > fabsf(sdf)) min_sdf = fabsf(sdf);
                }
                if (min_sdf < v->trunc)
                {
                    cube_t *c = find_cube(v, (int)i, (int)j, (int)k);
                    if (!c) c = add_cube(v, (int)i, (int)j, (int)k);
                    if (n == cap) { cap *= 2; list = (cube_t **)realloc(list, sizeof(cube_t *) * cap); }
                    list[n++] = c;
                }
            }
    *list_out = list;
    return n;
}
long orc_volume_prepare_cubes(orc_volume *v, const void *depth, int is_u16, const float *pose_cm, int32_t *ids, long cap)
{
    float pinv[16];
    orc_pose_inverse(pose_cm, pinv);
    cube_t **list;
    long n = prepare_cubes(v, depth, is_u16, pose_cm, pinv, &list);
    for (long i = 0; i < n && i < cap; ++i)
        for (int a = 0; a < 3; ++a) ids[3 * i + a] = list[i]->id[a];
    free(list);
    return n;
}

/* ------------------------------------------------------------------------------------------------------- */
/* voxel update                                                                                            */
/* ------------------------------------------------------------------------------------------------------- */
/* Integrator::IntegrateImage, src/Integration/Integrator.cpp:36-94; TSDFVoxel::operator+ TSDFVoxel.h:24-39 */
static void integrate_cube(const orc_volume *v, const void *depth, int is_u16, const uint8_t *bgr, const float *pinv,
                           cube_t *cube)
{
    const float ox = (float)cube->id[0] * CUBE * v->res, oy = (float)cube->id[1] * CUBE * v->res,
                oz = (float)cube->id[2] * CUBE * v->res; /* GetGlobalPoint, VoxelCube.h:75-80 */
    for (int n = 0; n < NVOX; ++n)
    {
        const float px = ox + v->centroid[n][0], py = oy + v->centroid[n][1], pz = oz + v->centroid[n][2];
        float X = row_xyz1(pinv, 0, px, py, pz), Y = row_xyz1(pinv, 1, px, py, pz), Z = row_xyz1(pinv, 2, px, py, pz);
        int u = cvtt_d(v->fx * X / Z + 0.5 + v->cx);
        int w = cvtt_d(v->fy * Y / Z + 0.5 + v->cy);
        if (w < 0 || w >= v->height || u < 0 || u >= v->width) continue;
        float d = depth_at(v, depth, is_u16, w, u);
        if (d <= 0) continue;
        float new_sdf = d - Z;
        if (fabsf(new_sdf) < v->trunc)
        {
            const uint8_t *px3 = bgr + 3 * ((size_t)w * v->width + u);
            float nc[3] = {px3[0] / 255.0f, px3[1] / 255.0f, px3[2] / 255.0f};
            voxel_t *vx = &cube->vox[n];
            int valid = !(vx->sdf >= 1 || vx->weight <= 0);
            if (valid && vx->weight != 0)
            {
                float W = vx->weight + 1.0f;
                voxel_t r = {999, W, {-1, -1, -1}};
                if (W != 0)
                {
                    r.sdf = (vx->weight * vx->sdf + 1.0f * new_sdf) / W;
                    for (int a = 0; a < 3; ++a) r.c[a] = (vx->weight * vx->c[a] + 1.0f * nc[a]) / W;
                }
                *vx = r;
            }
            else
            {
                vx->sdf = new_sdf; vx->weight = 1.0f;
                vx->c[0] = nc[0]; vx->c[1] = nc[1]; vx->c[2] = nc[2];
            }
        }
    }
}
/* CubeHandler::IntegrateImage, src/Integration/CubeHandler.cpp:197-210 */
long orc_volume_integrate(orc_volume *v, const void *depth, int is_u16, const uint8_t *bgr, const float *pose_cm)
{
    float pinv[16];
    orc_pose_inverse(pose_cm, pinv);
    cube_t **list;
    long n = prepare_cubes(v, depth, is_u16, pose_cm, pinv, &list);
    for (long i = 0; i < n; ++i) integrate_cube(v, depth, is_u16, bgr, pinv, list[i]);
    free(list);
    return n;
}
void orc_volume_download(const orc_volume *v, int32_t *ids, float *voxels)
{
    for (long c = 0; c < v->n_cubes; ++c)
    {
        for (int a = 0; a < 3; ++a) ids[3 * c + a] = v->cubes[c]->id[a];
        memcpy(voxels + (size_t)c * NVOX * 5, v->cubes[c]->vox, sizeof(voxel_t) * NVOX);
    }
}
void orc_volume_upload(orc_volume *v, const int32_t *ids, const float *voxels, long n)
{
    orc_volume_clear(v);
    for (long c = 0; c < n; ++c)
    {
        cube_t *cube = find_cube(v, ids[3 * c], ids[3 * c + 1], ids[3 * c + 2]);
        if (!cube) cube = add_cube(v, ids[3 * c], ids[3 * c + 1], ids[3 * c + 2]);
        memcpy(cube->vox, voxels + (size_t)c * NVOX * 5, sizeof(voxel_t) * NVOX);
    }
}

/* ------------------------------------------------------------------------------------------------------- */
/* Volume resampling and merging                                                    */
/* ------------------------------------------------------------------------------------------------------- */
static const voxel_t kDefaultVoxel = {999, 0, {-1, -1, -1}};
/* TSDFVoxel::operator+ (TSDFVoxel.h:24-39) */
static voxel_t vox_plus(voxel_t a, voxel_t b)
{
    if (a.weight == 0) return b;
    if (b.weight == 0) return a;
    voxel_t r = kDefaultVoxel;
    r.weight = a.weight + b.weight;
    if (r.weight != 0)
    {
        r.sdf = (a.weight * a.sdf + b.weight * b.sdf) / r.weight;
        for (int c = 0; c < 3; ++c) r.c[c] = (a.weight * a.c[c] + b.weight * b.c[c]) / r.weight;
    }
    return r;
}
/* TSDFVoxel::add (TSDFVoxel.h:40-53): direct addition */
static voxel_t vox_add(voxel_t a, voxel_t b)
{
    if (a.weight == 0) return b;
    if (b.weight == 0) return a;
    voxel_t r;
    r.weight = a.weight + b.weight;
    r.sdf = a.sdf + b.sdf;
    for (int c = 0; c < 3; ++c) r.c[c] = a.c[c] + b.c[c];
    return r;
}
/* TSDFVoxel::operator* (TSDFVoxel.h:58-70) */
static voxel_t vox_mul(voxel_t a, float w)
{
    if (w == 0 || a.weight == 0) return kDefaultVoxel;
    voxel_t r;
    r.weight = a.weight * w;
    r.sdf = a.sdf * w;
    for (int c = 0; c < 3; ++c) r.c[c] = a.c[c] * w;
    return r;
}
/* TSDFVoxel::operator/ (TSDFVoxel.h:71-74): multiplication by the reciprocal */
static voxel_t vox_div(voxel_t a, float w) { return vox_mul(a, 1 / w); }
/* one level of ReadVoxelInterpolate (VoxelCube.cpp:16-47): blend of a and b along one axis with weight t */
static voxel_t vox_lerp(voxel_t a, voxel_t b, float t)
{
    voxel_t r = kDefaultVoxel;
    if (a.weight != 0 || b.weight != 0)
        r = vox_div(vox_add(vox_mul(a, 1 - t), vox_mul(b, t)), (1 - t) * (float)(a.weight != 0) + t * (float)(b.weight != 0));
    return r;
}
/* ReadVoxelInterpolate (src/Integration/VoxelCube.cpp:6-50): n0 = voxel coordinates of neighbour 0 */
static voxel_t read_voxel_interpolate(const int *n0, const voxel_t *v8, const float *pos, float res)
{
    const float xw = (pos[0] - n0[0] * res) / res, yw = (pos[1] - n0[1] * res) / res, zw = (pos[2] - n0[2] * res) / res;
    const voxel_t r1 = vox_lerp(v8[0], v8[1], xw), r2 = vox_lerp(v8[2], v8[3], xw);
    const voxel_t z1 = vox_lerp(r1, r2, yw);
    const voxel_t r3 = vox_lerp(v8[4], v8[5], xw), r4 = vox_lerp(v8[6], v8[7], xw);
    const voxel_t z2 = vox_lerp(r3, r4, yw);
    return vox_lerp(z1, z2, zw);
}
/* CubePara::GetCubeID(Point3i) / GetVoxelID(Point3i) (VoxelCube.h:63-67,81-86) */
static int cube_of_voxel(int p) { return (int)floor((p + 0.0) / CUBE); }
/* trans * Vector4(x,y,z,1) then head<3>() / w (CubeHandler.h:205-208): Eigen gemv order for all four rows */
static void transform_h(const float *m, float x, float y, float z, float *out)
{
    const float w = row_xyz1(m, 3, x, y, z);
    out[0] = row_xyz1(m, 0, x, y, z) / w;
    out[1] = row_xyz1(m, 1, x, y, z) / w;
    out[2] = row_xyz1(m, 2, x, y, z) / w;
}
static void global_point(const int *id, int n, float res, float *out)
{
    /* CubePara::GetGlobalPoint + VoxelCentroidOffSet (VoxelCube.h:48-61,75-80) for an arbitrary resolution */
    const int xyz[3] = {n & 7, (n >> 3) & 7, n >> 6};
    const float half = res / 2;
    for (int a = 0; a < 3; ++a) out[a] = (float)id[a] * CUBE * res + (xyz[a] * res + half);
}
static const voxel_t *voxel_at(const orc_volume *v, const int *p, voxel_t *tmp)
{
    const int c[3] = {cube_of_voxel(p[0]), cube_of_voxel(p[1]), cube_of_voxel(p[2])};
    const cube_t *cube = find_cube(v, c[0], c[1], c[2]);
    *tmp = kDefaultVoxel;
    if (!cube) return tmp;
    return &cube->vox[(p[0] - c[0] * CUBE) + (p[1] - c[1] * CUBE) * CUBE + (p[2] - c[2] * CUBE) * CUBE * CUBE];
}
/* CubeHandler::Transform (CubeHandler.h:242-298, AddTransformedCube :199-225) and TransformNearest (:299-338,
 * AddTransformedCubeNearest :226-241).  alloc_res is the VoxelResolution of the RESULT handler, which the reference uses
 * for the allocation pass: Transform copies c_para (alloc_res = source resolution); TransformNearest forgets to
 * (:299-305), so its result handler keeps CubePara's default 0.01 whatever the source resolution was. */
orc_volume *orc_volume_transform(const orc_volume *v, const float *trans_cm, int nearest, float alloc_res)
{
    orc_volume *r = orc_volume_create(v->fx, v->fy, v->cx, v->cy, v->width, v->height, v->depth_scale, alloc_res, v->trunc,
                                      v->near_plane, v->far_plane);
    for (long c = 0; c < v->n_cubes; ++c)
        for (int n = 0; n < NVOX; ++n)
        {
            float g[3], p[3];
            global_point(v->cubes[c]->id, n, alloc_res, g);
            transform_h(trans_cm, g[0], g[1], g[2], p);
            if (!nearest)
                for (int a = 0; a < 3; ++a) p[a] = p[a] - alloc_res / 2;
            const int n0[3] = {cvtt_f(floorf(p[0] / alloc_res)), cvtt_f(floorf(p[1] / alloc_res)), cvtt_f(floorf(p[2] / alloc_res))};
            for (int i = 0; i < (nearest ? 1 : 8); ++i)
            {
                const int q[3] = {cube_of_voxel(n0[0] + (i & 1)), cube_of_voxel(n0[1] + ((i >> 1) & 1)), cube_of_voxel(n0[2] + (i >> 2))};
                if (!find_cube(r, q[0], q[1], q[2])) add_cube(r, q[0], q[1], q[2]);
            }
        }
    float inv[16];
    orc_pose_inverse(trans_cm, inv); /* trans.inverse(), evaluated per voxel in the reference */
    const float res = v->res;        /* the second pass runs with the SOURCE handler's c_para (a member function of it) */
    for (long c = 0; c < r->n_cubes; ++c)
        for (int n = 0; n < NVOX; ++n)
        {
            float g[3], p[3];
            voxel_t tmp[8], got;
            global_point(r->cubes[c]->id, n, res, g);
            transform_h(inv, g[0], g[1], g[2], p);
            if (nearest)
            {
                const int q[3] = {cvtt_f(floorf(p[0] / res)), cvtt_f(floorf(p[1] / res)), cvtt_f(floorf(p[2] / res))};
                got = *voxel_at(v, q, &tmp[0]);
            }
            else
            {
                for (int a = 0; a < 3; ++a) p[a] = p[a] - res / 2;
                const int n0[3] = {cvtt_f(floorf(p[0] / res)), cvtt_f(floorf(p[1] / res)), cvtt_f(floorf(p[2] / res))};
                voxel_t v8[8];
                for (int i = 0; i < 8; ++i)
                {
                    const int q[3] = {n0[0] + (i & 1), n0[1] + ((i >> 1) & 1), n0[2] + (i >> 2)};
                    v8[i] = *voxel_at(v, q, &tmp[i]);
                }
                got = read_voxel_interpolate(n0, v8, p, res);
            }
            r->cubes[c]->vox[n] = vox_plus(r->cubes[c]->vox[n], got); /* += on a fresh voxel */
        }
    return r;
}
/* CubeHandler::Merge(another) (CubeHandler.h:145-167); returns 0, or -1 for the resolution mismatch the reference only warns about */
int orc_volume_merge(orc_volume *v, const orc_volume *other)
{
    if (v->res != other->res) return -1;
    for (long c = 0; c < other->n_cubes; ++c)
    {
        const cube_t *o = other->cubes[c];
        cube_t *mine = find_cube(v, o->id[0], o->id[1], o->id[2]);
        if (!mine)
        {
            mine = add_cube(v, o->id[0], o->id[1], o->id[2]);
            memcpy(mine->vox, o->vox, sizeof(o->vox));
        }
        else
            for (int n = 0; n < NVOX; ++n) mine->vox[n] = vox_plus(mine->vox[n], o->vox[n]);
    }
    return 0;
}
float orc_volume_resolution(const orc_volume *v) { return v->res; }

/* ------------------------------------------------------------------------------------------------------- */
/* Marching Cubes                                                                                          */
/* ------------------------------------------------------------------------------------------------------- */
/* integration::MarchingCube, src/Integration/MarchingCube.cpp:9-74 */
int orc_marching_cube_cell(const float *corners, const float *sdf, const float *colors, float *xyz, float *rgb)
{
    int cs = 0;
    for (int i = 0; i < 8; ++i)
        if (sdf[i] > 0) cs |= 1 << i; /* DetermineCase, :18-29 */
    uint64_t row = kMcCases[cs];
    int n = 0;
    for (int k = 0; k < 15; ++k)
    {
        int e = (int)((row >> (4 * k)) & 0xF);
        if (e == 0xF) break;
        int a = kMcEdgeCorners[e][0], b = kMcEdgeCorners[e][1];
        /* InterpolateEdgeVetex, :9-16 */
        float diff = sdf[b] - sdf[a];
        float t = sdf[a] / diff;
        for (int c = 0; c < 3; ++c)
        {
            xyz[3 * n + c] = corners[3 * a + c] - t * (corners[3 * b + c] - corners[3 * a + c]);
            rgb[3 * n + c] = (colors[3 * a + c] + colors[3 * b + c]) / 2;
        }
        ++n;
    }
    return n;
}
/* CubeHandler::ExtractTriangleMesh / GenerateMeshByCube, src/Integration/CubeHandler.cpp:9-44,70-114 */
long orc_volume_extract_mesh(const orc_volume *v, float **xyz_out, float **rgb_out)
{
    long n = 0, cap = 1 << 16;
    float *xyz = (float *)malloc(sizeof(float) * 3 * cap), *rgb = (float *)malloc(sizeof(float) * 3 * cap);
    const float cube_res = CUBE * v->res; /* VoxelCube::GetOrigin, VoxelCube.h:143-147 */
    for (long ci = 0; ci < v->n_cubes; ++ci)
    {
        const cube_t *cube = v->cubes[ci];
        /* the (up to) 8 cubes a boundary cell reaches into, indexed by offset bits x | y<<1 | z<<2 */
        const cube_t *nb[8];
        for (int o = 0; o < 8; ++o)
            nb[o] = o == 0 ? cube : find_cube(v, cube->id[0] + (o & 1), cube->id[1] + ((o >> 1) & 1), cube->id[2] + ((o >> 2) & 1));
        for (int x = 0; x < CUBE; ++x)
            for (int y = 0; y < CUBE; ++y)
                for (int z = 0; z < CUBE; ++z)
                {
                    const int ex = x == CUBE - 1, ey = y == CUBE - 1, ez = z == CUBE - 1;
                    float corners[24], sdf[8], colors[24];
                    int ok = 1;
                    for (int i = 0; i < 8 && ok; ++i)
                    {
                        const int dx = kMcCornerOffset[i][0], dy = kMcCornerOffset[i][1], dz = kMcCornerOffset[i][2];
                        const cube_t *nc = nb[(dx & ex) | ((dy & ey) << 1) | ((dz & ez) << 2)];
                        if (!nc) { ok = 0; break; }
                        const int vid = (x + dx) % CUBE + ((y + dy) % CUBE) * CUBE + ((z + dz) % CUBE) * CUBE * CUBE;
                        const voxel_t *vx = &nc->vox[vid];
                        if (vx->sdf >= 1 || vx->weight <= 0) { ok = 0; break; } /* !IsValid() */
                        for (int a = 0; a < 3; ++a)
                        {
                            corners[3 * i + a] = nc->id[a] * cube_res + v->centroid[vid][a];
                            colors[3 * i + a] = vx->c[a];
                        }
                        sdf[i] = vx->sdf;
                    }
                    if (!ok) continue;
                    if (n + 15 > cap)
                    {
                        cap *= 2;
                        xyz = (float *)realloc(xyz, sizeof(float) * 3 * cap);
                        rgb = (float *)realloc(rgb, sizeof(float) * 3 * cap);
                    }
                    n += orc_marching_cube_cell(corners, sdf, colors, xyz + 3 * n, rgb + 3 * n);
                }
    }
    *xyz_out = xyz;
    *rgb_out = rgb;
    return n;
}

/* ------------------------------------------------------------------------------------------------------- */
/* registration                                                                                            */
/* ------------------------------------------------------------------------------------------------------- */
/* cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (row-major); eigenvalues on the diagonal of A */
static void jacobi_eig(double *A, double *V, int n)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = i == j;
    for (int sweep = 0; sweep < 60; ++sweep)
    {
        double off = 0;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q)
            {
                if (A[p * n + q] == 0) continue;
                double th = (A[q * n + q] - A[p * n + p]) / (2 * A[p * n + q]);
                double t = (th >= 0 ? 1 : -1) / (fabs(th) + sqrt(th * th + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; ++k)
                {
                    double a = A[k * n + p], b = A[k * n + q];
                    A[k * n + p] = c * a - s * b; A[k * n + q] = s * a + c * b;
                }
                for (int k = 0; k < n; ++k)
                {
                    double a = A[p * n + k], b = A[q * n + k];
                    A[p * n + k] = c * a - s * b; A[q * n + k] = s * a + c * b;
                }
                for (int k = 0; k < n; ++k)
                {
                    double a = V[k * n + p], b = V[k * n + q];
                    V[k * n + p] = c * a - s * b; V[k * n + q] = s * a + c * b;
                }
            }
    }
}
/* JacobiSVD(JTJ).solve(b), ICP.cpp:137-138: minimum-norm least squares; singular values below
 * epsilon(float) * 6 * max are treated as zero (Eigen's default rank threshold for the float build) */
static void solve_pinv6(const double *A_in, const double *b, double *x)
{
    double A[36], V[36];
    memcpy(A, A_in, sizeof(A));
    jacobi_eig(A, V, 6);
    double lmax = 0;
    for (int i = 0; i < 6; ++i) lmax = fmax(lmax, fabs(A[i * 6 + i]));
    for (int i = 0; i < 6; ++i) x[i] = 0;
    for (int k = 0; k < 6; ++k)
    {
        double l = A[k * 6 + k];
        if (!(fabs(l) > 6 * 1.1920929e-7 * lmax)) continue;
        double pr = 0;
        for (int i = 0; i < 6; ++i) pr += V[i * 6 + k] * b[i];
        for (int i = 0; i < 6; ++i) x[i] += V[i * 6 + k] * pr / l;
    }
}
/* geometry::Se3ToSE3 -> Sophus::SE3Group::exp (se3.hpp:468-489, so3.hpp:388-412); T row-major here */
static void se3_exp_rm(const double *x, double *T)
{
    const double *w = x + 3;
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}, O2[9], R[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            O2[i * 3 + j] = 0;
            for (int k = 0; k < 3; ++k) O2[i * 3 + j] += O[i * 3 + k] * O[k * 3 + j];
        }
    double a, b, c; /* R = I + a O + b O^2 ; V = I + b O + c O^2 */
    if (th < 1e-10) { a = 1; b = 0.5; c = 1.0 / 6; }
    else { a = sin(th) / th; b = (1 - cos(th)) / th2; c = (th - sin(th)) / (th2 * th); }
    for (int i = 0; i < 9; ++i)
    {
        R[i] = (i % 4 == 0) + a * O[i] + b * O2[i];
        V[i] = (i % 4 == 0) + b * O[i] + c * O2[i];
    }
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = V[i * 3] * x[0] + V[i * 3 + 1] * x[1] + V[i * 3 + 2] * x[2];
    }
    T[12] = T[13] = T[14] = 0; T[15] = 1;
}
void orc_se3_exp(const double *x6, double *T_cm)
{
    double T[16];
    se3_exp_rm(x6, T);
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) T_cm[c * 4 + r] = T[r * 4 + c];
}
/* geometry::EstimateRigidTransformation (Geometry.cpp:107-151) from pair lists, double; T row-major */
static void kabsch(const float *a, const float *b, long n, double *T)
{
    double ma[3] = {0, 0, 0}, mb[3] = {0, 0, 0}, W[9] = {0};
    for (long i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) { ma[c] += a[3 * i + c]; mb[c] += b[3 * i + c]; }
    for (int c = 0; c < 3; ++c) { ma[c] /= n; mb[c] /= n; }
    for (long i = 0; i < n; ++i)
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) W[p * 3 + q] += (a[3 * i + p] - ma[p]) * (b[3 * i + q] - mb[q]);
    /* SVD of W through the eigen-decomposition of W^T W */
    double WtW[9], V[9], U[9], sig[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            WtW[i * 3 + j] = 0;
            for (int k = 0; k < 3; ++k) WtW[i * 3 + j] += W[k * 3 + i] * W[k * 3 + j];
        }
    jacobi_eig(WtW, V, 3);
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (WtW[ord[j] * 4] > WtW[ord[i] * 4]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double Vs[9];
    for (int c = 0; c < 3; ++c)
    {
        sig[c] = sqrt(fmax(WtW[ord[c] * 4], 0));
        for (int i = 0; i < 3; ++i) Vs[i * 3 + c] = V[i * 3 + ord[c]];
    }
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 3; ++i)
        {
            double u = W[i * 3] * Vs[c] + W[i * 3 + 1] * Vs[3 + c] + W[i * 3 + 2] * Vs[6 + c];
            U[i * 3 + c] = sig[c] > 0 ? u / sig[c] : 0;
        }
    if (!(sig[2] > 1e-14 * sig[0]))
    {   /* rank 2: third left vector = u0 x u1 */
        U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
    }
    double R[9];
    for (int pass = 0; pass < 2; ++pass)
    {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
            {
                R[i * 3 + j] = 0;
                for (int k = 0; k < 3; ++k) R[i * 3 + j] += Vs[i * 3 + k] * U[j * 3 + k]; /* R = V U^T */
            }
        double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
        if (det >= 0 || pass) break;
        for (int i = 0; i < 3; ++i) Vs[i * 3 + 2] = -Vs[i * 3 + 2];
    }
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = mb[i] - (R[i * 3] * ma[0] + R[i * 3 + 1] * ma[1] + R[i * 3 + 2] * ma[2]);
    }
    T[12] = T[13] = T[14] = 0; T[15] = 1;
}
/* KDTree::KnnSearch(k = 1) (KDTree.h:177-196): exact; distance as nanoflann's L2_Simple_Adaptor sums it */
void orc_nearest(const float *q, long nq, const float *t, long nt, int32_t *nn)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < nq; ++i)
    {
        float best = FLT_MAX;
        int bi = -1;
        const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
        for (long j = 0; j < nt; ++j)
        {
            float dx = qx - t[3 * j], dy = qy - t[3 * j + 1], dz = qz - t[3 * j + 2];
            float d = dx * dx + dy * dy + dz * dz;
            if (d < best) { best = d; bi = (int)j; }
        }
        nn[i] = bi;
    }
}
/* CountInliers (ICP.cpp:9-30); T column-major float */
static long count_inliers(const float *src, const float *tgt, const int32_t *nn, long ns, const float *T, double thr,
                          int32_t *pairs, double *rmse)
{
    double sum = 0, thr2 = thr * thr;
    long n = 0;
    for (long i = 0; i < ns; ++i)
    {
        if (nn[i] < 0) continue;
        const float *s = src + 3 * i, *t = tgt + 3 * nn[i];
        float e[3];
        for (int r = 0; r < 3; ++r) e[r] = ((T[r] * s[0] + (T[4 + r] * s[1] + T[8 + r] * s[2])) + T[12 + r]) - t[r];
        double err = e[0] * e[0] + (e[1] * e[1] + e[2] * e[2]);
        if (err < thr2) { pairs[2 * n] = (int32_t)i; pairs[2 * n + 1] = nn[i]; ++n; sum += err; }
    }
    *rmse = sqrt(sum / n);
    return n;
}
long orc_icp(const float *src_in, long ns, const float *tgt_in, const float *nrm, long nt, const float *init_T, int max_it,
             double thr, double scaling, double *out_T, double *out_T_iter, int32_t *pairs, double *rmse)
{
    float *src = (float *)malloc(sizeof(float) * 3 * ns), *tgt = (float *)malloc(sizeof(float) * 3 * nt);
    memcpy(src, src_in, sizeof(float) * 3 * ns);
    memcpy(tgt, tgt_in, sizeof(float) * 3 * nt);
    if (scaling != 1)
    {
        if (nrm) { free(src); free(tgt); return -1; } /* ICP.cpp:159-163 */
        for (long i = 0; i < 3 * ns; ++i) src[i] = src[i] * (float)scaling;
        for (long i = 0; i < 3 * nt; ++i) tgt[i] = tgt[i] * (float)scaling;
    }
    float T[16];
    memcpy(T, init_T, sizeof(T));
    float *tp = (float *)malloc(sizeof(float) * 3 * ns);
    int32_t *nn = (int32_t *)malloc(sizeof(int32_t) * ns);
    long n_in = 0;
    for (long i = 0; i < ns; ++i) nn[i] = -1; /* corresponding_index(source.points.size(), -1) (ICP.cpp:58,174) */
    for (int it = 0; it <= max_it; ++it)
    {
        if (it == max_it)
        {
            /* The closing CountInliers (ICP.cpp:90,206) re-tests the corresponding_index of the LAST loop iteration -- the
             * neighbours found under the previous pose, whatever their distance -- with the final pose; no new search. */
            n_in = count_inliers(src, tgt, nn, ns, T, thr, pairs, rmse);
            break;
        }
        /* geometry::TransformPoints (Geometry.cpp:19-27) */
        for (long i = 0; i < ns; ++i)
        {
            const float *s = src + 3 * i;
            float w = row_xyz1(T, 3, s[0], s[1], s[2]);
            for (int r = 0; r < 3; ++r) tp[3 * i + r] = row_xyz1(T, r, s[0], s[1], s[2]) / w;
        }
        orc_nearest(tp, ns, tgt, nt, nn);
        n_in = count_inliers(src, tgt, nn, ns, T, thr, pairs, rmse);
        double dT[16];
        if (nrm)
        {   /* EstimateRigidTransformationPointToPlane (ICP.cpp:108-144) */
            double JTJ[36] = {0}, nJTr[6] = {0}, x[6];
            for (long k = 0; k < n_in; ++k)
            {
                const float *p = tp + 3 * pairs[2 * k], *t = tgt + 3 * pairs[2 * k + 1], *n = nrm + 3 * pairs[2 * k + 1];
                float r = dot3(n, p) - dot3(n, t);
                float row[6] = {n[0], n[1], n[2], p[1] * n[2] - p[2] * n[1], p[2] * n[0] - p[0] * n[2], p[0] * n[1] - p[1] * n[0]};
                for (int a = 0; a < 6; ++a)
                {
                    for (int b = 0; b < 6; ++b) JTJ[a * 6 + b] += (double)(row[a] * row[b]);
                    nJTr[a] -= (double)(r * row[a]);
                }
            }
            solve_pinv6(JTJ, nJTr, x);
            for (int a = 0; a < 6; ++a) x[a] = (float)x[a];
            se3_exp_rm(x, dT);
        }
        else
        {   /* PointToPoint (ICP.cpp:78-86) */
            if (n_in == 0) break;
            float *a = (float *)malloc(sizeof(float) * 3 * n_in), *b = (float *)malloc(sizeof(float) * 3 * n_in);
            for (long k = 0; k < n_in; ++k)
                for (int c = 0; c < 3; ++c) { a[3 * k + c] = tp[3 * pairs[2 * k] + c]; b[3 * k + c] = tgt[3 * pairs[2 * k + 1] + c]; }
            kabsch(a, b, n_in, dT);
            free(a); free(b);
        }
        /* start_T = tmp_T * start_T in float */
        float dTf[16], Tn[16];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) dTf[c * 4 + r] = (float)dT[r * 4 + c];
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r)
                Tn[c * 4 + r] = ((dTf[r] * T[c * 4] + dTf[4 + r] * T[c * 4 + 1]) + dTf[8 + r] * T[c * 4 + 2]) + dTf[12 + r] * T[c * 4 + 3];
        memcpy(T, Tn, sizeof(T));
    }
    for (int i = 0; i < 16; ++i) out_T_iter[i] = T[i];
    /* result.T: Kabsch over the inlier pairs of the (un-scaled) clouds (ICP.cpp:93-105,208-221) */
    if (n_in > 0)
    {
        float *a = (float *)malloc(sizeof(float) * 3 * n_in), *b = (float *)malloc(sizeof(float) * 3 * n_in);
        for (long k = 0; k < n_in; ++k)
            for (int c = 0; c < 3; ++c)
            {
                float sv = src[3 * pairs[2 * k] + c], tv = tgt[3 * pairs[2 * k + 1] + c];
                if (scaling != 1) { sv = sv / (float)scaling; tv = tv / (float)scaling; }
                a[3 * k + c] = sv; b[3 * k + c] = tv;
            }
        double Tk[16];
        kabsch(a, b, n_in, Tk);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) out_T[c * 4 + r] = Tk[r * 4 + c];
        free(a); free(b);
    }
    else
        for (int i = 0; i < 16; ++i) out_T[i] = NAN;
    free(src); free(tgt); free(tp); free(nn);
    return n_in;
}

/* ------------------------------------------------------------------------------------------------------- */
/* dense RGB-D odometry                                                                                    */
/* ------------------------------------------------------------------------------------------------------- */
#define MAX_DIFF_DEPTH 0.05 /* OdometryPredefined.h:4-19 */
#define SOBEL_SCALE 0.125
#define LAMBDA_HYBRID_DEPTH 0.5
#define ODO_MAX_DEPTH 4
#define ODO_MIN_DEPTH 0.5
#define MAX_INLIER_RATIO_DENSE 0.9
#define MIN_INLIER_RATIO_DENSE 0.3
#define ODO_LEVELS 3

static int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
/* cv::cvtColor(CV_RGB2GRAY) on 8-bit data (ImageProcessing.cpp:21-24), OpenCV 4.x fixed point: the FIRST stored
 * byte takes the red coefficient (the reference hands it BGR data from imread) */
void orc_gray_u8(const uint8_t *bgr, int n, uint8_t *gray)
{
    for (int i = 0; i < n; ++i)
        gray[i] = (uint8_t)((9798 * bgr[3 * i] + 19235 * bgr[3 * i + 1] + 3735 * bgr[3 * i + 2] + 16384) >> 15);
}
/* cv::GaussianBlur(3x3, sigma 0) (ImageProcessing.cpp:43-46): separable [1/4 1/2 1/4], BORDER_REFLECT_101, rows first */
void orc_blur3(const float *src, int w, int h, float *dst)
{
    float *tmp = (float *)malloc(sizeof(float) * w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            tmp[y * w + x] = 0.5f * src[y * w + x] + 0.25f * (src[y * w + refl101(x - 1, w)] + src[y * w + refl101(x + 1, w)]);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            dst[y * w + x] = 0.5f * tmp[y * w + x] + 0.25f * (tmp[refl101(y - 1, h) * w + x] + tmp[refl101(y + 1, h) * w + x]);
    free(tmp);
}
/* cv::pyrDown (ImageProcessing.cpp:16): separable [1 4 6 4 1], BORDER_REFLECT_101, every second sample, /256 */
void orc_pyr_down(const float *src, int w, int h, float *dst)
{
    const int ow = w / 2, oh = h / 2;
    float *tmp = (float *)malloc(sizeof(float) * ow * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < ow; ++x)
        {
            const float *r = src + y * w;
            const int c = 2 * x;
            tmp[y * ow + x] = r[refl101(c, w)] * 6.0f + (r[refl101(c - 1, w)] + r[refl101(c + 1, w)]) * 4.0f + r[refl101(c - 2, w)] +
                              r[refl101(c + 2, w)];
        }
    for (int y = 0; y < oh; ++y)
        for (int x = 0; x < ow; ++x)
        {
            const int c = 2 * y;
            dst[y * ow + x] = (tmp[refl101(c, h) * ow + x] * 6.0f + (tmp[refl101(c - 1, h) * ow + x] + tmp[refl101(c + 1, h) * ow + x]) * 4.0f +
                               tmp[refl101(c - 2, h) * ow + x] + tmp[refl101(c + 2, h) * ow + x]) * (1.0f / 256.0f);
        }
    free(tmp);
}
/* cv::Sobel(CV_32F, ksize 3) (ImageProcessing.cpp:25-33): derivative [-1 0 1] along one axis, [1 2 1] along the other */
void orc_sobel3(const float *src, int w, int h, int dx, float *dst)
{
    float *tmp = (float *)malloc(sizeof(float) * w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const float l = src[y * w + refl101(x - 1, w)], r = src[y * w + refl101(x + 1, w)];
            tmp[y * w + x] = dx ? r - l : l + 2.0f * src[y * w + x] + r;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const float u = tmp[refl101(y - 1, h) * w + x], d = tmp[refl101(y + 1, h) * w + x];
            dst[y * w + x] = dx ? u + 2.0f * tmp[y * w + x] + d : d - u;
        }
    free(tmp);
}

struct orc_frame
{
    int w, h, is_u16, preprocessed;
    uint8_t *bgr;
    void *depth;
    float *img[6][ODO_LEVELS]; /* gray, depth, gray dx, gray dy, depth dx, depth dy */
};
orc_frame *orc_frame_create(const uint8_t *bgr, const void *depth, int is_u16, int w, int h)
{
    orc_frame *f = (orc_frame *)calloc(1, sizeof(orc_frame));
    f->w = w; f->h = h; f->is_u16 = is_u16;
    f->bgr = (uint8_t *)malloc((size_t)w * h * 3);
    memcpy(f->bgr, bgr, (size_t)w * h * 3);
    f->depth = malloc((size_t)w * h * (is_u16 ? 2 : 4));
    memcpy(f->depth, depth, (size_t)w * h * (is_u16 ? 2 : 4));
    return f;
}
void orc_frame_destroy(orc_frame *f)
{
    if (!f) return;
    for (int a = 0; a < 6; ++a)
        for (int l = 0; l < ODO_LEVELS; ++l) free(f->img[a][l]);
    free(f->bgr); free(f->depth); free(f);
}
long orc_frame_image(const orc_frame *f, int what, int level, float *out)
{
    const long n = (long)(f->w >> level) * (f->h >> level);
    if (out && f->img[what][level]) memcpy(out, f->img[what][level], sizeof(float) * n);
    return n;
}
/* Odometry::InitializeRGBDDenseTracking (Odometry.cpp:609-620): gray/255 and NaN-masked metric depth, both blurred */
static void odo_initialize(const uint8_t *bgr, const void *depth, int is_u16, int w, int h, float depth_scale, float *gray, float *d32)
{
    const int n = w * h;
    uint8_t *g8 = (uint8_t *)malloc(n);
    float *tmp = (float *)malloc(sizeof(float) * n);
    orc_gray_u8(bgr, n, g8);
    for (int i = 0; i < n; ++i) tmp[i] = g8[i] / 255.0f; /* ConvertColorToIntensity32F, DenseOdometryFunction.cpp:59-71 */
    orc_blur3(tmp, w, h, gray);
    /* ConvertDepthTo32FNaN, DenseOdometryFunction.cpp:26-57 */
    for (int i = 0; i < n; ++i)
    {
        if (is_u16)
        {
            const unsigned short v = ((const unsigned short *)depth)[i];
            tmp[i] = (v > ODO_MIN_DEPTH * depth_scale && v < ODO_MAX_DEPTH * depth_scale) ? v / depth_scale : NAN;
        }
        else
        {
            const float v = ((const float *)depth)[i];
            tmp[i] = (v > ODO_MIN_DEPTH && v < ODO_MAX_DEPTH) ? v : NAN;
        }
    }
    orc_blur3(tmp, w, h, d32);
    free(g8); free(tmp);
}
/* Odometry::CreateImagePyramid (Odometry.cpp:436-449) for the levels above 0 plus the Sobel images of all levels */
static void odo_pyramids(orc_frame *f)
{
    for (int l = 1; l < ODO_LEVELS; ++l)
        for (int a = 0; a < 2; ++a)
        {
            free(f->img[a][l]);
            f->img[a][l] = (float *)malloc(sizeof(float) * (f->w >> l) * (f->h >> l));
            orc_pyr_down(f->img[a][l - 1], f->w >> (l - 1), f->h >> (l - 1), f->img[a][l]);
        }
    for (int l = 0; l < ODO_LEVELS; ++l)
        for (int a = 0; a < 2; ++a)
            for (int dx = 1; dx >= 0; --dx)
            {
                const int slot = 2 + 2 * a + (1 - dx); /* gray dx, gray dy, depth dx, depth dy */
                free(f->img[slot][l]);
                f->img[slot][l] = (float *)malloc(sizeof(float) * (f->w >> l) * (f->h >> l));
                orc_sobel3(f->img[a][l], f->w >> l, f->h >> l, dx, f->img[slot][l]);
            }
}
static void mat3_mul(const float *A, const float *B, float *C) /* row-major; Eigen lazy product: a0b0 + (a1b1 + a2b2) */
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + (A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j]);
}
/* ComputeCorrespondencePixelWise + AddElementToCorrespondenceMap (DenseOdometryFunction.cpp:8-25,72-128).
 * T column-major float.  pairs: (v_s,u_s,v_t,u_t) in raster order of the source; returns the count. */
static long odo_correspondences(const float *sd, const float *td, int w, int h, float fx, float fy, float cx, float cy, const float *T,
                                uint32_t *pairs, long cap)
{
    /* K, K^-1 (closed form of Eigen's 3x3 inverse for an upper-triangular camera matrix is NOT assumed: the generic
     * cofactor formula is evaluated like Eigen's compute_inverse_size3), K R K^-1 and K t */
    const float K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    const float R[9] = {T[0], T[4], T[8], T[1], T[5], T[9], T[2], T[6], T[10]}, t[3] = {T[12], T[13], T[14]};
    float cof[9], Kinv[9], KR[9], M[9], Kt[3];
    /* cofactor_3x3<i,j> = m(i1,j1)*m(i2,j2) - m(i1,j2)*m(i2,j1) with (i1,i2) = ((i+1)%3,(i+2)%3); inverse(i,j) = cof(j,i)*invdet */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            cof[i * 3 + j] = K[i1 * 3 + j1] * K[i2 * 3 + j2] - K[i1 * 3 + j2] * K[i2 * 3 + j1];
        }
    /* determinant from the first column of cofactors: (cof(0,0), cof(1,0), cof(2,0)) . matrix.col(0) */
    const float det = (cof[0] * K[0] + cof[3] * K[3]) + cof[6] * K[6];
    const float invdet = 1.0f / det;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Kinv[i * 3 + j] = cof[j * 3 + i] * invdet;
    mat3_mul(K, R, KR);
    mat3_mul(KR, Kinv, M);
    for (int i = 0; i < 3; ++i) Kt[i] = K[i * 3] * t[0] + (K[i * 3 + 1] * t[1] + K[i * 3 + 2] * t[2]);

    float *wd = (float *)malloc(sizeof(float) * w * h);
    int *wm = (int *)malloc(sizeof(int) * w * h);
    for (int i = 0; i < w * h; ++i) { wd[i] = -1; wm[i] = -1; }
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j)
        {
            const float d_s = sd[i * w + j];
            if (isnan(d_s)) continue;
            float uv[3];
            for (int r = 0; r < 3; ++r)
                uv[r] = ((d_s * M[r * 3]) * (float)j + ((d_s * M[r * 3 + 1]) * (float)i + (d_s * M[r * 3 + 2]) * 1.0f)) + Kt[r];
            const float tds = uv[2];
            const int u_t = cvtt_d(uv[0] / tds + 0.5), v_t = cvtt_d(uv[1] / tds + 0.5);
            if (u_t >= 0 && u_t < w && v_t >= 0 && v_t < h)
            {
                const float d_t = td[v_t * w + u_t];
                if (!isnan(d_t) && fabsf(d_t - tds) < MAX_DIFF_DEPTH)
                {
                    const float existing = wd[v_t * w + u_t];
                    if (existing == -1 || existing > tds) { wm[i * w + j] = v_t * w + u_t; wd[i * w + j] = tds; }
                }
            }
        }
    long n = 0;
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j)
            if (wd[i * w + j] != -1)
            {
                if (pairs && n < cap)
                {
                    pairs[4 * n] = i; pairs[4 * n + 1] = j;
                    pairs[4 * n + 2] = wm[i * w + j] / w; pairs[4 * n + 3] = wm[i * w + j] % w;
                }
                ++n;
            }
    free(wd); free(wm);
    return n;
}
long orc_correspondences(const float *sd, const float *td, int w, int h, float fx, float fy, float cx, float cy, const float *T,
                         uint32_t *pairs, long cap)
{
    return odo_correspondences(sd, td, w, h, fx, fy, cx, cy, T, pairs, cap);
}
/* NormalizeIntensity (DenseOdometryFunction.cpp:129-144) + tool::LinearTransform (ImageProcessing.cpp:56-63) */
static void odo_normalize(float *sg, float *tg, int w, int h, const uint32_t *pairs, long n)
{
    float mean_s = 0, mean_t = 0;
    for (long k = 0; k < n; ++k)
    {
        mean_s += sg[pairs[4 * k] * w + pairs[4 * k + 1]];
        mean_t += tg[pairs[4 * k + 2] * w + pairs[4 * k + 3]];
    }
    mean_s /= (float)n;
    mean_t /= (float)n;
    const float ss = (float)(0.5 / mean_s), st = (float)(0.5 / mean_t);
    for (int i = 0; i < w * h; ++i) sg[i] = sg[i] * ss + 0.0f;
    for (int i = 0; i < w * h; ++i) tg[i] = tg[i] * st + 0.0f;
}
/* DoSingleIteration{,PhotoTerm,DepthTerm} (DenseOdometryFunction.cpp:146-475): Jacobian rows in float exactly as the
 * reference writes them; J^T J, J^T r summed in double (the reference sums sequentially in float32) */
static long odo_iteration(orc_frame *S, orc_frame *Tg, int l, float fx, float fy, float cx, float cy, float *T, int term,
                          uint32_t *pairs, long cap, double *sums_out /* 36 JTJ, 6 JTr, sum r^2; may be NULL */)
{
    const int w = S->w >> l, h = S->h >> l;
    const long n = odo_correspondences(S->img[1][l], Tg->img[1][l], w, h, fx, fy, cx, cy, T, pairs, cap);
    double JTJ[36] = {0}, nJTr[6] = {0}, r2 = 0;
    const float sq_dep = (float)sqrt(LAMBDA_HYBRID_DEPTH), sq_img = (float)sqrt(1.0 - LAMBDA_HYBRID_DEPTH);
    for (long k = 0; k < n; ++k)
    {
        const int v_s = pairs[4 * k], u_s = pairs[4 * k + 1], v_t = pairs[4 * k + 2], u_t = pairs[4 * k + 3];
        /* source_XYZ[v_s][u_s] (TransformToMatXYZ, Geometry.cpp:72-106) */
        const float z = S->img[1][l][v_s * w + u_s];
        float p[3] = {-1, -1, -1};
        if (z > 0) { p[0] = (u_s - cx) * z / fx; p[1] = (v_s - cy) * z / fy; p[2] = z; }
        float q[3];
        for (int r = 0; r < 3; ++r) q[r] = (T[r] * p[0] + (T[4 + r] * p[1] + T[8 + r] * p[2])) + T[12 + r];
        const float invz = (float)(1. / q[2]);
        float J[2][6], res[2];
        int rows = 0;
        if (term == 0 || term == 1)
        {
            const float diff = Tg->img[0][l][v_t * w + u_t] - S->img[0][l][v_s * w + u_s];
            const float gx = (float)(SOBEL_SCALE * Tg->img[2][l][v_t * w + u_t]), gy = (float)(SOBEL_SCALE * Tg->img[3][l][v_t * w + u_t]);
            const float c0 = gx * fx * invz, c1 = gy * fy * invz, c2 = -(c0 * q[0] + c1 * q[1]) * invz;
            const float s = term == 0 ? sq_img : 1.0f;
            float *j = J[rows];
            j[0] = c0; j[1] = c1; j[2] = c2;
            j[3] = -q[2] * c1 + q[1] * c2; j[4] = q[2] * c0 - q[0] * c2; j[5] = -q[1] * c0 + q[0] * c1;
            if (term == 0) for (int a = 0; a < 6; ++a) j[a] = s * j[a];
            res[rows] = term == 0 ? s * diff : diff;
            ++rows;
        }
        if (term == 0 || term == 2)
        {
            float gx = (float)(SOBEL_SCALE * Tg->img[4][l][v_t * w + u_t]), gy = (float)(SOBEL_SCALE * Tg->img[5][l][v_t * w + u_t]);
            if (isnan(gx)) gx = 0;
            if (isnan(gy)) gy = 0;
            const float diff = Tg->img[1][l][v_t * w + u_t] - q[2];
            const float d0 = gx * fx * invz, d1 = gy * fy * invz, d2 = -(d0 * q[0] + d1 * q[1]) * invz;
            float *j = J[rows];
            j[0] = d0; j[1] = d1; j[2] = d2 - 1.0f;
            j[3] = (-q[2] * d1 + q[1] * d2) - q[1]; j[4] = (q[2] * d0 - q[0] * d2) + q[0]; j[5] = -q[1] * d0 + q[0] * d1;
            if (term == 0) for (int a = 0; a < 6; ++a) j[a] = sq_dep * j[a];
            res[rows] = term == 0 ? sq_dep * diff : diff;
            ++rows;
        }
        for (int r = 0; r < rows; ++r)
            for (int a = 0; a < 6; ++a)
            {
                /* float rows, products and sums in double (the product of two floats is exact in double): the reference's float64
                 * build does the same with double rows, its float32 build sums float products sequentially in float */
                for (int b = 0; b < 6; ++b) JTJ[a * 6 + b] += (double)J[r][a] * (double)J[r][b];
                nJTr[a] -= (double)J[r][a] * (double)res[r];
            }
        for (int r = 0; r < rows; ++r) r2 += (double)res[r] * (double)res[r];
    }
    if (sums_out)
    {
        memcpy(sums_out, JTJ, sizeof(JTJ));
        for (int a = 0; a < 6; ++a) sums_out[36 + a] = -nJTr[a];
        sums_out[42] = r2;
    }
    /* delta = JTJ.ldlt().solve(-JTr); T = Se3ToSE3(delta) * T */
    double x[6], dT[16];
    solve_pinv6(JTJ, nJTr, x);
    for (int a = 0; a < 6; ++a) x[a] = (float)x[a];
    se3_exp_rm(x, dT);
    float dTf[16], Tn[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) dTf[c * 4 + r] = (float)dT[r * 4 + c];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
            Tn[c * 4 + r] = ((dTf[r] * T[c * 4] + dTf[4 + r] * T[c * 4 + 1]) + dTf[8 + r] * T[c * 4 + 2]) + dTf[12 + r] * T[c * 4 + 3];
    memcpy(T, Tn, sizeof(Tn));
    return n;
}
/* Odometry::MultiScaleComputing (Odometry.cpp:621-685) + the result assembly of DenseTracking (:596-607) */
static void odo_multiscale(orc_frame *S, orc_frame *Tg, float fx, float fy, float cx, float cy, const float *init_T, int term,
                           orc_tracking_result *out, uint32_t *pixel_pairs, long cap)
{
    static const int iters[ODO_LEVELS] = {4, 8, 16}; /* Odometry.h:170, indexed by level */
    const int w = S->w, h = S->h;
    float T[16];
    memcpy(T, init_T, sizeof(T));
    uint32_t *pairs = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)w * h);
    float cam[ODO_LEVELS][4] = {{fx, fy, cx, cy}};
    for (int l = 1; l < ODO_LEVELS; ++l)
        for (int a = 0; a < 4; ++a) cam[l][a] = cam[l - 1][a] / 2; /* GenerateNextPyramid, Camera.h:38-42 */
    long n = 0;
    out->iterations = 0;
    for (int l = ODO_LEVELS - 1; l >= 0; --l)
        for (int j = 0; j != iters[l]; ++j)
        {
            n = odo_iteration(S, Tg, l, cam[l][0], cam[l][1], cam[l][2], cam[l][3], T, term, pairs, (long)w * h, NULL);
            if (out->iterations < 64)
            {
                out->corr_per_iteration[out->iterations] = n;
                for (int e = 0; e < 16; ++e) out->T_per_iteration[out->iterations][e] = T[e];
            }
            out->iterations++;
            if ((float)n / (h * w) > MAX_INLIER_RATIO_DENSE) break;
        }
    /* correspondence_set pairs xyz_s[v_s][u_s] with xyz_t[v_s][u_s] -- the SAME source pixel on the target's XYZ image
     * (reference quirk, Odometry.cpp:672-682); rmse = ComputeReprojectionError3D (Geometry.cpp:45-59) */
    double sum = 0;
    const int wl = w >> (n ? 0 : 0);
    for (long k = 0; k < n; ++k)
    {
        const int v_s = pairs[4 * k], u_s = pairs[4 * k + 1];
        float a[3] = {-1, -1, -1}, b[3] = {-1, -1, -1};
        const float zs = S->img[1][0][v_s * wl + u_s], zt = Tg->img[1][0][v_s * wl + u_s];
        if (zs > 0) { a[0] = (u_s - cx) * zs / fx; a[1] = (v_s - cy) * zs / fy; a[2] = zs; }
        if (zt > 0) { b[0] = (u_s - cx) * zt / fx; b[1] = (v_s - cy) * zt / fy; b[2] = zt; }
        /* TransformPoint: T*(x,y,z,1) then /w, minus second, squaredNorm, accumulated in double */
        const float wv = row_xyz1(T, 3, a[0], a[1], a[2]);
        float e[3];
        for (int r = 0; r < 3; ++r) e[r] = row_xyz1(T, r, a[0], a[1], a[2]) / wv - b[r];
        sum += (double)(e[0] * e[0] + (e[1] * e[1] + e[2] * e[2]));
    }
    out->rmse = sqrt(sum / n);
    out->n_correspondences = n;
    out->tracking_success = (float)n / (h * w) >= MIN_INLIER_RATIO_DENSE;
    for (int e = 0; e < 16; ++e) out->T[e] = T[e];
    if (pixel_pairs) memcpy(pixel_pairs, pairs, sizeof(uint32_t) * 4 * (size_t)(n < cap ? n : cap));
    free(pairs);
}
static void odo_preprocess(orc_frame *f, float depth_scale)
{
    const int n = f->w * f->h;
    for (int a = 0; a < 2; ++a) { free(f->img[a][0]); f->img[a][0] = (float *)malloc(sizeof(float) * n); }
    odo_initialize(f->bgr, f->depth, f->is_u16, f->w, f->h, depth_scale, f->img[0][0], f->img[1][0]);
    odo_pyramids(f);
    f->preprocessed = 1;
}
void orc_dense_tracking_frames(orc_frame *S, orc_frame *Tg, float fx, float fy, float cx, float cy, float depth_scale,
                               const float *init_T, int term, orc_tracking_result *out, uint32_t *pixel_pairs, long cap)
{
    /* Odometry.cpp:571-587: pre-process once per frame (IsPreprocessedDense) */
    if (!S->preprocessed) odo_preprocess(S, depth_scale);
    if (!Tg->preprocessed) odo_preprocess(Tg, depth_scale);
    /* :588-595: correspondences at identity on the level-0 depth, then NormalizeIntensity on frame.gray IN PLACE -- which
     * is also pyramid level 0 (CreatePyramid pushes the same cv::Mat header), but not levels 1-2 nor any Sobel image */
    const int w = S->w, h = S->h;
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    uint32_t *pairs = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)w * h);
    const long n = odo_correspondences(S->img[1][0], Tg->img[1][0], w, h, fx, fy, cx, cy, I, pairs, (long)w * h);
    odo_normalize(S->img[0][0], Tg->img[0][0], w, h, pairs, n);
    free(pairs);
    odo_multiscale(S, Tg, fx, fy, cx, cy, init_T, term, out, pixel_pairs, cap);
}
void orc_dense_tracking(const uint8_t *sb, const uint8_t *tb, const void *sdp, const void *tdp, int is_u16, int w, int h, float fx,
                        float fy, float cx, float cy, float depth_scale, const float *init_T, int term, orc_tracking_result *out,
                        uint32_t *pixel_pairs, long cap)
{
    /* Odometry.cpp:463-523: initialise both, normalise the full-resolution gray images, THEN build the pyramids */
    orc_frame *S = orc_frame_create(sb, sdp, is_u16, w, h), *Tg = orc_frame_create(tb, tdp, is_u16, w, h);
    const int n_px = w * h;
    for (int a = 0; a < 2; ++a)
    {
        S->img[a][0] = (float *)malloc(sizeof(float) * n_px);
        Tg->img[a][0] = (float *)malloc(sizeof(float) * n_px);
    }
    odo_initialize(sb, sdp, is_u16, w, h, depth_scale, S->img[0][0], S->img[1][0]);
    odo_initialize(tb, tdp, is_u16, w, h, depth_scale, Tg->img[0][0], Tg->img[1][0]);
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    uint32_t *pairs = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)n_px);
    const long n = odo_correspondences(S->img[1][0], Tg->img[1][0], w, h, fx, fy, cx, cy, I, pairs, n_px);
    odo_normalize(S->img[0][0], Tg->img[0][0], w, h, pairs, n);
    free(pairs);
    odo_pyramids(S);
    odo_pyramids(Tg);
    odo_multiscale(S, Tg, fx, fy, cx, cy, init_T, term, out, pixel_pairs, cap);
    orc_frame_destroy(S);
    orc_frame_destroy(Tg);
}

/* one teacher-forced iteration at `level` (frames must be pre-processed): T_cm in/out, sums = 36 JTJ + 6 JTr + sum r^2 */
long orc_single_iteration(orc_frame *S, orc_frame *Tg, int level, float fx, float fy, float cx, float cy, float *T_cm, int term,
                          double *sums43, uint32_t *pairs, long cap)
{
    for (int l = 0; l < level; ++l) { fx /= 2; fy /= 2; cx /= 2; cy /= 2; }
    return odo_iteration(S, Tg, level, fx, fy, cx, cy, T_cm, term, pairs, cap, sums43);
}
/* pre-process without tracking (Odometry.cpp:571-587) */
void orc_frame_preprocess(orc_frame *f, float depth_scale)
{
    if (!f->preprocessed) odo_preprocess(f, depth_scale);
}

/* ---------------------------------------------------------------------------------------------------------
 * Caller-side depth pre-filter: what every fusion main runs right before IntegrateImage
 * (example/ImageSequenceIntegration.cpp:36-38, example/DenseFusion/DenseFusion.cpp:92-95).
 * ------------------------------------------------------------------------------------------------------- */
/* tool::ConvertDepthTo32F (src/Tool/ImageProcessing.cpp:68-91) */
void orc_convert_depth_32f(const void *depth, int is_u16, long n, float depth_scale, float *out)
{
    for (long i = 0; i < n; ++i)
    {
        if (is_u16)
        {
            out[i] = ((const unsigned short *)depth)[i] / depth_scale;
            if (out[i] < 0) out[i] = 0;
        }
        else
            out[i] = ((const float *)depth)[i];
    }
}
/* tool::BilateralFilter (ImageProcessing.cpp:64-67) = cv::bilateralFilter(src, dst, d, sigma_color, sigma_space) on a
 * CV_32FC1 image, BORDER_REFLECT_101.  OpenCV is a third-party dependency that is not in the reference tree (README pins
 * "OpenCV 3.4"); this restates its published float algorithm (modules/imgproc/src/bilateral_filter.dispatch.cpp,
 * bilateralFilter_32f): circular mask of radius d/2; spatial weights exp(-r^2/(2 sigma_space^2)); the range weight
 * exp(-dv^2/(2 sigma_color^2)) read from a (1<<12)-bin table over [0, max-min] with linear interpolation; weighted mean.
 * Summation order (neighbours in raster order, then the centre pixel with weight 1) and separately rounded float
 * operations are chosen to be closest to cv2 4.13's non-dispatched code path; OpenCV's own code paths differ from each other
 * in the last bits (tests/golden/gen_golden_filters.py records both), so this boundary is pinned to a tolerance only. */
void orc_bilateral_tables(float vmin, float vmax, int d, double sigma_color, double sigma_space, float *lut4098, float *scale_index,
                          float *space_weight, int *space_dy, int *space_dx, int *n_taps)
{
    const int K = 1 << 12, radius = d / 2;
    const double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
    const float len = (float)((double)vmax - (double)vmin);
    *scale_index = K / len;
    float last = 1.0f;
    for (int i = 0; i < K + 2; ++i)
    {
        if (last > 0.0f)
        {
            const double val = i / *scale_index;
            lut4098[i] = (float)exp(val * val * gc);
            last = lut4098[i];
        }
        else
            lut4098[i] = 0.0f;
    }
    int m = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j)
        {
            const double r = sqrt((double)i * i + (double)j * j);
            if (r > radius || (i == 0 && j == 0)) continue;
            space_weight[m] = (float)exp(r * r * gs);
            space_dy[m] = i;
            space_dx[m] = j;
            ++m;
        }
    *n_taps = m;
}
void orc_bilateral_filter(const float *src, int w, int h, int d, double sigma_color, double sigma_space, float *dst)
{
    /* OpenCV's argument normalisation */
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    int radius = d <= 0 ? (int)lround(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    d = radius * 2 + 1;
    float vmin = src[0], vmax = src[0];
    for (long i = 1; i < (long)w * h; ++i)
    {
        if (src[i] < vmin) vmin = src[i];
        if (src[i] > vmax) vmax = src[i];
    }
    if (fabs((double)vmin - (double)vmax) < FLT_EPSILON)
    {
        memcpy(dst, src, sizeof(float) * (size_t)w * h);
        return;
    }
    float *lut = (float *)malloc(sizeof(float) * 4098), *sw = (float *)malloc(sizeof(float) * d * d), scale;
    int *dy = (int *)malloc(sizeof(int) * d * d), *dx = (int *)malloc(sizeof(int) * d * d), taps;
    orc_bilateral_tables(vmin, vmax, d, sigma_color, sigma_space, lut, &scale, sw, dy, dx, &taps);
#pragma omp parallel for
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const float v0 = src[(long)y * w + x];
            float sum = 0.0f, wsum = 0.0f;
            for (int k = 0; k < taps; ++k)
            {
                const float v = src[(long)refl101(y + dy[k], h) * w + refl101(x + dx[k], w)];
                float alpha = fabsf(v - v0) * scale;
                const int idx = (int)alpha; /* cvFloor of a non-negative value */
                alpha -= (float)idx;
                const float wk = sw[k] * (lut[idx] + alpha * (lut[idx + 1] - lut[idx]));
                sum += v * wk;
                wsum += wk;
            }
            sum += v0;
            wsum += 1.0f;
            dst[(long)y * w + x] = sum / wsum;
        }
    free(lut); free(sw); free(dy); free(dx);
}

/* ---------------------------------------------------------------------------------------------------------
 * Mesh post-processing: what the fusion mains run on the Marching-Cubes output before writing the
 * PLY (example/DenseFusion/DenseFusion.cpp:105, example/ImageSequenceIntegration.cpp:56).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct
{
    int key[3];
    int rep;     /* grid_map[..].first: the first vertex (in triangle order) that fell into the cell */
    int count;   /* grid_map[..].second */
    float sum[3];/* grid_to_point[..] */
    int used;
} cell_t;
static unsigned long cell_hash(const int *k)
{
    return ((unsigned long)(long)k[0] * 73856093ul) ^ ((unsigned long)(long)k[1] * 19349663ul) ^ ((unsigned long)(long)k[2] * 83492791ul);
}
/* TriangleMesh::ClusteringSimplify (src/Geometry/TriangleMesh.cpp:53-58) = ClusteringSimplification
 * (src/Geometry/MeshSimplification.cpp:579-657) + UpdateMesh (:114-139) + CompactMesh (:314-343) on a mesh WITHOUT normals
 * (Marching-Cubes output; with normals CompactMesh would recompute them, see orc_compute_normals).
 * points/colors are updated in place and compacted, triangles likewise; returns the new counts.  colors may be NULL. */
int orc_clustering_simplify(float *points, float *colors, long *n_points, uint32_t *tri, long *n_tris, float grid_len)
{
    if (grid_len <= 0) return -1; /* "[ClusteringMeshSimplification]::[ERROR]::Grid length cannot be less than 0." */
    const long nt = *n_tris, nv = *n_points;
    long cap = 16;
    while (cap < 6 * nt + 16) cap <<= 1;
    cell_t *cells = (cell_t *)calloc(cap, sizeof(cell_t));
    unsigned char *deleted = (unsigned char *)calloc((size_t)(nt > 0 ? nt : 1), 1);
    for (long i = 0; i < nt; ++i)
    {
        cell_t *c3[3];
        const uint32_t vtx[3] = {tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]};
        for (int k = 0; k < 3; ++k)
        {
            const float *p = points + 3 * (long)vtx[k];
            /* GetGridIndex (:575-578) */
            const int key[3] = {(int)floorf(p[0] / grid_len), (int)floorf(p[1] / grid_len), (int)floorf(p[2] / grid_len)};
            unsigned long h = cell_hash(key) & (cap - 1);
            while (cells[h].used && (cells[h].key[0] != key[0] || cells[h].key[1] != key[1] || cells[h].key[2] != key[2])) h = (h + 1) & (cap - 1);
            cell_t *c = &cells[h];
            if (!c->used)
            {
                c->used = 1; c->key[0] = key[0]; c->key[1] = key[1]; c->key[2] = key[2];
                c->rep = (int)vtx[k]; c->count = 1;
                c->sum[0] = p[0]; c->sum[1] = p[1]; c->sum[2] = p[2];
            }
            else
            {
                c->sum[0] += p[0]; c->sum[1] += p[1]; c->sum[2] += p[2];
                c->count += 1;
            }
            c3[k] = c;
        }
        const int v1 = c3[0]->rep, v2 = c3[1]->rep, v3 = c3[2]->rep;
        if (v1 == v2 || v1 == v3 || v2 == v3) deleted[i] = 1;
        else { tri[3 * i] = v1; tri[3 * i + 1] = v2; tri[3 * i + 2] = v3; }
    }
    /* UpdateMesh: drop deleted triangles, move every representative to the mean of its cell */
    long kept = 0;
    for (long i = 0; i < nt; ++i)
    {
        if (deleted[i]) continue;
        tri[3 * kept] = tri[3 * i]; tri[3 * kept + 1] = tri[3 * i + 1]; tri[3 * kept + 2] = tri[3 * i + 2];
        ++kept;
    }
    for (long h = 0; h < cap; ++h)
        if (cells[h].used)
            for (int a = 0; a < 3; ++a) points[3 * (long)cells[h].rep + a] = cells[h].sum[a] / (float)cells[h].count;
    /* CompactMesh: keep the points some triangle still references, in index order */
    long *remap = (long *)malloc(sizeof(long) * (nv ? nv : 1));
    for (long i = 0; i < nv; ++i) remap[i] = -1;
    for (long i = 0; i < 3 * kept; ++i) remap[tri[i]] = 0;
    long re_local = 0;
    for (long i = 0; i < nv; ++i)
    {
        if (remap[i] < 0) continue;
        for (int a = 0; a < 3; ++a) points[3 * re_local + a] = points[3 * i + a];
        if (colors)
            for (int a = 0; a < 3; ++a) colors[3 * re_local + a] = colors[3 * i + a];
        remap[i] = re_local++;
    }
    for (long i = 0; i < 3 * kept; ++i) tri[i] = (uint32_t)remap[tri[i]];
    *n_points = re_local;
    *n_tris = kept;
    free(cells); free(deleted); free(remap);
    return 0;
}
/* Eigen Vector3f::normalize(): z = squaredNorm() (redux order a0 + (a1 + a2)); if (z > 0) v /= sqrt(z) */
static void normalize3(float *v)
{
    const float z = v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]);
    if (z > 0)
    {
        const float n = sqrtf(z);
        v[0] /= n; v[1] /= n; v[2] /= n;
    }
}
/* TriangleMesh::ComputeNormals (src/Geometry/TriangleMesh.cpp:95-127): unit face normals, vertex normal = normalised sum of
 * the normals of the faces that reference the vertex, in reference order (ascending triangle index, UpdateReferences
 * MeshSimplification.cpp:531-541 -- a triangle naming a vertex twice contributes twice). */
void orc_compute_normals(const float *points, long n_points, const uint32_t *tri, long n_tris, float *normals)
{
    float *fn = (float *)malloc(sizeof(float) * 3 * (n_tris ? n_tris : 1));
    for (long i = 0; i < n_tris; ++i)
    {
        const float *p1 = points + 3 * (long)tri[3 * i], *p2 = points + 3 * (long)tri[3 * i + 1], *p3 = points + 3 * (long)tri[3 * i + 2];
        const float a[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, b[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
        float *n = fn + 3 * i;
        n[0] = a[1] * b[2] - a[2] * b[1];
        n[1] = a[2] * b[0] - a[0] * b[2];
        n[2] = a[0] * b[1] - a[1] * b[0];
        normalize3(n);
    }
    for (long i = 0; i < 3 * n_points; ++i) normals[i] = 0;
    for (long i = 0; i < n_tris; ++i)
        for (int k = 0; k < 3; ++k)
        {
            float *v = normals + 3 * (long)tri[3 * i + k];
            v[0] += fn[3 * i]; v[1] += fn[3 * i + 1]; v[2] += fn[3 * i + 2];
        }
    for (long i = 0; i < n_points; ++i) normalize3(normals + 3 * i);
    free(fn);
}

/* PointCloud::DownSample (src/Geometry/PointCloud.cpp:145-189): voxel-grid down-sampling.  The first point (in input order) of
 * every grid cell opens an output slot, later points of the cell are added to it in input order (float sums), every slot is
 * finally divided by its population; colours and normals (optional) are treated alike.  Returns the number of output points;
 * outputs must hold n entries. */
long orc_downsample(const float *points, const float *colors, const float *normals, long n, float grid_len, float *out_points,
                    float *out_colors, float *out_normals)
{
    long cap = 16;
    while (cap < 2 * n + 16) cap <<= 1;
    cell_t *cells = (cell_t *)calloc(cap, sizeof(cell_t)); /* rep = output slot, count = population */
    long ptr = 0;
    for (long i = 0; i < n; ++i)
    {
        const float *p = points + 3 * i;
        const int key[3] = {(int)floorf(p[0] / grid_len), (int)floorf(p[1] / grid_len), (int)floorf(p[2] / grid_len)};
        unsigned long h = cell_hash(key) & (cap - 1);
        while (cells[h].used && (cells[h].key[0] != key[0] || cells[h].key[1] != key[1] || cells[h].key[2] != key[2])) h = (h + 1) & (cap - 1);
        cell_t *c = &cells[h];
        if (!c->used)
        {
            c->used = 1; c->key[0] = key[0]; c->key[1] = key[1]; c->key[2] = key[2];
            c->rep = (int)ptr; c->count = 1;
            for (int a = 0; a < 3; ++a)
            {
                out_points[3 * ptr + a] = p[a];
                if (colors) out_colors[3 * ptr + a] = colors[3 * i + a];
                if (normals) out_normals[3 * ptr + a] = normals[3 * i + a];
            }
            ++ptr;
        }
        else
        {
            for (int a = 0; a < 3; ++a)
            {
                out_points[3 * (long)c->rep + a] += p[a];
                if (colors) out_colors[3 * (long)c->rep + a] += colors[3 * i + a];
                if (normals) out_normals[3 * (long)c->rep + a] += normals[3 * i + a];
            }
            c->count += 1;
        }
    }
    for (long h = 0; h < cap; ++h)
        if (cells[h].used)
            for (int a = 0; a < 3; ++a)
            {
                out_points[3 * (long)cells[h].rep + a] /= (float)cells[h].count;
                if (colors) out_colors[3 * (long)cells[h].rep + a] /= (float)cells[h].count;
                if (normals) out_normals[3 * (long)cells[h].rep + a] /= (float)cells[h].count;
            }
    free(cells);
    return ptr;
}

/* ---------------------------------------------------------------------------------------------------------
 * PointCloud::EstimateNormals (src/Geometry/PointCloud.cpp:102-144): the step in front of registration::PointToPlane when
 * the clouds come without normals (example/ICPTest.cpp:27-33).  For every point: the knn nearest points (nanoflann
 * knnSearch: ascending squared distance, the point itself first), cut where the SQUARED distance exceeds `radius`
 * (KDTree.h:248-254 compares dist^2 with radius), then geometry::FitPlane (Geometry.cpp:172-199): float mean, float
 * covariance W, Eigen::JacobiSVD<MatrixX>(W), normal = third column of U, normalised.
 * Eigen 3.3.7 (vendored in the reference, 3rdparty/Eigen) is restated for the 3x3 float case: SVD/JacobiSVD.h:660-780,
 * misc/RealSvd2x2.h:19-50, Jacobi/Jacobi.h:83-113 (makeJacobi), :300-420 (apply_rotation_in_the_plane: x' = c x + s y,
 * y' = -s x + c y, no FMA in the reference's SSE build).  Neighbours at exactly equal distance are ordered by index here;
 * nanoflann orders them by tree traversal, so clouds with exact ties can differ in the last bits of W.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct { float c, s; } jrot_t;
/* JacobiRotation::makeJacobi(x, y, z) */
static jrot_t make_jacobi(float x, float y, float z)
{
    jrot_t r;
    const float deno = 2.0f * fabsf(y);
    if (deno < FLT_MIN) { r.c = 1.0f; r.s = 0.0f; return r; }
    const float tau = (x - z) / deno;
    const float w = sqrtf(tau * tau + 1.0f);
    const float t = tau > 0.0f ? 1.0f / (tau + w) : 1.0f / (tau - w);
    const float sign_t = t > 0.0f ? 1.0f : -1.0f;
    const float n = 1.0f / sqrtf(t * t + 1.0f);
    r.s = -sign_t * (y / fabsf(y)) * fabsf(t) * n;
    r.c = n;
    return r;
}
/* apply_rotation_in_the_plane on two strided 3-vectors (or 2-vectors) */
static void rot_apply(float *x, int incx, float *y, int incy, int n, jrot_t j)
{
    if (j.c == 1.0f && j.s == 0.0f) return;
    for (int i = 0; i < n; ++i)
    {
        const float xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = j.c * xi + j.s * yi;
        y[i * incy] = -j.s * xi + j.c * yi;
    }
}
/* JacobiSVD<MatrixXf>(W 3x3, ComputeThinU | ComputeThinV): U (row-major 3x3) and the singular values, sorted descending */
static void jacobi_svd3_uv(const float *Win, float *U, float *V, float *sv);
static void jacobi_svd3(const float *Win, float *U, float *sv) { jacobi_svd3_uv(Win, U, NULL, sv); }
/* the same with V (row-major 3x3) when V != NULL: m_matrixV.applyOnTheRight(p, q, j_right), JacobiSVD.h:~730 */
static void jacobi_svd3_uv(const float *Win, float *U, float *V, float *sv)
{
    float W[9]; /* row-major work matrix */
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) if (fabsf(Win[i]) > scale) scale = fabsf(Win[i]);
    if (scale == 0.0f) scale = 1.0f;
    for (int i = 0; i < 9; ++i) W[i] = Win[i] / scale;
    for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    if (V) for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    const float precision = 2.0f * FLT_EPSILON, consider_zero = FLT_MIN;
    float max_diag = fmaxf(fabsf(W[0]), fmaxf(fabsf(W[4]), fabsf(W[8])));
    int finished = 0;
    while (!finished)
    {
        finished = 1;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q)
            {
                const float threshold = fmaxf(consider_zero, precision * max_diag);
                if (fabsf(W[p * 3 + q]) > threshold || fabsf(W[q * 3 + p]) > threshold)
                {
                    finished = 0;
                    /* real_2x2_jacobi_svd */
                    float m[4] = {W[p * 3 + p], W[p * 3 + q], W[q * 3 + p], W[q * 3 + q]};
                    jrot_t rot1;
                    const float t = m[0] + m[3], d = m[2] - m[1];
                    if (fabsf(d) < FLT_MIN) { rot1.s = 0.0f; rot1.c = 1.0f; }
                    else
                    {
                        const float u = t / d;
                        const float tmp = sqrtf(1.0f + u * u);
                        rot1.s = 1.0f / tmp;
                        rot1.c = u / tmp;
                    }
                    rot_apply(&m[0], 1, &m[2], 1, 2, rot1);                 /* m.applyOnTheLeft(0, 1, rot1): rows 0 and 1 */
                    const jrot_t j_right = make_jacobi(m[0], m[1], m[3]);
                    const jrot_t jrt = {j_right.c, -j_right.s};             /* transpose() */
                    const jrot_t j_left = {rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c}; /* rot1 * j_right^T */
                    const jrot_t jlt = {j_left.c, -j_left.s};
                    rot_apply(&W[p * 3], 1, &W[q * 3], 1, 3, j_left);       /* workMatrix.applyOnTheLeft(p, q, j_left): rows */
                    rot_apply(&U[p], 3, &U[q], 3, 3, (jrot_t){jlt.c, -jlt.s}); /* U.applyOnTheRight(p, q, j_left^T): columns, with (j^T)^T */
                    rot_apply(&W[p], 3, &W[q], 3, 3, jrt);                  /* workMatrix.applyOnTheRight(p, q, j_right): columns, with j^T */
                    if (V) rot_apply(&V[p], 3, &V[q], 3, 3, jrt);           /* m_matrixV.applyOnTheRight(p, q, j_right) */
                    max_diag = fmaxf(max_diag, fmaxf(fabsf(W[p * 3 + p]), fabsf(W[q * 3 + q])));
                }
            }
    }
    for (int i = 0; i < 3; ++i)
    {
        const float a = W[i * 3 + i];
        sv[i] = fabsf(a);
        if (a < 0.0f)
            for (int r = 0; r < 3; ++r) U[r * 3 + i] = -U[r * 3 + i];
    }
    for (int i = 0; i < 3; ++i) sv[i] *= scale;
    for (int i = 0; i < 3; ++i)
    {
        int pos = i;
        for (int k = i + 1; k < 3; ++k) if (sv[k] > sv[pos]) pos = k;
        if (sv[pos] == 0.0f) break;
        if (pos != i)
        {
            const float tmp = sv[i]; sv[i] = sv[pos]; sv[pos] = tmp;
            for (int r = 0; r < 3; ++r) { const float u = U[r * 3 + i]; U[r * 3 + i] = U[r * 3 + pos]; U[r * 3 + pos] = u; }
            if (V) for (int r = 0; r < 3; ++r) { const float v = V[r * 3 + i]; V[r * 3 + i] = V[r * 3 + pos]; V[r * 3 + pos] = v; }
        }
    }
}
/* geometry::FitPlane (Geometry.cpp:172-199) on the listed points -> normal; zero for fewer than 3 points */
static void fit_plane_normal(const float *pts, const int *idx, int n, float *normal)
{
    normal[0] = normal[1] = normal[2] = 0.0f;
    if (n < 3) return;
    float sum[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) sum[a] += pts[3 * (long)idx[i] + a];
    const float mean[3] = {sum[0] / (float)n, sum[1] / (float)n, sum[2] / (float)n};
    float W[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i)
    {
        const float d[3] = {pts[3 * (long)idx[i]] - mean[0], pts[3 * (long)idx[i] + 1] - mean[1], pts[3 * (long)idx[i] + 2] - mean[2]};
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) W[r * 3 + c] += d[r] * d[c];
    }
    for (int i = 0; i < 9; ++i) W[i] = W[i] / (float)n;
    float U[9], sv[3];
    jacobi_svd3(W, U, sv);
    normal[0] = U[2]; normal[1] = U[5]; normal[2] = U[8];
    normalize3(normal);
}
void orc_estimate_normals(const float *pts, long n, float radius, int knn, float *normals);

/* ------------------------------------------------------------------------------------------------------------------
 * nanoflann kd-tree as the reference drives it (src/Geometry/KDTree.h:60-262 over the vendored, locally modified
 * 3rdparty/nanoflann/include/nanoflann.hpp): L2_Simple_Adaptor<float>, DIM 3, leaf_max_size 10.
 *   build    nanoflann.hpp:843-1010  (computeMinMax, divideTree, middleSplit_, planeSplit)
 *   search   nanoflann.hpp:1012-1028,1228-1295,1354-1417 (computeInitialDistances, findNeighbors, searchLevel)
 *   results  nanoflann.hpp:150-214 (KNNResultSet), :216-262 (RadiusResultSet with the max_neighbors early stop added
 *            by the reference), :1285-1295 (radiusSearch + std::sort by distance only)
 * Everything here is what decides WHICH neighbours come back and IN WHAT ORDER when distances tie or when the radius
 * search stops early, which the float sums downstream (FitPlane, FPFH) depend on.
 * ------------------------------------------------------------------------------------------------------------------ */
#define KD_MAX_DIM 33
typedef struct { int left, right, child1, child2, divfeat; float divlow, divhigh; } kd_node;
typedef struct
{
    const float *pts;
    long n;
    int *vind;
    kd_node *nodes;
    int n_nodes, leaf_max, dim; /* dim = 3 (points) or 33 (FPFH features, KDTree<33>) */
    float root_lo[KD_MAX_DIM], root_hi[KD_MAX_DIM];
} kd_tree;

static void kd_minmax(const kd_tree *t, const int *ind, int count, int e, float *mn, float *mx)
{
    *mn = *mx = t->pts[t->dim * ind[0] + e];
    for (int i = 1; i < count; ++i)
    {
        const float v = t->pts[t->dim * ind[i] + e];
        if (v < *mn) *mn = v;
        if (v > *mx) *mx = v;
    }
}
/* nanoflann.hpp:974-1010, unsigned IndexType semantics (the "right &&" guards) kept */
static void kd_plane_split(const kd_tree *t, int *ind, unsigned count, int cutfeat, float cutval, unsigned *lim1, unsigned *lim2)
{
    unsigned left = 0, right = count - 1;
    for (;;)
    {
        while (left <= right && t->pts[t->dim * ind[left] + cutfeat] < cutval) ++left;
        while (right && left <= right && t->pts[t->dim * ind[right] + cutfeat] >= cutval) --right;
        if (left > right || !right) break;
        int tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp;
        ++left; --right;
    }
    *lim1 = left;
    right = count - 1;
    for (;;)
    {
        while (left <= right && t->pts[t->dim * ind[left] + cutfeat] <= cutval) ++left;
        while (right && left <= right && t->pts[t->dim * ind[right] + cutfeat] > cutval) --right;
        if (left > right || !right) break;
        int tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp;
        ++left; --right;
    }
    *lim2 = left;
}
/* lo/hi: in = the loose box handed down by the parent, out = the tight box of the subtree (divideTree's bbox reference) */
static int kd_divide(kd_tree *t, int left, int right, float *lo, float *hi)
{
    const int id = t->n_nodes++;
    kd_node *node = &t->nodes[id];
    node->left = left; node->right = right;
    if (right - left <= t->leaf_max)
    {
        node->child1 = node->child2 = -1; node->divfeat = -1; node->divlow = node->divhigh = 0.0f;
        for (int i = 0; i < t->dim; ++i) lo[i] = hi[i] = t->pts[t->dim * t->vind[left] + i];
        for (int k = left + 1; k < right; ++k)
            for (int i = 0; i < t->dim; ++i)
            {
                const float v = t->pts[t->dim * t->vind[k] + i];
                if (lo[i] > v) lo[i] = v;
                if (hi[i] < v) hi[i] = v;
            }
        return id;
    }
    /* middleSplit_ :916-963 */
    int *ind = t->vind + left;
    const unsigned count = (unsigned)(right - left);
    const float eps = 0.00001f;
    float max_span = hi[0] - lo[0];
    for (int i = 1; i < t->dim; ++i) { const float span = hi[i] - lo[i]; if (span > max_span) max_span = span; }
    float max_spread = -1.0f;
    int cutfeat = 0;
    for (int i = 0; i < t->dim; ++i)
    {
        const float span = hi[i] - lo[i];
        if (span > (1 - eps) * max_span)
        {
            float mn, mx;
            kd_minmax(t, ind, (int)count, i, &mn, &mx);
            const float spread = mx - mn;
            if (spread > max_spread) { cutfeat = i; max_spread = spread; }
        }
    }
    const float split_val = (lo[cutfeat] + hi[cutfeat]) / 2;
    float mn, mx, cutval;
    kd_minmax(t, ind, (int)count, cutfeat, &mn, &mx);
    if (split_val < mn) cutval = mn; else if (split_val > mx) cutval = mx; else cutval = split_val;
    unsigned lim1, lim2, idx;
    kd_plane_split(t, ind, count, cutfeat, cutval, &lim1, &lim2);
    if (lim1 > count / 2) idx = lim1; else if (lim2 < count / 2) idx = lim2; else idx = count / 2;

    float llo[KD_MAX_DIM], lhi[KD_MAX_DIM], rlo[KD_MAX_DIM], rhi[KD_MAX_DIM];
    const size_t box_bytes = sizeof(float) * (size_t)t->dim;
    memcpy(llo, lo, box_bytes); memcpy(lhi, hi, box_bytes); memcpy(rlo, lo, box_bytes); memcpy(rhi, hi, box_bytes);
    lhi[cutfeat] = cutval;
    const int c1 = kd_divide(t, left, left + (int)idx, llo, lhi);
    rlo[cutfeat] = cutval;
    const int c2 = kd_divide(t, left + (int)idx, right, rlo, rhi);
    node = &t->nodes[id];
    node->child1 = c1; node->child2 = c2; node->divfeat = cutfeat;
    node->divlow = lhi[cutfeat]; node->divhigh = rlo[cutfeat];
    for (int i = 0; i < t->dim; ++i) { lo[i] = rlo[i] < llo[i] ? rlo[i] : llo[i]; hi[i] = lhi[i] < rhi[i] ? rhi[i] : lhi[i]; } /* std::min / std::max: the first argument wins against NaN */
    return id;
}
static kd_tree *kd_build_dim(const float *pts, long n, int leaf_max, int dim)
{
    kd_tree *t = (kd_tree *)calloc(1, sizeof *t);
    t->pts = pts; t->n = n; t->leaf_max = leaf_max; t->dim = dim;
    t->vind = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    t->nodes = (kd_node *)malloc(sizeof(kd_node) * (size_t)(2 * n + 1));
    for (long i = 0; i < n; ++i) t->vind[i] = (int)i;
    if (n == 0) return t;
    /* computeBoundingBox :1324-1350 */
    for (int i = 0; i < t->dim; ++i) t->root_lo[i] = t->root_hi[i] = pts[i];
    for (long k = 1; k < n; ++k)
        for (int i = 0; i < t->dim; ++i)
        {
            if (pts[t->dim * k + i] < t->root_lo[i]) t->root_lo[i] = pts[t->dim * k + i];
            if (pts[t->dim * k + i] > t->root_hi[i]) t->root_hi[i] = pts[t->dim * k + i];
        }
    kd_divide(t, 0, (int)n, t->root_lo, t->root_hi);
    return t;
}
static kd_tree *kd_build(const float *pts, long n, int leaf_max) { return kd_build_dim(pts, n, leaf_max, 3); }
static void kd_free(kd_tree *t) { free(t->vind); free(t->nodes); free(t); }

/* result sets: mode 0 = KNNResultSet(capacity), mode 1 = RadiusResultSet(radius, max_neighbors) */
typedef struct
{
    int mode, capacity, count;
    float radius;
    int *index;
    float *dist;
} kd_result;
static float kd_worst(const kd_result *r) { return r->mode == 0 ? r->dist[r->capacity - 1] : r->radius; }
static int kd_add(kd_result *r, float dist, int index)
{
    if (r->mode == 0)
    {
        int i;
        for (i = r->count; i > 0; --i)
        {
            if (r->dist[i - 1] > dist)
            {
                if (i < r->capacity) { r->dist[i] = r->dist[i - 1]; r->index[i] = r->index[i - 1]; }
            }
            else break;
        }
        if (i < r->capacity) { r->dist[i] = dist; r->index[i] = index; }
        if (r->count < r->capacity) r->count++;
        return 1;
    }
    if (r->capacity > 0 && r->count >= r->capacity) return 0; /* the reference's max_neighbors stop, :253-255 */
    if (dist < r->radius) { r->index[r->count] = index; r->dist[r->count] = dist; r->count++; }
    return 1;
}
static int kd_search_level(const kd_tree *t, kd_result *r, const float *q, int node_id, float mindistsq, float *dists, float eps_error)
{
    const kd_node *node = &t->nodes[node_id];
    if (node->child1 < 0)
    {
        const float worst = kd_worst(r);
        for (int i = node->left; i < node->right; ++i)
        {
            const int index = t->vind[i];
            float d = 0.0f;
            for (int k = 0; k < t->dim; ++k) { const float diff = q[k] - t->pts[t->dim * index + k]; d += diff * diff; }
            if (d < worst)
                if (!kd_add(r, d, index)) return 0;
        }
        return 1;
    }
    const int idx = node->divfeat;
    const float val = q[idx];
    const float diff1 = val - node->divlow, diff2 = val - node->divhigh;
    int best, other;
    float cut_dist;
    if ((diff1 + diff2) < 0) { best = node->child1; other = node->child2; cut_dist = (val - node->divhigh) * (val - node->divhigh); }
    else { best = node->child2; other = node->child1; cut_dist = (val - node->divlow) * (val - node->divlow); }
    if (!kd_search_level(t, r, q, best, mindistsq, dists, eps_error)) return 0;
    const float dst = dists[idx];
    mindistsq = mindistsq + cut_dist - dst;
    dists[idx] = cut_dist;
    if (mindistsq * eps_error <= kd_worst(r))
        if (!kd_search_level(t, r, q, other, mindistsq, dists, eps_error)) return 0;
    dists[idx] = dst;
    return 1;
}
static void kd_find(const kd_tree *t, kd_result *r, const float *q, float eps)
{
    if (t->n == 0) return;
    const float eps_error = 1 + eps;
    float dists[KD_MAX_DIM] = {0}, distsq = 0.0f;
    for (int i = 0; i < t->dim; ++i)
    {
        if (q[i] < t->root_lo[i]) { dists[i] = (q[i] - t->root_lo[i]) * (q[i] - t->root_lo[i]); distsq += dists[i]; }
        if (q[i] > t->root_hi[i]) { dists[i] = (q[i] - t->root_hi[i]) * (q[i] - t->root_hi[i]); distsq += dists[i]; }
    }
    kd_search_level(t, r, q, 0, distsq, dists, eps_error);
}

/* libstdc++ std::sort (bits/stl_algo.h: __introsort_loop, __move_median_to_first, __unguarded_partition,
 * __final_insertion_sort, threshold 16) on (index, distance) pairs compared by distance only -- IndexDist_Sorter,
 * nanoflann.hpp:206-214.  Not a stable sort: the order of equal distances is whatever this exact algorithm leaves.
 * When the depth limit 2*floor(log2 n) runs out the range is heap-sorted (__partial_sort(first, last, last):
 * __make_heap, __sort_heap over __adjust_heap / __push_heap, bits/stl_heap.h) -- reached about once per few thousand
 * 250-element radius searches. */
typedef struct { int index; float dist; } kd_pair;
#define KD_LESS(a, b) ((a).dist < (b).dist)
static void kd_swap(kd_pair *a, kd_pair *b) { kd_pair t = *a; *a = *b; *b = t; }
static void kd_adjust_heap(kd_pair *first, long hole, long len, kd_pair value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (KD_LESS(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    long parent = (hole - 1) / 2;
    while (hole > top && KD_LESS(first[parent], value))
    {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void kd_heap_sort(kd_pair *first, kd_pair *last)
{
    const long len = last - first;
    if (len >= 2)
        for (long parent = (len - 2) / 2;; --parent)
        {
            kd_adjust_heap(first, parent, len, first[parent]);
            if (parent == 0) break;
        }
    while (last - first > 1)
    {
        --last;
        kd_pair value = *last;
        *last = *first;
        kd_adjust_heap(first, 0, last - first, value);
    }
}
static int kd_introsort_loop(kd_pair *first, kd_pair *last, int depth_limit)
{
    while (last - first > 16)
    {
        if (depth_limit == 0) { kd_heap_sort(first, last); return 0; }
        --depth_limit;
        kd_pair *mid = first + (last - first) / 2, *a = first + 1, *b = mid, *c = last - 1;
        if (KD_LESS(*a, *b))
        {
            if (KD_LESS(*b, *c)) kd_swap(first, b);
            else if (KD_LESS(*a, *c)) kd_swap(first, c);
            else kd_swap(first, a);
        }
        else if (KD_LESS(*a, *c)) kd_swap(first, a);
        else if (KD_LESS(*b, *c)) kd_swap(first, c);
        else kd_swap(first, b);
        kd_pair *lo = first + 1, *hi = last;
        for (;;)
        {
            while (KD_LESS(*lo, *first)) ++lo;
            --hi;
            while (KD_LESS(*first, *hi)) --hi;
            if (!(lo < hi)) break;
            kd_swap(lo, hi);
            ++lo;
        }
        if (kd_introsort_loop(lo, last, depth_limit)) return -1;
        last = lo;
    }
    return 0;
}
static void kd_unguarded_linear_insert(kd_pair *last)
{
    kd_pair val = *last, *next = last - 1;
    while (KD_LESS(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void kd_insertion_sort(kd_pair *first, kd_pair *last)
{
    if (first == last) return;
    for (kd_pair *i = first + 1; i != last; ++i)
    {
        if (KD_LESS(*i, *first))
        {
            kd_pair val = *i;
            memmove(first + 1, first, sizeof(kd_pair) * (size_t)(i - first));
            *first = val;
        }
        else kd_unguarded_linear_insert(i);
    }
}
static int kd_std_sort(kd_pair *first, long n)
{
    if (n == 0) return 0;
    int lg = 0;
    for (long m = n; m > 1; m >>= 1) ++lg;
    if (kd_introsort_loop(first, first + n, 2 * lg)) return -1;
    if (n > 16)
    {
        kd_insertion_sort(first, first + 16);
        for (kd_pair *i = first + 16; i != first + n; ++i) kd_unguarded_linear_insert(i);
    }
    else kd_insertion_sort(first, first + n);
    return 0;
}

/* KDTree<3>::KnnSearch (mode 0, KDTree.h:176-195), RadiusSearch (mode 1, :125-143: the L2 "radius" is compared with
 * SQUARED distances, at most (size_t)(max_result * 2.5) hits are collected in traversal order, sorted, cut to
 * max_result) and KnnRadiusSearch (mode 2, :230-256).  idx/dist need room for max(k, (int)(k * 2.5)) entries. */
static int kd_query(const kd_tree *t, const float *q, int mode, int k, float radius, int *idx, float *dist)
{
    kd_result r;
    r.index = idx; r.dist = dist; r.count = 0; r.radius = radius;
    if (mode == 1)
    {
        r.mode = 1; r.capacity = (int)(size_t)(k * 2.5);
        kd_find(t, &r, q, 1e-8f);
        kd_pair *tmp = (kd_pair *)malloc(sizeof(kd_pair) * (size_t)(r.count + 1));
        for (int i = 0; i < r.count; ++i) { tmp[i].index = idx[i]; tmp[i].dist = dist[i]; }
        if (kd_std_sort(tmp, r.count)) { free(tmp); return -1; }
        int cnt = r.count > k ? k : r.count;
        for (int i = 0; i < cnt; ++i) { idx[i] = tmp[i].index; dist[i] = tmp[i].dist; }
        free(tmp);
        return cnt;
    }
    r.mode = 0; r.capacity = k;
    if (k > 0) dist[k - 1] = FLT_MAX;
    kd_find(t, &r, q, 0.0f);
    if (mode == 0) return r.count;
    int in_radius = 0;
    for (; in_radius != r.count; ++in_radius)
        if (dist[in_radius] > radius) break;
    return in_radius;
}
void orc_kdtree_search(const float *pts, long n, const float *queries, long nq, int mode, int k, float radius, long cap,
                       int32_t *out_index, float *out_dist, int32_t *out_count)
{
    kd_tree *t = kd_build(pts, n, 10);
    const int room = (k > (int)(k * 2.5) ? k : (int)(k * 2.5)) + 1;
#pragma omp parallel
    {
        int *idx = (int *)malloc(sizeof(int) * (size_t)room);
        float *dist = (float *)malloc(sizeof(float) * (size_t)room);
#pragma omp for schedule(dynamic, 64)
        for (long q = 0; q < nq; ++q)
        {
            const int cnt = kd_query(t, queries + 3 * q, mode, k, radius, idx, dist);
            out_count[q] = cnt;
            for (long j = 0; j < cap; ++j)
            {
                out_index[q * cap + j] = j < cnt ? idx[j] : -1;
                out_dist[q * cap + j] = j < cnt ? dist[j] : -1.0f;
            }
        }
        free(idx); free(dist);
    }
    kd_free(t);
}
/* the built tree, for checking a device build: vind (n ints) and the nodes in pre-order, 5 ints + 2 floats each
 * (left, right, child1, child2, divfeat; divlow, divhigh) with children renumbered in pre-order; returns node count */
long orc_kdtree_dump_dim(const float *pts, long n, int dim, int32_t *vind, int32_t *node_ints, float *node_floats, float *root_box);
long orc_kdtree_dump(const float *pts, long n, int32_t *vind, int32_t *node_ints, float *node_floats, float *root_box)
{
    return orc_kdtree_dump_dim(pts, n, 3, vind, node_ints, node_floats, root_box);
}
/* root_box: dim lows then dim highs */
long orc_kdtree_dump_dim(const float *pts, long n, int dim, int32_t *vind, int32_t *node_ints, float *node_floats, float *root_box)
{
    kd_tree *t = kd_build_dim(pts, n, 10, dim);
    for (long i = 0; i < n; ++i) vind[i] = t->vind[i];
    for (int i = 0; i < t->n_nodes; ++i)
    {
        const kd_node *nd = &t->nodes[i]; /* kd_divide numbers nodes in pre-order already */
        node_ints[5 * i] = nd->left; node_ints[5 * i + 1] = nd->right; node_ints[5 * i + 2] = nd->child1;
        node_ints[5 * i + 3] = nd->child2; node_ints[5 * i + 4] = nd->divfeat;
        node_floats[2 * i] = nd->divlow; node_floats[2 * i + 1] = nd->divhigh;
    }
    for (int i = 0; i < dim; ++i) { root_box[i] = t->root_lo[i]; root_box[dim + i] = t->root_hi[i]; }
    const long nn = t->n_nodes;
    kd_free(t);
    return nn;
}

/* ------------------------------------------------------------------------------------------------------------------
 * registration::ComputeFPFHFeature (src/Registration/3DFeature.cpp:7-131): pair descriptor, SPFH histograms over the
 * radius-search neighbours (first hit skipped), distance-weighted sum.  Eigen's fixed-size dot/norm reduce as
 * x + (y + z).  The angle goes through the double-precision ::atan2 (no <cmath> overload is visible unqualified).
 * ------------------------------------------------------------------------------------------------------------------ */
#ifndef M_PI
#define M_PI 3.14159265358979323846 /* math.h value; hidden by -std=c11 */
#endif
static void cross3(const float *a, const float *b, float *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static int fpfh_pair_bins(const float *ps, const float *ns, const float *pt, const float *nt, int *bins)
{
    float d[3] = {ps[0] - pt[0], ps[1] - pt[1], ps[2] - pt[2]};
    const float distance = sqrtf(dot3(d, d));
    float dir[3] = {(pt[0] - ps[0]) / distance, (pt[1] - ps[1]) / distance, (pt[2] - ps[2]) / distance};
    float v[3], w[3], desc[3];
    cross3(ns, dir, v);
    if (sqrtf(dot3(v, v)) == 0) { desc[0] = desc[1] = desc[2] = 0.0f; }
    else
    {
        cross3(ns, v, w);
        float diff[3] = {pt[0] - ps[0], pt[1] - ps[1], pt[2] - ps[2]};
        desc[1] = dot3(v, nt);
        desc[2] = dot3(ns, diff) / distance;
        desc[0] = (float)atan2(dot3(w, nt), dot3(ns, nt));
    }
    bins[0] = (int)floor(11 * (desc[0] + M_PI) / (2.0 * M_PI));
    bins[1] = (int)floor(11 * (desc[1] + 1) / 2.0);
    bins[2] = (int)floor(11 * (desc[2] + 1) / 2.0);
    for (int k = 0; k < 3; ++k) { if (bins[k] > 10) bins[k] = 10; if (bins[k] < 0) bins[k] = 0; }
    return 0;
}
int orc_fpfh(const float *pts, const float *normals, long n, int knn, float radius, float *features)
{
    kd_tree *t = kd_build(pts, n, 10);
    const int room = (knn > (int)(knn * 2.5) ? knn : (int)(knn * 2.5)) + 1;
    float *spfh = (float *)calloc((size_t)n * 33 + 1, sizeof(float));
    int *nbr = (int *)malloc(sizeof(int) * (size_t)n * (size_t)(knn > 0 ? knn : 1));
    int *nbr_count = (int *)calloc((size_t)n + 1, sizeof(int));
    int failed = 0;
#pragma omp parallel
    {
        int *idx = (int *)malloc(sizeof(int) * (size_t)room);
        float *dist = (float *)malloc(sizeof(float) * (size_t)room);
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < n; ++i)
        {
            int points_num = kd_query(t, pts + 3 * i, 1, knn, radius, idx, dist);
            if (points_num < 0) { failed = 1; continue; }
            float *h = spfh + 33 * i;
            if (points_num - 1 > 0)
            {
                if (points_num > knn) points_num = knn;
                const double each = 100 / (points_num - 1);
                nbr_count[i] = points_num - 1;
                for (int j = 1; j != points_num; ++j)
                {
                    nbr[i * knn + j - 1] = idx[j];
                    int bins[3];
                    fpfh_pair_bins(pts + 3 * i, normals + 3 * i, pts + 3 * idx[j], normals + 3 * idx[j], bins);
                    h[bins[0]] = (float)(h[bins[0]] + each);
                    h[bins[1] + 11] = (float)(h[bins[1] + 11] + each);
                    h[bins[2] + 22] = (float)(h[bins[2] + 22] + each);
                }
            }
        }
        free(idx); free(dist);
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < n; ++i)
    {
        double sum[3] = {0, 0, 0};
        float *f = features + 33 * i;
        for (int e = 0; e < 33; ++e) f[e] = 0.0f;
        for (int j = 0; j < nbr_count[i]; ++j)
        {
            const int nb = nbr[i * knn + j];
            float d[3] = {pts[3 * i] - pts[3 * nb], pts[3 * i + 1] - pts[3 * nb + 1], pts[3 * i + 2] - pts[3 * nb + 2]};
            const float dist = sqrtf(dot3(d, d));
            if (dist != 0.0)
            {
                const float w_d = 1 / dist;
                const float *s = spfh + 33 * nb;
                for (int e = 0; e < 33; ++e) f[e] = f[e] + w_d * s[e];
                for (int b = 0; b < 3; ++b)
                {
                    float bs = 0.0f;
                    for (int e = 0; e < 11; ++e) bs += s[11 * b + e]; /* integer-valued: exact in any order */
                    sum[b] += bs;
                }
            }
        }
        for (int b = 0; b < 3; ++b)
        {
            const float scale = (float)(100.0 / sum[b]);
            for (int e = 0; e < 11; ++e) f[11 * b + e] = f[11 * b + e] * scale;
        }
        for (int e = 0; e < 33; ++e) f[e] = f[e] + spfh[33 * i + e];
    }
    free(spfh); free(nbr); free(nbr_count);
    kd_free(t);
    return failed ? -1 : 0;
}

/* PointCloud::EstimateNormals (PointCloud.cpp:102-144): KnnRadiusSearch(knn, radius) in the kd-tree's own visiting order
 * (ties included), then FitPlane */
void orc_estimate_normals(const float *pts, long n, float radius, int knn, float *normals)
{
    kd_tree *t = kd_build(pts, n, 10);
#pragma omp parallel
    {
        float *bd = (float *)malloc(sizeof(float) * (size_t)(knn + 1));
        int *bi = (int *)malloc(sizeof(int) * (size_t)(knn + 1));
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < n; ++i)
        {
            const int in_radius = kd_query(t, pts + 3 * i, 2, knn, radius, bi, bd);
            fit_plane_normal(pts, bi, in_radius, normals + 3 * i);
        }
        free(bd); free(bi);
    }
    kd_free(t);
}

/* ------------------------------------------------------------------------------------------------------------------
 * registration::FeatureMatching3D (src/Registration/GlobalRegistration.cpp:29-73): KDTree<33> over the target features,
 * KnnSearch(k = 1) per source feature; a source whose search returns nothing (NaN feature: no distance compares below
 * the initial worst) is left out.  Same tree code as above with dim = 33, so ties and NaN rows behave as in nanoflann.
 * ------------------------------------------------------------------------------------------------------------------ */
long orc_feature_matching(const float *src_feat, long ns, const float *tgt_feat, long nt, int32_t *pairs)
{
    kd_tree *t = kd_build_dim(tgt_feat, nt, 10, 33);
    int32_t *nn = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns + 1));
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < ns; ++i)
    {
        int idx[2];
        float dist[2];
        kd_result r;
        r.mode = 0; r.capacity = 1; r.count = 0; r.radius = 0.0f; r.index = idx; r.dist = dist;
        dist[0] = FLT_MAX;
        kd_find(t, &r, src_feat + 33 * i, 0.0f);
        nn[i] = r.count ? idx[0] : -1;
    }
    long m = 0;
    for (long i = 0; i < ns; ++i)
        if (nn[i] >= 0) { pairs[2 * m] = (int32_t)i; pairs[2 * m + 1] = nn[i]; ++m; }
    free(nn);
    kd_free(t);
    return m;
}

/* registration::RejectMatchesRanSaPC (GlobalRegistration.cpp:75-108), `rounds` calls sharing one default-constructed
 * std::default_random_engine as in RansacRegistration (:168-172).  libstdc++: default_random_engine = minstd_rand0
 * (x <- 16807 x mod 2^31-1, seed 1, range [1, 2^31-2]); uniform_int_distribution<int>(0, N-1) on it takes the
 * "downscaling" branch of bits/uniform_int_dist.h: scaling = urange_of_engine / N, reject draws >= N * scaling, divide. */
static uint32_t minstd_next(uint32_t *state)
{
    *state = (uint32_t)(((uint64_t)*state * 16807u) % 2147483647u);
    return *state;
}
static int uniform_int_libstdcxx(uint32_t *state, int n_values)
{
    const uint32_t urngmin = 1u, urngrange = 2147483646u - 1u;
    const uint32_t uerange = (uint32_t)n_values; /* urange + 1 */
    if (urngrange > uerange - 1u)
    {
        const uint32_t scaling = urngrange / uerange, past = uerange * scaling;
        uint32_t ret;
        do ret = minstd_next(state) - urngmin; while (ret >= past);
        return (int)(ret / scaling);
    }
    return -1; /* more matches than the engine's range: not reachable for int counts below 2^31 - 2 */
}
long orc_reject_matches(const float *src, const float *tgt, int32_t *pairs, long n, int rounds, int candidate_num, float difference)
{
    uint32_t state = 1u;
    int32_t *kept = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n + 1));
    for (int round = 0; round < rounds; ++round)
    {
        /* uniform_int_distribution<int> uniform(0, N - 1) with N = 0 would be (0, -1): the loop below never draws then */
        long m = 0;
        for (long i = 0; i < n; ++i)
        {
            const float *rp = src + 3 * pairs[2 * i], *np = tgt + 3 * pairs[2 * i + 1];
            int keep = 0;
            for (int j = 0; j < candidate_num; ++j)
            {
                const int c = uniform_int_libstdcxx(&state, (int)n);
                const float *rq = src + 3 * pairs[2 * c], *nq = tgt + 3 * pairs[2 * c + 1];
                float a[3] = {rq[0] - rp[0], rq[1] - rp[1], rq[2] - rp[2]}, b[3] = {nq[0] - np[0], nq[1] - np[1], nq[2] - np[2]};
                const float d1 = sqrtf(dot3(a, a)), d2 = sqrtf(dot3(b, b));
                if (fabs(d1 - d2) <= difference * d1) { keep = 1; break; } /* fabs: double of the float difference */
            }
            if (keep) { kept[2 * m] = pairs[2 * i]; kept[2 * m + 1] = pairs[2 * i + 1]; ++m; }
        }
        memcpy(pairs, kept, sizeof(int32_t) * 2 * (size_t)m);
        n = m;
    }
    free(kept);
    return n;
}

/* ------------------------------------------------------------------------------------------------------------------
 * geometry::EstimateRigidTransformation (Geometry.cpp:107-151) in the float build, operation for operation: float means,
 * W += (a - ma)(b - mb)^T, JacobiSVD<MatrixXf>(W, ThinU | ThinV), R = V U^T through Eigen's coefficient-based product of
 * dynamic 3x3 matrices (sequential sum), determinant by cofactors (bruteforce_det3_helper), V's last column flipped if
 * det < 0, t = mb - R ma (fixed-size product: x + (y + z)).  T row-major 4x4.  Used where the result decides something
 * discrete: the RANSAC hypotheses of geometry::EstimateRigidTransformationRANSAC (Ransac.cpp:7-41).
 * ------------------------------------------------------------------------------------------------------------------ */
static void kabsch_f32(const float *a, const float *b, long n, float *T)
{
    float ma[3] = {0, 0, 0}, mb[3] = {0, 0, 0}, W[9] = {0};
    for (long i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) { ma[c] += a[3 * i + c]; mb[c] += b[3 * i + c]; }
    for (int c = 0; c < 3; ++c) { ma[c] /= (float)n; mb[c] /= (float)n; }
    for (long i = 0; i < n; ++i)
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) W[p * 3 + q] += (a[3 * i + p] - ma[p]) * (b[3 * i + q] - mb[q]);
    float U[9], V[9], sv[3], R[9];
    jacobi_svd3_uv(W, U, V, sv);
    for (int pass = 0; pass < 2; ++pass)
    {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i * 3 + j] = (V[i * 3] * U[j * 3] + V[i * 3 + 1] * U[j * 3 + 1]) + V[i * 3 + 2] * U[j * 3 + 2];
        const float det = (R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6])) + R[2] * (R[3] * R[7] - R[4] * R[6]);
        if (!(det < 0) || pass) break;
        V[2] = -V[2]; V[5] = -V[5]; V[8] = -V[8];
    }
    for (int i = 0; i < 16; ++i) T[i] = 0.0f;
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = mb[i] - (R[i * 3] * ma[0] + (R[i * 3 + 1] * ma[1] + R[i * 3 + 2] * ma[2]));
    }
    T[15] = 1.0f;
}
void orc_kabsch_f32(const float *a, const float *b, long n, float *T_rowmajor) { kabsch_f32(a, b, n, T_rowmajor); }
/* one RANSAC hypothesis: TransformationModel over the eight sampled pairs (TransformationModel.hpp:54-75), Evaluate over all
 * pairs (:77-94): inlier if |R a + t - b| < threshold (float norm against the double threshold) -> inlier count */
long orc_ransac_hypothesis(const float *a, const float *b, long n, const int32_t *sample8, double threshold, float *T_rowmajor, uint8_t *inlier)
{
    float sa[24], sb[24], T[16];
    for (int k = 0; k < 8; ++k)
        for (int c = 0; c < 3; ++c) { sa[3 * k + c] = a[3 * (long)sample8[k] + c]; sb[3 * k + c] = b[3 * (long)sample8[k] + c]; }
    kabsch_f32(sa, sb, 8, T);
    long cnt = 0;
    for (long i = 0; i < n; ++i)
    {
        float e[3];
        for (int r = 0; r < 3; ++r)
            e[r] = ((T[r * 4] * a[3 * i] + (T[r * 4 + 1] * a[3 * i + 1] + T[r * 4 + 2] * a[3 * i + 2])) + T[r * 4 + 3]) - b[3 * i + r];
        const float err = sqrtf(e[0] * e[0] + (e[1] * e[1] + e[2] * e[2]));
        const int in = (double)err < threshold;
        if (inlier) inlier[i] = (uint8_t)in;
        cnt += in;
    }
    if (T_rowmajor) memcpy(T_rowmajor, T, sizeof T);
    return cnt;
}

/* geometry::EstimateRigidTransformationRANSAC with the samples given (iterations x 8 pair indices): the first iteration with
 * the strictly largest inlier fraction wins (GRANSAC.hpp:113-122); returns its index (-1 if none has an inlier), its motion
 * (row-major 4x4) and its inlier flags */
long orc_ransac_select(const float *a, const float *b, long n, const int32_t *samples, long iterations, double threshold, float *T_rowmajor,
                       uint8_t *inlier, long *best_count)
{
    long best = -1, best_cnt = 0;
    long *cnt = (long *)malloc(sizeof(long) * (size_t)(iterations + 1));
#pragma omp parallel for schedule(dynamic, 16)
    for (long h = 0; h < iterations; ++h) cnt[h] = orc_ransac_hypothesis(a, b, n, samples + 8 * h, threshold, NULL, NULL);
    for (long h = 0; h < iterations; ++h)
        if ((double)cnt[h] / (double)n > (double)best_cnt / (double)n) { best_cnt = cnt[h]; best = h; }
    free(cnt);
    if (best >= 0) orc_ransac_hypothesis(a, b, n, samples + 8 * best, threshold, T_rowmajor, inlier);
    if (best_count) *best_count = best_cnt;
    return best;
}

/* ------------------------------------------------------------------------------------------------------------------
 * optimization::SimpleBA = Optimizer::FastBA (src/Optimization/SimpleBA.cpp:18-157): pose-graph refinement over frame pairs
 * linked by 3-D point pairs.  Per pair and point: r = (R1 p1 + t1) - (R2 p2 + t2), J_s = [I | -skew(R1 p1 + t1)],
 * J_t = [-I | skew(R2 p2 + t2)]; the 6x6 blocks J^T J and -J^T r are summed per frame pair, assembled into the sparse
 * normal equations over poses 1 .. n-1 (pose 0 is fixed), solved (Eigen SimplicialLDLT there, a dense LDL^T here) and every
 * pose is updated as Se3ToSE3(delta_i) * pose_i.  Sums are kept in float like the reference's; the solve is in double, so
 * the restatement is pinned by tolerance (blocks 1e-5 relative, poses 1e-5), not bit for bit.
 * out of orc_ba_blocks: JTJ_ss, JTJ_tt, JTJ_st, JTJ_ts (row-major 6x6), JTr_s, JTr_t: 156 floats.  Poses row-major here.
 * ------------------------------------------------------------------------------------------------------------------ */
static void ba_blocks_rm(const float *Ps, const float *Pt, const float *a, const float *b, long n, float *out)
{
    for (int i = 0; i < 156; ++i) out[i] = 0.0f;
    float *ss = out, *tt = out + 36, *st = out + 72, *ts = out + 108, *rs = out + 144, *rt = out + 150;
    for (long k = 0; k < n; ++k)
    {
        float q1[3], q2[3], r[3];
        for (int i = 0; i < 3; ++i)
        {
            q1[i] = (Ps[4 * i] * a[3 * k] + (Ps[4 * i + 1] * a[3 * k + 1] + Ps[4 * i + 2] * a[3 * k + 2])) + Ps[4 * i + 3];
            q2[i] = (Pt[4 * i] * b[3 * k] + (Pt[4 * i + 1] * b[3 * k + 1] + Pt[4 * i + 2] * b[3 * k + 2])) + Pt[4 * i + 3];
        }
        for (int i = 0; i < 3; ++i) r[i] = q1[i] - q2[i];
        float Js[18] = {1, 0, 0, 0, q1[2], -q1[1], 0, 1, 0, -q1[2], 0, q1[0], 0, 0, 1, q1[1], -q1[0], 0};   /* [I | -skew(q1)] */
        float Jt[18] = {-1, 0, 0, 0, -q2[2], q2[1], 0, -1, 0, q2[2], 0, -q2[0], 0, 0, -1, -q2[1], q2[0], 0}; /* [-I | skew(q2)] */
        for (int i = 0; i < 6; ++i)
        {
            for (int j = 0; j < 6; ++j)
            {
                float s1 = 0, s2 = 0, s3 = 0, s4 = 0;
                for (int m = 0; m < 3; ++m)
                {
                    s1 += Js[6 * m + i] * Js[6 * m + j]; s2 += Jt[6 * m + i] * Jt[6 * m + j];
                    s3 += Js[6 * m + i] * Jt[6 * m + j]; s4 += Jt[6 * m + i] * Js[6 * m + j];
                }
                ss[6 * i + j] += s1; tt[6 * i + j] += s2; st[6 * i + j] += s3; ts[6 * i + j] += s4;
            }
            float g1 = 0, g2 = 0;
            for (int m = 0; m < 3; ++m) { g1 += Js[6 * m + i] * r[m]; g2 += Jt[6 * m + i] * r[m]; }
            rs[i] -= g1; rt[i] -= g2;
        }
    }
}
void orc_ba_blocks(const float *pose_s_cm, const float *pose_t_cm, const float *a, const float *b, long n, float *out)
{
    float Ps[16], Pt[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { Ps[4 * r + c] = pose_s_cm[4 * c + r]; Pt[4 * r + c] = pose_t_cm[4 * c + r]; }
    ba_blocks_rm(Ps, Pt, a, b, n, out);
}
/* dense LDL^T solve of a symmetric positive definite system, in place (A n x n row-major, b -> x); 0 on success */
static int ldlt_solve(double *A, double *b, int n)
{
    for (int j = 0; j < n; ++j)
    {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k] * A[k * n + k];
        if (!(fabs(d) > 0)) return -1;
        A[j * n + j] = d;
        for (int i = j + 1; i < n; ++i)
        {
            double v = A[i * n + j];
            for (int k = 0; k < j; ++k) v -= A[i * n + k] * A[j * n + k] * A[k * n + k];
            A[i * n + j] = v / d;
        }
    }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < i; ++k) b[i] -= A[i * n + k] * b[k];
    for (int i = 0; i < n; ++i) b[i] /= A[i * n + i];
    for (int i = n - 1; i >= 0; --i)
        for (int k = i + 1; k < n; ++k) b[i] -= A[k * n + i] * b[k];
    return 0;
}
int orc_simple_ba(int n_poses, float *poses_cm, int n_corr, const int32_t *src_id, const int32_t *tgt_id, const int64_t *offset,
                  const float *a, const float *b, int max_iteration)
{
    if (n_poses < 3) return 0;              /* "Too few optimization variables" (:84-88) */
    if (n_corr < n_poses - 1) return -1;    /* "There are unconnected components" (:89-93) */
    const int nv = 6 * (n_poses - 1);
    double *A = (double *)malloc(sizeof(double) * (size_t)nv * nv), *g = (double *)malloc(sizeof(double) * (size_t)nv);
    float *P = (float *)malloc(sizeof(float) * 16 * (size_t)n_poses); /* row-major */
    for (int i = 0; i < n_poses; ++i)
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) P[16 * i + 4 * r + c] = poses_cm[16 * i + 4 * c + r];
    int rc = 0;
    for (int iter = 0; iter < max_iteration && !rc; ++iter)
    {
        float *gf = (float *)calloc((size_t)nv, sizeof(float)); /* JTr is a float vector in the reference */
        memset(A, 0, sizeof(double) * (size_t)nv * nv);
        for (int k = 0; k < n_corr; ++k)
        {
            float blk[156];
            const int s = src_id[k], t = tgt_id[k];
            ba_blocks_rm(P + 16 * s, P + 16 * t, a + 3 * offset[k], b + 3 * offset[k], (long)(offset[k + 1] - offset[k]), blk);
            for (int i = 0; i < 6; ++i)
            {
                for (int j = 0; j < 6; ++j)
                {
                    if (s != 0)
                    {
                        A[(size_t)((s - 1) * 6 + i) * nv + (s - 1) * 6 + j] += blk[6 * i + j];
                        A[(size_t)((s - 1) * 6 + i) * nv + (t - 1) * 6 + j] += blk[72 + 6 * i + j];
                        A[(size_t)((t - 1) * 6 + i) * nv + (s - 1) * 6 + j] += blk[108 + 6 * i + j];
                    }
                    A[(size_t)((t - 1) * 6 + i) * nv + (t - 1) * 6 + j] += blk[36 + 6 * i + j];
                }
                if (s != 0) gf[(s - 1) * 6 + i] += blk[144 + i];
                gf[(t - 1) * 6 + i] += blk[150 + i];
            }
        }
        for (int i = 0; i < nv; ++i) g[i] = gf[i];
        free(gf);
        if (ldlt_solve(A, g, nv)) { rc = -2; break; }
        for (int i = 1; i < n_poses; ++i)
        {
            double x[6], D[16];
            for (int e = 0; e < 6; ++e) x[e] = (float)g[(i - 1) * 6 + e];
            se3_exp_rm(x, D);
            float N[16];
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c)
                {
                    double v = 0;
                    for (int m = 0; m < 4; ++m) v += D[4 * r + m] * P[16 * i + 4 * m + c];
                    N[4 * r + c] = (float)v;
                }
            memcpy(P + 16 * i, N, sizeof N);
        }
    }
    for (int i = 0; i < n_poses; ++i)
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) poses_cm[16 * i + 4 * c + r] = P[16 * i + 4 * r + c];
    free(A); free(g); free(P);
    return rc;
}
