// geometry::PointCloud on the device: the clouds registration::PointToPlane / PointToPoint work on, built where they are used.
//
// Reference path rebuilt here (file:line relative to the reference tree):
//   PointCloud::LoadFromDepth   src/Geometry/PointCloud.cpp:72-100   z = depth (f32 m, or u16 / depth_scale); for z > 0:
//                                                                    x = (j - cx) * z / fx, y = (i - cy) * z / fy; raster order
//   PointCloud::LoadFromRGBD    src/Geometry/PointCloud.cpp:16-47    the same + colours / 255 (colours are not needed by ICP and
//                                                                    stay in the image the frame keeps for IntegrateImage)
// The reference's callers back-project every frame on the host and hand the 3.7 MB cloud to ICP, which copies it again
// (ICP.cpp:150-151); a frame is the SOURCE of one registration and the TARGET of the next.  Here the depth image is uploaded
// once (0.6 MB as u16), the cloud is compacted in raster order on the device and stays there for both registrations and for the
// integration of the same frame.  Three small kernels: valid pixels per tile of 1024, exclusive scan of the tile counts (one
// CTA), ordered write.  Arithmetic is the reference's, operation for operation, so the points are bit-identical to the host's.
#include <cstring>

#include "opb_cloud_host.h"
#include "opb_common.cuh"

namespace opb
{
constexpr int kCloudTile = 1024;

__device__ __forceinline__ float cloud_depth_at(const void *depth, int depth_u16, float depth_scale, int idx)
{
    if (depth_u16) return fdiv((float)__ldg((const unsigned short *)depth + idx), depth_scale);
    return __ldg((const float *)depth + idx);
}
__global__ void __launch_bounds__(kCloudTile) cloud_count_kernel(const void *depth, int depth_u16, float depth_scale, int n, unsigned int *tiles)
{
    const int idx = blockIdx.x * kCloudTile + threadIdx.x;
    const int c = __syncthreads_count(idx < n && cloud_depth_at(depth, depth_u16, depth_scale, idx) > 0.0f);
    if (threadIdx.x == 0) tiles[blockIdx.x] = (unsigned int)c;
}
// exclusive scan of the tile counts in place; total -> *count (mapped host memory)
__global__ void __launch_bounds__(1024) cloud_scan_kernel(unsigned int *tiles, int n_tiles, unsigned int *count)
{
    __shared__ unsigned int warp_sums[32];
    __shared__ unsigned int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024)
    {
        const int i = base + threadIdx.x;
        const unsigned int v = i < n_tiles ? tiles[i] : 0u;
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int m = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += m;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0)
        {
            unsigned int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned int m = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += m;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const unsigned int excl = carry + (warp ? warp_sums[warp - 1] : 0u) + inc - v;
        if (i < n_tiles) tiles[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}
__global__ void __launch_bounds__(kCloudTile) cloud_write_kernel(const void *depth, int depth_u16, float depth_scale, int width, int n, float fx,
                                                                 float fy, float cx, float cy, const unsigned int *tile_off, float *xyz)
{
    __shared__ unsigned int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * kCloudTile + threadIdx.x;
    float z = 0.0f;
    if (idx < n) z = cloud_depth_at(depth, depth_u16, depth_scale, idx);
    const bool keep = idx < n && z > 0.0f;
    const unsigned int ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_sums[warp] = (unsigned int)__popc(ballot);
    __syncthreads();
    if (warp == 0)
    {
        unsigned int w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int m = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += m;
        }
        warp_sums[lane] = w; // inclusive
    }
    __syncthreads();
    if (!keep) return;
    const size_t pos = (size_t)tile_off[blockIdx.x] + (warp ? warp_sums[warp - 1] : 0u) + __popc(ballot & ((1u << lane) - 1u));
    const int i = idx / width, j = idx - i * width;
    // float x = (j - cx) * z / fx (PointCloud.cpp:90-93): int -> float, subtract, multiply, divide, each rounded
    xyz[3 * pos] = fdiv(fmul(fsub((float)j, cx), z), fx);
    xyz[3 * pos + 1] = fdiv(fmul(fsub((float)i, cy), z), fy);
    xyz[3 * pos + 2] = z;
}

int cloud_wait(opb_cloud *c, size_t *n)
{
    OPB_CUDA(cudaSetDevice(c->device));
    OPB_CUDA(cudaEventSynchronize(c->ready));
    if (n) *n = c->count_on_device ? (size_t)*(volatile unsigned int *)c->h_count : c->n_host;
    return OPB_OK;
}

static int cloud_reserve_points(opb_cloud *c, size_t n)
{
    if (n <= c->cap_pts) return OPB_OK;
    OPB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_xyz);
    c->d_xyz = nullptr; c->cap_pts = 0;
    OPB_CUDA(cudaMalloc(&c->d_xyz, n * 3 * sizeof(float)));
    c->cap_pts = n;
    return OPB_OK;
}
static int cloud_reserve_normals(opb_cloud *c, size_t n)
{
    if (n <= c->cap_nrm) return OPB_OK;
    OPB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_nrm);
    c->d_nrm = nullptr; c->cap_nrm = 0;
    OPB_CUDA(cudaMalloc(&c->d_nrm, n * 3 * sizeof(float)));
    c->cap_nrm = n;
    return OPB_OK;
}
} // namespace opb

using namespace opb;

extern "C"
{
int opb_cloud_create(int device, void *stream, opb_cloud **out)
{
    if (!out) { set_error("out is NULL"); return OPB_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        set_error("no CUDA device: onepiece_b200 has no CPU path");
        return OPB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(device));
    opb_cloud *c = new opb_cloud();
    c->device = device;
    cudaError_t e = cudaSuccess;
    if (stream) c->stream = (cudaStream_t)stream;
    else { e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking); c->own_stream = e == cudaSuccess; }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc(&c->h_count, sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) { *c->h_count = 0; e = cudaHostGetDevicePointer(&c->d_count, c->h_count, 0); }
    if (e == cudaSuccess) e = cudaEventRecord(c->ready, c->stream);
    if (e != cudaSuccess)
    {
        set_error("cloud allocation failed: %s", cudaGetErrorString(e));
        opb_cloud_destroy(c);
        return OPB_ERR_CUDA;
    }
    *out = c;
    return OPB_OK;
}

void opb_cloud_destroy(opb_cloud *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_depth); cudaFree(c->d_bgr); cudaFree(c->d_xyz); cudaFree(c->d_nrm); cudaFree(c->d_tiles);
    if (c->h_count) cudaFreeHost(c->h_count);
    if (c->ready) cudaEventDestroy(c->ready);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

int opb_cloud_load_from_depth(opb_cloud *c, const void *depth, int depth_type, const uint8_t *bgr, float fx, float fy, float cx, float cy,
                              int width, int height, float depth_scale)
{
    if (!c || !depth) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (depth_type != OPB_DEPTH_F32 && depth_type != OPB_DEPTH_U16)
    {
        set_error("unknown depth type %d (expected OPB_DEPTH_F32=5 or OPB_DEPTH_U16=2)", depth_type);
        return OPB_ERR_INVALID;
    }
    if (width <= 0 || height <= 0 || width > 16384 || height > 16384) { set_error("bad image size %dx%d", width, height); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    const size_t npx = (size_t)width * height;
    const int n_tiles = (int)((npx + kCloudTile - 1) / kCloudTile);
    if (npx > c->cap_px)
    {
        OPB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_depth); cudaFree(c->d_bgr);
        c->d_depth = nullptr; c->d_bgr = nullptr; c->cap_px = 0;
        OPB_CUDA(cudaMalloc(&c->d_depth, npx * sizeof(float)));
        OPB_CUDA(cudaMalloc(&c->d_bgr, npx * 3));
        c->cap_px = npx;
    }
    if ((size_t)n_tiles + 1 > c->cap_tiles)
    {
        OPB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_tiles);
        c->d_tiles = nullptr; c->cap_tiles = 0;
        OPB_CUDA(cudaMalloc(&c->d_tiles, ((size_t)n_tiles + 1) * sizeof(unsigned int)));
        c->cap_tiles = (size_t)n_tiles + 1;
    }
    int rc = cloud_reserve_points(c, npx);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    OPB_CUDA(cudaMemcpyAsync(c->d_depth, depth, npx * (depth_type == OPB_DEPTH_U16 ? 2 : 4), cudaMemcpyDefault, s));
    if (bgr) OPB_CUDA(cudaMemcpyAsync(c->d_bgr, bgr, npx * 3, cudaMemcpyDefault, s));
    c->depth_type = depth_type; c->width = width; c->height = height;
    c->has_images = true; c->has_bgr = bgr != nullptr; c->has_normals = false;
    const int u16 = depth_type == OPB_DEPTH_U16;
    cloud_count_kernel<<<n_tiles, kCloudTile, 0, s>>>(c->d_depth, u16, depth_scale, (int)npx, c->d_tiles);
    cloud_scan_kernel<<<1, 1024, 0, s>>>(c->d_tiles, n_tiles, c->d_count);
    cloud_write_kernel<<<n_tiles, kCloudTile, 0, s>>>(c->d_depth, u16, depth_scale, width, (int)npx, fx, fy, cx, cy, c->d_tiles, c->d_xyz);
    OPB_CUDA(cudaGetLastError());
    c->count_on_device = true;
    OPB_CUDA(cudaEventRecord(c->ready, s));
    return OPB_OK;
}

int opb_cloud_set_points(opb_cloud *c, const float *xyz, size_t n)
{
    if (!c || (n && !xyz)) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    int rc = cloud_reserve_points(c, n ? n : 1);
    if (rc) return rc;
    if (n) OPB_CUDA(cudaMemcpyAsync(c->d_xyz, xyz, n * 3 * sizeof(float), cudaMemcpyDefault, c->stream));
    c->n_host = n; c->count_on_device = false; c->has_images = false; c->has_bgr = false; c->has_normals = false;
    OPB_CUDA(cudaEventRecord(c->ready, c->stream));
    return OPB_OK;
}

int opb_cloud_set_normals(opb_cloud *c, const float *normals, size_t n)
{
    if (!c || (n && !normals)) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    int rc = cloud_reserve_normals(c, n ? n : 1);
    if (rc) return rc;
    if (n) OPB_CUDA(cudaMemcpyAsync(c->d_nrm, normals, n * 3 * sizeof(float), cudaMemcpyDefault, c->stream));
    c->has_normals = n > 0;
    OPB_CUDA(cudaEventRecord(c->ready, c->stream));
    return OPB_OK;
}

int opb_cloud_size(opb_cloud *c, size_t *n)
{
    if (!c || !n) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    return cloud_wait(c, n);
}

int opb_cloud_download(opb_cloud *c, float *xyz, float *normals)
{
    if (!c) { set_error("cloud is NULL"); return OPB_ERR_INVALID; }
    size_t n = 0;
    int rc = cloud_wait(c, &n);
    if (rc) return rc;
    if (xyz && n) OPB_CUDA(cudaMemcpyAsync(xyz, c->d_xyz, n * 3 * sizeof(float), cudaMemcpyDefault, c->stream));
    if (normals && n)
    {
        if (!c->has_normals) { set_error("the cloud has no normals"); return OPB_ERR_INVALID; }
        OPB_CUDA(cudaMemcpyAsync(normals, c->d_nrm, n * 3 * sizeof(float), cudaMemcpyDefault, c->stream));
    }
    OPB_CUDA(cudaStreamSynchronize(c->stream));
    return OPB_OK;
}
} // extern "C"
