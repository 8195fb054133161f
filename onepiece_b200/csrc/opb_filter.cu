// Caller-side depth pre-filter on the device: the two calls every fusion main of the reference makes
// right before CubeHandler::IntegrateImage (example/ImageSequenceIntegration.cpp:36-38, example/DenseFusion/DenseFusion.cpp:92-95):
//   tool::ConvertDepthTo32F   src/Tool/ImageProcessing.cpp:68-91   u16 / depth_scale (clamped at 0), or float copy
//   tool::BilateralFilter     src/Tool/ImageProcessing.cpp:64-67   cv::bilateralFilter(src, dst, 7, 0.03, 4.5), CV_32FC1
// cv::bilateralFilter is OpenCV code (third party, not in the reference tree; README pins OpenCV 3.4): its published float
// algorithm is restated -- circular mask of radius d/2, spatial weights exp(-r^2 / 2 sigma_space^2), range weights from a
// 4096-bin table over [0, max - min] with linear interpolation, BORDER_REFLECT_101, weighted mean.  The arithmetic
// (neighbours in raster order then the centre, separately rounded float operations, IEEE division) is pinned to cv2 4.13 by
// tolerance in tests/ (OpenCV's own code paths differ from each other in the last bits); the device result is bit-identical
// to the CPU restatement the tests hold because the two transcendental tables are built on the host with libm and the kernel
// only multiplies and adds.
//
// Launch shape: one pass converts the frame and reduces min/max (the table depends on them: one 8-byte D2H + 16 KB H2D per
// frame whose range changed); the filter kernel works on 32x8 pixel tiles, stages tile + halo and the range table in shared
// memory, one thread per pixel, taps in constant kernel parameters.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "../../include/onepiece_b200.h"
#include "opb_common.cuh"

namespace opb
{
constexpr int kLutBins = 1 << 12;
constexpr int kLutSize = kLutBins + 2;
constexpr int kMaxRadius = 7;
constexpr int kMaxTaps = (2 * kMaxRadius + 1) * (2 * kMaxRadius + 1);
constexpr int kTileW = 32, kTileH = 8;

struct BilateralTaps
{
    int n;
    int radius;
    float scale_index;
    signed char dy[kMaxTaps], dx[kMaxTaps];
    float weight[kMaxTaps];
};

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// tool::ConvertDepthTo32F + min/max of the converted image (ordered-uint encoding; NaN never wins)
__global__ void __launch_bounds__(256) filter_convert_kernel(const void *raw, int is_u16, float depth_scale, int n, float *conv,
                                                             unsigned int *minmax)
{
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        float v;
        if (is_u16)
        {
            v = fdiv((float)((const unsigned short *)raw)[i], depth_scale);
            if (v < 0) v = 0;
        }
        else
            v = ((const float *)raw)[i];
        conv[i] = v;
        if (v < lo) lo = v;
        if (v > hi) hi = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicMin(&minmax[0], float_to_ordered(lo));
        atomicMax(&minmax[1], float_to_ordered(hi));
    }
}

__global__ void __launch_bounds__(kTileW *kTileH) bilateral_kernel(const float *__restrict__ src, int w, int h, const float *__restrict__ lut,
                                                                   const __grid_constant__ BilateralTaps taps, float *__restrict__ dst)
{
    extern __shared__ float smem[];
    float *s_lut = smem;                 // kLutSize
    float *s_tile = smem + kLutSize + 2; // (kTileH + 2r) x (kTileW + 2r)
    const int r = taps.radius, tw = kTileW + 2 * r, th = kTileH + 2 * r;
    const int t = threadIdx.y * kTileW + threadIdx.x;
    for (int i = t; i < kLutSize; i += kTileW * kTileH) s_lut[i] = lut[i];
    const int x0 = blockIdx.x * kTileW - r, y0 = blockIdx.y * kTileH - r;
    for (int i = t; i < tw * th; i += kTileW * kTileH)
    {
        const int ty = i / tw, tx = i - ty * tw;
        s_tile[i] = src[(size_t)reflect101(y0 + ty, h) * w + reflect101(x0 + tx, w)];
    }
    __syncthreads();
    const int x = blockIdx.x * kTileW + threadIdx.x, y = blockIdx.y * kTileH + threadIdx.y;
    if (x >= w || y >= h) return;
    const float *c = s_tile + (threadIdx.y + r) * tw + threadIdx.x + r;
    const float v0 = c[0];
    float sum = 0.0f, wsum = 0.0f;
    for (int k = 0; k < taps.n; ++k)
    {
        const float v = c[taps.dy[k] * tw + taps.dx[k]];
        float alpha = fmul(fabsf(fsub(v, v0)), taps.scale_index);
        // (int)alpha as the host computes it; a NaN or huge difference must not index outside the table
        int idx = alpha < (float)kLutBins ? __float2int_rz(alpha) : kLutBins;
        if (idx < 0) idx = 0;
        alpha = fsub(alpha, (float)idx);
        const float l0 = s_lut[idx];
        const float wk = fmul(taps.weight[k], fadd(l0, fmul(alpha, fsub(s_lut[idx + 1], l0))));
        sum = fadd(sum, fmul(v, wk));
        wsum = fadd(wsum, wk);
    }
    sum = fadd(sum, v0);
    wsum = fadd(wsum, 1.0f);
    dst[(size_t)y * w + x] = fdiv(sum, wsum);
}
} // namespace opb

using namespace opb;

struct opb_prefilter
{
    int device = 0, width = 0, height = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    void *d_raw = nullptr;
    float *d_conv = nullptr, *d_out = nullptr, *d_lut = nullptr;
    unsigned int *d_minmax = nullptr;
    unsigned int *h_minmax = nullptr; // pinned
    float *h_lut = nullptr;           // pinned
    // the table on the device was built for these arguments
    bool lut_valid = false;
    float lut_min = 0, lut_max = 0;
    double lut_sigma_color = 0;
    BilateralTaps taps;
    int taps_d = 0;
    double taps_sigma_space = 0;
};

// cv::bilateralFilter's tables (bilateralFilter_32f): see the file header
static void build_range_table(float vmin, float vmax, double sigma_color, float *lut, float *scale_index)
{
    const double gc = -0.5 / (sigma_color * sigma_color);
    const float len = (float)((double)vmax - (double)vmin);
    *scale_index = kLutBins / len;
    float last = 1.0f;
    for (int i = 0; i < kLutSize; ++i)
    {
        if (last > 0.0f)
        {
            const double val = i / *scale_index;
            lut[i] = (float)std::exp(val * val * gc);
            last = lut[i];
        }
        else
            lut[i] = 0.0f;
    }
}
static void build_taps(int radius, double sigma_space, BilateralTaps &t)
{
    const double gs = -0.5 / (sigma_space * sigma_space);
    t.n = 0;
    t.radius = radius;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j)
        {
            const double r = std::sqrt((double)i * i + (double)j * j);
            if (r > radius || (i == 0 && j == 0)) continue;
            t.weight[t.n] = (float)std::exp(r * r * gs);
            t.dy[t.n] = (signed char)i;
            t.dx[t.n] = (signed char)j;
            ++t.n;
        }
}

extern "C"
{
int opb_prefilter_create(int device, void *stream, int width, int height, opb_prefilter **out)
{
    if (!out) { set_error("out is NULL"); return OPB_ERR_INVALID; }
    *out = nullptr;
    if (width <= 0 || height <= 0 || width > 16384 || height > 16384) { set_error("bad image size %dx%d", width, height); return OPB_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        set_error("no CUDA device: onepiece_b200 has no CPU path");
        return OPB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(device));
    opb_prefilter *f = new opb_prefilter();
    f->device = device; f->width = width; f->height = height;
    cudaDeviceProp prop;
    OPB_CUDA(cudaGetDeviceProperties(&prop, device));
    f->sm_count = prop.multiProcessorCount;
    if (stream) f->stream = (cudaStream_t)stream;
    else { OPB_CUDA(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking)); f->own_stream = true; }
    const size_t n = (size_t)width * height;
    cudaError_t e = cudaMalloc(&f->d_raw, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_conv, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_out, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_lut, kLutSize * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_minmax, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaHostAlloc(&f->h_minmax, 2 * sizeof(unsigned int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(&f->h_lut, kLutSize * sizeof(float), cudaHostAllocDefault);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(bilateral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((kLutSize + 2 + (kTileW + 2 * kMaxRadius) * (kTileH + 2 * kMaxRadius)) * sizeof(float)));
    if (e != cudaSuccess)
    {
        set_error("pre-filter allocation failed: %s", cudaGetErrorString(e));
        opb_prefilter_destroy(f);
        return OPB_ERR_CUDA;
    }
    *out = f;
    return OPB_OK;
}

void opb_prefilter_destroy(opb_prefilter *f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    if (f->stream) cudaStreamSynchronize(f->stream);
    cudaFree(f->d_raw); cudaFree(f->d_conv); cudaFree(f->d_out); cudaFree(f->d_lut); cudaFree(f->d_minmax);
    if (f->h_minmax) cudaFreeHost(f->h_minmax);
    if (f->h_lut) cudaFreeHost(f->h_lut);
    if (f->own_stream && f->stream) cudaStreamDestroy(f->stream);
    cudaGetLastError();
    delete f;
}

int opb_prefilter_run(opb_prefilter *f, const void *depth, int depth_type, float depth_scale, int d, double sigma_color,
                      double sigma_space, float *converted, float *filtered)
{
    if (!f || !depth) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (depth_type != OPB_DEPTH_F32 && depth_type != OPB_DEPTH_U16)
    {
        // the reference prints "Unknown depth image type" and exits (ImageProcessing.cpp:86-90)
        set_error("[ImageProcessing]::[ERROR]::Unknown depth image type: %d", depth_type);
        return OPB_ERR_UNSUPPORTED;
    }
    // cv::bilateralFilter's argument normalisation
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    int radius = d <= 0 ? (int)std::lround(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    if (radius > kMaxRadius) { set_error("bilateral filter diameter %d exceeds the supported maximum %d", d, 2 * kMaxRadius + 1); return OPB_ERR_UNSUPPORTED; }
    OPB_CUDA(cudaSetDevice(f->device));
    cudaStream_t s = f->stream;
    const int n = f->width * f->height;
    const int is_u16 = depth_type == OPB_DEPTH_U16;
    OPB_CUDA(cudaMemcpyAsync(f->d_raw, depth, (size_t)n * (is_u16 ? 2 : 4), cudaMemcpyDefault, s));
    f->h_minmax[0] = 0xFFFFFFFFu; f->h_minmax[1] = 0u;
    OPB_CUDA(cudaMemcpyAsync(f->d_minmax, f->h_minmax, 2 * sizeof(unsigned int), cudaMemcpyHostToDevice, s));
    const int blocks = (n + 255) / 256 < f->sm_count * 4 ? (n + 255) / 256 : f->sm_count * 4;
    filter_convert_kernel<<<blocks, 256, 0, s>>>(f->d_raw, is_u16, depth_scale, n, f->d_conv, f->d_minmax);
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaMemcpyAsync(f->h_minmax, f->d_minmax, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    if (converted) OPB_CUDA(cudaMemcpyAsync(converted, f->d_conv, (size_t)n * sizeof(float), cudaMemcpyDefault, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    const float vmin = ordered_to_float(f->h_minmax[0]), vmax = ordered_to_float(f->h_minmax[1]);
    if (std::fabs((double)vmin - (double)vmax) < FLT_EPSILON)
    {
        // constant image: cv::bilateralFilter copies the source
        OPB_CUDA(cudaMemcpyAsync(f->d_out, f->d_conv, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    else
    {
        if (f->taps_d != radius || f->taps_sigma_space != sigma_space)
        {
            build_taps(radius, sigma_space, f->taps);
            f->taps_d = radius; f->taps_sigma_space = sigma_space;
        }
        if (!f->lut_valid || f->lut_min != vmin || f->lut_max != vmax || f->lut_sigma_color != sigma_color)
        {
            build_range_table(vmin, vmax, sigma_color, f->h_lut, &f->taps.scale_index);
            OPB_CUDA(cudaMemcpyAsync(f->d_lut, f->h_lut, kLutSize * sizeof(float), cudaMemcpyHostToDevice, s));
            f->lut_valid = true; f->lut_min = vmin; f->lut_max = vmax; f->lut_sigma_color = sigma_color;
        }
        const dim3 grid((f->width + kTileW - 1) / kTileW, (f->height + kTileH - 1) / kTileH), block(kTileW, kTileH);
        const size_t smem = (kLutSize + 2 + (size_t)(kTileW + 2 * radius) * (kTileH + 2 * radius)) * sizeof(float);
        bilateral_kernel<<<grid, block, smem, s>>>(f->d_conv, f->width, f->height, f->d_lut, f->taps, f->d_out);
        OPB_CUDA(cudaGetLastError());
    }
    if (filtered)
    {
        OPB_CUDA(cudaMemcpyAsync(filtered, f->d_out, (size_t)n * sizeof(float), cudaMemcpyDefault, s));
        OPB_CUDA(cudaStreamSynchronize(s));
    }
    return OPB_OK;
}

const float *opb_prefilter_device_result(opb_prefilter *f) { return f ? f->d_out : nullptr; }

int opb_prefilter_synchronize(opb_prefilter *f)
{
    if (!f) { set_error("prefilter is NULL"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(f->device));
    OPB_CUDA(cudaStreamSynchronize(f->stream));
    return OPB_OK;
}
} // extern "C"
