// Device-side image pre-processing of the training loader (movedepth/datasets/mono_dataset.py:104-126,
// 164, 206, 220-223): decoded uint8 frames -> horizontal flip -> 4-scale Lanczos pyramid -> colour jitter -> float tensors.
// Integer / byte work, bit-exact to what the reference executes in Pillow and torchvision:
//   * resample_u8: one pass of Pillow's ImagingResample (Resample.c) -- 22-bit fixed-point coefficients (computed on the
//     host exactly as Pillow does), int32 accumulation with the rounding bias 1 << 21, arithmetic shift, clip to uint8;
//     the horizontal pass reads a flipped image through a mirrored index when the item's flip flag is set;
//   * jitter_blend: ImageEnhance.Brightness / Contrast / Color = Image.blend(degenerate, image, factor) (Blend.c): float
//     multiply and add with separate roundings, truncation to uint8, clipping only when extrapolating;
//     L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16 (Convert.c), contrast mean = int(sum L / n + 0.5);
//   * jitter_hue: torchvision adjust_hue = Convert.c rgb2hsv -> h += uint8(factor * 255) -> hsv2rgb, float / double mix
//     reproduced operation by operation with round-to-nearest intrinsics (no FMA contraction);
//   * u8_to_tensor: ToTensor (x / 255.0f, HWC -> CHW).
// HBM traffic is trivial (a KITTI frame is 1.4 MB); the point is to take PIL's 12 worker processes off the critical path.
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {
namespace dp {

constexpr int PRECISION_BITS = 32 - 8 - 2;

// src [outer][n_in][inner] -> dst [outer][n_out][inner] (uint8); bounds [n_out][2] = (first tap, tap count), coeff [n_out][ksize].
// flip (nullable, one flag per image of `outer_per_image` outer rows): read tap x of the MIRRORED axis (horizontal pass only).
__global__ void __launch_bounds__(256)
resample_u8_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, const int* __restrict__ bounds,
                   const int* __restrict__ coeff, int ksize, long long outer, int n_in, int n_out, int inner,
                   const unsigned char* __restrict__ flip, long long outer_per_image) {
    const long long total = outer * n_out * inner;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(e % inner);
        long long r = e / inner;
        const int xx = static_cast<int>(r % n_out);
        const long long o = r / n_out;
        const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
        const bool mirrored = flip != nullptr && flip[o / outer_per_image] != 0;
        const unsigned char* row = src + o * n_in * static_cast<long long>(inner) + i;
        const int* k = coeff + static_cast<long long>(xx) * ksize;
        int acc = 1 << (PRECISION_BITS - 1);
        for (int x = 0; x < cnt; ++x) {
            const int p = mirrored ? (n_in - 1 - (xmin + x)) : (xmin + x);
            acc += static_cast<int>(row[static_cast<long long>(p) * inner]) * k[x];
        }
        const int v = acc >> PRECISION_BITS;
        dst[e] = static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

// plain copy with the optional mirrored x index (scale 0 when the native size already equals the target size)
__global__ void __launch_bounds__(256)
flip_copy_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int N, int H, int W, int C,
                 const unsigned char* __restrict__ flip) {
    const long long total = static_cast<long long>(N) * H * W * C;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % C);
        long long r = e / C;
        const int x = static_cast<int>(r % W);
        r /= W;
        const long long n = r / H;
        const int sx = (flip != nullptr && flip[n]) ? W - 1 - x : x;
        dst[e] = src[(r * W + sx) * C + c];
    }
}

__global__ void __launch_bounds__(256)
u8_to_tensor_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int N, int H, int W) {
    const long long total = static_cast<long long>(N) * 3 * H * W;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(e % W);
        long long r = e / W;
        const int y = static_cast<int>(r % H);
        r /= H;
        const int c = static_cast<int>(r % 3);
        const long long n = r / 3;
        dst[e] = __fdiv_rn(static_cast<float>(src[((n * H + y) * W + x) * 3 + c]), 255.0f);
    }
}

__device__ __forceinline__ int luma(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16; }

// per-image sum of L (for the contrast mean); sums zeroed by the caller
__global__ void __launch_bounds__(256)
luma_sum_kernel(const unsigned char* __restrict__ img, unsigned long long* __restrict__ sums, int hw) {
    const int n = blockIdx.y;
    const unsigned char* p = img + static_cast<long long>(n) * hw * 3;
    unsigned long long s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) s += luma(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sums + n, s);
}

// Blend.c: out = (UINT8)(d + a * (x - d)) when 0 <= a <= 1, else clipped
__device__ __forceinline__ unsigned char blend(float d, float x, float a, bool interpolate) {
    const float t = __fadd_rn(d, __fmul_rn(a, __fsub_rn(x, d)));
    if (interpolate) return static_cast<unsigned char>(t);
    if (t <= 0.0f) return 0;
    if (t >= 255.0f) return 255;
    return static_cast<unsigned char>(static_cast<int>(t));
}

// mode 0 brightness (degenerate = black), 1 contrast (degenerate = mean of L), 2 saturation (degenerate = L per pixel);
// active (nullable): images with active[n] == 0 are left untouched
__global__ void __launch_bounds__(256)
jitter_blend_kernel(unsigned char* __restrict__ img, int hw, int mode, const float* __restrict__ factor,
                    const unsigned long long* __restrict__ sums, const unsigned char* __restrict__ active) {
    const int n = blockIdx.y;
    if (active != nullptr && active[n] == 0) return;
    const float a = factor[n];
    const bool interp = a >= 0.0f && a <= 1.0f;
    float mean = 0.f;
    if (mode == 1) mean = static_cast<float>(static_cast<int>(static_cast<double>(sums[n]) / static_cast<double>(hw) + 0.5));
    unsigned char* p = img + static_cast<long long>(n) * hw * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const int r = p[3 * i], g = p[3 * i + 1], b = p[3 * i + 2];
        const float d = mode == 0 ? 0.f : (mode == 1 ? mean : static_cast<float>(luma(r, g, b)));
        p[3 * i] = blend(d, static_cast<float>(r), a, interp);
        p[3 * i + 1] = blend(d, static_cast<float>(g), a, interp);
        p[3 * i + 2] = blend(d, static_cast<float>(b), a, interp);
    }
}

__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Convert.c rgb2hsv_row, operation by operation (float where C computes in float, double where a double literal promotes)
__device__ __forceinline__ void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
    const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
    uv = maxc;
    if (minc == maxc) {
        uh = 0;
        us = 0;
        return;
    }
    const float cr = static_cast<float>(maxc - minc);
    const float s = __fdiv_rn(cr, static_cast<float>(maxc));
    const float rc = __fdiv_rn(static_cast<float>(maxc - r), cr);
    const float gc = __fdiv_rn(static_cast<float>(maxc - g), cr);
    const float bc = __fdiv_rn(static_cast<float>(maxc - b), cr);
    float h;
    if (r == maxc) h = __fsub_rn(bc, gc);
    else if (g == maxc) h = static_cast<float>(__dsub_rn(__dadd_rn(2.0, static_cast<double>(rc)), static_cast<double>(bc)));
    else h = static_cast<float>(__dsub_rn(__dadd_rn(4.0, static_cast<double>(gc)), static_cast<double>(rc)));
    h = static_cast<float>(fmod(__dadd_rn(__ddiv_rn(static_cast<double>(h), 6.0), 1.0), 1.0));
    uh = clip8(static_cast<int>(__dmul_rn(static_cast<double>(h), 255.0)));
    us = clip8(static_cast<int>(__dmul_rn(static_cast<double>(s), 255.0)));
}

// Convert.c hsv2rgb
__device__ __forceinline__ void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
    if (s == 0) {
        r = g = b = v;
        return;
    }
    const double hh = __ddiv_rn(__dmul_rn(static_cast<double>(static_cast<float>(h)), 6.0), 255.0);
    const int i = static_cast<int>(floorf(static_cast<float>(hh)));
    const float f = static_cast<float>(__dsub_rn(hh, static_cast<double>(static_cast<float>(i))));
    const float fs = static_cast<float>(__ddiv_rn(static_cast<double>(static_cast<float>(s)), 255.0));
    const double vd = static_cast<double>(static_cast<float>(v)), fsd = static_cast<double>(fs), fd = static_cast<double>(f);
    const int p = clip8(static_cast<int>(floor(__dadd_rn(__dmul_rn(vd, __dsub_rn(1.0, fsd)), 0.5))));
    const int q = clip8(static_cast<int>(floor(__dadd_rn(__dmul_rn(vd, __dsub_rn(1.0, __dmul_rn(fsd, fd))), 0.5))));
    const int t = clip8(static_cast<int>(floor(__dadd_rn(__dmul_rn(vd, __dsub_rn(1.0, __dmul_rn(fsd, __dsub_rn(1.0, fd)))), 0.5))));
    switch (i % 6) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}

__global__ void __launch_bounds__(256)
jitter_hue_kernel(unsigned char* __restrict__ img, int hw, const unsigned char* __restrict__ shift, const unsigned char* __restrict__ active) {
    const int n = blockIdx.y;
    if (active != nullptr && active[n] == 0) return;
    const int sh = shift[n];
    unsigned char* p = img + static_cast<long long>(n) * hw * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        int h, s, v, r, g, b;
        rgb2hsv(p[3 * i], p[3 * i + 1], p[3 * i + 2], h, s, v);
        h = (h + sh) & 255;                                  // uint8 addition with wrap-around
        hsv2rgb(h, s, v, r, g, b);
        p[3 * i] = static_cast<unsigned char>(r);
        p[3 * i + 1] = static_cast<unsigned char>(g);
        p[3 * i + 2] = static_cast<unsigned char>(b);
    }
}

static int blocks_for(long long n, int per_sm = 8) {
    const long long blocks = (n + 255) / 256, cap = static_cast<long long>(sm_count()) * per_sm;
    return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace dp
}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_resample_u8(const unsigned char* src, unsigned char* dst, const int* bounds, const int* coeff, int ksize, long long outer, int n_in,
                    int n_out, int inner, const unsigned char* flip, long long outer_per_image, void* stream) {
    MVD_REQUIRE(src && dst && bounds && coeff && ksize > 0 && outer > 0 && n_in > 0 && n_out > 0 && inner > 0, "bad argument");
    MVD_REQUIRE(flip == nullptr || outer_per_image > 0, "flip flags need outer_per_image");
    dp::resample_u8_kernel<<<dp::blocks_for(outer * n_out * inner), 256, 0, as_stream(stream)>>>(src, dst, bounds, coeff, ksize, outer, n_in,
                                                                                               n_out, inner, flip, outer_per_image);
    return check_launch("resample_u8");
}

int mvd_flip_copy_u8(const unsigned char* src, unsigned char* dst, int N, int H, int W, int C, const unsigned char* flip, void* stream) {
    MVD_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && C > 0, "bad argument");
    dp::flip_copy_kernel<<<dp::blocks_for(static_cast<long long>(N) * H * W * C), 256, 0, as_stream(stream)>>>(src, dst, N, H, W, C, flip);
    return check_launch("flip_copy_u8");
}

int mvd_u8_to_tensor(const unsigned char* src, float* dst, int N, int H, int W, void* stream) {
    MVD_REQUIRE(src && dst && N > 0 && H > 0 && W > 0, "bad argument");
    dp::u8_to_tensor_kernel<<<dp::blocks_for(static_cast<long long>(N) * 3 * H * W), 256, 0, as_stream(stream)>>>(src, dst, N, H, W);
    return check_launch("u8_to_tensor");
}

int mvd_jitter_blend_u8(unsigned char* img, int N, int H, int W, int mode, const float* factor, unsigned long long* sums,
                        const unsigned char* active, void* stream) {
    MVD_REQUIRE(img && factor && N > 0 && H > 0 && W > 0 && mode >= 0 && mode <= 2, "bad argument");
    MVD_REQUIRE(mode != 1 || sums != nullptr, "contrast needs the luminance sum buffer (N x uint64)");
    cudaStream_t st = as_stream(stream);
    const int hw = H * W;
    dim3 grid(static_cast<unsigned>(dp::blocks_for(hw, 1)), static_cast<unsigned>(N));
    if (mode == 1) {
        cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(unsigned long long) * N, st);
        if (e != cudaSuccess) return fail(static_cast<int>(e), "jitter memset: %s", cudaGetErrorString(e));
        dp::luma_sum_kernel<<<grid, 256, 0, st>>>(img, sums, hw);
        if (int rc = check_launch("luma_sum")) return rc;
    }
    dp::jitter_blend_kernel<<<grid, 256, 0, st>>>(img, hw, mode, factor, sums, active);
    return check_launch("jitter_blend");
}

int mvd_jitter_hue_u8(unsigned char* img, int N, int H, int W, const unsigned char* shift, const unsigned char* active, void* stream) {
    MVD_REQUIRE(img && shift && N > 0 && H > 0 && W > 0, "bad argument");
    const int hw = H * W;
    dim3 grid(static_cast<unsigned>(dp::blocks_for(hw, 1)), static_cast<unsigned>(N));
    dp::jitter_hue_kernel<<<grid, 256, 0, as_stream(stream)>>>(img, hw, shift, active);
    return check_launch("jitter_hue");
}

}
