// Loss glue of the training step that used to be ~25 eager tensor ops per scale (forward) and as many again backward:
//   * disp_to_depth:    bilinear upsample of a sigmoid disparity to full resolution (align_corners=False) fused with
//                       depth = 1 / (1/max + disp * (1/min - 1/max))        movedepth/trainer.py:512-515, layers.py:400-409
//   * smooth_loss:      mean-normalised disparity + edge-aware first-order smoothness           trainer.py:712-713, layers.py:630-643
//   * masked_smooth_l1: smooth-L1 between the depth of the box-masked and the plain reference frame over the pixels whose
//                       bilinearly resized (align_corners=True) box mask is non-zero              trainer.py:398-400, layers.py:52-69
// All three are HBM-bound streaming kernels over at most B*H*W floats; reductions are fp32 per thread, fp64 across blocks.
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {
namespace glue {

__device__ __forceinline__ float block_sum_to(float v, double* dst) {
    v = warp_sum(v);
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += static_cast<double>(red[w]);
        atomicAdd(dst, t);
    }
    return v;
}

// ATen's source index for align_corners=False (area_pixel_compute_source_index): max(scale * (dst + 0.5) - 0.5, 0)
__device__ __forceinline__ void src_index(float scale, int dst, int in, int& i0, int& i1, float& l0, float& l1) {
    float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = static_cast<int>(s);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - static_cast<float>(i0);
    l0 = 1.f - l1;
}

// ------------------------------------------------------------------------------------------------ disp -> full-res depth
__global__ void __launch_bounds__(256)
disp_to_depth_fwd_kernel(const float* __restrict__ disp, float* __restrict__ depth, int hs, int ws, int H, int W, float inv_far,
                         float range, float sy, float sx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const float* d = disp + static_cast<size_t>(b) * hs * ws;
    float v;
    if (hs == H && ws == W) {
        v = __ldg(d + static_cast<size_t>(y) * ws + x);
    } else {
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        src_index(sy, y, hs, y0, y1, ly0, ly1);
        src_index(sx, x, ws, x0, x1, lx0, lx1);
        const float* r0 = d + static_cast<size_t>(y0) * ws;
        const float* r1 = d + static_cast<size_t>(y1) * ws;
        v = ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) + ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
    }
    depth[(static_cast<size_t>(b) * H + y) * W + x] = 1.0f / __fadd_rn(inv_far, __fmul_rn(range, v));
}

// gather form of the adjoint: one thread per low-resolution pixel visits the full-resolution pixels whose bilinear footprint
// contains it (deterministic, no atomics); d depth / d disp_up = -depth^2 * range
__global__ void __launch_bounds__(128)
disp_to_depth_bwd_kernel(const float* __restrict__ gdepth, const float* __restrict__ depth, float* __restrict__ gdisp, int hs, int ws,
                         int H, int W, float range, float sy, float sx) {
    const int xi = blockIdx.x * blockDim.x + threadIdx.x, yi = blockIdx.y, b = blockIdx.z;
    if (xi >= ws) return;
    const float* g = gdepth + static_cast<size_t>(b) * H * W;
    const float* dp = depth + static_cast<size_t>(b) * H * W;
    float acc = 0.f;
    if (hs == H && ws == W) {
        const size_t i = static_cast<size_t>(yi) * W + xi;
        const float z = __ldg(dp + i);
        acc = -__ldg(g + i) * z * z * range;
    } else {
        const int fy = H / hs, fx = W / ws;              // integer factors (checked by the host)
        const int ylo = max(0, (yi - 1) * fy), yhi = min(H - 1, (yi + 2) * fy);
        const int xlo = max(0, (xi - 1) * fx), xhi = min(W - 1, (xi + 2) * fx);
        for (int y = ylo; y <= yhi; ++y) {
            int y0, y1;
            float ly0, ly1;
            src_index(sy, y, hs, y0, y1, ly0, ly1);
            const float wy = (y0 == yi ? ly0 : 0.f) + (y1 == yi ? ly1 : 0.f);
            if (wy == 0.f) continue;
            for (int x = xlo; x <= xhi; ++x) {
                int x0, x1;
                float lx0, lx1;
                src_index(sx, x, ws, x0, x1, lx0, lx1);
                const float wx = (x0 == xi ? lx0 : 0.f) + (x1 == xi ? lx1 : 0.f);
                if (wx == 0.f) continue;
                const size_t i = static_cast<size_t>(y) * W + x;
                const float z = __ldg(dp + i);
                acc += wy * wx * (-__ldg(g + i) * z * z * range);
            }
        }
    }
    gdisp[(static_cast<size_t>(b) * hs + yi) * ws + xi] = acc;
}

// ------------------------------------------------------------------------------------------------ edge-aware smoothness
// work layout (doubles): [0..B) per-item sum of disp, [B] sum of the x terms, [B+1] sum of the y terms
__global__ void __launch_bounds__(256)
plane_sum_kernel(const float* __restrict__ d, double* __restrict__ work, int hw) {
    const int b = blockIdx.y;
    const float* p = d + static_cast<size_t>(b) * hw;
    float s = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) s += __ldg(p + i);
    block_sum_to(s, work + b);
}

__device__ __forceinline__ float edge_weight(const float* __restrict__ img, size_t plane, size_t i, size_t j) {
    const float e = fabsf(__ldg(img + i) - __ldg(img + j)) + fabsf(__ldg(img + plane + i) - __ldg(img + plane + j)) +
                    fabsf(__ldg(img + 2 * plane + i) - __ldg(img + 2 * plane + j));
    return expf(-e / 3.0f);
}

__global__ void __launch_bounds__(256)
smooth_fwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, double* __restrict__ work, int B, int h, int w,
                  int normalize) {
    const int b = blockIdx.y;
    const size_t hw = static_cast<size_t>(h) * w;
    const float inv = normalize ? 1.0f / (static_cast<float>(work[b] / static_cast<double>(hw)) + 1e-7f) : 1.0f;
    const float* d = disp + b * hw;
    const float* im = img + 3 * b * hw;
    float sx = 0.f, sy = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < static_cast<int>(hw); i += gridDim.x * blockDim.x) {
        const int y = i / w, x = i - y * w;
        const float c = __ldg(d + i) * inv;
        if (x + 1 < w) sx += fabsf(c - __ldg(d + i + 1) * inv) * edge_weight(im, hw, i, i + 1);
        if (y + 1 < h) sy += fabsf(c - __ldg(d + i + w) * inv) * edge_weight(im, hw, i, i + w);
    }
    block_sum_to(sx, work + B);
    block_sum_to(sy, work + B + 1);
}

__global__ void smooth_finalize_kernel(const double* __restrict__ work, float* __restrict__ loss, int B, int h, int w) {
    const double nx = static_cast<double>(B) * h * (w - 1), ny = static_cast<double>(B) * (h - 1) * w;
    loss[0] = static_cast<float>(work[B] / nx + work[B + 1] / ny);
}

__device__ __forceinline__ float sgn(float v) { return static_cast<float>(v > 0.f) - static_cast<float>(v < 0.f); }

// pass 1 of the backward: g = d loss / d normalised disparity (gather over the 4 neighbour pairs), dot[b] = sum g * disp
__global__ void __launch_bounds__(256)
smooth_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, const double* __restrict__ work, double* __restrict__ dot,
                  float* __restrict__ g_out, int B, int h, int w, int normalize) {
    const int b = blockIdx.y;
    const size_t hw = static_cast<size_t>(h) * w;
    const float inv = normalize ? 1.0f / (static_cast<float>(work[b] / static_cast<double>(hw)) + 1e-7f) : 1.0f;
    const float cx = static_cast<float>(1.0 / (static_cast<double>(B) * h * (w - 1)));
    const float cy = static_cast<float>(1.0 / (static_cast<double>(B) * (h - 1) * w));
    const float* d = disp + b * hw;
    const float* im = img + 3 * b * hw;
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < static_cast<int>(hw); i += gridDim.x * blockDim.x) {
        const int y = i / w, x = i - y * w;
        const float c = __ldg(d + i) * inv;
        float g = 0.f;
        if (x + 1 < w) g += cx * sgn(c - __ldg(d + i + 1) * inv) * edge_weight(im, hw, i, i + 1);
        if (x > 0) g -= cx * sgn(__ldg(d + i - 1) * inv - c) * edge_weight(im, hw, i - 1, i);
        if (y + 1 < h) g += cy * sgn(c - __ldg(d + i + w) * inv) * edge_weight(im, hw, i, i + w);
        if (y > 0) g -= cy * sgn(__ldg(d + i - w) * inv - c) * edge_weight(im, hw, i - w, i);
        g_out[b * hw + i] = g;
        acc += g * __ldg(d + i);
    }
    if (normalize) block_sum_to(acc, dot + b);
}

// pass 2: d/d disp of n = disp / (mean + eps):  g/(m+eps) - dot / ((m+eps)^2 * hw), times the incoming scalar gradient
__global__ void __launch_bounds__(256)
smooth_bwd_finish_kernel(const float* __restrict__ gloss, const double* __restrict__ work, const double* __restrict__ dot,
                         float* __restrict__ g, int h, int w, int normalize) {
    const int b = blockIdx.y;
    const size_t hw = static_cast<size_t>(h) * w;
    const float gl = gloss[0];
    float a = gl, c = 0.f;
    if (normalize) {
        const float m = static_cast<float>(work[b] / static_cast<double>(hw)) + 1e-7f;
        a = gl / m;
        c = gl * static_cast<float>(dot[b]) / (m * m * static_cast<float>(hw));
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < static_cast<int>(hw); i += gridDim.x * blockDim.x)
        g[b * hw + i] = g[b * hw + i] * a - c;
}

// ------------------------------------------------------------------------------------------------ masked smooth-L1
__device__ __forceinline__ float box_mask(int y, int x, int by, int bx, int fh, int fw) {
    return (x >= bx && x < bx + fw && y >= by && y < by + fh) ? 0.f : 1.f;
}

__global__ void __launch_bounds__(256)
masked_sl1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ bb, const long long* __restrict__ box,
                      unsigned char* __restrict__ sel, double* __restrict__ sums, int B, int h, int w, int H, int W, int fh, int fw,
                      float sy, float sx) {
    const int bx = static_cast<int>(box[0]), by = static_cast<int>(box[1]);
    const int n = B * h * w;
    float s = 0.f, c = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = i % w, y = (i / w) % h;
        // ATen bilinear, align_corners=True: src = dst * (in-1)/(out-1)
        const float fy = sy * static_cast<float>(y), fx = sx * static_cast<float>(x);
        const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
        const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly1 = fy - static_cast<float>(y0), lx1 = fx - static_cast<float>(x0);
        const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
        const float v = ly0 * (lx0 * box_mask(y0, x0, by, bx, fh, fw) + lx1 * box_mask(y0, x1, by, bx, fh, fw)) +
                        ly1 * (lx0 * box_mask(y1, x0, by, bx, fh, fw) + lx1 * box_mask(y1, x1, by, bx, fh, fw));
        const bool on = v != 0.f;
        sel[i] = on ? 1 : 0;
        if (on) {
            const float d = fabsf(__ldg(a + i) - __ldg(bb + i));
            s += d < 1.f ? 0.5f * d * d : d - 0.5f;
            c += 1.f;
        }
    }
    block_sum_to(s, sums);
    block_sum_to(c, sums + 1);
}

__global__ void masked_sl1_finalize_kernel(const double* __restrict__ sums, float* __restrict__ loss, float weight) {
    loss[0] = static_cast<float>(sums[0] / sums[1]) * weight;            // no selected pixel -> NaN, like the mean of an empty tensor
}

__global__ void __launch_bounds__(256)
masked_sl1_bwd_kernel(const float* __restrict__ gloss, const float* __restrict__ a, const float* __restrict__ bb,
                      const unsigned char* __restrict__ sel, const double* __restrict__ sums, float* __restrict__ ga,
                      float* __restrict__ gb, int n, float weight) {
    const float k = gloss[0] * weight * static_cast<float>(1.0 / sums[1]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = 0.f;
        if (sel[i]) {
            const float d = __ldg(a + i) - __ldg(bb + i);
            g = k * (fabsf(d) < 1.f ? d : sgn(d));
        }
        ga[i] = g;
        if (gb != nullptr) gb[i] = -g;
    }
}

static int blocks_for(long long n, int threads, int per_sm = 8) {
    const long long blocks = (n + threads - 1) / threads, cap = static_cast<long long>(sm_count()) * per_sm;
    return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace glue
}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_disp_to_depth_fwd(const float* disp, float* depth, int B, int hs, int ws, int H, int W, float inv_far, float range, void* stream) {
    MVD_REQUIRE(disp && depth && B > 0 && hs > 0 && ws > 0 && H >= hs && W >= ws, "bad argument");
    MVD_REQUIRE(H % hs == 0 && W % ws == 0, "disp_to_depth: %dx%d is not an integer multiple of %dx%d", H, W, hs, ws);
    dim3 grid((W + 255) / 256, H, B);
    glue::disp_to_depth_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(disp, depth, hs, ws, H, W, inv_far, range,
                                                                         static_cast<float>(hs) / static_cast<float>(H),
                                                                         static_cast<float>(ws) / static_cast<float>(W));
    return check_launch("disp_to_depth_fwd");
}

int mvd_disp_to_depth_bwd(const float* gdepth, const float* depth, float* gdisp, int B, int hs, int ws, int H, int W, float range,
                          void* stream) {
    MVD_REQUIRE(gdepth && depth && gdisp && B > 0 && hs > 0 && ws > 0, "bad argument");
    MVD_REQUIRE(H % hs == 0 && W % ws == 0, "disp_to_depth: %dx%d is not an integer multiple of %dx%d", H, W, hs, ws);
    dim3 grid((ws + 127) / 128, hs, B);
    glue::disp_to_depth_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(gdepth, depth, gdisp, hs, ws, H, W, range,
                                                                         static_cast<float>(hs) / static_cast<float>(H),
                                                                         static_cast<float>(ws) / static_cast<float>(W));
    return check_launch("disp_to_depth_bwd");
}

long long mvd_smooth_loss_workspace_bytes(int B) { return static_cast<long long>(B + 2) * sizeof(double); }

int mvd_smooth_loss_fwd(const float* disp, const float* img, double* work, float* loss, int B, int h, int w, int normalize,
                        void* stream) {
    MVD_REQUIRE(disp && img && work && loss && B > 0 && h > 1 && w > 1, "bad argument");
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(work, 0, static_cast<size_t>(B + 2) * sizeof(double), st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "smooth_loss memset: %s", cudaGetErrorString(e));
    const int hw = h * w;
    const int nb = glue::blocks_for(hw, 256, 2);
    if (normalize) {
        glue::plane_sum_kernel<<<dim3(nb, B), 256, 0, st>>>(disp, work, hw);
        if (int rc = check_launch("smooth_loss plane_sum")) return rc;
    }
    glue::smooth_fwd_kernel<<<dim3(nb, B), 256, 0, st>>>(disp, img, work, B, h, w, normalize);
    if (int rc = check_launch("smooth_loss_fwd")) return rc;
    glue::smooth_finalize_kernel<<<1, 1, 0, st>>>(work, loss, B, h, w);
    return check_launch("smooth_loss finalize");
}

int mvd_smooth_loss_bwd(const float* gloss, const float* disp, const float* img, const double* work, double* dot, float* gdisp, int B,
                        int h, int w, int normalize, void* stream) {
    MVD_REQUIRE(gloss && disp && img && work && dot && gdisp && B > 0 && h > 1 && w > 1, "bad argument");
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(dot, 0, static_cast<size_t>(B) * sizeof(double), st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "smooth_loss memset: %s", cudaGetErrorString(e));
    const int nb = glue::blocks_for(h * w, 256, 2);
    glue::smooth_bwd_kernel<<<dim3(nb, B), 256, 0, st>>>(disp, img, work, dot, gdisp, B, h, w, normalize);
    if (int rc = check_launch("smooth_loss_bwd")) return rc;
    glue::smooth_bwd_finish_kernel<<<dim3(nb, B), 256, 0, st>>>(gloss, work, dot, gdisp, h, w, normalize);
    return check_launch("smooth_loss_bwd finish");
}

int mvd_masked_smooth_l1_fwd(const float* a, const float* b, const long long* box_xy, unsigned char* sel, double* sums, float* loss,
                             int B, int h, int w, int H, int W, int fh, int fw, float weight, void* stream) {
    MVD_REQUIRE(a && b && box_xy && sel && sums && loss && B > 0 && h > 1 && w > 1 && H > 1 && W > 1, "bad argument");
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "masked_smooth_l1 memset: %s", cudaGetErrorString(e));
    const long long n = static_cast<long long>(B) * h * w;
    glue::masked_sl1_fwd_kernel<<<glue::blocks_for(n, 256), 256, 0, st>>>(a, b, box_xy, sel, sums, B, h, w, H, W, fh, fw,
                                                                          static_cast<float>(H - 1) / static_cast<float>(h - 1),
                                                                          static_cast<float>(W - 1) / static_cast<float>(w - 1));
    if (int rc = check_launch("masked_smooth_l1_fwd")) return rc;
    glue::masked_sl1_finalize_kernel<<<1, 1, 0, st>>>(sums, loss, weight);
    return check_launch("masked_smooth_l1 finalize");
}

int mvd_masked_smooth_l1_bwd(const float* gloss, const float* a, const float* b, const unsigned char* sel, const double* sums, float* ga,
                             float* gb, long long n, float weight, void* stream) {
    MVD_REQUIRE(gloss && a && b && sel && sums && ga && n > 0, "bad argument");
    glue::masked_sl1_bwd_kernel<<<glue::blocks_for(n, 256), 256, 0, as_stream(stream)>>>(gloss, a, b, sel, sums, ga, gb,
                                                                                         static_cast<int>(n), weight);
    return check_launch("masked_smooth_l1_bwd");
}

}

// ------------------------------------------------------------------------------------------------ ResNet stem max-pool
// MaxPool2d(kernel 3, stride 2, padding 1) of the ResNet stem (torchvision resnet; movedepth/networks/resnet_encoder.py:113)
// on channels-last activations: forward keeps the window position of the maximum (first maximum in scan order, like ATen),
// backward is a gather over the <= 4 windows that contain an input pixel.  ATen's NHWC kernels need 130 + 200 us for the
// 47 MB stem activation; these are plain float4 streams.
namespace mvd {
namespace glue {

__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, uchar4* __restrict__ idx, int B, int H, int W, int C4, int Ho, int Wo) {
    const long long total = static_cast<long long>(B) * Ho * Wo * C4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C4);
        long long r = i / C4;
        const int ox = static_cast<int>(r % Wo);
        r /= Wo;
        const int oy = static_cast<int>(r % Ho), b = static_cast<int>(r / Ho);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        uchar4 k = make_uchar4(0, 0, 0, 0);
        bool first = true;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                if (ix < 0 || ix >= W) continue;
                const float4 v = __ldg(x + ((static_cast<long long>(b) * H + iy) * W + ix) * C4 + c);
                const unsigned char t = static_cast<unsigned char>(ky * 3 + kx);
                if (first || v.x > m.x || v.x != v.x) { m.x = v.x; k.x = t; }
                if (first || v.y > m.y || v.y != v.y) { m.y = v.y; k.y = t; }
                if (first || v.z > m.z || v.z != v.z) { m.z = v.z; k.z = t; }
                if (first || v.w > m.w || v.w != v.w) { m.w = v.w; k.w = t; }
                first = false;
            }
        }
        y[i] = m;
        idx[i] = k;
    }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float4* __restrict__ gy, const uchar4* __restrict__ idx, float4* __restrict__ gx, int B, int H, int W, int C4,
                   int Ho, int Wo) {
    const long long total = static_cast<long long>(B) * H * W * C4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C4);
        long long r = i / C4;
        const int ix = static_cast<int>(r % W);
        r /= W;
        const int iy = static_cast<int>(r % H), b = static_cast<int>(r / H);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int oy = iy / 2; oy <= (iy + 1) / 2 && oy < Ho; ++oy) {           // windows with 2*oy-1 <= iy <= 2*oy+1
            const int ky = iy - (2 * oy - 1);
            for (int ox = ix / 2; ox <= (ix + 1) / 2 && ox < Wo; ++ox) {
                const unsigned char t = static_cast<unsigned char>(ky * 3 + (ix - (2 * ox - 1)));
                const long long o = ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * C4 + c;
                const uchar4 k = idx[o];
                const float4 v = __ldg(gy + o);
                if (k.x == t) g.x += v.x;
                if (k.y == t) g.y += v.y;
                if (k.z == t) g.z += v.z;
                if (k.w == t) g.w += v.w;
            }
        }
        gx[i] = g;
    }
}

// ---- pose-net outputs -> 4x4 camera transform (movedepth/layers.py:412-429, 464-518) -----------------------------------
// transformation_from_parameters = rot_from_axisangle (Rodrigues with the reference's angle + 1e-7 guard) composed with the
// translation: M = [R t; 0 1], or for invert = [R^T  -R^T t; 0 1].  The reference spends ~50 tiny elementwise launches per call
// and direction on it; here it is one launch each.  The backward uses forward-mode differentiation of the same expression
// (a thread per (item, input component) evaluates the 12 outputs with a dual number seeded on its component and contracts
// them with the upstream gradient), so value and derivative can never drift apart.
__device__ __forceinline__ float tsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float tcos(float x) { return cosf(x); }
__device__ __forceinline__ float tsin(float x) { return sinf(x); }

struct Dual {
    float v, d;
    __device__ __forceinline__ Dual() : v(0.f), d(0.f) {}
    __device__ __forceinline__ explicit Dual(float v_) : v(v_), d(0.f) {}
    __device__ __forceinline__ Dual(float v_, float d_) : v(v_), d(d_) {}
};
__device__ __forceinline__ Dual operator+(const Dual& a, const Dual& b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(const Dual& a, const Dual& b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(const Dual& a) { return Dual(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(const Dual& a, const Dual& b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ Dual operator/(const Dual& a, const Dual& b) {
    const float q = a.v / b.v;
    return Dual(q, (a.d - q * b.d) / b.v);
}
__device__ __forceinline__ Dual tsqrt(const Dual& x) {
    const float r = sqrtf(x.v);
    return Dual(r, r > 0.f ? 0.5f * x.d / r : 0.f);    // torch's norm backward: zero sub-gradient at the origin
}
__device__ __forceinline__ Dual tcos(const Dual& x) { return Dual(cosf(x.v), -sinf(x.v) * x.d); }
__device__ __forceinline__ Dual tsin(const Dual& x) { return Dual(sinf(x.v), cosf(x.v) * x.d); }

template <typename T>
struct PoseEval {
    // out[12]: rows of [R | t] (row-major 3x4)
    static __device__ __forceinline__ void run(const T (&a)[3], const T (&t)[3], int invert, T (&out)[12]) {
        const T theta = tsqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        const T inv = T(1.f) / (theta + T(1e-7f));
        const T ux = a[0] * inv, uy = a[1] * inv, uz = a[2] * inv;
        const T ca = tcos(theta), sa = tsin(theta);
        const T c1 = T(1.f) - ca;
        const T tx = ux * c1, ty = uy * c1, tz = uz * c1;
        T R[9] = {ux * tx + ca, ux * ty - uz * sa, uz * tx + uy * sa,
                  ux * ty + uz * sa, uy * ty + ca, uy * tz - ux * sa,
                  uz * tx - uy * sa, uy * tz + ux * sa, uz * tz + ca};
        if (!invert) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) out[i * 4 + j] = R[i * 3 + j];
                out[i * 4 + 3] = t[i];
            }
        } else {                                        // R^T * T(-t)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) out[i * 4 + j] = R[j * 3 + i];
                out[i * 4 + 3] = (R[0 * 3 + i] * (-t[0]) + R[1 * 3 + i] * (-t[1])) + R[2 * 3 + i] * (-t[2]);
            }
        }
    }
};
__global__ void pose_matrix_fwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, float* __restrict__ M, int B, int invert) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float a[3] = {aa[3 * b], aa[3 * b + 1], aa[3 * b + 2]}, t[3] = {tr[3 * b], tr[3 * b + 1], tr[3 * b + 2]};
    float out[12];
    PoseEval<float>::run(a, t, invert, out);
    float* m = M + 16 * b;
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = out[i];
    m[12] = 0.f; m[13] = 0.f; m[14] = 0.f; m[15] = 1.f;
}

// thread = (item b, input component j): j < 3 axis-angle, j >= 3 translation
__global__ void pose_matrix_bwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, const float* __restrict__ gM,
                                       float* __restrict__ gaa, float* __restrict__ gtr, int B, int invert) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * B) return;
    const int b = i / 6, j = i - 6 * b;
    Dual a[3], t[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = Dual(aa[3 * b + k], j == k ? 1.f : 0.f);
        t[k] = Dual(tr[3 * b + k], j == 3 + k ? 1.f : 0.f);
    }
    Dual out[12];
    PoseEval<Dual>::run(a, t, invert, out);
    float g = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) g = fmaf(gM[16 * b + k], out[k].d, g);
    if (j < 3) gaa[3 * b + j] = g;
    else gtr[3 * b + j - 3] = g;
}

}  // namespace glue
}  // namespace mvd

extern "C" {

int mvd_maxpool3x3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C, void* stream) {
    MVD_REQUIRE(x && y && idx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "bad argument (C must be a multiple of 4)");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = static_cast<long long>(B) * Ho * Wo * (C / 4);
    glue::maxpool_fwd_kernel<<<glue::blocks_for(total, 256, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), reinterpret_cast<uchar4*>(idx), B, H, W, C / 4, Ho, Wo);
    return check_launch("maxpool3x3s2_fwd");
}

int mvd_maxpool3x3s2_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int H, int W, int C, void* stream) {
    MVD_REQUIRE(gy && gx && idx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "bad argument (C must be a multiple of 4)");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = static_cast<long long>(B) * H * W * (C / 4);
    glue::maxpool_bwd_kernel<<<glue::blocks_for(total, 256, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(gy), reinterpret_cast<const uchar4*>(idx), reinterpret_cast<float4*>(gx), B, H, W, C / 4, Ho, Wo);
    return check_launch("maxpool3x3s2_bwd");
}

int mvd_pose_matrix_fwd(const float* axisangle, const float* translation, float* M, int B, int invert, void* stream) {
    MVD_REQUIRE(axisangle && translation && M && B > 0, "bad argument");
    glue::pose_matrix_fwd_kernel<<<(B + 63) / 64, 64, 0, as_stream(stream)>>>(axisangle, translation, M, B, invert ? 1 : 0);
    return check_launch("pose_matrix_fwd");
}

int mvd_pose_matrix_bwd(const float* axisangle, const float* translation, const float* gM, float* g_axisangle, float* g_translation,
                        int B, int invert, void* stream) {
    MVD_REQUIRE(axisangle && translation && gM && g_axisangle && g_translation && B > 0, "bad argument");
    glue::pose_matrix_bwd_kernel<<<(6 * B + 63) / 64, 64, 0, as_stream(stream)>>>(axisangle, translation, gM, g_axisangle, g_translation, B,
                                                                                  invert ? 1 : 0);
    return check_launch("pose_matrix_bwd");
}

}
