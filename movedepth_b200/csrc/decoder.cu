// DepthDecoder glue (movedepth/networks/depth_decoder.py:72-101, layers.py:521-553, 624-627): everything between two
// 3x3 convolutions of the U-Net decoder in ONE pass per direction.  The reference runs, per decoder conv,
//   bias add -> ELU -> [nearest x2 upsample -> cat(skip)] -> ReflectionPad2d(1)      (+ the 3xTF32 operand split here)
// as 5-6 full-tensor kernels (and as many again backward).  `decoder_prep` reads the previous conv's raw output z
// (channels-last, bias not yet added) and the skip feature once and writes the next conv's input directly:
//   xp [B, H+2, W+2, C1+C2]  = reflect_pad(cat(up(elu(z + bias)), skip))            (kept for the weight gradient)
//   x3 [B, H+2, W+2, 3(C1+C2)] = [hi | lo | hi] TF32 split of xp                    (the tensor-core forward operand)
// The backward folds the padding adjoint, the upsampling adjoint (2x2 sum), ELU' and the bias reduction into one gather.
// HBM-bound: forward reads (C1/up^2 + C2) and writes 4(C1+C2) floats per output pixel.
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {
namespace dec {

struct Args {
    const float* z;      // [B,h,w,C1]
    const float* bias;   // [C1] or null
    const float* skip;   // [B,H,W,C2] or null
    float* xp;           // [B,H+2,W+2,C]
    float* x3;           // [B,H+2,W+2,3C] or null
    int B, h, w, C1, C2, up, act;
};

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float elu(float t) { return t > 0.f ? t : expm1f(t); }
__device__ __forceinline__ int reflect(int p, int n) { return p == 0 ? 1 : (p == n + 1 ? n - 2 : p - 1); }

__global__ void __launch_bounds__(256) prep_fwd_kernel(const Args a) {
    const int H = a.h * a.up, W = a.w * a.up, C = a.C1 + a.C2, C4 = C >> 2, Hp = H + 2, Wp = W + 2;
    const long long total = static_cast<long long>(a.B) * Hp * Wp * C4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(i % C4);
        long long r = i / C4;
        const int px = static_cast<int>(r % Wp);
        r /= Wp;
        const int py = static_cast<int>(r % Hp), b = static_cast<int>(r / Hp);
        const int y = reflect(py, H), x = reflect(px, W), c = c4 << 2;
        float4 v;
        if (c < a.C1) {
            v = __ldg(reinterpret_cast<const float4*>(a.z + ((static_cast<size_t>(b) * a.h + y / a.up) * a.w + x / a.up) * a.C1 + c));
            if (a.bias != nullptr) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + c));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            if (a.act) v = make_float4(elu(v.x), elu(v.y), elu(v.z), elu(v.w));
        } else {
            v = __ldg(reinterpret_cast<const float4*>(a.skip + ((static_cast<size_t>(b) * H + y) * W + x) * a.C2 + (c - a.C1)));
        }
        const size_t pix = (static_cast<size_t>(b) * Hp + py) * Wp + px;
        *reinterpret_cast<float4*>(a.xp + pix * C + c) = v;
        if (a.x3 != nullptr) {
            const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
            float4* o = reinterpret_cast<float4*>(a.x3 + pix * 3 * C + c);
            o[0] = hi;
            o[C4] = lo;
            o[2 * C4] = hi;
        }
    }
}

struct BArgs {
    const float* gxp;    // [B,H+2,W+2,C]
    const float* z;      // [B,h,w,C1]
    const float* bias;   // [C1] or null
    float* gz;           // [B,h,w,C1]
    float* gskip;        // [B,H,W,C2] or null
    float* gbias;        // [C1] (pre-zeroed) or null
    int B, h, w, C1, C2, up, act;
};

// adjoint of ReflectionPad2d(1) at interior pixel (y, x): its own padded position plus the mirrored border copies
__device__ __forceinline__ float4 fold(const float* __restrict__ g, int b, int y, int x, int H, int W, int C, int c) {
    const int Hp = H + 2, Wp = W + 2;
    // row 1 also receives padded row 0, row H-2 padded row H+1 (both when H == 3); same for the columns
    const int ys[3] = {y + 1, y == 1 ? 0 : -1, y == H - 2 ? H + 1 : -1};
    const int xs[3] = {x + 1, x == 1 ? 0 : -1, x == W - 2 ? W + 1 : -1};
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (ys[i] < 0) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (xs[j] < 0) continue;
            const float4 v = __ldg(reinterpret_cast<const float4*>(g + ((static_cast<size_t>(b) * Hp + ys[i]) * Wp + xs[j]) * C + c));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    return s;
}

// z part: thread = (low-res pixel, 4 channels), channel group fastest; blockDim is a multiple of C1/4 so that a thread keeps
// its channel group across the grid-stride loop and the bias gradient reduces over the threads of equal (tid % (C1/4))
__global__ void __launch_bounds__(256) prep_bwd_z_kernel(const BArgs a) {
    const int H = a.h * a.up, W = a.w * a.up, C = a.C1 + a.C2, G = a.C1 >> 2;
    const long long total = static_cast<long long>(a.B) * a.h * a.w * G;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % G) << 2;      // == (threadIdx.x % G) * 4 for every i when G divides 256 (bias path)
        long long r = i / G;
        const int xl = static_cast<int>(r % a.w);
        r /= a.w;
        const int yl = static_cast<int>(r % a.h), b = static_cast<int>(r / a.h);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int dy = 0; dy < a.up; ++dy)
            for (int dx = 0; dx < a.up; ++dx) {
                const float4 v = fold(a.gxp, b, yl * a.up + dy, xl * a.up + dx, H, W, C, c);
                g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
            }
        const size_t o = ((static_cast<size_t>(b) * a.h + yl) * a.w + xl) * a.C1 + c;
        if (a.act) {
            float4 t = __ldg(reinterpret_cast<const float4*>(a.z + o));
            if (a.bias != nullptr) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + c));
                t.x += bb.x; t.y += bb.y; t.z += bb.z; t.w += bb.w;
            }
            g.x *= t.x > 0.f ? 1.f : expf(t.x);
            g.y *= t.y > 0.f ? 1.f : expf(t.y);
            g.z *= t.z > 0.f ? 1.f : expf(t.z);
            g.w *= t.w > 0.f ? 1.f : expf(t.w);
        }
        *reinterpret_cast<float4*>(a.gz + o) = g;
        bsum.x += g.x; bsum.y += g.y; bsum.z += g.z; bsum.w += g.w;
    }
    if (a.gbias != nullptr) {
        __shared__ float4 red[256];
        red[threadIdx.x] = bsum;
        __syncthreads();
        if (threadIdx.x < G) {
            const int c = threadIdx.x << 2;
            float4 s = red[threadIdx.x];
            for (int t = threadIdx.x + G; t < blockDim.x; t += G) {
                const float4 v = red[t];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            atomicAdd(a.gbias + c + 0, s.x);
            atomicAdd(a.gbias + c + 1, s.y);
            atomicAdd(a.gbias + c + 2, s.z);
            atomicAdd(a.gbias + c + 3, s.w);
        }
    }
}

__global__ void __launch_bounds__(256) prep_bwd_skip_kernel(const BArgs a) {
    const int H = a.h * a.up, W = a.w * a.up, C = a.C1 + a.C2, G = a.C2 >> 2;
    const long long total = static_cast<long long>(a.B) * H * W * G;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % G) << 2;
        long long r = i / G;
        const int x = static_cast<int>(r % W);
        r /= W;
        const int y = static_cast<int>(r % H), b = static_cast<int>(r / H);
        *reinterpret_cast<float4*>(a.gskip + ((static_cast<size_t>(b) * H + y) * W + x) * a.C2 + c) = fold(a.gxp, b, y, x, H, W, C, a.C1 + c);
    }
}

static int blocks_for(long long n, int per_sm = 8) {
    const long long blocks = (n + 255) / 256, cap = static_cast<long long>(sm_count()) * per_sm;
    return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace dec
}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_decoder_prep_fwd(const float* z, const float* bias, const float* skip, float* xp, float* x3, int B, int h, int w, int C1, int C2,
                         int up, int act, void* stream) {
    MVD_REQUIRE(z && xp && B > 0 && h > 0 && w > 0, "bad argument");
    MVD_REQUIRE(C1 > 0 && C1 % 4 == 0 && C2 >= 0 && C2 % 4 == 0 && (C2 == 0) == (skip == nullptr), "channel counts must be multiples of 4 (C1=%d C2=%d)", C1, C2);
    MVD_REQUIRE(up == 1 || up == 2, "upsampling factor must be 1 or 2, got %d", up);
    MVD_REQUIRE(h * up >= 2 && w * up >= 2, "reflection padding by 1 needs at least 2x2 pixels");
    MVD_REQUIRE(aligned16(z) && aligned16(xp) && aligned16(skip) && aligned16(x3) && aligned16(bias), "pointers must be 16-byte aligned");
    dec::Args a{z, bias, skip, xp, x3, B, h, w, C1, C2, up, act};
    const long long total = static_cast<long long>(B) * (h * up + 2) * (w * up + 2) * ((C1 + C2) / 4);
    dec::prep_fwd_kernel<<<dec::blocks_for(total), 256, 0, as_stream(stream)>>>(a);
    return check_launch("decoder_prep_fwd");
}

int mvd_decoder_prep_bwd(const float* gxp, const float* z, const float* bias, float* gz, float* gskip, float* gbias, int B, int h, int w,
                         int C1, int C2, int up, int act, void* stream) {
    MVD_REQUIRE(gxp && z && gz && B > 0 && h > 0 && w > 0, "bad argument");
    MVD_REQUIRE(C1 > 0 && C1 % 4 == 0 && C2 >= 0 && C2 % 4 == 0 && (C2 == 0) == (gskip == nullptr), "channel counts must be multiples of 4 (C1=%d C2=%d)", C1, C2);
    MVD_REQUIRE(up == 1 || up == 2, "upsampling factor must be 1 or 2, got %d", up);
    const int G = C1 / 4;
    MVD_REQUIRE(gbias == nullptr || (G <= 256 && 256 % G == 0), "bias gradient: C1/4 must divide 256 (C1=%d)", C1);
    MVD_REQUIRE(aligned16(gxp) && aligned16(z) && aligned16(gz) && aligned16(gskip) && aligned16(bias), "pointers must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    if (gbias != nullptr) {
        cudaError_t e = cudaMemsetAsync(gbias, 0, sizeof(float) * C1, st);
        if (e != cudaSuccess) return fail(static_cast<int>(e), "decoder_prep_bwd memset: %s", cudaGetErrorString(e));
    }
    dec::BArgs a{gxp, z, bias, gz, gskip, gbias, B, h, w, C1, C2, up, act};
    dec::prep_bwd_z_kernel<<<dec::blocks_for(static_cast<long long>(B) * h * w * G, 4), 256, 0, st>>>(a);
    if (int rc = check_launch("decoder_prep_bwd(z)")) return rc;
    if (C2 > 0) {
        dec::prep_bwd_skip_kernel<<<dec::blocks_for(static_cast<long long>(B) * h * up * w * up * (C2 / 4)), 256, 0, st>>>(a);
        return check_launch("decoder_prep_bwd(skip)");
    }
    return 0;
}

}
