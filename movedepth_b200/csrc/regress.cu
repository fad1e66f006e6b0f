// K3 -- softmax over the depth axis + entropy + local-max depth regression, and K4 -- convex
// upsampling.  Both are small HBM-bound passes; one thread per low-resolution pixel keeps every
// access coalesced along x (the depth/mask-channel stride is h*w).
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {

// ------------------------------------------------------------------------------------------ K3
// movedepth/trainer.py:367 (softmax), layers.py:862-863 (entropy), layers.py:796-812 (localmax)
__global__ void __launch_bounds__(128)
regress_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ inv_a, const float* __restrict__ inv_b,
                   float* __restrict__ prob, float* __restrict__ entropy, float* __restrict__ depth,
                   int* __restrict__ amax, int B, int D, int hw, int radius) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * hw) return;
    const int b = idx / hw, p = idx - b * hw;
    const float* lp = logits + static_cast<size_t>(b) * D * hw + p;
    float m = -INFINITY;
    int im = 0;
    for (int d = 0; d < D; ++d) {
        const float v = __ldg(lp + static_cast<size_t>(d) * hw);
        if (v > m) {
            m = v;
            im = d;
        }
    }
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += expf(__ldg(lp + static_cast<size_t>(d) * hw) - m);
    const float inv_s = 1.f / s;
    float ent = 0.f;
    float* pp = prob ? prob + static_cast<size_t>(b) * D * hw + p : nullptr;
    for (int d = 0; d < D; ++d) {
        const float q = expf(__ldg(lp + static_cast<size_t>(d) * hw) - m) * inv_s;
        if (pp) pp[static_cast<size_t>(d) * hw] = q;
        ent -= q * logf(fminf(fmaxf(q, 1e-9f), 1.f));
    }
    float num = 0.f, den = 1e-6f;
    for (int k = -radius; k <= radius; ++k) {
        const int j = min(max(im + k, 0), D - 1);
        const float q = expf(__ldg(lp + static_cast<size_t>(j) * hw) - m) * inv_s;
        num += static_cast<float>(j) * q;
        den += q;
    }
    const float n = (num / den) / static_cast<float>(D - 1);
    const float a = __ldg(inv_a + idx), bv = __ldg(inv_b + idx);
    entropy[idx] = ent;
    depth[idx] = 1.f / (a + n * (bv - a));
    amax[idx] = im;
}

__global__ void __launch_bounds__(128)
regress_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ inv_a, const float* __restrict__ inv_b,
                   const int* __restrict__ amax, const float* __restrict__ g_entropy,
                   const float* __restrict__ g_depth, float* __restrict__ glogits, int B, int D, int hw, int radius) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * hw) return;
    const int b = idx / hw, p = idx - b * hw;
    const float* lp = logits + static_cast<size_t>(b) * D * hw + p;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(lp + static_cast<size_t>(d) * hw));
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += expf(__ldg(lp + static_cast<size_t>(d) * hw) - m);
    const float inv_s = 1.f / s;

    // local soft-argmax pieces (window indices are constants of the backward, like autograd)
    const int im = amax[idx];
    float num = 0.f, den = 1e-6f;
    for (int k = -radius; k <= radius; ++k) {
        const int j = min(max(im + k, 0), D - 1);
        const float q = expf(__ldg(lp + static_cast<size_t>(j) * hw) - m) * inv_s;
        num += static_cast<float>(j) * q;
        den += q;
    }
    const float soft = num / den;
    const float a = __ldg(inv_a + idx), bv = __ldg(inv_b + idx);
    const float dep = 1.f / (a + (soft / static_cast<float>(D - 1)) * (bv - a));
    const float gd = g_depth ? __ldg(g_depth + idx) : 0.f;
    const float ge = g_entropy ? __ldg(g_entropy + idx) : 0.f;
    const float gsoft = -gd * dep * dep * (bv - a) / static_cast<float>(D - 1);
    const int lo = im - radius, hi = im + radius;

    auto grad_p = [&](int d, float q) {
        float g = 0.f;
        if (ge != 0.f) g = ge * (-logf(fminf(fmaxf(q, 1e-9f), 1.f)) - ((q >= 1e-9f && q <= 1.f) ? 1.f : 0.f));
        // multiplicity of d in the clamped window
        int cnt = 0;
        if (d >= lo && d <= hi) cnt = 1;
        if (d == 0 && lo < 0) cnt = 1 - lo;               // indices lo..0 all clamp to 0
        if (d == D - 1 && hi > D - 1) cnt = hi - (D - 1) + 1;
        if (cnt) g += gsoft * static_cast<float>(cnt) * (static_cast<float>(d) - soft) / den;
        return g;
    };
    float dot = 0.f;
    for (int d = 0; d < D; ++d) {
        const float q = expf(__ldg(lp + static_cast<size_t>(d) * hw) - m) * inv_s;
        dot = fmaf(q, grad_p(d, q), dot);
    }
    float* gp = glogits + static_cast<size_t>(b) * D * hw + p;
    for (int d = 0; d < D; ++d) {
        const float q = expf(__ldg(lp + static_cast<size_t>(d) * hw) - m) * inv_s;
        gp[static_cast<size_t>(d) * hw] = q * (grad_p(d, q) - dot);
    }
}

// ------------------------------------------------------------------------------------------ K4
// movedepth/layers.py:200-214: mask.view(B,9,f,f,h,w) softmax over 9; unfold(depth,3,pad=1).
template <bool BWD>
__global__ void __launch_bounds__(256)
convex_up_kernel(const float* __restrict__ depth, const float* __restrict__ mask, float* __restrict__ out,
                 const float* __restrict__ gout, float* __restrict__ gdepth, float* __restrict__ gmask, int B, int h,
                 int w, int f) {
    const int hw = h * w, ff = f * f;
    const size_t total = static_cast<size_t>(B) * ff * hw;
    const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int p = static_cast<int>(idx % hw);
    const int ij = static_cast<int>((idx / hw) % ff);
    const int b = static_cast<int>(idx / (static_cast<size_t>(hw) * ff));
    const int y = p / w, x = p - y * w;
    const int i = ij / f, j = ij - i * f;
    const float* mp = mask + (static_cast<size_t>(b) * 9 * ff + ij) * hw + p;
    float mk[9], nb[9];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        mk[k] = __ldg(mp + static_cast<size_t>(k) * ff * hw);
        mx = fmaxf(mx, mk[k]);
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        nb[k] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(depth + static_cast<size_t>(b) * hw + yy * w + xx) : 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        mk[k] = expf(mk[k] - mx);
        s += mk[k];
    }
    const float inv = 1.f / s;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        mk[k] *= inv;
        acc = fmaf(mk[k], nb[k], acc);
    }
    const size_t oix = (static_cast<size_t>(b) * h * f + (y * f + i)) * (static_cast<size_t>(w) * f) + (x * f + j);
    if (!BWD) {
        out[oix] = acc;
    } else {
        const float go = __ldg(gout + oix);
        float* gm = gmask + (static_cast<size_t>(b) * 9 * ff + ij) * hw + p;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            gm[static_cast<size_t>(k) * ff * hw] = go * mk[k] * (nb[k] - acc);
            const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
            if (yy >= 0 && yy < h && xx >= 0 && xx < w)
                atomicAdd(gdepth + static_cast<size_t>(b) * hw + yy * w + xx, go * mk[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam (no amsgrad, no weight decay): p -= step_size * m / (sqrt(v)/bias2 + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n4, long long n, float b1, float b2, float eps, float step_size, float bias2, float gscale) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = reinterpret_cast<float*>(&pv);
        const float* ga = reinterpret_cast<const float*>(&gv);
        float* ma = reinterpret_cast<float*>(&mv);
        float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gg = ga[k] * gscale;
            ma[k] = ma[k] + (gg - ma[k]) * (1.f - b1);          // lerp form used by torch
            va[k] = va[k] * b2 + (1.f - b2) * gg * gg;
            const float denom = sqrtf(va[k]) / bias2 + eps;
            pa[k] = pa[k] - step_size * (ma[k] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail (n not a multiple of 4)
    for (long long i = n4 * 4 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const float gg = g[i] * gscale;
        const float mm = m[i] + (gg - m[i]) * (1.f - b1);
        const float vv = v[i] * b2 + (1.f - b2) * gg * gg;
        m[i] = mm;
        v[i] = vv;
        p[i] = p[i] - step_size * (mm / (sqrtf(vv) / bias2 + eps));
    }
}

}  // namespace mvd

extern "C" {

int mvd_regress_fwd(const float* logits, const float* inv_a, const float* inv_b, float* prob, float* entropy,
                    float* depth, int* amax, int B, int D, int hw, int radius, void* stream) {
    MVD_REQUIRE(logits && inv_a && inv_b && entropy && depth && amax, "null pointer argument");
    MVD_REQUIRE(B > 0 && D > 1 && hw > 0 && radius >= 0, "bad shape B=%d D=%d hw=%d radius=%d", B, D, hw, radius);
    const int n = B * hw;
    mvd::regress_fwd_kernel<<<(n + 127) / 128, 128, 0, mvd::as_stream(stream)>>>(logits, inv_a, inv_b, prob, entropy,
                                                                                  depth, amax, B, D, hw, radius);
    return mvd::check_launch("regress_fwd");
}

int mvd_regress_bwd(const float* logits, const float* inv_a, const float* inv_b, const int* amax,
                    const float* g_entropy, const float* g_depth, float* glogits, int B, int D, int hw, int radius,
                    void* stream) {
    MVD_REQUIRE(logits && inv_a && inv_b && amax && glogits, "null pointer argument");
    MVD_REQUIRE(B > 0 && D > 1 && hw > 0 && radius >= 0, "bad shape B=%d D=%d hw=%d radius=%d", B, D, hw, radius);
    const int n = B * hw;
    mvd::regress_bwd_kernel<<<(n + 127) / 128, 128, 0, mvd::as_stream(stream)>>>(logits, inv_a, inv_b, amax, g_entropy,
                                                                                  g_depth, glogits, B, D, hw, radius);
    return mvd::check_launch("regress_bwd");
}

int mvd_convex_up_fwd(const float* depth, const float* mask, float* out, int B, int h, int w, int f, void* stream) {
    MVD_REQUIRE(depth && mask && out, "null pointer argument");
    MVD_REQUIRE(B > 0 && h > 0 && w > 0 && f > 0, "bad shape");
    const size_t total = static_cast<size_t>(B) * f * f * h * w;
    mvd::convex_up_kernel<false><<<static_cast<unsigned>((total + 255) / 256), 256, 0, mvd::as_stream(stream)>>>(
        depth, mask, out, nullptr, nullptr, nullptr, B, h, w, f);
    return mvd::check_launch("convex_up_fwd");
}

int mvd_convex_up_bwd(const float* depth, const float* mask, const float* gout, float* gdepth, float* gmask, int B,
                      int h, int w, int f, void* stream) {
    MVD_REQUIRE(depth && mask && gout && gdepth && gmask, "null pointer argument");
    MVD_REQUIRE(B > 0 && h > 0 && w > 0 && f > 0, "bad shape");
    cudaStream_t st = mvd::as_stream(stream);
    cudaError_t e = cudaMemsetAsync(gdepth, 0, sizeof(float) * B * h * w, st);
    if (e != cudaSuccess) return mvd::fail(static_cast<int>(e), "convex_up_bwd memset: %s", cudaGetErrorString(e));
    const size_t total = static_cast<size_t>(B) * f * f * h * w;
    mvd::convex_up_kernel<true><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(depth, mask, nullptr, gout,
                                                                                            gdepth, gmask, B, h, w, f);
    return mvd::check_launch("convex_up_bwd");
}

int mvd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float beta1,
                  float beta2, float eps, float step_size, float bias2, float grad_scale, void* stream) {
    MVD_REQUIRE(param && grad && exp_avg && exp_avg_sq, "null pointer argument");
    MVD_REQUIRE(n >= 0, "negative length");
    MVD_REQUIRE(mvd::aligned16(param) && mvd::aligned16(grad) && mvd::aligned16(exp_avg) && mvd::aligned16(exp_avg_sq),
                "Adam arenas must be 16-byte aligned");
    if (n == 0) return 0;
    const long long n4 = n / 4;
    const int grid = static_cast<int>(min(static_cast<long long>(mvd::sm_count()) * 8, (n4 + 255) / 256 + 1));
    mvd::adam_kernel<<<grid, 256, 0, mvd::as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n4, n, beta1, beta2, eps,
                                                               step_size, bias2, grad_scale);
    return mvd::check_launch("adam_step");
}

}

// ------------------------------------------------------------------------------------------ 3xTF32 operand split
// out[r, 0:C | C:2C | 2C:3C] = pieces of x[r, 0:C] split as x = hi + lo with hi = x rounded to TF32's
// 10-bit mantissa.  pattern 0: [hi, lo, hi] (activations / gradients), pattern 1: [hi, hi, lo] (weights).
// One pass instead of round/sub/cat; rows are channels-last pixels (C innermost).
namespace mvd {
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, int C, int pattern) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const float v = __ldg(x + i);
        const float hi = tf32_hi(v);
        const float lo = v - hi;
        const long long r = i / C;
        const int c = static_cast<int>(i - r * C);
        float* o = out + r * (3LL * C) + c;
        o[0] = hi;
        o[C] = pattern == 0 ? lo : hi;
        o[2 * C] = pattern == 0 ? hi : lo;
    }
}

// C % 4 == 0: one thread per 4 channels, 16-byte loads and stores (the scalar kernel ran at 3.7 TB/s on the 283 MB volume)
__global__ void __launch_bounds__(256)
split_tf32_v4_kernel(const float4* __restrict__ x, float4* __restrict__ out, long long n4, int C4, int pattern) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(x + i);
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        const long long r = i / C4;
        const int c = static_cast<int>(i - r * C4);
        float4* o = out + r * (3LL * C4) + c;
        o[0] = hi;
        o[C4] = pattern == 0 ? lo : hi;
        o[2 * C4] = pattern == 0 ? hi : lo;
    }
}
}  // namespace mvd

extern "C" int mvd_split_tf32(const float* x, float* out, long long n, int C, int pattern, void* stream) {
    MVD_REQUIRE(x && out, "null pointer argument");
    MVD_REQUIRE(n >= 0 && C > 0 && n % C == 0 && (pattern == 0 || pattern == 1), "bad split arguments n=%lld C=%d", n, C);
    if (n == 0) return 0;
    if (C % 4 == 0 && mvd::aligned16(x) && mvd::aligned16(out)) {
        const long long n4 = n / 4;
        const int grid = static_cast<int>(min(static_cast<long long>(mvd::sm_count()) * 16, (n4 + 255) / 256));
        mvd::split_tf32_v4_kernel<<<grid, 256, 0, mvd::as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out),
                                                                            n4, C / 4, pattern);
        return mvd::check_launch("split_tf32_v4");
    }
    const int grid = static_cast<int>(min(static_cast<long long>(mvd::sm_count()) * 16, (n + 255) / 256));
    mvd::split_tf32_kernel<<<grid, 256, 0, mvd::as_stream(stream)>>>(x, out, n, C, pattern);
    return mvd::check_launch("split_tf32");
}

// ------------------------------------------------------------------------------------------------
// Gradient gather: after back-propagation the per-parameter gradient tensors (whatever layout autograd / cuDNN produced
// them in) are written into the flat gradient arena by ONE launch instead of one accumulate kernel per parameter (232 for
// the ResNet18 model).  table: int64 [nseg][14] = {src pointer (0: no gradient -> zeros), dst offset, numel, linear,
// dims[5] (destination physical order, outermost first, padded with 1), src strides[5] (elements, same order)};
// block_map: int32 [nblocks][2] = {segment, chunk of GATHER_CHUNK elements}.
namespace mvd {
constexpr int GATHER_CHUNK = 4096;

__global__ void __launch_bounds__(256)
gather_segments_kernel(const long long* __restrict__ table, const int* __restrict__ block_map, float* __restrict__ dst) {
    const int seg = block_map[2 * blockIdx.x], chunk = block_map[2 * blockIdx.x + 1];
    const long long* t = table + 14ll * seg;
    const float* src = reinterpret_cast<const float*>(t[0]);
    float* out = dst + t[1];
    const long long n = t[2], lo = static_cast<long long>(chunk) * GATHER_CHUNK;
    const long long hi = lo + GATHER_CHUNK < n ? lo + GATHER_CHUNK : n;
    if (src == nullptr) {
        for (long long i = lo + threadIdx.x; i < hi; i += 256) out[i] = 0.f;
    } else if (t[3]) {
        for (long long i = lo + threadIdx.x; i < hi; i += 256) out[i] = __ldg(src + i);
    } else {
        const long long d1 = t[5], d2 = t[6], d3 = t[7], d4 = t[8];
        const long long s0 = t[9], s1 = t[10], s2 = t[11], s3 = t[12], s4 = t[13];
        for (long long i = lo + threadIdx.x; i < hi; i += 256) {
            long long r = i;
            const long long i4 = r % d4; r /= d4;
            const long long i3 = r % d3; r /= d3;
            const long long i2 = r % d2; r /= d2;
            const long long i1 = r % d1; r /= d1;
            out[i] = __ldg(src + r * s0 + i1 * s1 + i2 * s2 + i3 * s3 + i4 * s4);
        }
    }
}
}  // namespace mvd

extern "C" int mvd_gather_chunk(void) { return mvd::GATHER_CHUNK; }

extern "C" int mvd_gather_segments(const long long* table, const int* block_map, int nblocks, float* dst, void* stream) {
    MVD_REQUIRE(table && block_map && dst && nblocks > 0, "bad argument");
    mvd::gather_segments_kernel<<<nblocks, 256, 0, mvd::as_stream(stream)>>>(table, block_map, dst);
    return mvd::check_launch("gather_segments");
}
