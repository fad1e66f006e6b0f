// Loss assembly after the photometric kernel (movedepth/trainer.py:687-709 mono, 621-662 multi-frame, 589-609 fused):
//   reproj = min over the source frames of the per-pixel reprojection losses
//   mask   = argmin(cat([reproj, identity + 1e-5 * noise])) == 0        (auto-masking; ones when there is no identity term)
//   loss   = sum(reproj * mask) / (sum(mask) + 1e-7)
// The reference spends ~11 elementwise / reduction launches per scale on this (cat, min, mul, add, le, float, mul, sum,
// sum, add, div) and as many again in the backward; here it is one streaming pass forward (per-pixel select + fp64 block
// sums), a one-thread finalize, and one pass backward that routes the gradient to the selected source.
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {

__global__ void __launch_bounds__(256)
reproj_select_fwd_kernel(const float* __restrict__ l0, const float* __restrict__ l1, const float* __restrict__ ident,
                         const float* __restrict__ noise, float* __restrict__ reproj, unsigned char* __restrict__ sel,
                         double* __restrict__ sums, long long n) {
    float s_loss = 0.f, s_mask = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float a = __ldg(l0 + i);
        unsigned src = 0u;
        if (l1 != nullptr) {
            const float b = __ldg(l1 + i);
            if (b < a) {                       // torch.min keeps the first of equal values
                a = b;
                src = 1u;
            }
        }
        unsigned m = 1u;
        if (ident != nullptr) {                // identity + noise * 1e-5, rounded like the two separate torch ops
            const float t = noise != nullptr ? __fadd_rn(__ldg(ident + i), __fmul_rn(__ldg(noise + i), 1e-5f)) : __ldg(ident + i);
            m = (a <= t) ? 1u : 0u;            // argmin over [reproj, identity]: ties go to index 0
        }
        reproj[i] = a;
        sel[i] = static_cast<unsigned char>(src | (m << 1));
        if (m) {
            s_loss += a;
            s_mask += 1.f;
        }
    }
    s_loss = warp_sum(s_loss);
    s_mask = warp_sum(s_mask);
    __shared__ float red[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = s_loss;
        red[1][warp] = s_mask;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += static_cast<double>(red[threadIdx.x][w]);
        atomicAdd(sums + threadIdx.x, t);
    }
}

__global__ void reproj_select_finalize_kernel(const double* __restrict__ sums, float* __restrict__ loss) {
    loss[0] = static_cast<float>(sums[0] / (sums[1] + 1e-7));
}

__global__ void __launch_bounds__(256)
reproj_select_bwd_kernel(const float* __restrict__ gloss, const double* __restrict__ sums, const unsigned char* __restrict__ sel,
                         float* __restrict__ g0, float* __restrict__ g1, long long n) {
    const float g = gloss[0] * static_cast<float>(1.0 / (sums[1] + 1e-7));
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const unsigned s = sel[i];
        const bool m = (s >> 1) != 0u;
        g0[i] = (m && (s & 1u) == 0u) ? g : 0.f;
        if (g1 != nullptr) g1[i] = (m && (s & 1u) == 1u) ? g : 0.f;
    }
}

static int grid_for(long long n) {
    const long long blocks = (n + 255) / 256, cap = static_cast<long long>(sm_count()) * 8;
    return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace mvd

extern "C" {

int mvd_reproj_select_fwd(const float* l0, const float* l1, const float* ident, const float* noise, float* reproj,
                          unsigned char* sel, double* sums, float* loss, long long n, void* stream) {
    MVD_REQUIRE(l0 && reproj && sel && sums && loss && n > 0, "bad argument");
    cudaStream_t st = mvd::as_stream(stream);
    cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return mvd::fail(static_cast<int>(e), "reproj_select memset: %s", cudaGetErrorString(e));
    mvd::reproj_select_fwd_kernel<<<mvd::grid_for(n), 256, 0, st>>>(l0, l1, ident, noise, reproj, sel, sums, n);
    if (int rc = mvd::check_launch("reproj_select_fwd")) return rc;
    mvd::reproj_select_finalize_kernel<<<1, 1, 0, st>>>(sums, loss);
    return mvd::check_launch("reproj_select_finalize");
}

int mvd_reproj_select_bwd(const float* gloss, const double* sums, const unsigned char* sel, float* g0, float* g1, long long n,
                          void* stream) {
    MVD_REQUIRE(gloss && sums && sel && g0 && n > 0, "bad argument");
    mvd::reproj_select_bwd_kernel<<<mvd::grid_for(n), 256, 0, mvd::as_stream(stream)>>>(gloss, sums, sel, g0, g1, n);
    return mvd::check_launch("reproj_select_bwd");
}

}
