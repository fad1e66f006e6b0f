// Direct 2-D convolutions for the skinny layers of the matching-feature network FPN4 (movedepth/networks/resnet_encoder.py:
// 311-391: conv0 = 3->8->8 at full resolution, conv1 = 8->16 (5x5, stride 2)->16->16 at half resolution) and of UncertNet.
// With 3..16 channels these layers are nowhere near a GEMM: 0.05-0.2 GFLOP per image against 30-70 MB of activations, i.e.
// HBM / FMA-issue bound work for the CUDA cores.  cuDNN spends 70-240 us per layer and pass on them (tensor-core kernels
// padded to 32 channels, a 3x wider channel dimension for the 3xTF32 emulation, legacy weight-gradient engines); these
// kernels are exact fp32 (no operand split needed), read every activation once and run at 15-40 us.
//   layout: activations channels-last [B,H,W,C]; weights [COUT][K][K][CIN] (= channels-last storage of the logical OIHW tensor)
//   forward      y[b,oy,ox,co] = sum_{ky,kx,ci} x[b, oy*S+ky-P, ox*S+kx-P, ci] * w[co,ky,kx,ci],  P = K/2, zero padding
//   dgrad (S=1)  the same kernel on gy with the weights flipped and transposed while they are staged in shared memory
//   dgrad (S=2)  parity-gathered transposed convolution (only taps with ky = iy+P mod 2 reach an input row)
//   wgrad        thread = (tap, ci) pair x pixel group, COUT accumulators in registers, persistent CTAs, per-CTA partials
//                reduced by a second kernel (deterministic)
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {
namespace c2s {

struct Args {
    const float* x;     // input activation  [B,H,W,CIN]
    const float* w;     // weights           [COUT][K][K][CIN]
    float* y;           // output            [B,Ho,Wo,COUT]
    int B, H, W, Ho, Wo;
    int flip;           // 1: stage w as the data-gradient filter (flipped taps, channels transposed; w is then [CIN][K][K][COUT])
    int pad;            // zero padding on every side: K/2 ("same"), 0 ("valid": the DepthDecoder's pre-padded inputs), K-1 (their dgrad)
};

// ------------------------------------------------------------------------------------------------ forward / dgrad (S = 1)
// 128 threads: lane = output column of a 32-wide tile, warp q = PR consecutive output rows.  The input tile (+halo) lives in
// shared memory as CIN planes, the filter as [tap][ci][co] (co fastest: float4 broadcast loads).
template <int CIN, int COUT, int K, int S, int PR>
struct FwdCfg {
    static constexpr int ROWS = 4 * PR;
    static constexpr int IH = (ROWS - 1) * S + K, IW = 31 * S + K, IWP = IW | 1;
    static constexpr int XS = (CIN * IH * IWP + 3) & ~3, WS = K * K * CIN * COUT;      // float4 alignment of the filter behind it
    static constexpr int SMEM = (XS + WS) * 4;
    static constexpr int COL = (PR - 1) * S + K;          // input rows one thread touches per (ci, kx)
};

template <int CIN, int COUT, int K, int S, int PR>
__global__ void __launch_bounds__(128) small_fwd_kernel(const Args a) {
    using CF = FwdCfg<CIN, COUT, K, S, PR>;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;
    float* ws = smem + CF::XS;
    const int tid = threadIdx.x, lane = tid & 31, q = tid >> 5;
    const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * CF::ROWS, b = blockIdx.z;
    const int P = a.pad;
    // filter -> ws[tap][ci][co]
    for (int e = tid; e < CF::WS; e += 128) {
        const int co = e % COUT, ci = (e / COUT) % CIN, tap = e / (COUT * CIN);
        float v;
        if (a.flip) {        // data gradient: w is [CIN][K][K][COUT] (the forward layer's [co'][ky][kx][ci'] with co' = ci, ci' = co)
            v = __ldg(a.w + ((static_cast<size_t>(ci) * K * K) + (K * K - 1 - tap)) * COUT + co);
        } else {
            v = __ldg(a.w + ((static_cast<size_t>(co) * K * K) + tap) * CIN + ci);
        }
        ws[e] = v;
    }
    // input tile -> xs[ci][r][c], zero padded
    const int iy0 = oy0 * S - P, ix0 = ox0 * S - P;
    if (CIN % 4 == 0) {                // 16-byte global loads, scattered to the channel planes
        constexpr int C4 = CIN / 4 > 0 ? CIN / 4 : 1;
        for (int e = tid; e < CF::IH * CF::IW * C4; e += 128) {
            const int c4 = e % C4, p = e / C4, c = p % CF::IW, r = p / CF::IW;
            const int gy = iy0 + r, gx = ix0 + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W)
                v = __ldg(reinterpret_cast<const float4*>(a.x + ((static_cast<size_t>(b) * a.H + gy) * a.W + gx) * CIN) + c4);
            float* d = xs + ((4 * c4) * CF::IH + r) * CF::IWP + c;
            d[0] = v.x;
            d[CF::IH * CF::IWP] = v.y;
            d[2 * CF::IH * CF::IWP] = v.z;
            d[3 * CF::IH * CF::IWP] = v.w;
        }
    } else {
        for (int e = tid; e < CF::IH * CF::IW * CIN; e += 128) {
            const int ci = e % CIN, p = e / CIN, c = p % CF::IW, r = p / CF::IW;
            const int gy = iy0 + r, gx = ix0 + c;
            float v = 0.f;
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v = __ldg(a.x + ((static_cast<size_t>(b) * a.H + gy) * a.W + gx) * CIN + ci);
            xs[(ci * CF::IH + r) * CF::IWP + c] = v;
        }
    }
    __syncthreads();
    uint64_t acc2[PR][COUT / 2];                 // packed output-channel pairs: one FFMA2 = two of the FMAs (Blackwell fp32x2)
#pragma unroll
    for (int p = 0; p < PR; ++p)
#pragma unroll
        for (int co = 0; co < COUT / 2; ++co) acc2[p][co] = 0ull;
    const float* xt = xs + (q * PR * S) * CF::IWP + lane * S;
#pragma unroll 1
    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            float xv[CF::COL];
#pragma unroll
            for (int j = 0; j < CF::COL; ++j) xv[j] = xt[(ci * CF::IH + j) * CF::IWP + kx];
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const float* wp = ws + ((ky * K + kx) * CIN + ci) * COUT;
#pragma unroll
                for (int c4 = 0; c4 < COUT; c4 += 4) {
                    const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(wp + c4);
#pragma unroll
                    for (int p = 0; p < PR; ++p) {
                        const uint64_t v = pk2(xv[p * S + ky], xv[p * S + ky]);
                        acc2[p][c4 / 2] = fma2(v, w4.x, acc2[p][c4 / 2]);
                        acc2[p][c4 / 2 + 1] = fma2(v, w4.y, acc2[p][c4 / 2 + 1]);
                    }
                }
            }
        }
    }
    const int ox = ox0 + lane;
    if (ox < a.Wo) {
#pragma unroll
        for (int p = 0; p < PR; ++p) {
            const int oy = oy0 + q * PR + p;
            if (oy < a.Ho) {
                float4* yp = reinterpret_cast<float4*>(a.y + ((static_cast<size_t>(b) * a.Ho + oy) * a.Wo + ox) * COUT);
#pragma unroll
                for (int c4 = 0; c4 < COUT; c4 += 4)
                    reinterpret_cast<ulonglong2*>(yp)[c4 >> 2] = make_ulonglong2(acc2[p][c4 / 2], acc2[p][c4 / 2 + 1]);
            }
        }
    }
}

// COUT = 3 would break the float4 paths; the 3-channel image is only ever an INPUT (CIN = 3), which the planes handle.

// ------------------------------------------------------------------------------------------------ dgrad, stride 2
// gx[b,iy,ix,ci] = sum over co and the taps with ky = (iy + P) mod 2 (+2, +4), kx likewise, of gy[b,(iy+P-ky)/2,(ix+P-kx)/2,co]
// * w[co,ky,kx,ci].  One warp per input row (uniform row parity), a lane computes the two pixels 2*lane, 2*lane+1 of a
// 64-wide tile one after the other, so the tap set (hence the broadcast filter address) is warp-uniform.
template <int CIN, int COUT, int K>
struct Dg2Cfg {
    static constexpr int ROWS = 8;                            // input rows per block (4 warps x 2)
    static constexpr int GH = ROWS / 2 + K / 2 + 1, GW = 32 + K / 2 + 1, GWP = GW | 1;
    static constexpr int GS = (COUT * GH * GWP + 3) & ~3, WS = K * K * COUT * CIN;
    static constexpr int SMEM = (GS + WS) * 4;
};

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(128) small_dgrad_s2_kernel(const Args a) {   // a.x = gy [B,Ho,Wo,COUT], a.y = gx [B,H,W,CIN]
    using CF = Dg2Cfg<CIN, COUT, K>;
    extern __shared__ __align__(16) float smem[];
    float* gs = smem;
    float* ws = smem + CF::GS;                                // [tap][co][ci], ci fastest
    const int tid = threadIdx.x, lane = tid & 31, q = tid >> 5;
    const int ix0 = blockIdx.x * 64, iy0 = blockIdx.y * CF::ROWS, b = blockIdx.z;
    constexpr int P = K / 2;
    for (int e = tid; e < CF::WS; e += 128) {
        const int ci = e % CIN, co = (e / CIN) % COUT, tap = e / (CIN * COUT);
        ws[e] = __ldg(a.w + (static_cast<size_t>(co) * K * K + tap) * CIN + ci);
    }
    // gy tile: rows oy in [(iy0 + P - (K-1)) / 2 floor, ...]: start at floor((iy0 - P) / 2) (iy0 is even)
    const int oy_base = (iy0 - P) / 2 - ((iy0 - P) < 0 && ((iy0 - P) & 1) ? 1 : 0);
    const int ox_base = (ix0 - P) / 2 - ((ix0 - P) < 0 && ((ix0 - P) & 1) ? 1 : 0);
    for (int e = tid; e < CF::GH * CF::GW * (COUT / 4); e += 128) {
        const int c4 = e % (COUT / 4), p = e / (COUT / 4), c = p % CF::GW, r = p / CF::GW;
        const int oy = oy_base + r, ox = ox_base + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (oy >= 0 && oy < a.Ho && ox >= 0 && ox < a.Wo)
            v = __ldg(reinterpret_cast<const float4*>(a.x + ((static_cast<size_t>(b) * a.Ho + oy) * a.Wo + ox) * COUT) + c4);
        float* d = gs + ((4 * c4) * CF::GH + r) * CF::GWP + c;
        d[0] = v.x;
        d[CF::GH * CF::GWP] = v.y;
        d[2 * CF::GH * CF::GWP] = v.z;
        d[3 * CF::GH * CF::GWP] = v.w;
    }
    __syncthreads();
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
        const int iy = iy0 + q * 2 + rr;
#pragma unroll 1
        for (int px = 0; px < 2; ++px) {
            const int ix = ix0 + 2 * lane + px;
            uint64_t acc2[CIN / 2];
#pragma unroll
            for (int ci = 0; ci < CIN / 2; ++ci) acc2[ci] = 0ull;
            for (int ky = (iy + P) & 1; ky < K; ky += 2) {
                const int r = (iy + P - ky) / 2 - oy_base;           // iy + P - ky is even and may be negative only outside the tile
                if ((iy + P - ky) < 0) continue;
                for (int kx = (ix + P) & 1; kx < K; kx += 2) {
                    if ((ix + P - kx) < 0) continue;
                    const int c = (ix + P - kx) / 2 - ox_base;
                    const float* wp = ws + (ky * K + kx) * COUT * CIN;
#pragma unroll 4
                    for (int co = 0; co < COUT; ++co) {
                        const float g1 = gs[(co * CF::GH + r) * CF::GWP + c];
                        const uint64_t g = pk2(g1, g1);
#pragma unroll
                        for (int c4 = 0; c4 < CIN; c4 += 4) {
                            const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(wp + co * CIN + c4);
                            acc2[c4 / 2] = fma2(g, w4.x, acc2[c4 / 2]);
                            acc2[c4 / 2 + 1] = fma2(g, w4.y, acc2[c4 / 2 + 1]);
                        }
                    }
                }
            }
            if (iy < a.H && ix < a.W) {
                float4* op = reinterpret_cast<float4*>(a.y + ((static_cast<size_t>(b) * a.H + iy) * a.W + ix) * CIN);
#pragma unroll
                for (int c4 = 0; c4 < CIN; c4 += 4) reinterpret_cast<ulonglong2*>(op)[c4 >> 2] = make_ulonglong2(acc2[c4 / 2], acc2[c4 / 2 + 1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// gw[co,ky,kx,ci] = sum_{b,oy,ox} gy[b,oy,ox,co] * x[b, oy*S+ky-P, ox*S+kx-P, ci].
// Thread = ((ky,kx,ci) pair, pixel group): COUT accumulators in registers; per output pixel one x load (own address) and
// COUT/4 broadcast float4 loads of gy.  Persistent CTAs walk the 8x32 output tiles; per-CTA partial sums go to `part`.
template <int CIN, int COUT, int K, int S>
struct WgCfg {
    static constexpr int NT = K * K * CIN;                                    // (tap, ci) pairs
    static constexpr int PG = (NT >= 256) ? 1 : (256 + NT - 1) / NT;          // pixel groups
    static constexpr int THREADS = NT * PG;
    static constexpr int TH = 8, TW = 32;
    static constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
    static constexpr int XN = IH * IW * CIN, XS = (XN + 3) & ~3, GS = TH * TW * COUT;
    static constexpr int SMEM = (XS + GS) * 4;
    static constexpr int NW = K * K * CIN * COUT;
};

struct WArgs {
    const float* x;     // [B,H,W,CIN]
    const float* gy;    // [B,Ho,Wo,COUT]
    float* part;        // [grid][NW], layout [co][tap][ci]
    int B, H, W, Ho, Wo, tiles_x, tiles_y, num_tiles;
    int pad;
};

template <int CIN, int COUT, int K, int S>
__global__ void __launch_bounds__(WgCfg<CIN, COUT, K, S>::THREADS) small_wgrad_kernel(const WArgs a) {
    using CF = WgCfg<CIN, COUT, K, S>;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                  // [r][c][ci]
    float* gs = smem + CF::XS;         // [pixel][co]
    const int tid = threadIdx.x;
    const int pair = tid % CF::NT, grp = tid / CF::NT;
    const int ci = pair % CIN, tap = pair / CIN, ky = tap / K, kx = tap % K;
    const int P = a.pad;
    uint64_t acc2[COUT / 2];                     // packed output-channel pairs
#pragma unroll
    for (int co = 0; co < COUT / 2; ++co) acc2[co] = 0ull;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, b = tile / (a.tiles_x * a.tiles_y);
        const int ox0 = tx * CF::TW, oy0 = ty * CF::TH;
        const int iy0 = oy0 * S - P, ix0 = ox0 * S - P;
        __syncthreads();               // previous tile fully consumed
        if (CIN % 4 == 0) {
            constexpr int C4 = CIN / 4 > 0 ? CIN / 4 : 1;
            for (int e = tid; e < CF::IH * CF::IW * C4; e += CF::THREADS) {
                const int c4 = e % C4, p = e / C4, c = p % CF::IW, r = p / CF::IW;
                const int gy_ = iy0 + r, gx_ = ix0 + c;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gy_ >= 0 && gy_ < a.H && gx_ >= 0 && gx_ < a.W)
                    v = __ldg(reinterpret_cast<const float4*>(a.x + ((static_cast<size_t>(b) * a.H + gy_) * a.W + gx_) * CIN) + c4);
                reinterpret_cast<float4*>(xs)[e] = v;
            }
        } else {
            for (int e = tid; e < CF::XN; e += CF::THREADS) {
                const int c_ = e % CIN, p = e / CIN, c = p % CF::IW, r = p / CF::IW;
                const int gy_ = iy0 + r, gx_ = ix0 + c;
                float v = 0.f;
                if (gy_ >= 0 && gy_ < a.H && gx_ >= 0 && gx_ < a.W) v = __ldg(a.x + ((static_cast<size_t>(b) * a.H + gy_) * a.W + gx_) * CIN + c_);
                xs[e] = v;
            }
        }
        for (int e = tid; e < CF::TH * CF::TW * (COUT / 4); e += CF::THREADS) {
            const int c4 = e % (COUT / 4), p = e / (COUT / 4), c = p % CF::TW, r = p / CF::TW;
            const int oy = oy0 + r, ox = ox0 + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);                        // pixels outside the output contribute nothing
            if (oy < a.Ho && ox < a.Wo) v = __ldg(reinterpret_cast<const float4*>(a.gy + ((static_cast<size_t>(b) * a.Ho + oy) * a.Wo + ox) * COUT) + c4);
            reinterpret_cast<float4*>(gs)[e] = v;
        }
        __syncthreads();
        const float* xb = xs + (ky * CF::IW + kx) * CIN + ci;
#pragma unroll 2
        for (int p = grp; p < CF::TH * CF::TW; p += CF::PG) {
            const int r = p / CF::TW, c = p % CF::TW;                           // TW = 32: shift / mask
            const float v = xb[((r * S) * CF::IW + c * S) * CIN];
            const uint64_t vv = pk2(v, v);
            const float* gp = gs + p * COUT;
#pragma unroll
            for (int c4 = 0; c4 < COUT; c4 += 4) {
                const ulonglong2 g4 = *reinterpret_cast<const ulonglong2*>(gp + c4);
                acc2[c4 / 2] = fma2(vv, g4.x, acc2[c4 / 2]);
                acc2[c4 / 2 + 1] = fma2(vv, g4.y, acc2[c4 / 2 + 1]);
            }
        }
    }
    // reduce the pixel groups through shared memory, then one partial per CTA
    __syncthreads();
    float* red = smem;                 // [PG][NT][COUT] <= THREADS * COUT floats; fits: checked by the host
    for (int co = 0; co < COUT / 2; ++co) unpk2(acc2[co], red[(grp * CF::NT + pair) * COUT + 2 * co], red[(grp * CF::NT + pair) * COUT + 2 * co + 1]);
    __syncthreads();
    for (int e = tid; e < CF::NT * COUT; e += CF::THREADS) {
        const int co = e % COUT, pr = e / COUT;
        float s = 0.f;
        for (int g = 0; g < CF::PG; ++g) s += red[(g * CF::NT + pr) * COUT + co];
        const int t_ = pr / CIN, c_ = pr % CIN;
        a.part[static_cast<size_t>(blockIdx.x) * CF::NW + (static_cast<size_t>(co) * K * K + t_) * CIN + c_] = s;
    }
}

// one warp per weight element: lanes stride over the per-CTA partials, shuffle reduction (fixed order: deterministic)
__global__ void __launch_bounds__(256) small_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw, int nw, int ctas) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= nw) return;
    float s = 0.f;
    for (int c = lane; c < ctas; c += 32) s += __ldg(part + static_cast<size_t>(c) * nw + i);
    s = warp_sum(s);
    if (lane == 0) gw[i] = s;
}

// ------------------------------------------------------------------------------------------------ host dispatch
template <int CIN, int COUT, int K, int S, int PR>
static int launch_fwd(const Args& a, cudaStream_t st) {
    using CF = FwdCfg<CIN, COUT, K, S, PR>;
    cudaFuncSetAttribute(small_fwd_kernel<CIN, COUT, K, S, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    dim3 grid((a.Wo + 31) / 32, (a.Ho + CF::ROWS - 1) / CF::ROWS, a.B);
    small_fwd_kernel<CIN, COUT, K, S, PR><<<grid, 128, CF::SMEM, st>>>(a);
    return check_launch("conv2d_small_fwd");
}

template <int CIN, int COUT, int K, int S>
static int wgrad_ctas(int B, int Ho, int Wo) {
    using CF = WgCfg<CIN, COUT, K, S>;
    const int tiles = B * ((Ho + CF::TH - 1) / CF::TH) * ((Wo + CF::TW - 1) / CF::TW);
    const int cap = sm_count() * 3;
    return tiles < cap ? tiles : cap;
}

template <int CIN, int COUT, int K, int S>
static int launch_wgrad(const float* x, const float* gy, float* gw, float* ws, long long ws_bytes, int B, int H, int W, int Ho, int Wo,
                        int pad, cudaStream_t st) {
    using CF = WgCfg<CIN, COUT, K, S>;
    static_assert(CF::THREADS * COUT * 4 <= CF::SMEM || CF::THREADS * COUT <= CF::XS + CF::GS, "group reduction scratch fits");
    WArgs a{x, gy, ws, B, H, W, Ho, Wo, (Wo + CF::TW - 1) / CF::TW, (Ho + CF::TH - 1) / CF::TH, 0, pad};
    a.num_tiles = a.tiles_x * a.tiles_y * B;
    const int ctas = wgrad_ctas<CIN, COUT, K, S>(B, Ho, Wo);
    MVD_REQUIRE(ws_bytes >= static_cast<long long>(ctas) * CF::NW * 4, "conv2d_small wgrad workspace too small");
    const int smem = (CF::SMEM > CF::THREADS * COUT * 4) ? CF::SMEM : CF::THREADS * COUT * 4;
    cudaFuncSetAttribute(small_wgrad_kernel<CIN, COUT, K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    small_wgrad_kernel<CIN, COUT, K, S><<<ctas, CF::THREADS, smem, st>>>(a);
    if (int rc = check_launch("conv2d_small_wgrad")) return rc;
    small_wgrad_reduce_kernel<<<(CF::NW + 7) / 8, 256, 0, st>>>(ws, gw, CF::NW, ctas);
    return check_launch("conv2d_small_wgrad_reduce");
}

}  // namespace c2s
}  // namespace mvd

using namespace mvd;

// supported (CIN, COUT, K, S): the skinny FPN4 / UncertNet layers and the DepthDecoder's finest stage (16 -> 16, and its
// 16 -> 1 disparity head zero-padded to 4 output channels by the caller)
#define C2S_DISPATCH(M)                       \
    M(3, 8, 3, 1)                             \
    M(8, 8, 3, 1)                             \
    M(8, 16, 5, 2)                            \
    M(16, 16, 3, 1)                           \
    M(16, 4, 3, 1)

extern "C" {

// pad: K/2 ("same") for every supported layer; 0 ("valid", pre-padded input) for the stride-1 layers
static int c2s_check(int cin, int cout, int k, int stride, int pad) {
    MVD_REQUIRE(mvd_conv2d_small_supported(cin, cout, k, stride), "conv2d_small: unsupported layer %d->%d k%d s%d", cin, cout, k, stride);
    MVD_REQUIRE(pad == k / 2 || (pad == 0 && stride == 1), "conv2d_small: padding %d not supported for k%d s%d", pad, k, stride);
    return 0;
}

int mvd_conv2d_small_supported(int cin, int cout, int k, int stride) {
#define M(CI, CO, KK, SS) if (cin == CI && cout == CO && k == KK && stride == SS) return 1;
    C2S_DISPATCH(M)
#undef M
    return 0;
}

int mvd_conv2d_small_fwd(const float* x, const float* w, float* y, int B, int H, int W, int cin, int cout, int k, int stride, int pad,
                         void* stream) {
    MVD_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0, "bad argument");
    MVD_REQUIRE(aligned16(y), "output must be 16-byte aligned");
    if (int rc = c2s_check(cin, cout, k, stride, pad)) return rc;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    MVD_REQUIRE(Ho > 0 && Wo > 0, "conv2d_small: empty output");
    c2s::Args a{x, w, y, B, H, W, Ho, Wo, 0, pad};
    cudaStream_t st = as_stream(stream);
#define M(CI, CO, KK, SS) if (cin == CI && cout == CO && k == KK && stride == SS) return c2s::launch_fwd<CI, CO, KK, SS, (SS == 1 ? 4 : 2)>(a, st);
    C2S_DISPATCH(M)
#undef M
    return fail(-1, "conv2d_small: unsupported layer %d->%d k%d s%d", cin, cout, k, stride);
}

// gx = d loss / d x given gy = d loss / d y; w as for the forward ([cout][k][k][cin]); H, W = the INPUT's size
int mvd_conv2d_small_dgrad(const float* gy, const float* w, float* gx, int B, int H, int W, int cin, int cout, int k, int stride, int pad,
                           void* stream) {
    MVD_REQUIRE(gy && w && gx && B > 0 && H > 0 && W > 0, "bad argument");
    MVD_REQUIRE(aligned16(gx), "output must be 16-byte aligned");
    if (int rc = c2s_check(cin, cout, k, stride, pad)) return rc;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    cudaStream_t st = as_stream(stream);
    if (stride == 1) {          // a stride-1 convolution of gy (padding k-1-pad) with the flipped, transposed filter: forward kernel, roles swapped
        c2s::Args a{gy, w, gx, B, Ho, Wo, H, W, 1, k - 1 - pad};
        if (cin == 8 && cout == 8 && k == 3) return c2s::launch_fwd<8, 8, 3, 1, 4>(a, st);
        if (cin == 16 && cout == 16 && k == 3) return c2s::launch_fwd<16, 16, 3, 1, 4>(a, st);
        if (cin == 16 && cout == 4 && k == 3) return c2s::launch_fwd<4, 16, 3, 1, 4>(a, st);
    } else if (cin == 8 && cout == 16 && k == 5 && stride == 2) {
        using CF = c2s::Dg2Cfg<8, 16, 5>;
        c2s::Args a{gy, w, gx, B, H, W, Ho, Wo, 0, pad};
        cudaFuncSetAttribute(c2s::small_dgrad_s2_kernel<8, 16, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
        dim3 grid((W + 63) / 64, (H + CF::ROWS - 1) / CF::ROWS, B);
        c2s::small_dgrad_s2_kernel<8, 16, 5><<<grid, 128, CF::SMEM, st>>>(a);
        return check_launch("conv2d_small_dgrad_s2");
    }
    return fail(-1, "conv2d_small dgrad: unsupported layer %d->%d k%d s%d", cin, cout, k, stride);
}

long long mvd_conv2d_small_wgrad_workspace_bytes(int B, int H, int W, int cin, int cout, int k, int stride, int pad) {
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
#define M(CI, CO, KK, SS) \
    if (cin == CI && cout == CO && k == KK && stride == SS) \
        return static_cast<long long>(c2s::wgrad_ctas<CI, CO, KK, SS>(B, Ho, Wo)) * c2s::WgCfg<CI, CO, KK, SS>::NW * 4;
    C2S_DISPATCH(M)
#undef M
    return 0;
}

int mvd_conv2d_small_wgrad(const float* x, const float* gy, float* gw, void* workspace, long long workspace_bytes, int B, int H, int W,
                           int cin, int cout, int k, int stride, int pad, void* stream) {
    MVD_REQUIRE(x && gy && gw && workspace && B > 0 && H > 0 && W > 0, "bad argument");
    if (int rc = c2s_check(cin, cout, k, stride, pad)) return rc;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    cudaStream_t st = as_stream(stream);
#define M(CI, CO, KK, SS) \
    if (cin == CI && cout == CO && k == KK && stride == SS) \
        return c2s::launch_wgrad<CI, CO, KK, SS>(x, gy, gw, static_cast<float*>(workspace), workspace_bytes, B, H, W, Ho, Wo, pad, st);
    C2S_DISPATCH(M)
#undef M
    return fail(-1, "conv2d_small wgrad: unsupported layer %d->%d k%d s%d", cin, cout, k, stride);
}

}
