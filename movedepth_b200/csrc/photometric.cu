// K5 + K6 -- photometric reprojection loss of one source frame, forward and backward (sm_100a).
//
// Reference chain per (scale, source frame): BackprojectDepth -> Project3D -> F.grid_sample(border,
// bilinear, align_corners=True) (movedepth/trainer.py:519-529, layers.py:556-621), then SSIM over
// reflection-padded 3x3 windows (layers.py:646-677) mixed with L1 (trainer.py:535-550): ~45 ATen
// launches and ~25 full-resolution temporaries per call.
//
// Here: one CTA per 32x8 pixel tile.  Phase 1 warps the source image for the tile plus its 1-pixel
// SSIM halo (reflection handled by index mapping) into shared memory next to the target tile;
// phase 2 is the 3x3 stencil out of shared memory.  HBM traffic: depth + 3-channel source taps +
// target in, loss (+ warped, kept for the backward) out.
//
// Backward: the SSIM adjoint is separable per window: d s_p / d x_q = a_p + b_p x_q + c_p y_q, so
// the kernel builds the (a,b,c) coefficient maps for tile+halo in shared memory, gathers them per
// pixel (reflection adjoint included), then differentiates the bilinear warp and the projection to
// get d loss/d depth per pixel and d loss/d (K T)[3x4] per batch item (block-reduced, 12 atomics).
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {

constexpr int PH_TW = 32, PH_TH = 8, PH_THREADS = PH_TW * PH_TH;
constexpr float SSIM_C1 = 0.0001f, SSIM_C2 = 0.0009f;

struct PhGeo {   // shared per block
    float P[12];
    float iK[9];
};

__device__ __forceinline__ int reflect_idx(int i, int n) {   // ReflectionPad2d(1) index map, valid for i in [-2, n+1]
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__device__ __forceinline__ void load_geo(PhGeo* geo, const float* K, const float* invK, const float* T, int b) {
    const int tid = threadIdx.x;
    if (tid < 12) {
        const int i = tid >> 2, j = tid & 3;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s = fmaf(K[b * 16 + i * 4 + k], T[b * 16 + k * 4 + j], s);
        geo->P[tid] = s;
    } else if (tid < 21) {
        const int q = tid - 12;
        geo->iK[q] = invK[b * 16 + (q / 3) * 4 + (q % 3)];
    }
}


// Window statistics -> SSIM numerator / denominator factors.  Products are rounded explicitly (no FMA
// contraction asymmetry) so that x == y gives n == d bit for bit, like the reference's pooled form.
struct SsimTerms {
    float mx, my, A1, A2, B1, B2;
};
__device__ __forceinline__ SsimTerms ssim_terms(float sx, float sy, float sxx, float syy, float sxy) {
    const float k9 = 1.f / 9.f;
    SsimTerms t;
    t.mx = __fmul_rn(sx, k9);
    t.my = __fmul_rn(sy, k9);
    const float mxx = __fmul_rn(t.mx, t.mx), myy = __fmul_rn(t.my, t.my), mxy = __fmul_rn(t.mx, t.my);
    const float vx = __fsub_rn(__fmul_rn(sxx, k9), mxx), vy = __fsub_rn(__fmul_rn(syy, k9), myy);
    const float cxy = __fsub_rn(__fmul_rn(sxy, k9), mxy);
    t.A1 = __fadd_rn(__fmul_rn(2.f, mxy), SSIM_C1);
    t.A2 = __fadd_rn(__fmul_rn(2.f, cxy), SSIM_C2);
    t.B1 = __fadd_rn(__fadd_rn(mxx, myy), SSIM_C1);
    t.B2 = __fadd_rn(__fadd_rn(vx, vy), SSIM_C2);
    return t;
}

struct WarpPt {
    float u, v;            // clamped source coordinates
    float X[3];            // camera point depth * ray
    float r[3];            // ray = inv_K[:3,:3] (x, y, 1)
    float inv_z;           // 1 / (z + eps)
    float ur, vr;          // unclamped coordinates
    bool gu, gv;           // gradient passes the border clamp
};

__device__ __forceinline__ WarpPt warp_point(const PhGeo& g, float depth, int x, int y, int W, int H) {
    WarpPt w;
    const float xf = static_cast<float>(x), yf = static_cast<float>(y);
    const float rx = fmaf(g.iK[0], xf, fmaf(g.iK[1], yf, g.iK[2]));
    const float ry = fmaf(g.iK[3], xf, fmaf(g.iK[4], yf, g.iK[5]));
    const float rz = fmaf(g.iK[6], xf, fmaf(g.iK[7], yf, g.iK[8]));
    w.r[0] = rx;
    w.r[1] = ry;
    w.r[2] = rz;
    w.X[0] = depth * rx;
    w.X[1] = depth * ry;
    w.X[2] = depth * rz;
    const float px = fmaf(g.P[0], w.X[0], fmaf(g.P[1], w.X[1], fmaf(g.P[2], w.X[2], g.P[3])));
    const float py = fmaf(g.P[4], w.X[0], fmaf(g.P[5], w.X[1], fmaf(g.P[6], w.X[2], g.P[7])));
    const float pz = fmaf(g.P[8], w.X[0], fmaf(g.P[9], w.X[1], fmaf(g.P[10], w.X[2], g.P[11]))) + 1e-7f;
    w.inv_z = __frcp_rn(pz);
    w.ur = px * w.inv_z;
    w.vr = py * w.inv_z;
    const float mu = static_cast<float>(W - 1), mv = static_cast<float>(H - 1);
    // ATen clip_coordinates: min(max, max(in, 0)); gradient only strictly inside (0, max)
    w.u = fminf(mu, fmaxf(w.ur, 0.f));
    w.v = fminf(mv, fmaxf(w.vr, 0.f));
    w.gu = (w.ur > 0.f) && (w.ur < mu);
    w.gv = (w.vr > 0.f) && (w.vr < mv);
    return w;
}

struct Taps {
    int o00, o01, o10, o11;
    float fx0, fx1, fy0, fy1;   // fx0 = x1 - u, fx1 = u - x0 (ATen grid_sampler weight factors)
    bool x1ok, y1ok;
};
__device__ __forceinline__ Taps taps_at(float u, float v, int W, int H) {
    Taps t;
    const float x0 = floorf(u), y0 = floorf(v);
    t.fx1 = u - x0;
    t.fx0 = (x0 + 1.f) - u;
    t.fy1 = v - y0;
    t.fy0 = (y0 + 1.f) - v;
    const int ix = static_cast<int>(x0), iy = static_cast<int>(y0);
    t.x1ok = ix + 1 < W;
    t.y1ok = iy + 1 < H;
    t.o00 = iy * W + ix;
    t.o01 = t.o00 + (t.x1ok ? 1 : 0);
    t.o10 = t.o00 + (t.y1ok ? W : 0);
    t.o11 = t.o10 + (t.x1ok ? 1 : 0);
    return t;
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(PH_THREADS)
photometric_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ src, const float* __restrict__ tgt,
                       const float* __restrict__ K, const float* __restrict__ invK, const float* __restrict__ T,
                       float* __restrict__ warped, float* __restrict__ loss, int H, int W, float ssim_w, int identity) {
    constexpr int SW = PH_TW + 2, SH = PH_TH + 2;
    __shared__ float xs[3][SH][SW + 1];
    __shared__ float ys[3][SH][SW + 1];
    __shared__ PhGeo geo;
    const int b = blockIdx.z, tx0 = blockIdx.x * PH_TW, ty0 = blockIdx.y * PH_TH;
    const int tid = threadIdx.x;
    const size_t hw = static_cast<size_t>(H) * W;
    if (!identity) load_geo(&geo, K, invK, T, b);
    __syncthreads();
    const float* sb = src + static_cast<size_t>(b) * 3 * hw;
    const float* tb = tgt + static_cast<size_t>(b) * 3 * hw;
    const bool need_halo = ssim_w != 0.f;
    for (int s = tid; s < SW * SH; s += PH_THREADS) {
        const int sy = s / SW, sx = s - sy * SW;
        const bool center = sy >= 1 && sy <= PH_TH && sx >= 1 && sx <= PH_TW;
        if (!need_halo && !center) continue;
        int y = ty0 + sy - 1, x = tx0 + sx - 1;
        if (y > H || x > W) {       // beyond the padded image: never read by a valid pixel
            continue;
        }
        y = reflect_idx(y, H);
        x = reflect_idx(x, W);
        const size_t pix = static_cast<size_t>(y) * W + x;
        float v0, v1, v2;
        if (identity) {
            v0 = __ldg(sb + pix);
            v1 = __ldg(sb + hw + pix);
            v2 = __ldg(sb + 2 * hw + pix);
        } else {
            const float d = __ldg(depth + b * hw + pix);
            const WarpPt w = warp_point(geo, d, x, y, W, H);
            const Taps t = taps_at(w.u, w.v, W, H);
            const float w00 = t.fx0 * t.fy0, w01 = t.fx1 * t.fy0, w10 = t.fx0 * t.fy1, w11 = t.fx1 * t.fy1;
            float o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* sc = sb + c * hw;
                // out-of-range neighbours (only at u == W-1 / v == H-1) carry weight 0 and are skipped by ATen
                const float a00 = __ldg(sc + t.o00);
                const float a01 = t.x1ok ? __ldg(sc + t.o01) : 0.f;
                const float a10 = t.y1ok ? __ldg(sc + t.o10) : 0.f;
                const float a11 = (t.x1ok && t.y1ok) ? __ldg(sc + t.o11) : 0.f;
                o[c] = fmaf(a11, w11, fmaf(a10, w10, fmaf(a01, w01, a00 * w00)));
            }
            v0 = o[0];
            v1 = o[1];
            v2 = o[2];
            if (center && warped != nullptr && (ty0 + sy - 1) < H && (tx0 + sx - 1) < W) {
                float* wp = warped + static_cast<size_t>(b) * 3 * hw + pix;
                wp[0] = v0;
                wp[hw] = v1;
                wp[2 * hw] = v2;
            }
        }
        xs[0][sy][sx] = v0;
        xs[1][sy][sx] = v1;
        xs[2][sy][sx] = v2;
        ys[0][sy][sx] = __ldg(tb + pix);
        ys[1][sy][sx] = __ldg(tb + hw + pix);
        ys[2][sy][sx] = __ldg(tb + 2 * hw + pix);
    }
    __syncthreads();
    const int lx = tid % PH_TW, ly = tid / PH_TW;
    const int x = tx0 + lx, y = ty0 + ly;
    if (x >= W || y >= H) return;
    float l1 = 0.f, ss = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        l1 += fabsf(ys[c][ly + 1][lx + 1] - xs[c][ly + 1][lx + 1]);
        if (need_halo) {
            float sx_ = 0.f, sy_ = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float a = xs[c][ly + dy][lx + dx], bb = ys[c][ly + dy][lx + dx];
                    sx_ += a;
                    sy_ += bb;
                    sxx = fmaf(a, a, sxx);
                    syy = fmaf(bb, bb, syy);
                    sxy = fmaf(a, bb, sxy);
                }
            const SsimTerms st = ssim_terms(sx_, sy_, sxx, syy, sxy);
            const float n = __fmul_rn(st.A1, st.A2), d = __fmul_rn(st.B1, st.B2);
            ss += fminf(fmaxf((1.f - n / d) * 0.5f, 0.f), 1.f);
        }
    }
    const float third = 1.f / 3.f;
    loss[b * hw + static_cast<size_t>(y) * W + x] =
        need_halo ? ssim_w * (ss * third) + (1.f - ssim_w) * (l1 * third) : l1 * third;
}

// ------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(PH_THREADS)
photometric_bwd_kernel(const float* __restrict__ gloss, const float* __restrict__ depth, const float* __restrict__ src,
                       const float* __restrict__ tgt, const float* __restrict__ warped, const float* __restrict__ K,
                       const float* __restrict__ invK, const float* __restrict__ T, float* __restrict__ gdepth,
                       float* __restrict__ gP, int H, int W, float ssim_w) {
    constexpr int XW = PH_TW + 4, XH = PH_TH + 4;     // values: tile + 2 halo
    constexpr int CW = PH_TW + 2, CH = PH_TH + 2;     // coefficients: tile + 1 halo
    __shared__ float xs[3][XH][XW + 1];
    __shared__ float ys[3][XH][XW + 1];
    __shared__ float ca[3][CH][CW + 1];
    __shared__ float cb[3][CH][CW + 1];
    __shared__ float cc[3][CH][CW + 1];
    __shared__ PhGeo geo;
    __shared__ float red[PH_THREADS / 32][12];
    const int b = blockIdx.z, tx0 = blockIdx.x * PH_TW, ty0 = blockIdx.y * PH_TH;
    const int tid = threadIdx.x;
    const size_t hw = static_cast<size_t>(H) * W;
    load_geo(&geo, K, invK, T, b);
    const float* wb = warped + static_cast<size_t>(b) * 3 * hw;
    const float* tb = tgt + static_cast<size_t>(b) * 3 * hw;
    const bool use_ssim = ssim_w != 0.f;
    if (use_ssim) {
        for (int s = tid; s < XW * XH; s += PH_THREADS) {
            const int sy = s / XW, sx = s - sy * XW;
            int y = ty0 + sy - 2, x = tx0 + sx - 2;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
            if (y >= -1 && y <= H && x >= -1 && x <= W) {
                y = reflect_idx(y, H);
                x = reflect_idx(x, W);
                const size_t pix = static_cast<size_t>(y) * W + x;
                a0 = __ldg(wb + pix);
                a1 = __ldg(wb + hw + pix);
                a2 = __ldg(wb + 2 * hw + pix);
                b0 = __ldg(tb + pix);
                b1 = __ldg(tb + hw + pix);
                b2 = __ldg(tb + 2 * hw + pix);
            }
            xs[0][sy][sx] = a0;
            xs[1][sy][sx] = a1;
            xs[2][sy][sx] = a2;
            ys[0][sy][sx] = b0;
            ys[1][sy][sx] = b1;
            ys[2][sy][sx] = b2;
        }
        __syncthreads();
        // coefficient maps: d(ssim loss_p)/d x_q = a_p + b_p x_q + c_p y_q for q in window(p), scaled by gloss_p * w/3
        for (int s = tid; s < CW * CH; s += PH_THREADS) {
            const int sy = s / CW, sx = s - sy * CW;
            const int y = ty0 + sy - 1, x = tx0 + sx - 1;
            const bool in_img = y >= 0 && y < H && x >= 0 && x < W;
            const float gl = in_img ? __ldg(gloss + b * hw + static_cast<size_t>(y) * W + x) * (ssim_w * (1.f / 3.f)) : 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float A = 0.f, Bc = 0.f, Cc = 0.f;
                if (in_img && gl != 0.f) {
                    float sx_ = 0.f, sy_ = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float a = xs[c][sy + dy][sx + dx], bb = ys[c][sy + dy][sx + dx];
                            sx_ += a;
                            sy_ += bb;
                            sxx = fmaf(a, a, sxx);
                            syy = fmaf(bb, bb, syy);
                            sxy = fmaf(a, bb, sxy);
                        }
                    const float k9 = 1.f / 9.f;
                    const SsimTerms st = ssim_terms(sx_, sy_, sxx, syy, sxy);
                    const float mx = st.mx, my = st.my, A1 = st.A1, A2 = st.A2, B1 = st.B1, B2 = st.B2;
                    const float n = __fmul_rn(A1, A2), d = __fmul_rn(B1, B2);
                    const float val = (1.f - n / d) * 0.5f;
                    if (val >= 0.f && val <= 1.f) {          // clamp passes gradient on the closed interval
                        const float inv_d = 1.f / d, k = gl * k9 * inv_d;
                        Cc = -A1 * k;
                        Bc = n * B1 * k * inv_d;
                        A = -(my * (A2 - A1) - n * inv_d * mx * (B2 - B1)) * k;
                    }
                }
                ca[c][sy][sx] = A;
                cb[c][sy][sx] = Bc;
                cc[c][sy][sx] = Cc;
            }
        }
    }
    __syncthreads();

    const int lx = tid % PH_TW, ly = tid / PH_TW;
    const int x = tx0 + lx, y = ty0 + ly;
    float part[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) part[k] = 0.f;
    if (x < W && y < H) {
        const size_t pix = static_cast<size_t>(y) * W + x;
        const float gl = __ldg(gloss + b * hw + pix);
        float gx[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float xv = __ldg(wb + c * hw + pix), yv = __ldg(tb + c * hw + pix);
            const float df = yv - xv;
            const float sgn = (df > 0.f) ? 1.f : ((df < 0.f) ? -1.f : 0.f);
            const float l1w = use_ssim ? (1.f - ssim_w) : 1.f;
            float g = -gl * l1w * (1.f / 3.f) * sgn;
            if (use_ssim) {
                // padded positions that hold x_q: q itself, plus the reflected pad rows/cols when q is 1 or n-2
                float sa = 0.f, sb_ = 0.f, sc_ = 0.f;
#pragma unroll
                for (int jy = 0; jy < 3; ++jy) {
                    int py0;     // centre row (pixel coords) of the 3 windows containing padded position j
                    if (jy == 0) py0 = y;
                    else if (jy == 1) { if (y != 1) continue; py0 = -1; }
                    else { if (y != H - 2) continue; py0 = H; }
#pragma unroll
                    for (int jx = 0; jx < 3; ++jx) {
                        int px0;
                        if (jx == 0) px0 = x;
                        else if (jx == 1) { if (x != 1) continue; px0 = -1; }
                        else { if (x != W - 2) continue; px0 = W; }
#pragma unroll
                        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int py = py0 + dy, px = px0 + dx;
                                if (py < 0 || py >= H || px < 0 || px >= W) continue;
                                const int cy = py - ty0 + 1, cx = px - tx0 + 1;     // always inside tile + 1 halo
                                sa += ca[c][cy][cx];
                                sb_ += cb[c][cy][cx];
                                sc_ += cc[c][cy][cx];
                            }
                    }
                }
                g += sa + sb_ * xv + sc_ * yv;
            }
            gx[c] = g;
        }
        // bilinear warp adjoint (border clamp: zero gradient where the coordinate was clipped)
        const float d = __ldg(depth + b * hw + pix);
        const WarpPt wp = warp_point(geo, d, x, y, W, H);
        const Taps t = taps_at(wp.u, wp.v, W, H);
        float gu = 0.f, gv = 0.f;
        const float* sbp = src + static_cast<size_t>(b) * 3 * hw;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* sc = sbp + c * hw;
            const float a00 = __ldg(sc + t.o00);
            const float a01 = t.x1ok ? __ldg(sc + t.o01) : 0.f;
            const float a10 = t.y1ok ? __ldg(sc + t.o10) : 0.f;
            const float a11 = (t.x1ok && t.y1ok) ? __ldg(sc + t.o11) : 0.f;
            // ATen: gix = sum val * d(weight)/dx with weights (x1-u)(y1-v) etc.
            gu += gx[c] * ((a01 - a00) * t.fy0 + (a11 - a10) * t.fy1);
            gv += gx[c] * ((a10 - a00) * t.fx0 + (a11 - a01) * t.fx1);
        }
        if (!wp.gu) gu = 0.f;
        if (!wp.gv) gv = 0.f;
        // u = px / z, v = py / z
        const float gpx = gu * wp.inv_z, gpy = gv * wp.inv_z;
        const float gpz = -(gu * wp.ur + gv * wp.vr) * wp.inv_z;
        // p_i = sum_j P_ij X_j (X_3 = 1), X = depth * ray
        const float gX0 = geo.P[0] * gpx + geo.P[4] * gpy + geo.P[8] * gpz;
        const float gX1 = geo.P[1] * gpx + geo.P[5] * gpy + geo.P[9] * gpz;
        const float gX2 = geo.P[2] * gpx + geo.P[6] * gpy + geo.P[10] * gpz;
        gdepth[b * hw + pix] = gX0 * wp.r[0] + gX1 * wp.r[1] + gX2 * wp.r[2];
        const float gp[3] = {gpx, gpy, gpz};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            part[i * 4 + 0] = gp[i] * wp.X[0];
            part[i * 4 + 1] = gp[i] * wp.X[1];
            part[i * 4 + 2] = gp[i] * wp.X[2];
            part[i * 4 + 3] = gp[i];
        }
    }
    if (gP != nullptr) {
#pragma unroll
        for (int k = 0; k < 12; ++k) part[k] = warp_sum(part[k]);
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) red[tid >> 5][k] = part[k];
        }
        __syncthreads();
        if (tid < 12) {
            float s = 0.f;
#pragma unroll
            for (int wi = 0; wi < PH_THREADS / 32; ++wi) s += red[wi][tid];
            atomicAdd(gP + b * 12 + tid, s);
        }
    }
}

}  // namespace mvd

extern "C" {

int mvd_photometric_fwd(const float* depth, const float* src, const float* tgt, const float* K, const float* invK,
                        const float* T, float* warped, float* loss, int B, int H, int W, float ssim_w, int identity,
                        void* stream) {
    MVD_REQUIRE(src && tgt && loss, "null pointer argument");
    MVD_REQUIRE(identity || (depth && K && invK && T), "warp mode needs depth, K, invK, T");
    MVD_REQUIRE(B > 0 && H >= 4 && W >= 4, "bad shape B=%d H=%d W=%d", B, H, W);
    dim3 grid((W + mvd::PH_TW - 1) / mvd::PH_TW, (H + mvd::PH_TH - 1) / mvd::PH_TH, B);
    mvd::photometric_fwd_kernel<<<grid, mvd::PH_THREADS, 0, mvd::as_stream(stream)>>>(depth, src, tgt, K, invK, T, warped,
                                                                                       loss, H, W, ssim_w, identity);
    return mvd::check_launch("photometric_fwd");
}

int mvd_photometric_bwd(const float* gloss, const float* depth, const float* src, const float* tgt,
                        const float* warped, const float* K, const float* invK, const float* T, float* gdepth,
                        float* gP, int B, int H, int W, float ssim_w, void* stream) {
    MVD_REQUIRE(gloss && depth && src && tgt && warped && K && invK && T && gdepth, "null pointer argument");
    MVD_REQUIRE(B > 0 && H >= 4 && W >= 4, "bad shape B=%d H=%d W=%d", B, H, W);
    cudaStream_t st = mvd::as_stream(stream);
    if (gP != nullptr) {
        cudaError_t e = cudaMemsetAsync(gP, 0, sizeof(float) * B * 12, st);
        if (e != cudaSuccess) return mvd::fail(static_cast<int>(e), "photometric_bwd memset: %s", cudaGetErrorString(e));
    }
    dim3 grid((W + mvd::PH_TW - 1) / mvd::PH_TW, (H + mvd::PH_TH - 1) / mvd::PH_TH, B);
    mvd::photometric_bwd_kernel<<<grid, mvd::PH_THREADS, 0, st>>>(gloss, depth, src, tgt, warped, K, invK, T, gdepth, gP, H,
                                                                   W, ssim_w);
    return mvd::check_launch("photometric_bwd");
}

}
