// K1 / K1b -- fused homography-warp cost volume with group correlation (sm_100a).
//
// What the reference does (movedepth/layers.py:778-794 + movedepth/trainer.py:359): per batch item
// repeat the source feature map D times, back-project D depth maps, project, grid_sample
// (bilinear, zeros, align_corners=True), multiply by the reference feature, stack to
// [B,D,C,h,w] (566 MB at B=6, D=96) and only then average channel pairs into G groups.
//
// What this kernel does instead: one persistent CTA per tile of R x 32 reference pixels.
//   * the bounding box of all source positions the tile can touch (over every hypothesis) is
//     found with a warp-shuffle min/max reduction of the projected segment end points;
//   * that source box is staged once into shared memory by TMA (channels-last, 128B swizzle,
//     out-of-bounds zero fill == padding_mode='zeros'), the reference tile likewise;
//   * every thread owns one pixel and a contiguous chunk of hypotheses.  Because consecutive
//     hypotheses fall into the same bilinear cell most of the time, the thread keeps the
//     per-tap group correlations  P_tap[g] = 1/2 * sum_{c in group g} ref[c] * src_tap[c]  of its
//     current cell in registers; a hypothesis then costs 64 FMAs and 16 coalesced stores, and
//     shared memory is only read when the cell changes;
//   * tiles whose box does not fit (degenerate poses) gather straight from global/L2.
// The C-channel volume is never materialised: HBM traffic is ref + src + prior in, the grouped
// volume out (4*B*h*w*(2C + 1 + D*G) bytes).
//
// Backward: Q_tap[g] += w_tap * gout[g] is accumulated per bilinear cell in registers and flushed with vector atomics
// when the cell changes.  The production kernel is costvol_grouped_bwd_v5_kernel (warp-autonomous, a lane pair per pixel,
// the gradient volume streamed through a per-warp bulk-copy ring; see its header below); the round-1 kernel on the forward's
// tiles is kept behind MVD_FLAG_BWD_V2 for A/B measurements.
#include "common.cuh"
#include "../../include/movedepth_b200.h"

#include <limits.h>

namespace mvd {

constexpr int CV_C = 32;
constexpr int CV_G = 16;
constexpr int CV_TW = 32;   // tile width  (one lane per pixel column)
constexpr int CV_R = 2;     // tile rows
constexpr int CV_NCH = 4;   // hypothesis chunks per CTA
constexpr int CV_WARPS = CV_R * CV_NCH;
constexpr int CV_THREADS = 32 * CV_WARPS;
constexpr int CV_BOX_W = 48;          // staged source box (pixels)
constexpr int CV_BOX_H = CV_R + 8;
constexpr int CV_MAXD = 1024;

struct CvArgs {
    const float* ref;
    const float* src;
    const float* prior;
    const float* ratio;
    const float* hyps;
    const float* K;
    const float* invK;
    const float* T;
    float* out;          // fwd
    const float* gout;   // bwd
    float* gref;
    float* gsrc;
    int B, h, w, D;
    int layout;
    int use_tma;
    int use_table;
    int dbg_nostore;
    int tma_store;
    int tiles_x, tiles_y, num_tiles;
    int DC;              // hypotheses per chunk
};

struct CvSmem {
    // offsets into dynamic shared memory (base aligned to 1024 B for the 128B swizzle)
    static constexpr int SRC_BYTES = CV_BOX_W * CV_BOX_H * 128;
    static constexpr int REF_BYTES = CV_R * CV_TW * 128;
    static constexpr int OFF_SRC = 0;
    static constexpr int OFF_REF = (SRC_BYTES + 1023) / 1024 * 1024;
    static constexpr int OFF_RATIO = OFF_REF + REF_BYTES;
    static constexpr int OFF_GEO = OFF_RATIO + CV_MAXD * 4;
    static constexpr int OFF_RED = OFF_GEO + 32 * 4;
    static constexpr int OFF_BAR = OFF_RED + 2 * CV_WARPS * 4 * 4;
    static constexpr int OFF_STAGE = (OFF_BAR + 16 + 127) / 128 * 128;      // per-warp double-buffered output rows
    static constexpr int STAGE_BYTES = CV_WARPS * 2 * CV_G * CV_TW * 4;
    static constexpr int TOTAL = OFF_STAGE + STAGE_BYTES;
    static constexpr int ALLOC = TOTAL + 1024;   // slack for manual 1024 B alignment
};

struct PixelCtx {
    float mrx, mry, mrz, tx, ty, tz;   // p(depth) = depth * mr + t
    int b, x, y, d0, d1;
    bool active;
    bool boxed;         // source box staged in shared memory for this tile
    int ox, oy;         // box origin in source pixels
};

__device__ __forceinline__ void project_uv(const PixelCtx& c, float depth, float& u, float& v, float& pz) {
    pz = fmaf(depth, c.mrz, c.tz) + 1e-7f;
    const float inv = __frcp_rn(pz);
    u = fmaf(depth, c.mrx, c.tx) * inv;
    v = fmaf(depth, c.mry, c.ty) * inv;
}

__device__ __forceinline__ float4 lds128(const unsigned char* base, int pix, int j) {
    return *reinterpret_cast<const float4*>(base + pix * 128 + ((j ^ (pix & 7)) << 4));
}

// Fetch channels [4j,4j+4) and [16+4j,16+4j+4) of source pixel (px,py); zero outside the image.
__device__ __forceinline__ void fetch_tap(const CvArgs& a, const PixelCtx& c, const unsigned char* sbox, bool in_box,
                                          int rel_pix, int px, int py, int j, float4& lo, float4& hi) {
    if (in_box) {
        lo = lds128(sbox, rel_pix, j);
        hi = lds128(sbox, rel_pix, j + 4);
    } else if (px >= 0 && px < a.w && py >= 0 && py < a.h) {
        const float4* g = reinterpret_cast<const float4*>(a.src + (static_cast<size_t>(c.b * a.h + py) * a.w + px) * CV_C);
        lo = __ldg(g + j);
        hi = __ldg(g + j + 4);
    } else {
        lo = make_float4(0.f, 0.f, 0.f, 0.f);
        hi = lo;
    }
}

// Per-tile prologue shared by forward and backward.  Returns the pixel context; leaves the
// reference tile (and, when it fits, the source box) in shared memory.
__device__ __forceinline__ void tile_prologue(const CvArgs& a, const CUtensorMap* map_src, const CUtensorMap* map_ref,
                                              unsigned char* smem, int tile, int iter, uint32_t& phase, PixelCtx& c) {
    float* ratio_s = reinterpret_cast<float*>(smem + CvSmem::OFF_RATIO);
    float* geo = reinterpret_cast<float*>(smem + CvSmem::OFF_GEO);
    float* red = reinterpret_cast<float*>(smem + CvSmem::OFF_RED) + (iter & 1) * CV_WARPS * 4;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + CvSmem::OFF_BAR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_b = a.tiles_x * a.tiles_y;
    const int b = tile / per_b;
    const int trem = tile - b * per_b;
    const int tyi = trem / a.tiles_x, txi = trem - tyi * a.tiles_x;

    __syncthreads();   // everyone is done with the previous tile's shared data
    if (tid < 12) {    // P = (K @ T)[:3, :]   (movedepth/layers.py:609)
        const int i = tid >> 2, j = tid & 3;
        const float* Kb = a.K + b * 16;
        const float* Tb = a.T + b * 16;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s = fmaf(Kb[i * 4 + k], Tb[k * 4 + j], s);
        geo[tid] = s;
    } else if (tid < 21) {   // inv_K[:3,:3]   (movedepth/layers.py:582)
        const int q = tid - 12;
        geo[12 + q] = a.invK[b * 16 + (q / 3) * 4 + (q % 3)];
    }
    if (a.ratio != nullptr)
        for (int d = tid; d < a.D; d += CV_THREADS) ratio_s[d] = a.ratio[b * a.D + d];
    __syncthreads();

    const int row = warp % CV_R, chunk = warp / CV_R;
    c.b = b;
    c.x = txi * CV_TW + lane;
    c.y = tyi * CV_R + row;
    c.d0 = chunk * a.DC;
    c.d1 = min(a.D, c.d0 + a.DC);
    c.active = (c.x < a.w) && (c.y < a.h) && (c.d0 < c.d1);
    {
        const float xf = static_cast<float>(c.x), yf = static_cast<float>(c.y);
        const float rx = fmaf(geo[12], xf, fmaf(geo[13], yf, geo[14]));
        const float ry = fmaf(geo[15], xf, fmaf(geo[16], yf, geo[17]));
        const float rz = fmaf(geo[18], xf, fmaf(geo[19], yf, geo[20]));
        c.mrx = fmaf(geo[0], rx, fmaf(geo[1], ry, geo[2] * rz));
        c.mry = fmaf(geo[4], rx, fmaf(geo[5], ry, geo[6] * rz));
        c.mrz = fmaf(geo[8], rx, fmaf(geo[9], ry, geo[10] * rz));
        c.tx = geo[3];
        c.ty = geo[7];
        c.tz = geo[11];
    }

    // ---- bounding box of the source footprint: the projection of a depth interval is the
    // segment between the projections of its end points as long as z stays positive.
    float lo_u = 3.0e38f, hi_u = -3.0e38f, lo_v = 3.0e38f, hi_v = -3.0e38f;
    int bad = 0;
    if (c.active && a.use_tma) {
        float dmin, dmax;
        if (a.hyps != nullptr) {
            dmin = 3.0e38f;
            dmax = -3.0e38f;
            const float* hp = a.hyps + (static_cast<size_t>(b) * a.D * a.h + c.y) * a.w + c.x;
            for (int d = c.d0; d < c.d1; ++d) {
                const float dv = __ldg(hp + static_cast<size_t>(d) * a.h * a.w);
                dmin = fminf(dmin, dv);
                dmax = fmaxf(dmax, dv);
            }
        } else {
            const float pr = __ldg(a.prior + static_cast<size_t>(b * a.h + c.y) * a.w + c.x);
            const float d_a = pr * ratio_s[c.d0], d_b = pr * ratio_s[c.d1 - 1];
            dmin = fminf(d_a, d_b);
            dmax = fmaxf(d_a, d_b);
        }
        float u0, v0, z0, u1, v1, z1;
        project_uv(c, dmin, u0, v0, z0);
        project_uv(c, dmax, u1, v1, z1);
        if (!(z0 > 1e-6f) || !(z1 > 1e-6f) || !(fabsf(u0) < 1e8f) || !(fabsf(u1) < 1e8f) || !(fabsf(v0) < 1e8f) ||
            !(fabsf(v1) < 1e8f)) {
            bad = 1;
        } else {
            const float wf = static_cast<float>(a.w), hf = static_cast<float>(a.h);
            const float mnu = fminf(u0, u1), mxu = fmaxf(u0, u1), mnv = fminf(v0, v1), mxv = fmaxf(v0, v1);
            if (mxu > -1.f && mnu < wf && mxv > -1.f && mnv < hf) {   // segment touches the image
                lo_u = fmaxf(mnu, -1.f);
                hi_u = fminf(mxu, wf);
                lo_v = fmaxf(mnv, -1.f);
                hi_v = fminf(mxv, hf);
            }
        }
    }
    lo_u = warp_min(lo_u);
    hi_u = warp_max(hi_u);
    lo_v = warp_min(lo_v);
    hi_v = warp_max(hi_v);
    if (lane == 0) {
        red[warp * 4 + 0] = lo_u;
        red[warp * 4 + 1] = hi_u;
        red[warp * 4 + 2] = lo_v;
        red[warp * 4 + 3] = hi_v;
    }
    const int any_bad = __syncthreads_or(bad);
#pragma unroll
    for (int wi = 0; wi < CV_WARPS; ++wi) {
        lo_u = fminf(lo_u, red[wi * 4 + 0]);
        hi_u = fmaxf(hi_u, red[wi * 4 + 1]);
        lo_v = fminf(lo_v, red[wi * 4 + 2]);
        hi_v = fmaxf(hi_v, red[wi * 4 + 3]);
    }
    c.boxed = false;
    c.ox = 0;
    c.oy = 0;
    if (a.use_tma && !any_bad) {
        if (hi_u < lo_u) {   // nothing of this tile lands inside the image: empty box is fine
            c.boxed = true;
        } else {
            const int ix0 = static_cast<int>(floorf(lo_u)), ix1 = min(static_cast<int>(floorf(hi_u)), a.w - 1);
            const int iy0 = static_cast<int>(floorf(lo_v)), iy1 = min(static_cast<int>(floorf(hi_v)), a.h - 1);
            c.ox = ix0 - 1;
            c.oy = iy0 - 1;
            c.boxed = (ix1 + 2 - c.ox + 1 <= CV_BOX_W) && (iy1 + 2 - c.oy + 1 <= CV_BOX_H);
        }
    }
    if (tid == 0) {
        const uint32_t bytes = CvSmem::REF_BYTES + (c.boxed ? CvSmem::SRC_BYTES : 0);
        mbar_expect_tx(bar, bytes);
        tma_load_4d(smem + CvSmem::OFF_REF, map_ref, bar, 0, txi * CV_TW, tyi * CV_R, b);
        if (c.boxed) tma_load_4d(smem + CvSmem::OFF_SRC, map_src, bar, 0, c.ox, c.oy, b);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
}

__device__ __forceinline__ void load_half_ref(const unsigned char* smem, int warp, int lane, float (&rh)[CV_C]) {
    const unsigned char* rbox = smem + CvSmem::OFF_REF;
    const int pr = (warp % CV_R) * CV_TW + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 v = lds128(rbox, pr, j);
        rh[4 * j + 0] = 0.5f * v.x;
        rh[4 * j + 1] = 0.5f * v.y;
        rh[4 * j + 2] = 0.5f * v.z;
        rh[4 * j + 3] = 0.5f * v.w;
    }
}

struct Bilinear {
    float w00, w01, w10, w11;
    int ix, iy;
    bool ok;
};

// ATen grid_sampler_2d (bilinear, align_corners=True, zeros): nw=(x1-u)(y1-v), ne=(u-x0)(y1-v), ...
__device__ __forceinline__ Bilinear bilinear_at(const CvArgs& a, float u, float v) {
    Bilinear s;
    s.ok = (u > -1.f) && (u < static_cast<float>(a.w)) && (v > -1.f) && (v < static_cast<float>(a.h));
    s.w00 = s.w01 = s.w10 = s.w11 = 0.f;
    s.ix = 0;
    s.iy = 0;
    if (s.ok) {
        const float x0 = floorf(u), y0 = floorf(v);
        const float fx1 = u - x0, fx0 = (x0 + 1.f) - u;
        const float fy1 = v - y0, fy0 = (y0 + 1.f) - v;
        s.w00 = fx0 * fy0;
        s.w01 = fx1 * fy0;
        s.w10 = fx0 * fy1;
        s.w11 = fx1 * fy1;
        s.ix = static_cast<int>(x0);
        s.iy = static_cast<int>(y0);
    }
    return s;
}

__device__ __forceinline__ bool cell_in_box(const PixelCtx& c, int ix, int iy, int& rel) {
    const int rx = ix - c.ox, ry = iy - c.oy;
    rel = ry * CV_BOX_W + rx;
    return c.boxed && rx >= 0 && rx <= CV_BOX_W - 2 && ry >= 0 && ry <= CV_BOX_H - 2;
}

// ------------------------------------------------------------------------------------------ forward
__device__ __forceinline__ ulonglong2 lds128p(const unsigned char* base, int pix, int j) {
    return *reinterpret_cast<const ulonglong2*>(base + pix * 128 + ((j ^ (pix & 7)) << 4));
}

// 5-D TMA tiled store shared::cta -> global (SASS: UTMASTG), bulk-group completion.
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3,
                                             int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- forward, v4 ---------------------------------------------------------------------------------
// Tile = one row of 32 reference pixels x all D; warp = chunk of D/8 hypotheses, lane = pixel.
//  * Over the whole hypothesis range a pixel's epipolar segment is only a few source pixels long, so
//    the per-tap group correlations  P_tap[g] = 1/2 * sum_{c in {g,g+16}} ref[c] * src_tap[c]  of
//    every tap the pixel can touch are computed ONCE per tile (8 warps share the taps of a pixel) into
//    a shared-memory table; the hypothesis loop refills its register-resident bilinear cell from that
//    table with 16 LDS.128 instead of recomputing 128 products whenever one lane of the warp crosses
//    a cell boundary.  Pixels whose footprint does not fit the table (degenerate poses) fall back to
//    the direct computation from global memory.
//  * The persistent tile loop is software pipelined: while the warps run the hypothesis loop of tile
//    i, the source box and reference row of tile i+1 are already in flight (their footprints are
//    computed right after the table of tile i is finished, warp 0 issues the TMA loads from behind a
//    named barrier the other warps only arrive at), and the prior / ratio values of tile i+1 are
//    prefetched into registers one stage earlier.
namespace f3 {
constexpr int TW = 32;
constexpr int NCH = 8;                      // hypothesis chunks = warps per CTA
constexpr int WARPS = NCH;
constexpr int THREADS = 32 * WARPS;
constexpr int BOX_W = 48;                   // staged source box (pixels)
constexpr int BOX_H = 6;
constexpr int NT = 16;                      // table capacity: taps per pixel
constexpr int PT_STRIDE = NT * 64 + 16;     // bytes per pixel; the 16 B pad makes 8 consecutive lanes hit 8 bank groups
constexpr int STAGE_ROW = CV_G * TW * 4;    // one hypothesis of the tile row: 16 groups x 32 pixels
constexpr int MAXB = 16;                    // batch items per launch (their geometry lives in shared memory)
constexpr int GEO_STRIDE = 24;              // floats per batch item: (K@T)[:3,:4] row-major, then inv_K[:3,:3]

struct Smem {
    static constexpr int SRC_BYTES = BOX_W * BOX_H * 128;
    static constexpr int REF_BYTES = TW * 128;
    static constexpr int OFF_SRC = 0;                                   // 1024 B aligned (128B swizzle)
    static constexpr int OFF_REF = OFF_SRC + SRC_BYTES;
    static constexpr int OFF_PTAB = OFF_REF + REF_BYTES;
    static constexpr int OFF_STAGE = OFF_PTAB + TW * PT_STRIDE;         // per-warp double-buffered output rows
    static constexpr int OFF_FOOT = OFF_STAGE + WARPS * 2 * STAGE_ROW;  // [2][WARPS][32] short4 cell ranges
    static constexpr int OFF_RED = OFF_FOOT + 2 * WARPS * 32 * 8;       // [2][WARPS][8] ints
    static constexpr int OFF_GEO = OFF_RED + 2 * WARPS * 8 * 4;         // [MAXB][GEO_STRIDE] floats
    static constexpr int OFF_BAR = OFF_GEO + MAXB * GEO_STRIDE * 4;
    static constexpr int TOTAL = OFF_BAR + 16;
    static constexpr int ALLOC = TOTAL + 1024;                          // slack for manual 1024 B alignment
};
static_assert(Smem::OFF_REF % 1024 == 0 && Smem::OFF_STAGE % 512 == 0, "TMA alignment");
static_assert(2 * (Smem::ALLOC + 1024) <= 233472, "two CTAs per SM");

__device__ __forceinline__ float rcp_approx(float x) {       // MUFU.RCP, 1 ulp
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Direct computation of the four per-tap group correlations of cell (cx,cy) from global memory (no table).
__device__ __forceinline__ void load_cell_direct(const CvArgs& a, int b, int ref_pix_global, int cx, int cy,
                                                 uint64_t (&P2)[4][8]) {
    const ulonglong2* rg = reinterpret_cast<const ulonglong2*>(a.ref + static_cast<size_t>(ref_pix_global) * CV_C);
    const uint64_t half2 = pk2(0.5f, 0.5f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ulonglong2 rl = __ldg(rg + j), rh = __ldg(rg + j + 4);
        rl.x = mul2(rl.x, half2);
        rl.y = mul2(rl.y, half2);
        rh.x = mul2(rh.x, half2);
        rh.y = mul2(rh.y, half2);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int px = cx + (t & 1), py = cy + (t >> 1);
            ulonglong2 lo = make_ulonglong2(0ull, 0ull), hi = lo;
            if (px >= 0 && px < a.w && py >= 0 && py < a.h) {
                const ulonglong2* g =
                    reinterpret_cast<const ulonglong2*>(a.src + (static_cast<size_t>(b * a.h + py) * a.w + px) * CV_C);
                lo = __ldg(g + j);
                hi = __ldg(g + j + 4);
            }
            P2[t][2 * j] = fma2(rl.x, lo.x, mul2(rh.x, hi.x));
            P2[t][2 * j + 1] = fma2(rl.y, lo.y, mul2(rh.y, hi.y));
        }
    }
}

// Refill the register cell from the pixel's table row: taps t00, t00+1, t00+nxp, t00+nxp+1.
__device__ __forceinline__ void load_cell_table(const unsigned char* ptab_pix, int t00, int nxp, uint64_t (&P2)[4][8]) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(ptab_pix + (t00 + (t & 1) + (t >> 1) * nxp) * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const ulonglong2 v = p[q];
            P2[t][2 * q] = v.x;
            P2[t][2 * q + 1] = v.y;
        }
    }
}

struct Geom {
    float mrx, mry, mrz, tx, ty, tz;   // p(depth) = depth * mr + t
};

// r = inv_K[:3,:3] (x,y,1);  mr = (K T)[:3,:3] r;  t = (K T)[:3,3]    (movedepth/layers.py:581-586, 601-621)
__device__ __forceinline__ Geom pixel_geom(const float* geo, int x, int y) {
    const float xf = static_cast<float>(x), yf = static_cast<float>(y);
    const float rx = fmaf(geo[12], xf, fmaf(geo[13], yf, geo[14]));
    const float ry = fmaf(geo[15], xf, fmaf(geo[16], yf, geo[17]));
    const float rz = fmaf(geo[18], xf, fmaf(geo[19], yf, geo[20]));
    Geom c;
    c.mrx = fmaf(geo[0], rx, fmaf(geo[1], ry, geo[2] * rz));
    c.mry = fmaf(geo[4], rx, fmaf(geo[5], ry, geo[6] * rz));
    c.mrz = fmaf(geo[8], rx, fmaf(geo[9], ry, geo[10] * rz));
    c.tx = geo[3];
    c.ty = geo[7];
    c.tz = geo[11];
    return c;
}

struct TileXY {
    int b, y, tx0;
};
__device__ __forceinline__ TileXY tile_xy(const CvArgs& a, int tile) {
    const int per_b = a.tiles_x * a.h;
    TileXY t;
    t.b = tile / per_b;
    const int trem = tile - t.b * per_b;
    t.y = trem / a.tiles_x;
    t.tx0 = (trem - t.y * a.tiles_x) * TW;
    return t;
}

// Tile box / per-pixel table footprint, derived identically by every thread from the per-warp
// reductions (`red`) and per-(chunk, pixel) cell ranges (`foot`) that footprint_stage left in shared memory.
struct Footprint {
    int fx0, fx1, fy0, fy1, nxp;   // cells the pixel can touch (inclusive); taps per table row
    int ox, oy;                    // staged source box origin
    bool boxed, tab_ok;
    int ntaps;
};
__device__ __forceinline__ void tile_box(const CvArgs& a, const int* red, int& ox, int& oy, bool& boxed, bool& any_bad) {
    int bx0 = 32767, bx1 = -32768, by0 = 32767, by1 = -32768, bad = 0;
#pragma unroll
    for (int wi = 0; wi < WARPS; ++wi) {
        bx0 = min(bx0, red[wi * 8 + 0]);
        bx1 = max(bx1, red[wi * 8 + 1]);
        by0 = min(by0, red[wi * 8 + 2]);
        by1 = max(by1, red[wi * 8 + 3]);
        bad |= red[wi * 8 + 4];
    }
    any_bad = bad != 0;
    boxed = a.use_tma && !any_bad && (bx1 >= bx0) && (bx1 - bx0 + 2 <= BOX_W) && (by1 - by0 + 2 <= BOX_H);
    ox = bx0;
    oy = by0;
}

// Stage F of the pipeline: which source cells can (pixel, chunk) touch?  The projection of a depth
// interval is the segment between the projections of its end points as long as z stays positive.
__device__ __forceinline__ void footprint_stage(const CvArgs& a, const CUtensorMap* map_src, const CUtensorMap* map_ref,
                                                unsigned char* smem, int tile, int par, float prior_v, float ra, float rb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileXY t = tile_xy(a, tile);
    const int x = t.tx0 + lane;
    const bool lane_ok = x < a.w;
    const int d0 = warp * a.DC, d1 = min(a.D, d0 + a.DC);
    short4* foot = reinterpret_cast<short4*>(smem + Smem::OFF_FOOT) + par * WARPS * 32;
    int* red = reinterpret_cast<int*>(smem + Smem::OFF_RED) + par * WARPS * 8;
    const Geom c = pixel_geom(reinterpret_cast<const float*>(smem + Smem::OFF_GEO) + t.b * GEO_STRIDE, x, t.y);

    short4 ft = make_short4(32767, -32768, 32767, -32768);     // empty
    int bad = 0;
    if (lane_ok && d0 < d1) {
        float dmin, dmax;
        if (a.hyps != nullptr) {
            const int hw = a.h * a.w;
            const float* hp = a.hyps + static_cast<size_t>(t.b) * a.D * hw + t.y * a.w + x;
            dmin = 3.0e38f;
            dmax = -3.0e38f;
            for (int d = d0; d < d1; ++d) {
                const float dv = __ldg(hp + static_cast<size_t>(d) * hw);
                dmin = fminf(dmin, dv);
                dmax = fmaxf(dmax, dv);
            }
        } else {
            const float d_a = prior_v * ra, d_b = prior_v * rb;
            dmin = fminf(d_a, d_b);
            dmax = fmaxf(d_a, d_b);
        }
        const float z0 = fmaf(dmin, c.mrz, c.tz) + 1e-7f, z1 = fmaf(dmax, c.mrz, c.tz) + 1e-7f;   // same arithmetic as the
        const float i0 = rcp_approx(z0), i1 = rcp_approx(z1);                                     // hypothesis loop
        const float u0 = fmaf(dmin, c.mrx, c.tx) * i0, v0 = fmaf(dmin, c.mry, c.ty) * i0;
        const float u1 = fmaf(dmax, c.mrx, c.tx) * i1, v1 = fmaf(dmax, c.mry, c.ty) * i1;
        if (!(z0 > 1e-6f) || !(z1 > 1e-6f) || !(fabsf(u0) < 1e8f) || !(fabsf(u1) < 1e8f) || !(fabsf(v0) < 1e8f) ||
            !(fabsf(v1) < 1e8f)) {
            bad = 1;
        } else {
            const float wf = static_cast<float>(a.w), hf = static_cast<float>(a.h);
            const float mnu = fminf(u0, u1), mxu = fmaxf(u0, u1), mnv = fminf(v0, v1), mxv = fmaxf(v0, v1);
            if (mxu > -1.f && mnu < wf && mxv > -1.f && mnv < hf) {   // segment touches the image
                ft.x = static_cast<short>(floorf(fmaxf(mnu, -1.f)));
                ft.y = static_cast<short>(min(static_cast<int>(floorf(mxu)), a.w - 1));
                ft.z = static_cast<short>(floorf(fmaxf(mnv, -1.f)));
                ft.w = static_cast<short>(min(static_cast<int>(floorf(mxv)), a.h - 1));
            }
        }
    }
    foot[warp * 32 + lane] = ft;
    const int lo_x = __reduce_min_sync(0xffffffffu, static_cast<int>(ft.x));
    const int hi_x = __reduce_max_sync(0xffffffffu, static_cast<int>(ft.y));
    const int lo_y = __reduce_min_sync(0xffffffffu, static_cast<int>(ft.z));
    const int hi_y = __reduce_max_sync(0xffffffffu, static_cast<int>(ft.w));
    const int wbad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        red[warp * 8 + 0] = lo_x;
        red[warp * 8 + 1] = hi_x;
        red[warp * 8 + 2] = lo_y;
        red[warp * 8 + 3] = hi_y;
        red[warp * 8 + 4] = wbad;
    }
    // Hand-off: warps 1..7 only arrive at named barrier 1 and go on; warp 0 waits for all of them and issues the loads.
    if (warp != 0) {
        asm volatile("bar.arrive 1, %0;" ::"n"(THREADS) : "memory");
    } else {
        asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
        int ox, oy;
        bool boxed, any_bad;
        tile_box(a, red, ox, oy, boxed, any_bad);
        if (lane == 0) {
            uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Smem::OFF_BAR);
            mbar_expect_tx(bar, Smem::REF_BYTES + (boxed ? Smem::SRC_BYTES : 0));
            tma_load_4d(smem + Smem::OFF_REF, map_ref, bar, 0, t.tx0, t.y, t.b);
            if (boxed) tma_load_4d(smem + Smem::OFF_SRC, map_src, bar, 0, ox, oy, t.b);
        }
    }
}
}  // namespace f3

__global__ void __launch_bounds__(f3::THREADS, 2)
costvol_grouped_fwd_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_ref,
                           const __grid_constant__ CUtensorMap map_out, const CvArgs a) {
    using namespace f3;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Smem::OFF_BAR);
    const unsigned char* sbox = smem + Smem::OFF_SRC;
    const unsigned char* rbox = smem + Smem::OFF_REF;
    float* geo_all = reinterpret_cast<float*>(smem + Smem::OFF_GEO);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* ptab_pix = smem + Smem::OFF_PTAB + lane * PT_STRIDE;
    float* stage = reinterpret_cast<float*>(smem + Smem::OFF_STAGE + warp * 2 * STAGE_ROW);
    const int hw = a.h * a.w;
    const int d0 = warp * a.DC, d1 = min(a.D, d0 + a.DC), n = d1 - d0;

    // ---- geometry of every batch item, once per CTA
    for (int e = tid; e < a.B * 21; e += THREADS) {
        const int bb = e / 21, q = e - bb * 21;
        float v;
        if (q < 12) {                        // P = (K @ T)[:3, :]   (movedepth/layers.py:609)
            const int i = q >> 2, j = q & 3;
            const float* Kb = a.K + bb * 16;
            const float* Tb = a.T + bb * 16;
            v = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) v = fmaf(Kb[i * 4 + k], Tb[k * 4 + j], v);
        } else {                             // inv_K[:3,:3]   (movedepth/layers.py:582)
            const int r = q - 12;
            v = a.invK[bb * 16 + (r / 3) * 4 + (r % 3)];
        }
        geo_all[bb * GEO_STRIDE + q] = v;
    }
    if (tid == 0) {
        tma_prefetch_desc(&map_src);
        tma_prefetch_desc(&map_ref);
        tma_prefetch_desc(&map_out);
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // prior / chunk-end ratios of a tile, fetched one pipeline stage ahead of their use
    float pf_prior = 0.f, pf_ra = 0.f, pf_rb = 0.f;
    auto prefetch = [&](int tile) {
        if (a.hyps != nullptr || n <= 0) return;
        const TileXY t = tile_xy(a, tile);
        const int xx = min(t.tx0 + lane, a.w - 1);
        pf_prior = __ldg(a.prior + static_cast<size_t>(t.b) * hw + t.y * a.w + xx);
        pf_ra = __ldg(a.ratio + static_cast<size_t>(t.b) * a.D + d0);
        pf_rb = __ldg(a.ratio + static_cast<size_t>(t.b) * a.D + d1 - 1);
    };

    int tile = blockIdx.x;
    if (tile < a.num_tiles) {
        prefetch(tile);
        footprint_stage(a, &map_src, &map_ref, smem, tile, 0, pf_prior, pf_ra, pf_rb);
    }
    __syncthreads();

    uint32_t phase = 0, sbuf = 0;
    for (int iter = 0; tile < a.num_tiles; tile += gridDim.x, ++iter) {
        const int par = iter & 1;
        const int next = tile + gridDim.x;
        const TileXY t = tile_xy(a, tile);
        const int b = t.b, y = t.y, tx0 = t.tx0, x = tx0 + lane;
        const bool lane_ok = x < a.w;
        const int pix = y * a.w + (lane_ok ? x : a.w - 1);
        const float prior_v = pf_prior;
        if (next < a.num_tiles) prefetch(next);
        const Geom c = pixel_geom(geo_all + b * GEO_STRIDE, x, y);
        const float* hp = (a.hyps != nullptr) ? a.hyps + static_cast<size_t>(b) * a.D * hw + pix : nullptr;
        const float* rp = (a.hyps == nullptr) ? a.ratio + static_cast<size_t>(b) * a.D : nullptr;

        // ---- this tile's box and this pixel's table footprint (union over the 8 chunks)
        int ox, oy;
        bool boxed, any_bad;
        tile_box(a, reinterpret_cast<const int*>(smem + Smem::OFF_RED) + par * WARPS * 8, ox, oy, boxed, any_bad);
        int fx0 = 32767, fx1 = -32768, fy0 = 32767, fy1 = -32768;
        {
            const short4* foot = reinterpret_cast<const short4*>(smem + Smem::OFF_FOOT) + par * WARPS * 32;
#pragma unroll
            for (int wi = 0; wi < WARPS; ++wi) {
                const short4 f = foot[wi * 32 + lane];
                fx0 = min(fx0, static_cast<int>(f.x));
                fx1 = max(fx1, static_cast<int>(f.y));
                fy0 = min(fy0, static_cast<int>(f.z));
                fy1 = max(fy1, static_cast<int>(f.w));
            }
        }
        const int nxp = fx1 - fx0 + 2, nyp = fy1 - fy0 + 2;
        const bool foot_empty = fx1 < fx0;
        const bool tab_ok = a.use_table && !any_bad && (foot_empty || nxp * nyp <= NT);
        const int ntaps = (tab_ok && !foot_empty && lane_ok) ? nxp * nyp : 0;

        mbar_wait(bar, phase);               // reference row (+ source box) of this tile have landed
        phase ^= 1u;

        // ---- build the table: warp w computes taps w, w+8 of every pixel
        if (__any_sync(0xffffffffu, warp < ntaps)) {
            ulonglong2 rl[4], rh[4];
            const uint64_t half2 = pk2(0.5f, 0.5f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                rl[j] = lds128p(rbox, lane, j);
                rh[j] = lds128p(rbox, lane, j + 4);
                rl[j].x = mul2(rl[j].x, half2);
                rl[j].y = mul2(rl[j].y, half2);
                rh[j].x = mul2(rh[j].x, half2);
                rh[j].y = mul2(rh[j].y, half2);
            }
            for (int tp = warp; tp < ntaps; tp += WARPS) {
                const int ty = tp / nxp, tx = tp - ty * nxp;
                const int sx = fx0 + tx, sy = fy0 + ty;
                const int rx = sx - ox, ry = sy - oy;
                const bool in_box = boxed && rx >= 0 && rx < BOX_W && ry >= 0 && ry < BOX_H;
                const bool in_img = sx >= 0 && sx < a.w && sy >= 0 && sy < a.h;
                const ulonglong2* g = reinterpret_cast<const ulonglong2*>(
                    a.src + (static_cast<size_t>(b * a.h + (in_img ? sy : 0)) * a.w + (in_img ? sx : 0)) * CV_C);
                ulonglong2* dst = reinterpret_cast<ulonglong2*>(ptab_pix + tp * 64);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ulonglong2 lo, hi;
                    if (in_box) {
                        lo = lds128p(sbox, ry * BOX_W + rx, j);
                        hi = lds128p(sbox, ry * BOX_W + rx, j + 4);
                    } else if (in_img) {
                        lo = __ldg(g + j);
                        hi = __ldg(g + j + 4);
                    } else {
                        lo = make_ulonglong2(0ull, 0ull);
                        hi = lo;
                    }
                    dst[j] = make_ulonglong2(fma2(rl[j].x, lo.x, mul2(rh[j].x, hi.x)), fma2(rl[j].y, lo.y, mul2(rh[j].y, hi.y)));
                }
            }
        }
        __syncthreads();                     // table complete; source box and reference row are free again

        // ---- stage F for the next tile: its loads fly while this tile's hypothesis loop runs
        if (next < a.num_tiles) footprint_stage(a, &map_src, &map_ref, smem, next, par ^ 1, pf_prior, pf_ra, pf_rb);

        if (n > 0) {                         // warp-uniform (idle chunk when D < 8 * DC)
            uint64_t P2[4][8];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt)
#pragma unroll
                for (int k = 0; k < 8; ++k) P2[tt][k] = 0ull;
            int cx = INT_MIN, cy = INT_MIN;
            float* obase = a.out + static_cast<size_t>(b) * CV_G * a.D * hw;
            const float cu = 0.5f * static_cast<float>(a.w - 1), ru = 0.5f * static_cast<float>(a.w + 1);
            const float cv = 0.5f * static_cast<float>(a.h - 1), rv = 0.5f * static_cast<float>(a.h + 1);
            float ratio_l = 0.f;             // lane l holds the ratio of hypothesis d0 + (i & ~31) + l

            for (int i = 0; i < n; ++i) {
                const int d = d0 + i;
                if (hp == nullptr && (i & 31) == 0) ratio_l = (i + lane < n) ? __ldg(rp + d + lane) : 0.f;
                const float depth = hp ? __ldg(hp + static_cast<size_t>(d) * hw) : prior_v * __shfl_sync(0xffffffffu, ratio_l, i & 31);
                // p = depth * (KT K^-1 x) + t ;  u = p.x / (p.z + 1e-7)   (movedepth/layers.py:601-621)
                const float inv = rcp_approx(fmaf(depth, c.mrz, c.tz) + 1e-7f);
                const float u = fmaf(depth, c.mrx, c.tx) * inv, v = fmaf(depth, c.mry, c.ty) * inv;
                // ATen grid_sampler_2d (bilinear, zeros, align_corners=True): taps outside the image contribute 0; a
                // sample at u <= -1 or u >= w has no in-image tap.  |u - (w-1)/2| < (w+1)/2  <=>  -1 < u < w.
                const bool ok = (fabsf(u - cu) < ru) && (fabsf(v - cv) < rv);
                const float x0 = floorf(u), y0 = floorf(v);
                const float wx1 = u - x0, wx0 = (x0 + 1.f) - u, wy1 = v - y0, wy0 = (y0 + 1.f) - v;
                const int ix = static_cast<int>(x0), iy = static_cast<int>(y0);
                if (ok && (ix != cx || iy != cy)) {
                    cx = ix;
                    cy = iy;
                    if (tab_ok && cx >= fx0 && cx <= fx1 && cy >= fy0 && cy <= fy1)
                        load_cell_table(ptab_pix, (cy - fy0) * nxp + (cx - fx0), nxp, P2);
                    else
                        load_cell_direct(a, b, b * hw + pix, cx, cy, P2);
                }
                const float a00 = ok ? wx0 * wy0 : 0.f, a01 = ok ? wx1 * wy0 : 0.f, a10 = ok ? wx0 * wy1 : 0.f,
                            a11 = ok ? wx1 * wy1 : 0.f;
                const uint64_t w00 = pk2(a00, a00), w01 = pk2(a01, a01), w10 = pk2(a10, a10), w11 = pk2(a11, a11);
                uint64_t o2[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    o2[k] = fma2(w11, P2[3][k], fma2(w10, P2[2][k], fma2(w01, P2[1][k], mul2(w00, P2[0][k]))));

                if (a.tma_store) {
                    float* buf = stage + sbuf * (STAGE_ROW / 4);
                    if (lane == 0) bulk_wait_read<1>();          // the store issued two hypotheses ago has drained this buffer
                    __syncwarp();
                    if (a.layout == MVD_LAYOUT_BGDHW) {           // buf[g][lane]
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float p, q;
                            unpk2(o2[k], p, q);
                            buf[(2 * k) * TW + lane] = p;
                            buf[(2 * k + 1) * TW + lane] = q;
                        }
                    } else {                                      // buf[lane][g], 64B-swizzled like the output descriptor
                        ulonglong2* bp = reinterpret_cast<ulonglong2*>(buf + lane * CV_G);
                        const int sw = (lane >> 1) & 3;
#pragma unroll
                        for (int q = 0; q < 4; ++q) bp[q ^ sw] = make_ulonglong2(o2[2 * q], o2[2 * q + 1]);
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0 && !a.dbg_nostore) {
                        if (a.layout == MVD_LAYOUT_BGDHW) tma_store_5d(&map_out, buf, tx0, y, d, 0, b);
                        else tma_store_5d(&map_out, buf, 0, tx0, y, d, b);
                        bulk_commit();
                    }
                    sbuf ^= 1u;
                } else if (lane_ok) {
                    if (a.layout == MVD_LAYOUT_BGDHW) {
                        const int off = d * hw + pix, gs = a.D * hw;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float p, q;
                            unpk2(o2[k], p, q);
                            obase[off + (2 * k) * gs] = p;
                            obase[off + (2 * k + 1) * gs] = q;
                        }
                    } else {
                        ulonglong2* op = reinterpret_cast<ulonglong2*>(obase + (static_cast<size_t>(d) * hw + pix) * CV_G);
#pragma unroll
                        for (int q = 0; q < 4; ++q) op[q] = make_ulonglong2(o2[2 * q], o2[2 * q + 1]);
                    }
                }
            }
        }
        __syncthreads();                     // everyone is done with this tile's table; next tile's footprints are visible
    }
    if (lane == 0) bulk_wait_read<0>();       // shared memory must outlive the in-flight stores
}

// ------------------------------------------------------------------------------------------ backward
__device__ __forceinline__ void flush_cell(const CvArgs& a, const PixelCtx& c, const unsigned char* sbox,
                                           const float (&rh)[CV_C], float (&Q)[4][CV_G], float (&gr)[CV_C], int cx,
                                           int cy) {
    int rel;
    const bool in_box = cell_in_box(c, cx, cy, rel);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int px = cx + (t & 1), py = cy + (t >> 1);
        const int rp = rel + (t & 1) + (t >> 1) * CV_BOX_W;
        if (px >= 0 && px < a.w && py >= 0 && py < a.h) {
            float4* gs = reinterpret_cast<float4*>(a.gsrc + (static_cast<size_t>(c.b * a.h + py) * a.w + px) * CV_C);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 lo, hi;
                fetch_tap(a, c, sbox, in_box, rp, px, py, j, lo, hi);
                const float q0 = Q[t][4 * j + 0], q1 = Q[t][4 * j + 1], q2 = Q[t][4 * j + 2], q3 = Q[t][4 * j + 3];
                gr[4 * j + 0] = fmaf(q0, lo.x, gr[4 * j + 0]);
                gr[4 * j + 1] = fmaf(q1, lo.y, gr[4 * j + 1]);
                gr[4 * j + 2] = fmaf(q2, lo.z, gr[4 * j + 2]);
                gr[4 * j + 3] = fmaf(q3, lo.w, gr[4 * j + 3]);
                gr[16 + 4 * j + 0] = fmaf(q0, hi.x, gr[16 + 4 * j + 0]);
                gr[16 + 4 * j + 1] = fmaf(q1, hi.y, gr[16 + 4 * j + 1]);
                gr[16 + 4 * j + 2] = fmaf(q2, hi.z, gr[16 + 4 * j + 2]);
                gr[16 + 4 * j + 3] = fmaf(q3, hi.w, gr[16 + 4 * j + 3]);
                atomicAdd(gs + j, make_float4(q0 * rh[4 * j + 0], q1 * rh[4 * j + 1], q2 * rh[4 * j + 2], q3 * rh[4 * j + 3]));
                atomicAdd(gs + j + 4, make_float4(q0 * rh[16 + 4 * j + 0], q1 * rh[16 + 4 * j + 1],
                                                  q2 * rh[16 + 4 * j + 2], q3 * rh[16 + 4 * j + 3]));
            }
        }
#pragma unroll
        for (int g = 0; g < CV_G; ++g) Q[t][g] = 0.f;
    }
}

__global__ void __launch_bounds__(CV_THREADS, 1)
costvol_grouped_bwd_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_ref,
                           const CvArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + CvSmem::OFF_BAR);
    const float* ratio_s = reinterpret_cast<const float*>(smem + CvSmem::OFF_RATIO);
    const unsigned char* sbox = smem + CvSmem::OFF_SRC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        tma_prefetch_desc(&map_src);
        tma_prefetch_desc(&map_ref);
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    uint32_t phase = 0;
    const size_t hw = static_cast<size_t>(a.h) * a.w;
    int iter = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++iter) {
        PixelCtx c;
        tile_prologue(a, &map_src, &map_ref, smem, tile, iter, phase, c);
        if (!c.active) continue;

        float rh[CV_C], gr[CV_C];
        load_half_ref(smem, warp, lane, rh);
#pragma unroll
        for (int k = 0; k < CV_C; ++k) gr[k] = 0.f;
        float Q[4][CV_G];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int g = 0; g < CV_G; ++g) Q[t][g] = 0.f;
        int cx = INT_MIN, cy = INT_MIN;

        const size_t pix = static_cast<size_t>(c.y) * a.w + c.x;
        const float prior_v = (a.hyps == nullptr) ? __ldg(a.prior + c.b * hw + pix) : 0.f;
        const float* hp = (a.hyps != nullptr) ? a.hyps + static_cast<size_t>(c.b) * a.D * hw + pix : nullptr;

        for (int d = c.d0; d < c.d1; ++d) {
            float go[CV_G];
            if (a.layout == MVD_LAYOUT_BGDHW) {
                const float* gp = a.gout + (static_cast<size_t>(c.b) * CV_G * a.D + d) * hw + pix;
                const size_t gs = static_cast<size_t>(a.D) * hw;
#pragma unroll
                for (int g = 0; g < CV_G; ++g) go[g] = __ldg(gp + g * gs);
            } else {
                const float4* gp = reinterpret_cast<const float4*>(a.gout + ((static_cast<size_t>(c.b) * a.D + d) * hw + pix) * CV_G);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t4 = __ldg(gp + q);
                    go[4 * q] = t4.x;
                    go[4 * q + 1] = t4.y;
                    go[4 * q + 2] = t4.z;
                    go[4 * q + 3] = t4.w;
                }
            }
            const float depth = hp ? __ldg(hp + d * hw) : prior_v * ratio_s[d];
            float u, v, pz;
            project_uv(c, depth, u, v, pz);
            const Bilinear s = bilinear_at(a, u, v);
            if (!s.ok) continue;
            if (s.ix != cx || s.iy != cy) {
                if (cx != INT_MIN) flush_cell(a, c, sbox, rh, Q, gr, cx, cy);
                cx = s.ix;
                cy = s.iy;
            }
#pragma unroll
            for (int g = 0; g < CV_G; ++g) {
                Q[0][g] = fmaf(s.w00, go[g], Q[0][g]);
                Q[1][g] = fmaf(s.w01, go[g], Q[1][g]);
                Q[2][g] = fmaf(s.w10, go[g], Q[2][g]);
                Q[3][g] = fmaf(s.w11, go[g], Q[3][g]);
            }
        }
        if (cx != INT_MIN) flush_cell(a, c, sbox, rh, Q, gr, cx, cy);
        // d out / d ref_c = 1/2 * sum_t w_t * src_tap_c ; chunks of the same pixel meet in global memory
        float4* grp = reinterpret_cast<float4*>(a.gref + (c.b * hw + pix) * CV_C);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            atomicAdd(grp + j, make_float4(0.5f * gr[4 * j], 0.5f * gr[4 * j + 1], 0.5f * gr[4 * j + 2], 0.5f * gr[4 * j + 3]));
    }
}

// ---- backward, v5: warp-autonomous ----------------------------------------------------------------
// No block barriers: a warp owns 16 consecutive pixels of a row and a chunk of hypotheses, a LANE PAIR owns a pixel
// (lane = 2*pixel + half, half h = groups 8h..8h+7 = the lane's 32 B of every hypothesis' gradient): 32 accumulator registers
// per thread (v2: 64 + ref + d ref = 253 registers, one 8-warp CTA per SM, one hypothesis in flight per thread).
//  * In the channels-last layout the gradients of a warp's 16 pixels for one hypothesis are ONE contiguous 1 KB run: each warp
//    streams them through its own shared-memory ring with bulk async copies (cp.async.bulk + mbarrier, issued by lane 0,
//    RING hypotheses ahead), so the loads cost no registers and no issue slots and 16 warps keep 100+ KB in flight per SM --
//    the v2 loop ran at the latency of one 64 B load per thread.  The [B,G,D,h,w] layout loads through registers, four
//    hypotheses at a time.
//  * A cell change flushes Q_tap[g] straight to global memory like v2 (vector atomics on d src, d ref partial in registers),
//    but a thread walks a long chunk of the pixel's epipolar segment, so cells -- and flushes -- per pixel are ~3x fewer than with
//    v2's 8 chunks; ref / src taps are read at flush time from the L1/L2.
namespace b5 {
constexpr int WARPS = 8, THREADS = 32 * WARPS;
constexpr int PIXW = 16;                              // pixels per warp
constexpr int GOB = 4;                                // register path: hypotheses loaded before the first is consumed
constexpr int RING = 8;                               // ring path: stages of 1 KB per warp
constexpr int SMEM = WARPS * RING * (PIXW * 64) + WARPS * RING * 8;

struct Task {
    int b, y, x0, d0, d1;
};
__device__ __forceinline__ Task task_of(const CvArgs& a, int t, int nch, int xblocks) {
    Task k;
    const int ch = t % nch;
    t /= nch;
    const int xb = t % xblocks;
    t /= xblocks;
    k.y = t % a.h;
    k.b = t / a.h;
    k.x0 = xb * PIXW;
    k.d0 = ch * a.DC;
    k.d1 = min(a.D, k.d0 + a.DC);
    return k;
}
// 1-D bulk async copy global -> shared::cta with byte-count completion on an mbarrier; 32-bit shared addresses throughout (the
// generic -> shared conversions stay out of the hypothesis loop)
__device__ __forceinline__ void bulk_load_s(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar_s) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(src),
                 "r"(bytes), "r"(bar_s)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_s, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t"
        "}" ::"r"(bar_s),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void lds_32B(uint32_t addr_s, uint64_t (&v)[4]) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "r"(addr_s) : "memory");
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2+16];" : "=l"(v[2]), "=l"(v[3]) : "r"(addr_s) : "memory");
}
}  // namespace b5

template <bool USE_RING, bool HAS_HYPS>
__global__ void __launch_bounds__(b5::THREADS, 2)
costvol_grouped_bwd_v5_kernel(const CvArgs a, int nch, int xblocks, int ntasks, float cu, float ru, float cv, float rv) {
    using namespace b5;
    using f3::rcp_approx;
    extern __shared__ __align__(128) unsigned char smem5[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane & 1;
    constexpr bool use_ring = USE_RING;
    const uint32_t ring_s = smem_u32(smem5 + warp * (RING * PIXW * 64));
    const uint32_t bars_s = smem_u32(smem5 + WARPS * RING * PIXW * 64) + warp * RING * 8;
    if (use_ring) {
        if (lane == 0) {
            for (int s0 = 0; s0 < RING; ++s0) mbar_init(reinterpret_cast<uint64_t*>(smem5 + WARPS * RING * PIXW * 64) + warp * RING + s0, 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    uint32_t st = 0, par = 0;                                // ring stage / parity of the next hypothesis to consume
    const int hw = a.h * a.w;
    for (int task = blockIdx.x * WARPS + warp; task < ntasks; task += gridDim.x * WARPS) {
        const Task k = task_of(a, task, nch, xblocks);
        const int b = k.b, x = k.x0 + (lane >> 1);
        const bool lane_ok = x < a.w;
        const int pix = k.y * a.w + (lane_ok ? x : a.w - 1);
        const int n = k.d1 - k.d0;
        const float* gbase = a.gout + static_cast<size_t>(b) * CV_G * a.D * hw;
        // ring: the warp's run of hypothesis d starts at pixel (y, x0): npx pixels of 64 B; lane 0 keeps the issue cursor
        const uint32_t run_bytes = static_cast<uint32_t>(min(PIXW, a.w - k.x0)) * 64u;
        const float* run_next = gbase + (static_cast<size_t>(k.d0) * hw + k.y * a.w + k.x0) * CV_G;
        const size_t run_stride = static_cast<size_t>(hw) * CV_G;
        int to_issue = n;
        if (use_ring && lane == 0) {
            uint32_t s1 = st;
            for (int j = 0; j < RING && j < n; ++j) {
                bulk_load_s(ring_s + s1 * (PIXW * 64), run_next, run_bytes, bars_s + s1 * 8);
                run_next += run_stride;
                --to_issue;
                s1 = (s1 + 1 == RING) ? 0 : s1 + 1;
            }
        }

        // p(depth) = depth * (K T)[:3,:3] inv_K[:3,:3] (x,y,1) + (K T)[:3,3]   (movedepth/layers.py:581-621), as the forward
        f3::Geom c;
        {
            float P[12];
            const float* Kb = a.K + b * 16;
            const float* Tb = a.T + b * 16;
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const int i = q >> 2, j = q & 3;
                float v = 0.f;
#pragma unroll
                for (int m = 0; m < 4; ++m) v = fmaf(__ldg(Kb + i * 4 + m), __ldg(Tb + m * 4 + j), v);
                P[q] = v;
            }
            const float* iK = a.invK + b * 16;
            const float xf = static_cast<float>(x), yf = static_cast<float>(k.y);
            const float rx = fmaf(__ldg(iK + 0), xf, fmaf(__ldg(iK + 1), yf, __ldg(iK + 2)));
            const float ry = fmaf(__ldg(iK + 4), xf, fmaf(__ldg(iK + 5), yf, __ldg(iK + 6)));
            const float rz = fmaf(__ldg(iK + 8), xf, fmaf(__ldg(iK + 9), yf, __ldg(iK + 10)));
            c.mrx = fmaf(P[0], rx, fmaf(P[1], ry, P[2] * rz));
            c.mry = fmaf(P[4], rx, fmaf(P[5], ry, P[6] * rz));
            c.mrz = fmaf(P[8], rx, fmaf(P[9], ry, P[10] * rz));
            c.tx = P[3];
            c.ty = P[7];
            c.tz = P[11];
        }
        const float prior_v = HAS_HYPS ? 0.f : __ldg(a.prior + static_cast<size_t>(b) * hw + pix);
        // depth of hypothesis d: hyps[b, d, pixel] or prior * ratio[b, d]; dp walks it with stride dstep
        const float* dp = HAS_HYPS ? a.hyps + (static_cast<size_t>(b) * a.D + k.d0) * hw + pix : a.ratio + static_cast<size_t>(b) * a.D + k.d0;
        const int dstep = HAS_HYPS ? hw : 1;
        const float4* rg = reinterpret_cast<const float4*>(a.ref + (static_cast<size_t>(b) * hw + pix) * CV_C) + 2 * half;

        uint64_t Q2[4][4];
#pragma unroll
        for (int tt = 0; tt < 4; ++tt)
#pragma unroll
            for (int q = 0; q < 4; ++q) Q2[tt][q] = 0ull;
        float gr[16];                                        // d ref of channels 8h..8h+7 and 16+8h..16+8h+7 (x2, applied at the end)
#pragma unroll
        for (int q = 0; q < 16; ++q) gr[q] = 0.f;
        int cx = INT_MIN, cy = INT_MIN;
        auto flush = [&]() {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int px = cx + (t & 1), py = cy + (t >> 1);
                float Q[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    unpk2(Q2[t][q], Q[2 * q], Q[2 * q + 1]);
                    Q2[t][q] = 0ull;
                }
                if (px < 0 || px >= a.w || py < 0 || py >= a.h) continue;         // zero padding: no source pixel behind this tap
                const size_t sp = (static_cast<size_t>(b * a.h + py) * a.w + px) * CV_C;
                const float4* g = reinterpret_cast<const float4*>(a.src + sp) + 2 * half;
                float4* gs = reinterpret_cast<float4*>(a.gsrc + sp) + 2 * half;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 rl = __ldg(rg + j), rh = __ldg(rg + j + 4), lo = __ldg(g + j), hi = __ldg(g + j + 4);
                    const float q0 = Q[4 * j], q1 = Q[4 * j + 1], q2 = Q[4 * j + 2], q3 = Q[4 * j + 3];
                    gr[4 * j + 0] = fmaf(q0, lo.x, gr[4 * j + 0]);
                    gr[4 * j + 1] = fmaf(q1, lo.y, gr[4 * j + 1]);
                    gr[4 * j + 2] = fmaf(q2, lo.z, gr[4 * j + 2]);
                    gr[4 * j + 3] = fmaf(q3, lo.w, gr[4 * j + 3]);
                    gr[8 + 4 * j + 0] = fmaf(q0, hi.x, gr[8 + 4 * j + 0]);
                    gr[8 + 4 * j + 1] = fmaf(q1, hi.y, gr[8 + 4 * j + 1]);
                    gr[8 + 4 * j + 2] = fmaf(q2, hi.z, gr[8 + 4 * j + 2]);
                    gr[8 + 4 * j + 3] = fmaf(q3, hi.w, gr[8 + 4 * j + 3]);
                    const float h0 = 0.5f * q0, h1 = 0.5f * q1, h2 = 0.5f * q2, h3 = 0.5f * q3;
                    atomicAdd(gs + j, make_float4(h0 * rl.x, h1 * rl.y, h2 * rl.z, h3 * rl.w));
                    atomicAdd(gs + j + 4, make_float4(h0 * rh.x, h1 * rh.y, h2 * rh.z, h3 * rh.w));
                }
            }
        };
        // one hypothesis: projection, cell bookkeeping, Q += w (x) gout
        auto consume = [&](float depth, bool valid, const uint64_t (&go)[4]) {
            const float inv = rcp_approx(fmaf(depth, c.mrz, c.tz) + 1e-7f);
            const float uu = fmaf(depth, c.mrx, c.tx) * inv, vv = fmaf(depth, c.mry, c.ty) * inv;
            const bool ok = lane_ok && valid && (fabsf(uu - cu) < ru) && (fabsf(vv - cv) < rv);
            if (ok) {
                const float x0 = floorf(uu), y0 = floorf(vv);
                const float wx1 = uu - x0, wx0 = (x0 + 1.f) - uu, wy1 = vv - y0, wy0 = (y0 + 1.f) - vv;
                const int ix = static_cast<int>(x0), iy = static_cast<int>(y0);
                if (ix != cx || iy != cy) {
                    if (cx != INT_MIN) flush();
                    cx = ix;
                    cy = iy;
                }
                const float a00 = wx0 * wy0, a01 = wx1 * wy0, a10 = wx0 * wy1, a11 = wx1 * wy1;
                const uint64_t w00 = pk2(a00, a00), w01 = pk2(a01, a01), w10 = pk2(a10, a10), w11 = pk2(a11, a11);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    Q2[0][q] = fma2(w00, go[q], Q2[0][q]);
                    Q2[1][q] = fma2(w01, go[q], Q2[1][q]);
                    Q2[2][q] = fma2(w10, go[q], Q2[2][q]);
                    Q2[3][q] = fma2(w11, go[q], Q2[3][q]);
                }
            }
        };

        if constexpr (use_ring) {
            const uint32_t lane_off = static_cast<uint32_t>(lane) * 32u;
#pragma unroll 1
            for (int i = 0; i < n; ++i) {
                const float dv = __ldg(dp);
                dp += dstep;
                const float depth = HAS_HYPS ? dv : prior_v * dv;
                mbar_wait_s(bars_s + st * 8, par);
                uint64_t go[4];
                lds_32B(ring_s + st * (PIXW * 64) + lane_off, go);   // lanes past the row end read stale bytes; never used (lane_ok)
                // The stage may be refilled only when every lane's loads have RETURNED: a warp barrier alone orders their issue,
                // not their completion, and a 1 KB copy from the L2 can land first (measured: ~7 % of runs had corrupted pixels in
                // the L2-resident part of the volume).  The vote consumes one register of each 16 B load, so it cannot execute
                // before the data is there; the proxy fence orders those generic-proxy reads before the async-proxy write.
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    ".reg .b32 r;\n\t"
                    ".reg .b64 t;\n\t"
                    "or.b64 t, %0, %1;\n\t"
                    "setp.ne.b64 p, t, 0;\n\t"
                    "vote.sync.ballot.b32 r, p, 0xffffffff;\n\t"
                    "}" ::"l"(go[0]),
                    "l"(go[2])
                    : "memory");
                if (lane == 0 && to_issue > 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    bulk_load_s(ring_s + st * (PIXW * 64), run_next, run_bytes, bars_s + st * 8);
                    run_next += run_stride;
                    --to_issue;
                }
                if (++st == RING) {
                    st = 0;
                    par ^= 1u;
                }
                consume(depth, true, go);
            }
        } else {
#pragma unroll 1
            for (int i0 = 0; i0 < n; i0 += GOB) {
                uint64_t go2[GOB][4];
                float depth[GOB];
#pragma unroll
                for (int u = 0; u < GOB; ++u) {
                    const int du = min(i0 + u, n - 1);
                    const float* gp = gbase + (static_cast<size_t>(8 * half) * a.D + k.d0 + du) * hw + pix;
                    const size_t gs = static_cast<size_t>(a.D) * hw;
#pragma unroll
                    for (int q = 0; q < 4; ++q) go2[u][q] = pk2(__ldg(gp + (2 * q) * gs), __ldg(gp + (2 * q + 1) * gs));
                    const float dv = __ldg(dp + static_cast<size_t>(du) * dstep);
                    depth[u] = HAS_HYPS ? dv : prior_v * dv;
                }
#pragma unroll
                for (int u = 0; u < GOB; ++u) consume(depth[u], i0 + u < n, go2[u]);
            }
        }
        if (cx != INT_MIN) flush();
        if (lane_ok) {                                       // chunks of one pixel meet in global memory
            float4* grp = reinterpret_cast<float4*>(a.gref + (static_cast<size_t>(b) * hw + pix) * CV_C) + 2 * half;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                atomicAdd(grp + j, make_float4(0.5f * gr[4 * j], 0.5f * gr[4 * j + 1], 0.5f * gr[4 * j + 2], 0.5f * gr[4 * j + 3]));
                atomicAdd(grp + j + 4, make_float4(0.5f * gr[8 + 4 * j], 0.5f * gr[8 + 4 * j + 1], 0.5f * gr[8 + 4 * j + 2], 0.5f * gr[8 + 4 * j + 3]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ host
static int costvol_grouped_launch(bool bwd, CvArgs a, int C, int G, int flags, cudaStream_t st) {
    MVD_REQUIRE(C == CV_C && G == CV_G, "grouped cost volume is built for C=32, G=16 (got C=%d, G=%d)", C, G);
    MVD_REQUIRE(a.B > 0 && a.h > 0 && a.w > 0 && a.D > 0, "empty shape B=%d h=%d w=%d D=%d", a.B, a.h, a.w, a.D);
    MVD_REQUIRE(a.D <= CV_MAXD, "D=%d exceeds the supported maximum %d", a.D, CV_MAXD);
    MVD_REQUIRE(a.h < 32000 && a.w < 32000, "feature map %dx%d too large", a.h, a.w);
    MVD_REQUIRE(a.layout == MVD_LAYOUT_BGDHW || a.layout == MVD_LAYOUT_BDHWG, "unknown out_layout %d", a.layout);
    MVD_REQUIRE(a.hyps != nullptr || (a.prior != nullptr && a.ratio != nullptr), "need hyps or (prior, ratio)");
    MVD_REQUIRE(aligned16(a.ref) && aligned16(a.src) && aligned16(a.out) && aligned16(a.gout) && aligned16(a.gref) &&
                    aligned16(a.gsrc),
                "feature / volume pointers must be 16-byte aligned");
    a.use_tma = (flags & MVD_FLAG_NO_TMA) ? 0 : 1;
    a.use_table = (flags & MVD_FLAG_NO_TABLE) ? 0 : 1;
    a.tiles_x = (a.w + CV_TW - 1) / CV_TW;

    CUtensorMap map_src, map_ref;
    int rc;
    if (bwd && (flags & MVD_FLAG_BWD_V2)) {          // the round-1 kernel (A/B measurements)
        a.tiles_y = (a.h + CV_R - 1) / CV_R;
        a.num_tiles = a.tiles_x * a.tiles_y * a.B;
        a.DC = (a.D + CV_NCH - 1) / CV_NCH;
        rc = make_nhwc32_tensor_map(&map_src, a.src, a.B, a.h, a.w, CV_BOX_W, CV_BOX_H);
        if (rc) return rc;
        rc = make_nhwc32_tensor_map(&map_ref, a.ref, a.B, a.h, a.w, CV_TW, CV_R);
        if (rc) return rc;
        const int smem = CvSmem::ALLOC;
        const int grid = min(a.num_tiles, sm_count());
        cudaError_t e = cudaMemsetAsync(a.gref, 0, sizeof(float) * a.B * a.h * a.w * CV_C, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(a.gsrc, 0, sizeof(float) * a.B * a.h * a.w * CV_C, st);
        if (e != cudaSuccess) return fail(static_cast<int>(e), "costvol bwd memset: %s", cudaGetErrorString(e));
        cudaFuncSetAttribute(costvol_grouped_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        costvol_grouped_bwd_kernel<<<grid, CV_THREADS, smem, st>>>(map_src, map_ref, a);
        return check_launch("costvol_grouped_bwd");
    }
    if (bwd) {                                       // v5: warp-autonomous
        const int nch = a.D >= 64 ? ((flags >> 8) & 15 ? (flags >> 8) & 15 : 3) : 1;      // bits 8..11: chunk count override (tuning)
        a.DC = (a.D + nch - 1) / nch;
        const int xblocks = (a.w + b5::PIXW - 1) / b5::PIXW;
        const long long ntasks = static_cast<long long>(a.B) * a.h * xblocks * nch;
        MVD_REQUIRE(ntasks < (1ll << 31), "too many tiles");
        const size_t gcount = static_cast<size_t>(a.B) * a.h * a.w * CV_C;
        cudaError_t e;
        if (a.gsrc == a.gref + gcount) {             // adjacent gradient buffers (what ops.py allocates): one memset node
            e = cudaMemsetAsync(a.gref, 0, sizeof(float) * 2 * gcount, st);
        } else {
            e = cudaMemsetAsync(a.gref, 0, sizeof(float) * gcount, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(a.gsrc, 0, sizeof(float) * gcount, st);
        }
        if (e != cudaSuccess) return fail(static_cast<int>(e), "costvol bwd memset: %s", cudaGetErrorString(e));
        const int blocks = static_cast<int>((ntasks + b5::WARPS - 1) / b5::WARPS);
        const int grid = min(blocks, sm_count() * 2);
        // |u - (w-1)/2| < (w+1)/2  <=>  -1 < u < w: the forward's in-image test, constants passed as kernel parameters
        const float cu = 0.5f * static_cast<float>(a.w - 1), ru = 0.5f * static_cast<float>(a.w + 1);
        const float cv = 0.5f * static_cast<float>(a.h - 1), rv = 0.5f * static_cast<float>(a.h + 1);
        const int nt = static_cast<int>(ntasks);
        const bool ring = a.layout == MVD_LAYOUT_BDHWG;      // channels-last gradient: 1 KB runs through the per-warp bulk-copy ring
        const size_t smem = ring ? b5::SMEM : 0;
#define MVD_LAUNCH_V5(R, H)                                                                                                    \
    do {                                                                                                                       \
        if (R) cudaFuncSetAttribute(costvol_grouped_bwd_v5_kernel<R, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, b5::SMEM); \
        costvol_grouped_bwd_v5_kernel<R, H><<<grid, b5::THREADS, smem, st>>>(a, nch, xblocks, nt, cu, ru, cv, rv);             \
    } while (0)
        if (ring && a.hyps) MVD_LAUNCH_V5(true, true);
        else if (ring) MVD_LAUNCH_V5(true, false);
        else if (a.hyps) MVD_LAUNCH_V5(false, true);
        else MVD_LAUNCH_V5(false, false);
#undef MVD_LAUNCH_V5
        return check_launch("costvol_grouped_bwd_v5");
    }
    // forward: the kernel keeps the geometry of every batch item in shared memory -> at most MAXB items per launch
    if (a.B > f3::MAXB) {
        for (int b0 = 0; b0 < a.B; b0 += f3::MAXB) {
            CvArgs s = a;
            s.B = min(f3::MAXB, a.B - b0);
            const size_t hw = static_cast<size_t>(a.h) * a.w;
            s.ref = a.ref + b0 * hw * CV_C;
            s.src = a.src + b0 * hw * CV_C;
            s.prior = a.prior ? a.prior + b0 * hw : nullptr;
            s.ratio = a.ratio ? a.ratio + static_cast<size_t>(b0) * a.D : nullptr;
            s.hyps = a.hyps ? a.hyps + b0 * hw * a.D : nullptr;
            s.K = a.K + b0 * 16;
            s.invK = a.invK + b0 * 16;
            s.T = a.T + b0 * 16;
            s.out = a.out + b0 * hw * a.D * CV_G;
            const int r = costvol_grouped_launch(false, s, C, G, flags, st);
            if (r) return r;
        }
        return 0;
    }
    // one tile = one row of 32 pixels, 8 hypothesis chunks
    a.tiles_y = a.h;
    a.num_tiles = a.tiles_x * a.h * a.B;
    a.DC = (a.D + f3::NCH - 1) / f3::NCH;
    rc = make_nhwc32_tensor_map(&map_src, a.src, a.B, a.h, a.w, f3::BOX_W, f3::BOX_H);
    if (rc) return rc;
    rc = make_nhwc32_tensor_map(&map_ref, a.ref, a.B, a.h, a.w, f3::TW, 1);
    if (rc) return rc;
    // output descriptor for the per-warp TMA stores (one hypothesis x one tile row x 16 groups per store)
    CUtensorMap map_out;
    const uint64_t W = a.w, H = a.h, Dd = a.D;
    a.tma_store = a.use_tma && !(flags & MVD_FLAG_PLAIN_STORE);
    a.dbg_nostore = (flags & MVD_FLAG_DBG_NO_STORE) ? 1 : 0;
    if (a.layout == MVD_LAYOUT_BGDHW) {
        if (a.w % 4 != 0) a.tma_store = 0;      // TMA needs 16-byte global strides; ragged widths use plain stores
        const uint64_t dims[5] = {W, H, Dd, CV_G, static_cast<uint64_t>(a.B)};
        const uint64_t str[4] = {W * 4, W * H * 4, W * H * Dd * 4, W * H * Dd * CV_G * 4};
        const uint32_t box[5] = {f3::TW, 1, 1, CV_G, 1};
        if (a.tma_store) rc = make_f32_tensor_map(&map_out, a.out, 5, dims, str, box, 0);
    } else {
        const uint64_t dims[5] = {CV_G, W, H, Dd, static_cast<uint64_t>(a.B)};
        const uint64_t str[4] = {CV_G * 4, W * CV_G * 4, W * H * CV_G * 4, W * H * Dd * CV_G * 4};
        const uint32_t box[5] = {CV_G, f3::TW, 1, 1, 1};
        rc = make_f32_tensor_map(&map_out, a.out, 5, dims, str, box, 64);   // 64B swizzle: conflict-free staging writes
    }
    if (rc) return rc;
    if (!a.tma_store) map_out = map_ref;        // unused by the kernel, but must be a valid descriptor
    const int smem = f3::Smem::ALLOC;
    const int grid = min(a.num_tiles, sm_count() * 2);
    bool attr_done = false;       // per call: the attribute is per device, a process-wide latch is not
    if (!attr_done) {
        cudaFuncSetAttribute(costvol_grouped_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_done = true;
    }
    costvol_grouped_fwd_kernel<<<grid, f3::THREADS, smem, st>>>(map_src, map_ref, map_out, a);
    return check_launch("costvol_grouped_fwd");
}

// ---- reference-layout volume [B,D,C,h,w] (public generate_costvol API; any C, NCHW) ---------
struct FullGeom {
    float mrx, mry, mrz, tx, ty, tz;
};
__device__ __forceinline__ FullGeom full_geom(const float* K, const float* invK, const float* T, int b, int x, int y) {
    const float* Kb = K + b * 16;
    const float* Tb = T + b * 16;
    const float* Ib = invK + b * 16;
    float P[12];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = fmaf(Kb[i * 4 + k], Tb[k * 4 + j], s);
            P[i * 4 + j] = s;
        }
    const float xf = static_cast<float>(x), yf = static_cast<float>(y);
    const float rx = fmaf(Ib[0], xf, fmaf(Ib[1], yf, Ib[2]));
    const float ry = fmaf(Ib[4], xf, fmaf(Ib[5], yf, Ib[6]));
    const float rz = fmaf(Ib[8], xf, fmaf(Ib[9], yf, Ib[10]));
    FullGeom g;
    g.mrx = fmaf(P[0], rx, fmaf(P[1], ry, P[2] * rz));
    g.mry = fmaf(P[4], rx, fmaf(P[5], ry, P[6] * rz));
    g.mrz = fmaf(P[8], rx, fmaf(P[9], ry, P[10] * rz));
    g.tx = P[3];
    g.ty = P[7];
    g.tz = P[11];
    return g;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
costvol_full_kernel(const float* __restrict__ ref, const float* __restrict__ src, const float* __restrict__ hyps,
                    const float* __restrict__ K, const float* __restrict__ invK, const float* __restrict__ T,
                    float* __restrict__ out, const float* __restrict__ gout, float* __restrict__ gref,
                    float* __restrict__ gsrc, int B, int C, int h, int w, int D) {
    const size_t hw = static_cast<size_t>(h) * w;
    const size_t total = static_cast<size_t>(B) * D * hw;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(idx % w);
        const int y = static_cast<int>((idx / w) % h);
        const int d = static_cast<int>((idx / hw) % D);
        const int b = static_cast<int>(idx / (hw * D));
        const FullGeom g = full_geom(K, invK, T, b, x, y);
        const float depth = __ldg(hyps + idx);
        const float pz = fmaf(depth, g.mrz, g.tz) + 1e-7f;
        const float inv = __frcp_rn(pz);
        const float u = fmaf(depth, g.mrx, g.tx) * inv, v = fmaf(depth, g.mry, g.ty) * inv;
        const bool ok = (u > -1.f) && (u < static_cast<float>(w)) && (v > -1.f) && (v < static_cast<float>(h));
        float wt[4] = {0.f, 0.f, 0.f, 0.f};
        int off[4] = {0, 0, 0, 0};
        bool inb[4] = {false, false, false, false};
        if (ok) {
            const float x0 = floorf(u), y0 = floorf(v);
            const float fx1 = u - x0, fx0 = (x0 + 1.f) - u, fy1 = v - y0, fy0 = (y0 + 1.f) - v;
            wt[0] = fx0 * fy0;
            wt[1] = fx1 * fy0;
            wt[2] = fx0 * fy1;
            wt[3] = fx1 * fy1;
            const int ix = static_cast<int>(x0), iy = static_cast<int>(y0);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int px = ix + (t & 1), py = iy + (t >> 1);
                inb[t] = px >= 0 && px < w && py >= 0 && py < h;
                off[t] = py * w + px;
            }
        }
        const size_t pix = static_cast<size_t>(y) * w + x;
        for (int c = 0; c < C; ++c) {
            const float* sp = src + (static_cast<size_t>(b) * C + c) * hw;
            float tap[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) tap[t] = inb[t] ? __ldg(sp + off[t]) : 0.f;
            const float warped = fmaf(tap[3], wt[3], fmaf(tap[2], wt[2], fmaf(tap[1], wt[1], tap[0] * wt[0])));
            const size_t rix = (static_cast<size_t>(b) * C + c) * hw + pix;
            const size_t oix = ((static_cast<size_t>(b) * D + d) * C + c) * hw + pix;
            if (!BWD) {
                out[oix] = warped * __ldg(ref + rix);
            } else {
                const float go = __ldg(gout + oix);
                atomicAdd(gref + rix, go * warped);
                const float gw = go * __ldg(ref + rix);
                float* gp = gsrc + (static_cast<size_t>(b) * C + c) * hw;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (inb[t]) atomicAdd(gp + off[t], gw * wt[t]);
            }
        }
    }
}

}  // namespace mvd

extern "C" {

int mvd_costvol_grouped_fwd(const float* ref, const float* src, const float* prior, const float* ratio,
                            const float* hyps, const float* K, const float* invK, const float* T, float* out, int B,
                            int C, int G, int h, int w, int D, int out_layout, int flags, void* stream) {
    MVD_REQUIRE(ref && src && K && invK && T && out, "null pointer argument");
    mvd::CvArgs a{};
    a.ref = ref; a.src = src; a.prior = prior; a.ratio = ratio; a.hyps = hyps;
    a.K = K; a.invK = invK; a.T = T; a.out = out;
    a.B = B; a.h = h; a.w = w; a.D = D; a.layout = out_layout;
    return mvd::costvol_grouped_launch(false, a, C, G, flags, mvd::as_stream(stream));
}

int mvd_costvol_grouped_bwd(const float* gout, const float* ref, const float* src, const float* prior,
                            const float* ratio, const float* hyps, const float* K, const float* invK, const float* T,
                            float* gref, float* gsrc, int B, int C, int G, int h, int w, int D, int out_layout,
                            int flags, void* stream) {
    MVD_REQUIRE(gout && ref && src && K && invK && T && gref && gsrc, "null pointer argument");
    mvd::CvArgs a{};
    a.ref = ref; a.src = src; a.prior = prior; a.ratio = ratio; a.hyps = hyps;
    a.K = K; a.invK = invK; a.T = T; a.gout = gout; a.gref = gref; a.gsrc = gsrc;
    a.B = B; a.h = h; a.w = w; a.D = D; a.layout = out_layout;
    return mvd::costvol_grouped_launch(true, a, C, G, flags, mvd::as_stream(stream));
}

int mvd_costvol_full_fwd(const float* ref, const float* src, const float* hyps, const float* K, const float* invK,
                         const float* T, float* out, int B, int C, int h, int w, int D, void* stream) {
    MVD_REQUIRE(ref && src && hyps && K && invK && T && out, "null pointer argument");
    MVD_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && D > 0, "empty shape");
    const size_t total = static_cast<size_t>(B) * D * h * w;
    const int grid = static_cast<int>(min(static_cast<size_t>(mvd::sm_count()) * 8, (total + 255) / 256));
    mvd::costvol_full_kernel<false><<<grid, 256, 0, mvd::as_stream(stream)>>>(ref, src, hyps, K, invK, T, out, nullptr,
                                                                               nullptr, nullptr, B, C, h, w, D);
    return mvd::check_launch("costvol_full_fwd");
}

int mvd_costvol_full_bwd(const float* gout, const float* ref, const float* src, const float* hyps, const float* K,
                         const float* invK, const float* T, float* gref, float* gsrc, int B, int C, int h, int w, int D,
                         void* stream) {
    MVD_REQUIRE(gout && ref && src && hyps && K && invK && T && gref && gsrc, "null pointer argument");
    MVD_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && D > 0, "empty shape");
    cudaStream_t st = mvd::as_stream(stream);
    const size_t fbytes = sizeof(float) * B * C * h * w;
    cudaError_t e = cudaMemsetAsync(gref, 0, fbytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(gsrc, 0, fbytes, st);
    if (e != cudaSuccess) return mvd::fail(static_cast<int>(e), "costvol_full_bwd memset: %s", cudaGetErrorString(e));
    const size_t total = static_cast<size_t>(B) * D * h * w;
    const int grid = static_cast<int>(min(static_cast<size_t>(mvd::sm_count()) * 8, (total + 255) / 256));
    mvd::costvol_full_kernel<true><<<grid, 256, 0, st>>>(ref, src, hyps, K, invK, T, nullptr, gout, gref, gsrc, B, C, h,
                                                          w, D);
    return mvd::check_launch("costvol_full_bwd");
}

}
