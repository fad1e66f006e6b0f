// reg3d's output head `prob` = Conv3d(16 -> 1, 3x3x3, pad 1, no bias) on the full-resolution
// channels-last volume (movedepth/networks/resnet_encoder.py:254, 279), forward + dgrad + wgrad.
//
// At BASELINE config 2 the layer reads a 283 MB activation ([6,96,48,160,16] fp32) and produces
// 17.7 MB of logits: 432 MACs per output, i.e. an HBM-bound stencil, not a GEMM -- a tensor-core
// implicit GEMM with N = 1 wastes the MMA tile (cuDNN: 1.2 ms forward, 0.5 ms dgrad + 0.7 ms of
// layout conversions, 2.6 ms wgrad).  Here all three passes are exact-fp32 CUDA-core kernels:
//   * the activation is staged slice by slice (one depth slice of a 8x32 tile + halo = 10x34
//     positions x 64 B) by TMA with the 64B swizzle, out-of-bounds zero fill == the conv's zero
//     padding; a 3-deep mbarrier ring keeps two slices in flight;
//   * each thread marches along the depth axis with the three partial sums of the outputs the
//     current slice contributes to (kd = 0,1,2) in registers, so every activation byte is read
//     from shared memory once per (kh,kw) tap and from HBM ~1.3 times in total;
//   * the 432 weights sit in constant memory: every FMA takes its weight as a constant operand.
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {
namespace c16 {

constexpr int C = 16;
constexpr int TH = 8, TW = 32;                 // output tile (rows x columns) per depth slice
constexpr int HH = TH + 2, HW = TW + 2;        // with the 3x3 halo
constexpr int SLICE_POS = HH * HW;             // 340 positions of 64 B
constexpr int SLICE_BYTES = 22016;             // 340 * 64 rounded up to the 512 B swizzle period
constexpr int NBUF = 3;
constexpr int THREADS = TH * TW;               // 256
constexpr int NW = C * 27;                     // 432 weights, index c*27 + kd*9 + kh*3 + kw

__constant__ float c_w[NW];

struct Args {
    const float* x;
    const float* gy;
    float* y;
    float* gx;
    float* part;
    int B, D, H, W;
    int tiles_h, tiles_w, dsplit, dlen;
};

struct Item {
    int b, d0, d1, h0, w0;
};
__device__ __forceinline__ Item decode_item(const Args& a, int item) {
    Item it;
    const int tw = item % a.tiles_w;
    int r = item / a.tiles_w;
    const int th = r % a.tiles_h;
    r /= a.tiles_h;
    const int dc = r % a.dsplit;
    it.b = r / a.dsplit;
    it.d0 = dc * a.dlen;
    it.d1 = min(a.D, it.d0 + a.dlen);
    it.h0 = th * TH;
    it.w0 = tw * TW;
    return it;
}

// 5-D TMA tiled load global -> shared::cta, completion on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}

// channels [4q, 4q+4) of halo-tile position p (64B swizzle: 16 B chunk index ^= address bits [7:8])
__device__ __forceinline__ float4 lds_x(const unsigned char* buf, int p, int q) {
    return *reinterpret_cast<const float4*>(buf + p * 64 + ((q ^ ((p >> 1) & 3)) << 4));
}

__device__ __forceinline__ void issue_slice(const CUtensorMap* map, unsigned char* buf, uint64_t* bar, const Item& it, int s) {
    mbar_expect_tx(bar, SLICE_POS * 64);
    tma_load_5d(buf, map, bar, 0, it.w0 - 1, it.h0 - 1, s, it.b);
}

// ------------------------------------------------------------------------------------------ forward
// y[b,d,h,w] = sum_{kd,kh,kw,c} x[b,d+kd-1,h+kh-1,w+kw-1,c] * W[c,kd,kh,kw]
__global__ void __launch_bounds__(THREADS, 2)
conv3d_c16o1_fwd_kernel(const __grid_constant__ CUtensorMap map_x, const Args a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NBUF * SLICE_BYTES);
    const int tid = threadIdx.x, hh = tid >> 5, ww = tid & 31;
    const Item it = decode_item(a, blockIdx.x);
    const int count = it.d1 - it.d0 + 2;                       // x slices d0-1 .. d1
    if (tid == 0) {
        tma_prefetch_desc(&map_x);
        for (int i = 0; i < NBUF; ++i) mbar_init(full + i, 1);
        mbar_fence_init();
        for (int i = 0; i < NBUF && i < count; ++i) issue_slice(&map_x, smem + i * SLICE_BYTES, full + i, it, it.d0 - 1 + i);
    }
    __syncthreads();
    const int h = it.h0 + hh, w = it.w0 + ww;
    const bool inside = h < a.H && w < a.W;
    float* yp = a.y + (static_cast<size_t>(it.b) * a.D * a.H + h) * a.W + w;       // + d*H*W
    const size_t dstride = static_cast<size_t>(a.H) * a.W;
    float acc_prev = 0.f, acc_cur = 0.f, acc_next = 0.f;       // outputs d = s-1 (kd=2), s (kd=1), s+1 (kd=0)
    for (int i = 0; i < count; ++i) {
        const int s = it.d0 - 1 + i, bi = i % NBUF;
        const unsigned char* buf = smem + bi * SLICE_BYTES;
        mbar_wait(full + bi, (i / NBUF) & 1);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int p = (hh + kh) * HW + ww + kw, t = kh * 3 + kw;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = lds_x(buf, p, q);
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = 4 * q + k;
                        acc_prev = fmaf(e[k], c_w[c * 27 + 18 + t], acc_prev);
                        acc_cur = fmaf(e[k], c_w[c * 27 + 9 + t], acc_cur);
                        acc_next = fmaf(e[k], c_w[c * 27 + t], acc_next);
                    }
                }
            }
        if (s - 1 >= it.d0 && inside) yp[(s - 1) * dstride] = acc_prev;
        acc_prev = acc_cur;
        acc_cur = acc_next;
        acc_next = 0.f;
        __syncthreads();                                       // everyone is done with this buffer
        if (tid == 0 && i + NBUF < count) issue_slice(&map_x, smem + bi * SLICE_BYTES, full + bi, it, s + NBUF);
    }
}

// ------------------------------------------------------------------------------------------ dgrad
// gx[b,d,h,w,c] = sum_{kd,kh,kw} gy[b,d-kd+1,h-kh+1,w-kw+1] * W[c,kd,kh,kw]
constexpr int GY_POS = HH * HW;                // one halo'd gy slice (floats)

__device__ __forceinline__ void load_gy_slice(const Args& a, const Item& it, int d, float* dst, int tid, int dlo, int dhi) {
    // rows h0-1 .. h0+8, columns w0-1 .. w0+32; zero outside the image / outside [dlo, dhi)
    const bool dok = d >= dlo && d < dhi;
    for (int e = tid; e < GY_POS; e += THREADS) {
        const int r = e / HW, cc = e - r * HW;
        const int h = it.h0 - 1 + r, w = it.w0 - 1 + cc;
        float v = 0.f;
        if (dok && h >= 0 && h < a.H && w >= 0 && w < a.W)
            v = __ldg(a.gy + ((static_cast<size_t>(it.b) * a.D + d) * a.H + h) * a.W + w);
        dst[e] = v;
    }
}

__global__ void __launch_bounds__(THREADS, 2)
conv3d_c16o1_dgrad_kernel(const Args a) {
    __shared__ float gys[4][GY_POS];                           // ring of gy slices (slot = (d + 1) & 3)
    __shared__ __align__(16) float stage[THREADS / 32][32 * C];
    const int tid = threadIdx.x, hh = tid >> 5, ww = tid & 31, warp = hh, lane = ww;
    const Item it = decode_item(a, blockIdx.x);
    load_gy_slice(a, it, it.d0 - 1, gys[(it.d0) & 3], tid, 0, a.D);
    load_gy_slice(a, it, it.d0, gys[(it.d0 + 1) & 3], tid, 0, a.D);
    const int h = it.h0 + hh;
    for (int d = it.d0; d < it.d1; ++d) {
        load_gy_slice(a, it, d + 1, gys[(d + 2) & 3], tid, 0, a.D);
        __syncthreads();
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
            const float* g = gys[(d - kd + 2) & 3];            // slice d - kd + 1
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float v = g[(hh + 2 - kh) * HW + ww + 2 - kw];   // (h - kh + 1, w - kw + 1) in halo coordinates
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[c] = fmaf(v, c_w[c * 27 + kd * 9 + kh * 3 + kw], acc[c]);
                }
        }
        // transpose through shared memory so that the warp writes its 2 KB row segment with fully coalesced 16 B stores
        float4* st = reinterpret_cast<float4*>(stage[warp]);
        const int sw = (lane >> 1) & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) st[lane * 4 + (q ^ sw)] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        __syncwarp();
        if (h < a.H) {
            float4* gp = reinterpret_cast<float4*>(a.gx + (((static_cast<size_t>(it.b) * a.D + d) * a.H + h) * a.W + it.w0) * C);
            const int valid = min(TW, a.W - it.w0) * 4;        // float4 chunks of this row segment inside the image
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = lane + 32 * j, pos = chunk >> 2, q = chunk & 3;
                if (chunk < valid) gp[chunk] = st[pos * 4 + (q ^ ((pos >> 1) & 3))];
            }
        }
        __syncthreads();                                       // gy ring slot (d - 1 + 1) & 3 is rewritten next iteration
    }
}

// ------------------------------------------------------------------------------------------ wgrad
// gw[c,kd,kh,kw] = sum_{b,d,h,w} gy[b,d,h,w] * x[b,d+kd-1,h+kh-1,w+kw-1,c]
// thread = (channel quad c4, kd, segment of 17 halo-tile positions); it walks its segment of the x slice with
// the 3x3 (kh,kw) window of gy values in registers: 36 FMAs per 16-byte shared-memory load.
constexpr int GP_H = TH + 4, GP_W = TW + 4;    // gy tile zero-padded by 2 on every side
constexpr int GP_POS = GP_H * GP_W;
constexpr int NSEG = 20, SEG_LEN = 17;

__device__ __forceinline__ void load_gy_padded(const Args& a, const Item& it, int d, float* dst, int tid) {
    const bool dok = d >= it.d0 && d < it.d1;                  // only this work item's own outputs count
    for (int e = tid; e < GP_POS; e += THREADS) {
        const int r = e / GP_W - 2, cc = e % GP_W - 2;
        const int h = it.h0 + r, w = it.w0 + cc;
        float v = 0.f;
        if (dok && r >= 0 && r < TH && cc >= 0 && cc < TW && h < a.H && w < a.W)
            v = __ldg(a.gy + ((static_cast<size_t>(it.b) * a.D + d) * a.H + h) * a.W + w);
        dst[e] = v;
    }
}

__global__ void __launch_bounds__(THREADS, 2)
conv3d_c16o1_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const Args a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NBUF * SLICE_BYTES);
    float* gys = reinterpret_cast<float*>(smem + NBUF * SLICE_BYTES + 64);          // [4][GP_POS], slot = (d + 1) & 3
    const int tid = threadIdx.x;
    const Item it = decode_item(a, blockIdx.x);
    const int count = it.d1 - it.d0 + 2;
    if (tid == 0) {
        tma_prefetch_desc(&map_x);
        for (int i = 0; i < NBUF; ++i) mbar_init(full + i, 1);
        mbar_fence_init();
        for (int i = 0; i < NBUF && i < count; ++i) issue_slice(&map_x, smem + i * SLICE_BYTES, full + i, it, it.d0 - 1 + i);
    }
    // x slice s pairs with gy slices s+1 (kd=0), s (kd=1), s-1 (kd=2)
    load_gy_padded(a, it, it.d0 - 2, gys + ((it.d0 - 1) & 3) * GP_POS, tid);
    load_gy_padded(a, it, it.d0 - 1, gys + ((it.d0) & 3) * GP_POS, tid);
    const int c4 = tid & 3, kd = (tid >> 2) % 3, seg = tid / 12;
    const bool active = seg < NSEG;
    const int hq = seg >> 1, wq0 = (seg & 1) * SEG_LEN;
    float acc[3][3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

    for (int i = 0; i < count; ++i) {
        const int s = it.d0 - 1 + i, bi = i % NBUF;
        const unsigned char* buf = smem + bi * SLICE_BYTES;
        load_gy_padded(a, it, s + 1, gys + ((s + 2) & 3) * GP_POS, tid);
        __syncthreads();
        mbar_wait(full + bi, (i / NBUF) & 1);
        if (active) {
            // gy tile-local index for x halo position (hq,wq) and tap (kh,kw): row hq-kh, column wq-kw (+2 padding)
            const float* g = gys + ((s - kd + 2) & 3) * GP_POS;     // slice s - kd + 1
            float win[3][3];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                win[kh][1] = g[(hq - kh + 2) * GP_W + wq0 - 1 + 2];
                win[kh][2] = g[(hq - kh + 2) * GP_W + wq0 - 2 + 2];
            }
#pragma unroll 1
            for (int j = 0; j < SEG_LEN; ++j) {
                const int wq = wq0 + j;
                const float4 xv = lds_x(buf, hq * HW + wq, c4);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    win[kh][0] = g[(hq - kh + 2) * GP_W + wq + 2];  // kw = 0: column wq
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float gv = win[kh][kw];
                        acc[kh][kw][0] = fmaf(gv, xv.x, acc[kh][kw][0]);
                        acc[kh][kw][1] = fmaf(gv, xv.y, acc[kh][kw][1]);
                        acc[kh][kw][2] = fmaf(gv, xv.z, acc[kh][kw][2]);
                        acc[kh][kw][3] = fmaf(gv, xv.w, acc[kh][kw][3]);
                    }
                    win[kh][2] = win[kh][1];
                    win[kh][1] = win[kh][0];
                }
            }
        }
        __syncthreads();
        if (tid == 0 && i + NBUF < count) issue_slice(&map_x, smem + bi * SLICE_BYTES, full + bi, it, s + NBUF);
    }
    // ---- reduce the 20 segment partials of this work item (re-using the x ring as scratch)
    float* part = reinterpret_cast<float*>(smem);              // [NSEG][NW]
    if (active) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int k = 0; k < 4; ++k) part[seg * NW + (c4 * 4 + k) * 27 + kd * 9 + kh * 3 + kw] = acc[kh][kw][k];
    }
    __syncthreads();
    for (int o = tid; o < NW; o += THREADS) {
        float sum = 0.f;
#pragma unroll 4
        for (int sgi = 0; sgi < NSEG; ++sgi) sum += part[sgi * NW + o];
        a.part[static_cast<size_t>(blockIdx.x) * NW + o] = sum;
    }
}

// deterministic final reduction over the work items
__global__ void __launch_bounds__(128) conv3d_c16o1_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw,
                                                                         int items) {
    const int o = blockIdx.x, lane = threadIdx.x;
    float s = 0.f;
    for (int i = lane; i < items; i += 128) s += part[static_cast<size_t>(i) * NW + o];
    __shared__ float red[4];
    s = warp_sum(s);
    if ((lane & 31) == 0) red[lane >> 5] = s;
    __syncthreads();
    if (lane == 0) gw[o] = red[0] + red[1] + red[2] + red[3];
}

// ------------------------------------------------------------------------------------------ host
static int plan(Args& a) {
    a.tiles_h = (a.H + TH - 1) / TH;
    a.tiles_w = (a.W + TW - 1) / TW;
    const int base = a.B * a.tiles_h * a.tiles_w;
    const int slots = 2 * sm_count();
    int ds = 1;
    while (base * ds < 2 * slots && a.D / (ds + 1) >= 8) ++ds;  // enough work items for ~2 waves, >= 8 slices each
    a.dsplit = ds;
    a.dlen = (a.D + ds - 1) / ds;
    a.dsplit = (a.D + a.dlen - 1) / a.dlen;
    return base * a.dsplit;
}

static int make_x_map(CUtensorMap* map, const float* x, const Args& a) {
    const uint64_t W = a.W, H = a.H, D = a.D;
    const uint64_t dims[5] = {C, W, H, D, static_cast<uint64_t>(a.B)};
    const uint64_t str[4] = {C * 4, W * C * 4, W * H * C * 4, W * H * D * C * 4};
    const uint32_t box[5] = {C, HW, HH, 1, 1};
    return make_f32_tensor_map(map, x, 5, dims, str, box, 64);
}

static int check_shape(int B, int D, int H, int W) {
    MVD_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "empty shape B=%d D=%d H=%d W=%d", B, D, H, W);
    return 0;
}

static int upload_weights(const float* w, cudaStream_t st) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_w, w, sizeof(float) * NW, 0, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "conv3d_c16o1 weight upload: %s", cudaGetErrorString(e));
    return 0;
}

constexpr int FWD_SMEM = NBUF * SLICE_BYTES + 64 + 1024;
constexpr int WGRAD_SMEM = NBUF * SLICE_BYTES + 64 + 4 * GP_POS * 4 + 1024;
static_assert(NSEG * NW * 4 <= NBUF * SLICE_BYTES, "wgrad scratch fits the x ring");

}  // namespace c16
}  // namespace mvd

extern "C" {

int mvd_conv3d_c16o1_fwd(const float* x, const float* w, float* y, int B, int D, int H, int W, void* stream) {
    using namespace mvd::c16;
    MVD_REQUIRE(x && w && y, "null pointer argument");
    if (int rc = check_shape(B, D, H, W)) return rc;
    MVD_REQUIRE(mvd::aligned16(x), "x must be 16-byte aligned");
    cudaStream_t st = mvd::as_stream(stream);
    Args a{};
    a.x = x; a.y = y; a.B = B; a.D = D; a.H = H; a.W = W;
    const int items = plan(a);
    CUtensorMap map;
    if (int rc = make_x_map(&map, x, a)) return rc;
    if (int rc = upload_weights(w, st)) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv3d_c16o1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
        attr_done = true;
    }
    conv3d_c16o1_fwd_kernel<<<items, THREADS, FWD_SMEM, st>>>(map, a);
    return mvd::check_launch("conv3d_c16o1_fwd");
}

int mvd_conv3d_c16o1_dgrad(const float* gy, const float* w, float* gx, int B, int D, int H, int W, void* stream) {
    using namespace mvd::c16;
    MVD_REQUIRE(gy && w && gx, "null pointer argument");
    if (int rc = check_shape(B, D, H, W)) return rc;
    MVD_REQUIRE(mvd::aligned16(gx), "gx must be 16-byte aligned");
    cudaStream_t st = mvd::as_stream(stream);
    Args a{};
    a.gy = gy; a.gx = gx; a.B = B; a.D = D; a.H = H; a.W = W;
    const int items = plan(a);
    if (int rc = upload_weights(w, st)) return rc;
    conv3d_c16o1_dgrad_kernel<<<items, THREADS, 0, st>>>(a);
    return mvd::check_launch("conv3d_c16o1_dgrad");
}

long long mvd_conv3d_c16o1_wgrad_workspace_bytes(int B, int D, int H, int W) {
    using namespace mvd::c16;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    Args a{};
    a.B = B; a.D = D; a.H = H; a.W = W;
    return static_cast<long long>(plan(a)) * NW * sizeof(float);
}

int mvd_conv3d_c16o1_wgrad(const float* gy, const float* x, float* gw, void* workspace, long long workspace_bytes, int B,
                           int D, int H, int W, void* stream) {
    using namespace mvd::c16;
    MVD_REQUIRE(gy && x && gw && workspace, "null pointer argument");
    if (int rc = check_shape(B, D, H, W)) return rc;
    MVD_REQUIRE(mvd::aligned16(x), "x must be 16-byte aligned");
    cudaStream_t st = mvd::as_stream(stream);
    Args a{};
    a.x = x; a.gy = gy; a.part = static_cast<float*>(workspace); a.B = B; a.D = D; a.H = H; a.W = W;
    const int items = plan(a);
    MVD_REQUIRE(workspace_bytes >= static_cast<long long>(items) * NW * 4, "workspace too small: %lld < %lld", workspace_bytes,
                static_cast<long long>(items) * NW * 4);
    CUtensorMap map;
    if (int rc = make_x_map(&map, x, a)) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv3d_c16o1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WGRAD_SMEM);
        attr_done = true;
    }
    conv3d_c16o1_wgrad_kernel<<<items, THREADS, WGRAD_SMEM, st>>>(map, a);
    if (int rc = mvd::check_launch("conv3d_c16o1_wgrad")) return rc;
    conv3d_c16o1_wgrad_reduce_kernel<<<NW, 128, 0, st>>>(a.part, gw, items);
    return mvd::check_launch("conv3d_c16o1_wgrad_reduce");
}

}
