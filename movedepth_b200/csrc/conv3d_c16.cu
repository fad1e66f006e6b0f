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

#include <stdlib.h>

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
__constant__ __align__(16) float c_wt[NW];         // the same weights tap-major, index tap*16 + c: channel pairs for the packed FMAs
__device__ float g_wt[NW];                          // staging for c_wt (written by transpose_w_kernel)

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

// the same 16 B as two packed channel pairs, through an explicit shared-window address (a generic pointer costs a generic LD)
__device__ __forceinline__ ulonglong2 lds_x2(uint32_t buf_s, int p, int q) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(buf_s + p * 64 + ((q ^ ((p >> 1) & 3)) << 4)));
    return v;
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
    // Packed fp32 math: an accumulator is the pair (even channels' sum, odd channels' sum); one FFMA2 multiplies a channel pair
    // of x (a 64-bit half of the 16 B shared-memory load) by the weight pair held in a uniform register (LDCU from c_wt):
    // 216 FFMA2 + 108 LDCU per output instead of 432 FFMA.
    const uint64_t* w2 = reinterpret_cast<const uint64_t*>(c_wt);   // [tap][8 channel pairs]
    const uint32_t smem_s = smem_u32(smem);
    uint64_t acc_prev = 0ull, acc_cur = 0ull, acc_next = 0ull;      // outputs d = s-1 (kd=2), s (kd=1), s+1 (kd=0)
    for (int i = 0; i < count; ++i) {
        const int s = it.d0 - 1 + i, bi = i % NBUF;
        const uint32_t buf_s = smem_s + bi * SLICE_BYTES;
        mbar_wait(full + bi, (i / NBUF) & 1);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int p = (hh + kh) * HW + ww + kw, t = kh * 3 + kw;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const ulonglong2 v = lds_x2(buf_s, p, q);
                    acc_prev = fma2(v.x, w2[(18 + t) * 8 + 2 * q], acc_prev);
                    acc_prev = fma2(v.y, w2[(18 + t) * 8 + 2 * q + 1], acc_prev);
                    acc_cur = fma2(v.x, w2[(9 + t) * 8 + 2 * q], acc_cur);
                    acc_cur = fma2(v.y, w2[(9 + t) * 8 + 2 * q + 1], acc_cur);
                    acc_next = fma2(v.x, w2[t * 8 + 2 * q], acc_next);
                    acc_next = fma2(v.y, w2[t * 8 + 2 * q + 1], acc_next);
                }
            }
        if (s - 1 >= it.d0 && inside) {
            float even, odd;
            unpk2(acc_prev, even, odd);
            yp[(s - 1) * dstride] = even + odd;
        }
        acc_prev = acc_cur;
        acc_cur = acc_next;
        acc_next = 0ull;
        __syncthreads();                                       // everyone is done with this buffer
        if (tid == 0 && i + NBUF < count) issue_slice(&map_x, smem + bi * SLICE_BYTES, full + bi, it, s + NBUF);
    }
}

// ------------------------------------------------------------------------------------------ dgrad
// gx[b,d,h,w,c] = sum_{kd,kh,kw} gy[b,d-kd+1,h-kh+1,w-kw+1] * W[c,kd,kh,kw]
constexpr int GY_POS = HH * HW;                // one halo'd gy slice (floats)

// one halo'd gy slice = 340 floats = up to 2 per thread: rows h0-1 .. h0+8, columns w0-1 .. w0+32; zero outside the image /
// outside [dlo, dhi).  Split into the global loads (issued a whole slice ahead of their use) and the shared-memory stores.
struct GyRegs {
    float v[2];
};
__device__ __forceinline__ GyRegs fetch_gy_slice(const Args& a, const Item& it, int d, int tid, int dlo, int dhi) {
    GyRegs g;
    const bool dok = d >= dlo && d < dhi;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int e = tid + k * THREADS;
        const int r = e / HW, cc = e - r * HW;
        const int h = it.h0 - 1 + r, w = it.w0 - 1 + cc;
        g.v[k] = 0.f;
        if (e < GY_POS && dok && h >= 0 && h < a.H && w >= 0 && w < a.W)
            g.v[k] = __ldg(a.gy + ((static_cast<size_t>(it.b) * a.D + d) * a.H + h) * a.W + w);
    }
    return g;
}
__device__ __forceinline__ void store_gy_slice(const GyRegs& g, float* dst, int tid) {
    dst[tid] = g.v[0];
    if (tid + THREADS < GY_POS) dst[tid + THREADS] = g.v[1];
}
static_assert(GY_POS <= 2 * THREADS, "two values per thread cover a gy slice");

__global__ void __launch_bounds__(THREADS, 2)
conv3d_c16o1_dgrad_kernel(const Args a) {
    __shared__ float gys[4][GY_POS];                           // ring of gy slices (slot = (d + 1) & 3)
    __shared__ __align__(16) float stage[THREADS / 32][32 * C];
    const int tid = threadIdx.x, hh = tid >> 5, ww = tid & 31, warp = hh, lane = ww;
    const Item it = decode_item(a, blockIdx.x);
    store_gy_slice(fetch_gy_slice(a, it, it.d0 - 1, tid, 0, a.D), gys[(it.d0) & 3], tid);
    store_gy_slice(fetch_gy_slice(a, it, it.d0, tid, 0, a.D), gys[(it.d0 + 1) & 3], tid);
    GyRegs ahead = fetch_gy_slice(a, it, it.d0 + 1, tid, 0, a.D);
    const int h = it.h0 + hh;
    for (int d = it.d0; d < it.d1; ++d) {
        // slice d+1 (requested during the previous iteration) goes to its slot; slice d+2 is requested now.  One barrier per
        // slice: the slot written here, (d+2)&3, is read by nobody in iteration d-1 (slots (d-1)&3, d&3, (d+1)&3).
        store_gy_slice(ahead, gys[(d + 2) & 3], tid);
        ahead = fetch_gy_slice(a, it, d + 2, tid, 0, a.D);
        __syncthreads();
        const uint64_t* w2 = reinterpret_cast<const uint64_t*>(c_wt);   // [tap][8 channel pairs]
        uint64_t acc2[C / 2];                                  // packed channel pairs: 216 FFMA2 per position instead of 432 FFMA
#pragma unroll
        for (int c = 0; c < C / 2; ++c) acc2[c] = 0ull;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
            const float* g = gys[(d - kd + 2) & 3];            // slice d - kd + 1
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float v = g[(hh + 2 - kh) * HW + ww + 2 - kw];   // (h - kh + 1, w - kw + 1) in halo coordinates
                    const uint64_t vv = pk2(v, v);
#pragma unroll
                    for (int c = 0; c < C / 2; ++c) acc2[c] = fma2(vv, w2[(kd * 9 + kh * 3 + kw) * 8 + c], acc2[c]);
                }
        }
        float acc[C];
#pragma unroll
        for (int c = 0; c < C / 2; ++c) unpk2(acc2[c], acc[2 * c], acc[2 * c + 1]);
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
    }
}

// ------------------------------------------------------------------------------------------ wgrad
// gw[c,kd,kh,kw] = sum_{b,d,h,w} gy[b,d,h,w] * x[b,d+kd-1,h+kh-1,w+kw-1,c]
// thread = (channel quad c4, kd, segment of 17 halo-tile positions); it walks its segment of the x slice with
// the 3x3 (kh,kw) window of gy values in registers: 36 FMAs per 16-byte shared-memory load.
constexpr int GP_H = TH + 4, GP_W = TW + 4;    // gy tile zero-padded by 2 on every side
constexpr int GP_POS = GP_H * GP_W;
constexpr int NSEG = 20, SEG_LEN = 17;

// the zero-padded gy tile of one slice = 432 floats = up to 2 per thread, split into global loads (a slice ahead) and stores
__device__ __forceinline__ GyRegs fetch_gy_padded(const Args& a, const Item& it, int d, int tid) {
    GyRegs g;
    const bool dok = d >= it.d0 && d < it.d1;                  // only this work item's own outputs count
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int e = tid + k * THREADS;
        const int r = e / GP_W - 2, cc = e % GP_W - 2;
        const int h = it.h0 + r, w = it.w0 + cc;
        g.v[k] = 0.f;
        if (e < GP_POS && dok && r >= 0 && r < TH && cc >= 0 && cc < TW && h < a.H && w < a.W)
            g.v[k] = __ldg(a.gy + ((static_cast<size_t>(it.b) * a.D + d) * a.H + h) * a.W + w);
    }
    return g;
}
__device__ __forceinline__ void store_gy_padded(const GyRegs& g, float* dst, int tid) {
    dst[tid] = g.v[0];
    if (tid + THREADS < GP_POS) dst[tid + THREADS] = g.v[1];
}
static_assert(GP_POS <= 2 * THREADS, "two values per thread cover a padded gy tile");

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
    store_gy_padded(fetch_gy_padded(a, it, it.d0 - 2, tid), gys + ((it.d0 - 1) & 3) * GP_POS, tid);
    store_gy_padded(fetch_gy_padded(a, it, it.d0 - 1, tid), gys + ((it.d0) & 3) * GP_POS, tid);
    GyRegs ahead = fetch_gy_padded(a, it, it.d0, tid);
    const int c4 = tid & 3, kd = (tid >> 2) % 3, seg = tid / 12;
    const bool active = seg < NSEG;
    const int hq = seg >> 1, wq0 = (seg & 1) * SEG_LEN;
    uint64_t acc2[3][3][2];                                    // packed (x,y) / (z,w) channel pairs of the thread's channel quad
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) acc2[i][j][k] = 0ull;

    for (int i = 0; i < count; ++i) {
        const int s = it.d0 - 1 + i, bi = i % NBUF;
        const uint32_t buf_s = smem_u32(smem) + bi * SLICE_BYTES;
        // gy slice s+1 (requested during the previous iteration) goes to its slot, slice s+2 is requested now; after the barrier
        // everybody has also finished with the x buffer of iteration i-1, which thread 0 refills (one barrier per slice)
        store_gy_padded(ahead, gys + ((s + 2) & 3) * GP_POS, tid);
        ahead = fetch_gy_padded(a, it, s + 2, tid);
        __syncthreads();
        if (tid == 0 && i >= 1 && i - 1 + NBUF < count) issue_slice(&map_x, smem + ((i - 1) % NBUF) * SLICE_BYTES, full + (i - 1) % NBUF, it, s - 1 + NBUF);
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
#pragma unroll 4                                                     // the loads of the next positions overlap this one's FMAs
            for (int j = 0; j < SEG_LEN; ++j) {
                const int wq = wq0 + j;
                const ulonglong2 xv = lds_x2(buf_s, hq * HW + wq, c4);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    win[kh][0] = g[(hq - kh + 2) * GP_W + wq + 2];  // kw = 0: column wq
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const uint64_t gv = pk2(win[kh][kw], win[kh][kw]);
                        acc2[kh][kw][0] = fma2(gv, xv.x, acc2[kh][kw][0]);
                        acc2[kh][kw][1] = fma2(gv, xv.y, acc2[kh][kw][1]);
                    }
                    win[kh][2] = win[kh][1];
                    win[kh][1] = win[kh][0];
                }
            }
        }
    }
    __syncthreads();
    // ---- reduce the 20 segment partials of this work item (re-using the x ring as scratch)
    float* part = reinterpret_cast<float*>(smem);              // [NSEG][NW]
    if (active) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                float acc[4];
                unpk2(acc2[kh][kw][0], acc[0], acc[1]);
                unpk2(acc2[kh][kw][1], acc[2], acc[3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) part[seg * NW + (c4 * 4 + k) * 27 + kd * 9 + kh * 3 + kw] = acc[k];
            }
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

// ================================================================================================
// reg3d's first layer: Conv3d(16 -> 16, 3x3x3, pad 1) on the full-resolution volume, as an implicit GEMM on
// the warp-level tensor-core path (mma.sync m16n8k8 TF32; the tcgen05 version of this layer is the next step,
// DESIGN.md section 7).  One kernel serves the forward (3xTF32 operand split in registers: hi*hi + lo*hi + hi*lo,
// the precision policy of movedepth_b200/precision.py) and the data gradient (single-pass TF32 with the
// flipped / transposed filter).  Same slice-marching structure as the 16->1 head: the current input slice of a
// 8x32 tile (+halo) sits in shared memory (TMA, 64B swizzle), each warp owns 2 rows x 32 columns = four m16
// tiles and keeps the accumulators of the three output slices the input slice contributes to in registers,
// so an A fragment (16 B per lane, both k-steps) is loaded once per (kh,kw) and used for the three kd taps.
// The K (channel) order inside a k-step is permuted so that a lane's four channels are contiguous.
constexpr int G_WARPS = 4, G_THREADS = 32 * G_WARPS;
constexpr int BF_BYTES = 27 * 2 * 32 * 16;          // B fragments: [tap][k-step][lane] x float4

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct GArgs {
    const float* w;      // [16][16][27]  (co, ci, tap)
    float* out;          // [B,D,H,W,16]
    int mode;            // 0: forward, 1: data gradient
    int B, D, H, W;
    int tiles_h, tiles_w, dsplit, dlen;
};

template <int PASSES>
struct GState {
    float acc[3][4][2][4];
};

template <int PASSES, int R>
__device__ __forceinline__ void igemm_slice(GState<PASSES>& st, const unsigned char* buf, const float4* bfrag, int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            uint32_t ahi[4][2][4], alo[PASSES == 3 ? 4 : 1][2][4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int row = 2 * warp + (m >> 1), col = (m & 1) * 16;
                const int p0 = (row + kh) * HW + col + kw + g;
                const float4 v = lds_x(buf, p0, t), v2 = lds_x(buf, p0 + 8, t);
                const float e[2][4] = {{v.x, v2.x, v.y, v2.y}, {v.z, v2.z, v.w, v2.w}};
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        ahi[m][ks][q] = to_tf32(e[ks][q]);
                        if (PASSES == 3) alo[m][ks][q] = to_tf32(e[ks][q] - __uint_as_float(ahi[m][ks][q]));
                    }
            }
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                const int tap = kd * 9 + kh * 3 + kw;
                constexpr int dummy = 0;
                (void)dummy;
                const int slot = (R + 1 - kd + 3) % 3;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const float4 bf = bfrag[(tap * 2 + ks) * 32 + lane];
                    const float bv[4] = {bf.x, bf.y, bf.z, bf.w};
                    uint32_t bhi[4], blo[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        bhi[q] = to_tf32(bv[q]);
                        if (PASSES == 3) blo[q] = to_tf32(bv[q] - __uint_as_float(bhi[q]));
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int n = 0; n < 2; ++n) {
                            mma_tf32(st.acc[slot][m][n], ahi[m][ks], bhi[2 * n], bhi[2 * n + 1]);
                            if (PASSES == 3) {
                                mma_tf32(st.acc[slot][m][n], alo[m][ks], bhi[2 * n], bhi[2 * n + 1]);
                                mma_tf32(st.acc[slot][m][n], ahi[m][ks], blo[2 * n], blo[2 * n + 1]);
                            }
                        }
                }
            }
        }
}

// write the finished output slice (accumulator set `slot`) and clear it
template <int PASSES>
__device__ __forceinline__ void igemm_store(GState<PASSES>& st, int slot, const GArgs& a, const Item& it, int d, int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int h = it.h0 + 2 * warp + (m >> 1), w0 = it.w0 + (m & 1) * 16 + g;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            float (&c)[4] = slot == 0 ? st.acc[0][m][n] : (slot == 1 ? st.acc[1][m][n] : st.acc[2][m][n]);
            if (d >= it.d0 && h < a.H) {
                float* op = a.out + (((static_cast<size_t>(it.b) * a.D + d) * a.H + h) * a.W) * C + n * 8 + 2 * t;
                if (w0 < a.W) *reinterpret_cast<float2*>(op + static_cast<size_t>(w0) * C) = make_float2(c[0], c[1]);
                if (w0 + 8 < a.W) *reinterpret_cast<float2*>(op + static_cast<size_t>(w0 + 8) * C) = make_float2(c[2], c[3]);
            }
            c[0] = c[1] = c[2] = c[3] = 0.f;
        }
    }
}

template <int PASSES>
__global__ void __launch_bounds__(G_THREADS, 2)
conv3d_c16c16_kernel(const __grid_constant__ CUtensorMap map_in, const GArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float4* bfrag = reinterpret_cast<float4*>(smem + NBUF * SLICE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NBUF * SLICE_BYTES + BF_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Args pa{};
    pa.B = a.B; pa.D = a.D; pa.H = a.H; pa.W = a.W;
    pa.tiles_h = a.tiles_h; pa.tiles_w = a.tiles_w; pa.dsplit = a.dsplit; pa.dlen = a.dlen;
    const Item it = decode_item(pa, blockIdx.x);
    const int count = it.d1 - it.d0 + 2;
    if (tid == 0) {
        tma_prefetch_desc(&map_in);
        for (int i = 0; i < NBUF; ++i) mbar_init(full + i, 1);
        mbar_fence_init();
        for (int i = 0; i < NBUF && i < count; ++i) issue_slice(&map_in, smem + i * SLICE_BYTES, full + i, it, it.d0 - 1 + i);
    }
    // B fragments.  lane (g,t), k-step ks: k = t -> channel 4t+2ks, k = t+4 -> channel 4t+2ks+1; n = g (tile 0), g+8 (tile 1)
    for (int e = tid; e < 27 * 2 * 32; e += G_THREADS) {
        const int ln = e & 31, ks = (e >> 5) & 1, tap = e >> 6;
        const int g = ln >> 2, t = ln & 3;
        const int k0 = 4 * t + 2 * ks, k1 = k0 + 1;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = (q & 1) ? k1 : k0, n = g + ((q >> 1) ? 8 : 0);
            // forward: B[k=ci][n=co] = W[co][ci][tap];  data gradient: B[k=co][n=ci] = W[co][ci][26 - tap]
            v[q] = (a.mode == 0) ? __ldg(a.w + (n * 16 + k) * 27 + tap) : __ldg(a.w + (k * 16 + n) * 27 + (26 - tap));
        }
        bfrag[e] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();

    GState<PASSES> st;
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3)
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int q = 0; q < 4; ++q) st.acc[s3][m][n][q] = 0.f;

    // slice i (x slice s = d0-1+i) adds tap kd to output slice s+1-kd, kept in accumulator set (i+1-kd) mod 3;
    // afterwards output slice s-1 (set (i+2) mod 3) is complete.
    for (int i0 = 0; i0 < count; i0 += 3) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int i = i0 + r;
            if (i < count) {
                const int s = it.d0 - 1 + i, bi = i % NBUF;
                mbar_wait(full + bi, (i / NBUF) & 1);
                const unsigned char* buf = smem + bi * SLICE_BYTES;
                if (r == 0) igemm_slice<PASSES, 0>(st, buf, bfrag, warp, lane);
                else if (r == 1) igemm_slice<PASSES, 1>(st, buf, bfrag, warp, lane);
                else igemm_slice<PASSES, 2>(st, buf, bfrag, warp, lane);
                igemm_store<PASSES>(st, (r + 2) % 3, a, it, s - 1, warp, lane);
                __syncthreads();
                if (tid == 0 && i + NBUF < count) issue_slice(&map_in, smem + bi * SLICE_BYTES, full + bi, it, s + NBUF);
            }
        }
    }
}

constexpr int G_SMEM = NBUF * SLICE_BYTES + BF_BYTES + 64 + 1024;

// ================================================================================================
// Weight gradient of the 16 -> 16 layer on the CUDA cores, exact fp32 with packed FFMA2:
//   gw[co,ci,kd,kh,kw] = sum_{b,d,h,w} gy[b,d,h,w,co] * x[b,d+kd-1,h+kh-1,w+kw-1,ci]
// (measured here: mma.sync TF32 m16n8k8 sustains ~130 MAC/clk/SM on this part, no more than the FP32 pipe, and
// cuDNN's legacy sm80 wgrad kernel needs 2.6 ms for this layer.)
// thread role = (4 co) x (4 ci) x (kd,kh); it keeps the 3 kw taps x 16 products in 24 packed accumulators and
// walks 4 rows x 32 columns of the tile per depth slice: 2 LDS.128 + 24 FFMA2 per position.  x slices march
// through a 2-deep TMA ring, gy slices through a 4-deep one (x slice s pairs with gy slice s+1-kd).
constexpr int WG_ROLES = 144, WG_GROUPS = 2, WG_THREADS = WG_ROLES * WG_GROUPS;   // 288
constexpr int WG_XBUF = 2, WG_GBUF = 4;
constexpr int GY_BYTES = TH * TW * 64;                // 16 KB, multiple of 512
constexpr int NW16 = 16 * 16 * 27;                    // 6912

__device__ __forceinline__ uint64_t pk2f(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ ulonglong2 lds_x2(const unsigned char* buf, int p, int q) {
    return *reinterpret_cast<const ulonglong2*>(buf + p * 64 + ((q ^ ((p >> 1) & 3)) << 4));
}

struct WArgs {
    float* part;         // [items][6912]
    int B, D, H, W;
    int tiles_h, tiles_w, dsplit, dlen;
};

__global__ void __launch_bounds__(WG_THREADS, 2)
conv3d_c16c16_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_gy, const WArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* xs = smem;                                      // [WG_XBUF][SLICE_BYTES]
    unsigned char* gs = smem + WG_XBUF * SLICE_BYTES;              // [WG_GBUF][GY_BYTES]
    uint64_t* xfull = reinterpret_cast<uint64_t*>(gs + WG_GBUF * GY_BYTES);
    uint64_t* gfull = xfull + WG_XBUF;
    const int tid = threadIdx.x;
    Args pa{};
    pa.B = a.B; pa.D = a.D; pa.H = a.H; pa.W = a.W;
    pa.tiles_h = a.tiles_h; pa.tiles_w = a.tiles_w; pa.dsplit = a.dsplit; pa.dlen = a.dlen;
    const Item it = decode_item(pa, blockIdx.x);
    const int count = it.d1 - it.d0 + 2;                           // x slices d0-1 .. d1

    auto issue_gy = [&](int d) {                                   // gy slice d (only this item's own slices are ever used)
        if (d >= it.d0 && d < it.d1) {
            uint64_t* bar = gfull + (d & 3);
            mbar_expect_tx(bar, GY_BYTES);
            tma_load_5d(gs + (d & 3) * GY_BYTES, &map_gy, bar, 0, it.w0, it.h0, d, it.b);
        }
    };
    if (tid == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_gy);
        for (int i = 0; i < WG_XBUF; ++i) mbar_init(xfull + i, 1);
        for (int i = 0; i < WG_GBUF; ++i) mbar_init(gfull + i, 1);
        mbar_fence_init();
        for (int i = 0; i < WG_XBUF && i < count; ++i) issue_slice(&map_x, xs + i * SLICE_BYTES, xfull + i, it, it.d0 - 1 + i);
        issue_gy(it.d0);
        issue_gy(it.d0 + 1);
    }
    __syncthreads();

    const int grp = tid / WG_ROLES, role = tid - grp * WG_ROLES;
    const int co4 = role & 3, ci4 = (role >> 2) & 3, kdkh = role >> 4;
    const int kd = kdkh / 3, kh = kdkh - kd * 3;
    uint64_t acc[3][4][2];                                         // [kw][co][ci pair]
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0ull;

    for (int i = 0; i < count; ++i) {
        const int s = it.d0 - 1 + i, xb = i % WG_XBUF;
        mbar_wait(xfull + xb, (i / WG_XBUF) & 1);
        const int d = s + 1 - kd;                                  // the gy slice this thread's kd pairs with x slice s
        if (d >= it.d0 && d < it.d1) {
            const int first = it.d0 + (((d & 3) - (it.d0 & 3)) & 3);   // first slice of this item that used ring slot d&3
            mbar_wait(gfull + (d & 3), ((d - first) >> 2) & 1);
            const unsigned char* xbuf = xs + xb * SLICE_BYTES;
            const unsigned char* gbuf = gs + (d & 3) * GY_BYTES;
#pragma unroll 1
            for (int rr = 0; rr < 4; ++rr) {
                const int r = grp * 4 + rr;                        // tile row
                const int xrow = (r + kh) * HW;                    // halo row r+kh; halo column = c + kw
                ulonglong2 xw0 = lds_x2(xbuf, xrow, ci4), xw1 = lds_x2(xbuf, xrow + 1, ci4);
#pragma unroll 4
                for (int c = 0; c < TW; ++c) {
                    const ulonglong2 xw2 = lds_x2(xbuf, xrow + c + 2, ci4);
                    const float4 g = lds_x(gbuf, r * TW + c, co4);
                    const uint64_t g2[4] = {pk2f(g.x, g.x), pk2f(g.y, g.y), pk2f(g.z, g.z), pk2f(g.w, g.w)};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[0][j][0] = ffma2(g2[j], xw0.x, acc[0][j][0]);
                        acc[0][j][1] = ffma2(g2[j], xw0.y, acc[0][j][1]);
                        acc[1][j][0] = ffma2(g2[j], xw1.x, acc[1][j][0]);
                        acc[1][j][1] = ffma2(g2[j], xw1.y, acc[1][j][1]);
                        acc[2][j][0] = ffma2(g2[j], xw2.x, acc[2][j][0]);
                        acc[2][j][1] = ffma2(g2[j], xw2.y, acc[2][j][1]);
                    }
                    xw0 = xw1;
                    xw1 = xw2;
                }
            }
        }
        __syncthreads();                                           // x buffer xb and gy slice s-1 are free
        if (tid == 0) {
            if (i + WG_XBUF < count) issue_slice(&map_x, xs + xb * SLICE_BYTES, xfull + xb, it, s + WG_XBUF);
            issue_gy(s + 3);
        }
    }
    // ---- reduce the two position groups, write this item's partial
    float* part = reinterpret_cast<float*>(smem);                  // [WG_GROUPS][6912] over the dead rings
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[kw][j][pr]));
                const int co = co4 * 4 + j, ci = ci4 * 4 + 2 * pr, tap = kd * 9 + kh * 3 + kw;
                part[grp * NW16 + (co * 16 + ci) * 27 + tap] = lo;
                part[grp * NW16 + (co * 16 + ci + 1) * 27 + tap] = hi;
            }
    __syncthreads();
    for (int o = tid; o < NW16; o += WG_THREADS) a.part[static_cast<size_t>(blockIdx.x) * NW16 + o] = part[o] + part[NW16 + o];
}

__global__ void __launch_bounds__(128) conv3d_c16c16_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw,
                                                                          int items) {
    const int o = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;    // one warp per output
    if (o >= NW16) return;
    float s = 0.f;
    for (int i = lane; i < items; i += 32) s += part[static_cast<size_t>(i) * NW16 + o];
    s = warp_sum(s);
    if (lane == 0) gw[o] = s;
}

constexpr int WG_SMEM = WG_XBUF * SLICE_BYTES + WG_GBUF * GY_BYTES + 64 + 1024;
static_assert(WG_GROUPS * NW16 * 4 <= WG_XBUF * SLICE_BYTES + WG_GBUF * GY_BYTES, "wgrad scratch fits the rings");

// ================================================================================================
// The 16 -> 16 layer on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) with TMEM accumulators.
//   D[128 positions x N] += A[128 x 16 in-ch] * B[16 x N]
// * The input slice of a 7x32 tile (+halo = 9x34 positions x 64 B) is TMA-loaded with the 64B swizzle.  Output
//   positions are flattened with the halo pitch, v = r*34 + c, so that EVERY (kh,kw) tap's A operand is the same
//   shared-memory tile at a start address shifted by (kh*34 + kw) rows -- a K-major SWIZZLE_64B canonical layout (rows
//   64 B apart, 8-row groups 512 B apart); the two garbage columns per row are computed and dropped (2 M-blocks of
//   128 rows cover the 238 flattened positions).
// * Slice marching: input slice s adds tap kd to output slice s+1-kd.  With N = 16 an MMA is bound by the shared-memory
//   read of its A operand (4 KB per 8 tensor cycles), so the three kd taps -- same A, different filters, different
//   output slices -- are folded into ONE MMA of N = 48: the filter bank is stored per (kh,kw) as 48 K-major rows
//   [kd][out-ch], and the accumulators of consecutive output slices sit in consecutive TMEM column groups of a ring of
//   four (slice d -> group (-d) & 3), so slices s+1, s, s-1 are a contiguous column range (split in two MMAs when the
//   range wraps).  Every MMA accumulates; a group is zeroed (tcgen05.st) right after the epilogue drained it.
// * 3xTF32 (forward): the slice is split in place into hi = x & ~0x1fff and lo = x - hi (own buffer).  A_hi meets the
//   interleaved bank [kd][hi|lo][out-ch] (N = 96: hi*hi and hi*lo land in adjacent 16-column halves of a 32-column
//   group), A_lo meets the hi bank (N = 48, second ring); the epilogue adds the three partial sums.  A is read from
//   shared memory twice per (kh,kw) instead of nine times.
// * One thread issues the MMAs; tcgen05.commit signals an mbarrier; the four warps drain their TMEM lane quadrant with
//   tcgen05.ld (one position = 16 channels = 64 B per thread) while the MMAs of the next slice run.
namespace tc {
constexpr int TH = 7, TW = 32, HH = TH + 2, HW = TW + 2;     // 9 x 34 halo tile
constexpr int SLICE_POS = HH * HW;                            // 306 positions written by TMA
constexpr int SLICE_BYTES = 21504;                            // >= (255 + 2*34 + 2 + 1) * 64 = 20864, multiple of 512
constexpr int NBUF = 3;
constexpr int THREADS = 128;
constexpr int BH_KHW_BYTES = 48 * 64;                         // hi bank per (kh,kw): rows [kd][n]
constexpr int BI_KHW_BYTES = 96 * 64;                         // interleaved bank per (kh,kw): rows [kd][hi|lo][n]

template <int PASSES>
struct Cfg {
    static constexpr int OFF_A = 0;
    static constexpr int OFF_LO = OFF_A + NBUF * SLICE_BYTES;                        // [2] lo halves of the slice (3xTF32 only)
    static constexpr int OFF_BH = OFF_LO + (PASSES == 3 ? 2 * SLICE_BYTES : 0);      // hi bank (raw fp32 when PASSES == 1)
    static constexpr int OFF_BI = OFF_BH + 9 * BH_KHW_BYTES;                         // interleaved hi|lo bank (3xTF32 only)
    static constexpr int OFF_BAR = OFF_BI + (PASSES == 3 ? 9 * BI_KHW_BYTES : 0);    // full[NBUF], done[2], tmem base
    static constexpr int TOTAL = OFF_BAR + 64;
    static constexpr int ALLOC = TOTAL + 1024;
    // TMEM columns per M-block: ring 1 (4 groups of GW1 columns), ring 2 (4 groups of 16; 3xTF32 only)
    static constexpr uint32_t GW1 = PASSES == 3 ? 32 : 16;
    static constexpr uint32_t MB_COLS = 4 * GW1 + (PASSES == 3 ? 64 : 0);
    static constexpr uint32_t TMEM_COLS = PASSES == 3 ? 512 : 128;
};

// K-major, SWIZZLE_64B shared-memory matrix descriptor: LBO = 1 (16 B units, unused for swizzled K-major), SBO = 512 B
// (8 rows x 64 B), version 1 (Blackwell), base offset 0 -- measured: the swizzle phase comes from the absolute
// shared-memory address, so a start address shifted by whole 64 B rows needs NO base-offset correction (setting
// (addr >> 7) & 7 there gives wrong results).  Advancing by `bytes` = adding bytes >> 4 to the low word.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __forceinline__ constexpr uint32_t idesc(uint32_t n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t id) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(id)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a fully converged warp; code dominated by this predicate is single-threaded, so the compiler can keep
// descriptors / TMEM addresses in uniform registers instead of wrapping every UTCHMMA in an ELECT + R2UR.BROADCAST loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "@P mov.u32 %0, 1;\n\t"
        "}\n"
        : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
                 : "memory");
}

struct TArgs {
    const float* w;
    float* out;
    double* bn_sums;     // nullable: [sum y (16), sum y^2 (16)] accumulated over the whole output (fused BatchNorm statistics)
    int mode, dbg;
    int B, D, H, W;
    int tiles_h, tiles_w, dsplit, dlen;
};

__device__ __forceinline__ Item tc_item(const TArgs& a, int item) {
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

// MMAs of one input slice for the kd range [KD0, KD0 + NKD) landing in ring group G (a contiguous TMEM column range)
template <int PASSES, int KD0, int NKD, int G>
__device__ __forceinline__ void issue_part(uint32_t tmem_base, uint64_t a_desc, uint64_t lo_desc, uint64_t bh_desc, uint64_t bi_desc,
                                           int warp) {
    using CF = Cfg<PASSES>;
#pragma unroll
    for (int khw = 0; khw < 9; ++khw) {
        if ((khw * 4) / 9 != warp) continue;                     // the (kh,kw) taps are dealt to the four issuing warps
        const int kh = khw / 3, kw = khw - kh * 3;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint64_t ao = static_cast<uint64_t>(((kh * HW + kw) * 64 + mb * 128 * 64 + ks * 32) >> 4);
                const uint32_t d1 = tmem_base + mb * CF::MB_COLS + G * CF::GW1;
                if (PASSES == 3) {
                    const uint64_t bio = static_cast<uint64_t>((khw * BI_KHW_BYTES + KD0 * 32 * 64 + ks * 32) >> 4);
                    const uint64_t bho = static_cast<uint64_t>((khw * BH_KHW_BYTES + KD0 * 16 * 64 + ks * 32) >> 4);
                    umma_tf32(d1, a_desc + ao, bi_desc + bio, idesc(32 * NKD));                    // hi * [hi | lo]
                    umma_tf32(tmem_base + mb * CF::MB_COLS + 4 * CF::GW1 + G * 16, lo_desc + ao, bh_desc + bho, idesc(16 * NKD));   // lo * hi
                } else {
                    const uint64_t bho = static_cast<uint64_t>((khw * BH_KHW_BYTES + KD0 * 16 * 64 + ks * 32) >> 4);
                    umma_tf32(d1, a_desc + ao, bh_desc + bho, idesc(16 * NKD));
                }
            }
        }
    }
}

template <int PASSES>
__global__ void __launch_bounds__(THREADS, PASSES == 3 ? 1 : 2)
conv3d_c16c16_tc_kernel(const __grid_constant__ CUtensorMap map_in, const TArgs a) {
    using CF = Cfg<PASSES>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + CF::OFF_BAR);
    uint64_t* done = full + NBUF;                                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + NBUF + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const Item it = tc_item(a, blockIdx.x);
    const int count = it.d1 - it.d0 + 2;

    auto issue = [&](int i) {                                   // TMA load of x slice d0-1+i into ring buffer i % NBUF
        uint64_t* bar = full + (i % NBUF);
        mbar_expect_tx(bar, SLICE_POS * 64);
        tma_load_5d(smem + CF::OFF_A + (i % NBUF) * SLICE_BYTES, &map_in, bar, 0, it.w0 - 1, it.h0 - 1, it.d0 - 1 + i, it.b);
    };
    if (tid == 0) {
        tma_prefetch_desc(&map_in);
        for (int i = 0; i < NBUF; ++i) mbar_init(full + i, 1);
        mbar_init(done, THREADS / 32);                          // one commit per issuing warp
        mbar_init(done + 1, THREADS / 32);
        mbar_fence_init();
        for (int i = 0; i < NBUF && i < count; ++i) issue(i);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(CF::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // filter banks, K-major SWIZZLE_64B: row r, k -> r*64 + ((k/4) ^ ((r>>1)&3))*16 + (k%4)*4 (bank bases are 512 B aligned)
    for (int e = tid; e < 27 * 256; e += THREADS) {
        const int tap = e >> 8, n = (e >> 4) & 15, k = e & 15;
        const int kd = tap / 9, khw = tap - kd * 9;
        // forward: B[n=co][k=ci] = W[co][ci][tap];  data gradient: B[n=ci][k=co] = W[co][ci][26 - tap]
        const float v = (a.mode == 0) ? __ldg(a.w + (n * 16 + k) * 27 + tap) : __ldg(a.w + (k * 16 + n) * 27 + (26 - tap));
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        auto put = [&](int base, int row, float val) {
            *reinterpret_cast<float*>(smem + base + row * 64 + (((k >> 2) ^ ((row >> 1) & 3)) << 4) + ((k & 3) << 2)) = val;
        };
        put(CF::OFF_BH + khw * BH_KHW_BYTES, kd * 16 + n, PASSES == 3 ? hi : v);
        if (PASSES == 3) {
            put(CF::OFF_BI + khw * BI_KHW_BYTES, kd * 32 + n, hi);
            put(CF::OFF_BI + khw * BI_KHW_BYTES, kd * 32 + 16 + n, v - hi);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the MMA's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (uint32_t c = 0; c < 2 * CF::MB_COLS; c += 16) tmem_zero16(lane_addr + c);   // every MMA accumulates
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const uint64_t bh_desc = smem_desc(smem_u32(smem + CF::OFF_BH)), bi_desc = smem_desc(smem_u32(smem + CF::OFF_BI));

    float bn_s[16], bn_ss[16];                                   // this thread's share of the BatchNorm statistics of the output
#pragma unroll
    for (int q = 0; q < 16; ++q) bn_s[q] = bn_ss[q] = 0.f;
    // epilogue of input slice j: output slice d = (d0-1+j) - 1 is complete (it received kd = 2 from slice j)
    auto epilogue = [&](int j) {
        const int d = it.d0 - 2 + j;
        const uint32_t grp = static_cast<uint32_t>(-d) & 3u;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const uint32_t t1 = lane_addr + mb * CF::MB_COLS + grp * CF::GW1;
            const uint32_t t2 = lane_addr + mb * CF::MB_COLS + 4 * CF::GW1 + grp * 16;
            if (d >= it.d0) {
                uint32_t r[16];
                tmem_ld16(t1, r);
                float o[16];
                if (PASSES == 3) {
                    uint32_t r2[16], r3[16];
                    tmem_ld16(t1 + 16, r2);
                    tmem_ld16(t2, r3);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int q = 0; q < 16; ++q) o[q] = __uint_as_float(r[q]) + (__uint_as_float(r2[q]) + __uint_as_float(r3[q]));
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int q = 0; q < 16; ++q) o[q] = __uint_as_float(r[q]);
                }
                const int v = mb * 128 + tid, rr = v / HW, cc = v - rr * HW;
                const int h = it.h0 + rr, w = it.w0 + cc;
                if (a.bn_sums != nullptr && rr < TH && cc < TW && h < a.H && w < a.W) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        bn_s[q] += o[q];
                        bn_ss[q] = fmaf(o[q], o[q], bn_ss[q]);
                    }
                }
                if (rr < TH && cc < TW && h < a.H && w < a.W && !(a.dbg & 4)) {
                    float4* op = reinterpret_cast<float4*>(a.out + (((static_cast<size_t>(it.b) * a.D + d) * a.H + h) * a.W + w) * C);
                    op[0] = make_float4(o[0], o[1], o[2], o[3]);
                    op[1] = make_float4(o[4], o[5], o[6], o[7]);
                    op[2] = make_float4(o[8], o[9], o[10], o[11]);
                    op[3] = make_float4(o[12], o[13], o[14], o[15]);
                }
            }
            tmem_zero16(t1);                                     // the group is the fresh accumulator of output slice d + 4
            if (PASSES == 3) {
                tmem_zero16(t1 + 16);
                tmem_zero16(t2);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };

    // The MMAs of slice i run asynchronously while the four warps drain the accumulators finished by slice i-1.
    for (int i = 0; i < count; ++i) {
        const int s = it.d0 - 1 + i, bi = i % NBUF;
        unsigned char* buf = smem + CF::OFF_A + bi * SLICE_BYTES;
        unsigned char* lob = smem + CF::OFF_LO + (i & 1) * SLICE_BYTES;
        mbar_wait(full + bi, (i / NBUF) & 1);
        if (PASSES == 3 && !(a.dbg & 8)) {                       // split the slice: hi in place, lo into its own buffer
            float4* hp = reinterpret_cast<float4*>(buf);
            float4* lp = reinterpret_cast<float4*>(lob);
            for (int e = tid; e < SLICE_POS * 4; e += THREADS) {
                const float4 v = hp[e];
                float4 h;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                hp[e] = h;
                lp[e] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();                                         // slice i is ready; everyone is done with the epilogue of slice i-2
        tc_fence_after();
        if (elect_one()) {                                       // one lane per warp issues its share of the taps
            const uint64_t ad = smem_desc(smem_u32(buf)), ld = smem_desc(smem_u32(lob));
            // output slices s+1, s, s-1 (kd = 0,1,2) live in ring groups g0, g0+1, g0+2 (mod 4), g0 = (-(s+1)) & 3
            switch ((a.dbg & 2) ? 99u : (static_cast<uint32_t>(-(s + 1)) & 3u)) {
                case 99: break;
                case 0: issue_part<PASSES, 0, 3, 0>(tmem_base, ad, ld, bh_desc, bi_desc, warp); break;
                case 1: issue_part<PASSES, 0, 3, 1>(tmem_base, ad, ld, bh_desc, bi_desc, warp); break;
                case 2:
                    issue_part<PASSES, 0, 2, 2>(tmem_base, ad, ld, bh_desc, bi_desc, warp);
                    issue_part<PASSES, 2, 1, 0>(tmem_base, ad, ld, bh_desc, bi_desc, warp);
                    break;
                default:
                    issue_part<PASSES, 0, 1, 3>(tmem_base, ad, ld, bh_desc, bi_desc, warp);
                    issue_part<PASSES, 1, 2, 0>(tmem_base, ad, ld, bh_desc, bi_desc, warp);
                    break;
            }
            umma_commit(done + (i & 1));                         // arrives when every MMA this thread issued has completed
        }
        if (i >= 1) {
            mbar_wait(done + ((i - 1) & 1), ((i - 1) >> 1) & 1);
            tc_fence_after();
            if (tid == 0 && i - 1 + NBUF < count) issue(i - 1 + NBUF);     // ring buffer of slice i-1 is free again
            epilogue(i - 1);
        }
    }
    mbar_wait(done + ((count - 1) & 1), ((count - 1) >> 1) & 1);
    tc_fence_after();
    epilogue(count - 1);
    tc_fence_before();
    __syncthreads();
    if (a.bn_sums != nullptr) {                                  // block reduction in the (now idle) slice ring, 32 fp64 atomics per item
        float* red = reinterpret_cast<float*>(smem + CF::OFF_A);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            red[q * THREADS + tid] = bn_s[q];
            red[(16 + q) * THREADS + tid] = bn_ss[q];
        }
        __syncthreads();
        if (tid < 32) {
            double t = 0.0;
            for (int k = 0; k < THREADS; ++k) t += static_cast<double>(red[tid * THREADS + ((k + tid) & (THREADS - 1))]);
            atomicAdd(a.bn_sums + tid, t);
        }
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(CF::TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of the 16 -> 16 layer on tcgen05:  gw[co,ci,kd,kh,kw] = sum_pos gy[pos,co] * x[pos + off(kd,kh,kw), ci]
// with the reduction over positions as the MMA's K dimension, i.e. both operands MN-major (channels contiguous).
// Measured on this part: kind::tf32 accepts MN-major operands ONLY in the SWIZZLE_128B_BASE32B layout (128 B rows, 4-row
// atoms, 32 B chunks XOR-ed with the row index); with the 32/64/128 B swizzles or no swizzle the MMA is a silent no-op.
// A 128 B row of a position-major buffer (64 B per position) is a PAIR of positions, so the tiles are TMA-loaded through
// a paired view of the tensors ([.., W/2, 32] fp32, inner box = 128 B = the swizzle span, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
// with a flattened pitch of 36 positions (x tile origin at column w0-2, gy tile at w0; both even).  Then
//   A row k, M-group m = [x(A0 + 2k + 2m), x(A0 + 2k + 2m + 1)]   (leading-dimension offset 128 B = the next row: groups overlap),
//   B row k            = [gy(2k), gy(2k + 1)],
//   D[(shift s, ci), (j, co)] = sum_k x[A0 + 2k + s, ci] * gy[2k + j, co]      (M = 128: s = 0..7, N = 32: j = 0,1),
// with A0 = kh * 36, and the wanted tap is  gw[.., kw] = D[s = kw + 1][j = 0] + D[s = kw + 2][j = 1]  (even + odd positions).
// One MMA covers 16 positions; nine accumulators (kd,kh) x 32 columns stay in TMEM for the whole work item.
constexpr int WP = 36;                                         // flattened pitch (positions)
constexpr int WX_POS = HH * WP, WG_POS = TH * WP;              // 324 / 252 positions per slice
constexpr int GY_BYTES = 16384;                                // 256 positions = 16 k-steps of 16
constexpr int WG_KSTEPS = 16;
constexpr int WG_XBUF = 3, WG_GBUF = 4;
constexpr uint32_t WG_TMEM_COLS = 512;                         // 9 x 32 columns in use

struct WgCfg {
    static constexpr int OFF_X = 0;
    static constexpr int OFF_G = OFF_X + WG_XBUF * SLICE_BYTES;
    static constexpr int OFF_BAR = OFF_G + WG_GBUF * GY_BYTES;  // xfull[3], gfull[4], done[2], tmem base
    static constexpr int TOTAL = OFF_BAR + 96;
    static constexpr int ALLOC = TOTAL + 1024;
};
static_assert(SLICE_BYTES % 1024 == 0 && GY_BYTES % 1024 == 0, "128 B-row swizzle needs 1024 B aligned tiles");
static_assert((2 * (WP / 2) + (WG_KSTEPS - 1) * 8 + 8 + 3) * 128 <= SLICE_BYTES, "A operand stays inside the x slice buffer");

// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = byte stride between 32-element (128 B) groups along M/N,
// SBO = stride between groups of 4 K rows (512 B)
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes = 128) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(1) << 61;
    return d;
}
// D = F32, A = B = TF32, both MN-major, M = 128, N = n (32, 64 or 96: one 32-column group per gy slice)
__device__ __forceinline__ constexpr uint32_t wg_idesc(uint32_t n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct WTArgs {
    float* part;         // [2 * items][6912]: even-position and odd-position halves of every item
    int B, D, H, W;
    int tiles_h, tiles_w, dsplit, dlen;
};

__global__ void __launch_bounds__(THREADS, 1)
conv3d_c16c16_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_gy, const WTArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* xfull = reinterpret_cast<uint64_t*>(smem + WgCfg::OFF_BAR);
    uint64_t* gfull = xfull + WG_XBUF;
    uint64_t* done = gfull + WG_GBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    TArgs ta{};
    ta.B = a.B; ta.D = a.D; ta.H = a.H; ta.W = a.W;
    ta.tiles_h = a.tiles_h; ta.tiles_w = a.tiles_w; ta.dsplit = a.dsplit; ta.dlen = a.dlen;
    const Item it = tc_item(ta, blockIdx.x);
    const int count = it.d1 - it.d0 + 2;                           // x slices d0-1 .. d1

    auto issue_x = [&](int i) {                                    // paired view: coordinate 1 counts pairs of columns
        uint64_t* bar = xfull + (i % WG_XBUF);
        mbar_expect_tx(bar, WX_POS * 64);
        tma_load_5d(smem + WgCfg::OFF_X + (i % WG_XBUF) * SLICE_BYTES, &map_x, bar, 0, it.w0 / 2 - 1, it.h0 - 1, it.d0 - 1 + i, it.b);
    };
    auto issue_gy = [&](int d) {                                   // only this item's own gy slices are ever used
        if (d >= it.d0 && d < it.d1) {
            uint64_t* bar = gfull + (d & 3);
            mbar_expect_tx(bar, WG_POS * 64);
            tma_load_5d(smem + WgCfg::OFF_G + (d & 3) * GY_BYTES, &map_gy, bar, 0, it.w0 / 2, it.h0, d, it.b);
        }
    };
    if (tid == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_gy);
        for (int i = 0; i < WG_XBUF; ++i) mbar_init(xfull + i, 1);
        for (int i = 0; i < WG_GBUF; ++i) mbar_init(gfull + i, 1);
        mbar_init(done, THREADS / 32);
        mbar_init(done + 1, THREADS / 32);
        mbar_fence_init();
        for (int i = 0; i < WG_XBUF && i < count; ++i) issue_x(i);
        issue_gy(it.d0);
        issue_gy(it.d0 + 1);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(WG_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (uint32_t c = 0; c < 9 * 32; c += 16) tmem_zero16(lane_addr + c);     // every MMA accumulates
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");

    for (int i = 0; i < count; ++i) {
        const int s = it.d0 - 1 + i, xb = i % WG_XBUF;
        mbar_wait(xfull + xb, (i / WG_XBUF) & 1);
        // gy slice s+1 is new in this iteration: wait for it and zero what must not contribute -- the four pitch
        // columns 32..35 (they hold the neighbouring tile's gradients) and the four pad positions at the end
        const int dn = s + 1;
        if (dn >= it.d0 && dn < it.d1) {
            const int first = it.d0 + (((dn & 3) - (it.d0 & 3)) & 3);
            mbar_wait(gfull + (dn & 3), ((dn - first) >> 2) & 1);
            unsigned char* gb = smem + WgCfg::OFF_G + (dn & 3) * GY_BYTES;
            {                                                          // 32 positions x 4 pieces of 16 B = 128 threads
                const int j = tid >> 2, q = tid & 3;
                const int p = j < 28 ? (j >> 2) * WP + TW + (j & 3) : WG_POS + (j - 28);
                const int row = p >> 1, chunk = ((2 * (p & 1) + (q >> 1)) ^ (row & 3));      // 32 B chunks are XOR-ed with the row
                *reinterpret_cast<float4*>(gb + row * 128 + chunk * 32 + (q & 1) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (elect_one()) {
            const uint32_t xaddr = smem_u32(smem + WgCfg::OFF_X + xb * SLICE_BYTES);
            // x slice s meets the gy slices d = s-1, s, s+1 (kd = 2, 1, 0).  The kd taps are folded into the MMA's N dimension:
            // the gy ring slots are GY_BYTES apart (= the descriptor's leading-dimension offset between 32-column groups), so
            // slices in consecutive slots are ONE MMA of N = 32 * run and the 4 KB x operand is read once for up to three kd
            // (the operand reads, not the tensor pipe, bound this kernel).  Accumulator of (kh): columns kh*96 + (2-kd)*32.
            const int lo = max(s - 1, it.d0), hi = min(s + 1, it.d1 - 1);
            if (warp < 3 && lo <= hi) {                                // kh = warp
                const int kh = warp;
                const uint64_t ad = smem_desc_mn(xaddr + static_cast<uint32_t>(kh * WP) * 64u);       // kh tile rows down = 18 smem rows
                for (int d = lo; d <= hi;) {
                    const int slot = d & 3, run = min(hi - d + 1, 4 - slot);
                    const uint32_t col = static_cast<uint32_t>(kh * 96 + (d - (s - 1)) * 32);
                    const uint64_t bd = smem_desc_mn(smem_u32(smem + WgCfg::OFF_G + slot * GY_BYTES), GY_BYTES);
                    const uint32_t id = wg_idesc(32u * run);
#pragma unroll 4
                    for (int ks = 0; ks < WG_KSTEPS; ++ks)             // 8 rows = 16 positions = 1 KB of x per MMA
                        umma_tf32(tmem_base + col, ad + static_cast<uint64_t>(ks * 64), bd + static_cast<uint64_t>(ks * 64), id);
                    d += run;
                }
            }
            umma_commit(done + (i & 1));
        }
        if (i >= 1) {
            mbar_wait(done + ((i - 1) & 1), ((i - 1) >> 1) & 1);
            tc_fence_after();
            if (tid == 0) {                                            // slice i-1's MMAs are complete: its buffers are free
                if (i - 1 + WG_XBUF < count) issue_x(i - 1 + WG_XBUF);
                issue_gy(s + 2);                                       // slot of gy slice s-2, last used by x slice s-1
            }
        }
    }
    mbar_wait(done + ((count - 1) & 1), ((count - 1) >> 1) & 1);
    tc_fence_after();
    // ---- epilogue: accumulator (kd,kh): row (shift s, ci) = TMEM lane 16*s + ci, column j*16 + co
    {
        float* p0 = a.part + static_cast<size_t>(2 * blockIdx.x) * NW16;       // j = 0: kw = s - 1
        float* p1 = p0 + NW16;                                                 // j = 1: kw = s - 2
        const int sh = tid >> 4, ci = tid & 15;
#pragma unroll 1
        for (int acc = 0; acc < 9; ++acc) {                            // acc = kd*3 + kh lives in columns kh*96 + (2-kd)*32
            uint32_t r0[16], r1[16];
            const uint32_t col = static_cast<uint32_t>((acc % 3) * 96 + (2 - acc / 3) * 32);
            tmem_ld16(lane_addr + col, r0);
            tmem_ld16(lane_addr + col + 16, r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int co = 0; co < 16; ++co) {
                if (sh >= 1 && sh <= 3) p0[(co * 16 + ci) * 27 + acc * 3 + sh - 1] = __uint_as_float(r0[co]);
                if (sh >= 2 && sh <= 4) p1[(co * 16 + ci) * 27 + acc * 3 + sh - 2] = __uint_as_float(r1[co]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(WG_TMEM_COLS) : "memory");
}
}  // namespace tc

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

__global__ void transpose_w_kernel(const float* __restrict__ w) {           // [c][27] -> [tap][16]
    const int o = threadIdx.x;
    if (o < NW) g_wt[o] = w[(o & 15) * 27 + (o >> 4)];
}

static int upload_weights(const float* w, cudaStream_t st) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_w, w, sizeof(float) * NW, 0, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "conv3d_c16o1 weight upload: %s", cudaGetErrorString(e));
    transpose_w_kernel<<<1, 448, 0, st>>>(w);
    void* staging = nullptr;
    e = cudaGetSymbolAddress(&staging, g_wt);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_wt, staging, sizeof(float) * NW, 0, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "conv3d_c16o1 weight upload (tap-major): %s", cudaGetErrorString(e));
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
    bool attr_done = false;       // set on every call (a few hundred ns): the attribute is per device, a process-wide latch is not
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
    bool attr_done = false;       // set on every call (a few hundred ns): the attribute is per device, a process-wide latch is not
    if (!attr_done) {
        cudaFuncSetAttribute(conv3d_c16o1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WGRAD_SMEM);
        attr_done = true;
    }
    conv3d_c16o1_wgrad_kernel<<<items, THREADS, WGRAD_SMEM, st>>>(map, a);
    if (int rc = mvd::check_launch("conv3d_c16o1_wgrad")) return rc;
    conv3d_c16o1_wgrad_reduce_kernel<<<NW, 128, 0, st>>>(a.part, gw, items);
    return mvd::check_launch("conv3d_c16o1_wgrad_reduce");
}


int mvd_conv3d_c16c16(const float* in, const float* w, float* out, int B, int D, int H, int W, int mode, int passes,
                      void* stream) {
    using namespace mvd::c16;
    MVD_REQUIRE(in && w && out, "null pointer argument");
    if (int rc = check_shape(B, D, H, W)) return rc;
    MVD_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (forward) or 1 (data gradient), got %d", mode);
    MVD_REQUIRE(passes == 1 || passes == 3, "passes must be 1 (TF32) or 3 (3xTF32), got %d", passes);
    MVD_REQUIRE(mvd::aligned16(in) && mvd::aligned16(out), "activation pointers must be 16-byte aligned");
    cudaStream_t st = mvd::as_stream(stream);
    Args pa{};
    pa.B = B; pa.D = D; pa.H = H; pa.W = W;
    const int items = plan(pa);
    GArgs a{};
    a.w = w; a.out = out; a.mode = mode; a.B = B; a.D = D; a.H = H; a.W = W;
    a.tiles_h = pa.tiles_h; a.tiles_w = pa.tiles_w; a.dsplit = pa.dsplit; a.dlen = pa.dlen;
    pa.x = in;
    CUtensorMap map;
    if (int rc = make_x_map(&map, in, pa)) return rc;
    bool attr_done = false;       // set on every call (a few hundred ns): the attribute is per device, a process-wide latch is not
    if (!attr_done) {
        cudaFuncSetAttribute(conv3d_c16c16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
        cudaFuncSetAttribute(conv3d_c16c16_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
        attr_done = true;
    }
    if (passes == 3) conv3d_c16c16_kernel<3><<<items, G_THREADS, G_SMEM, st>>>(map, a);
    else conv3d_c16c16_kernel<1><<<items, G_THREADS, G_SMEM, st>>>(map, a);
    return mvd::check_launch("conv3d_c16c16");
}


long long mvd_conv3d_c16c16_wgrad_workspace_bytes(int B, int D, int H, int W) {
    using namespace mvd::c16;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    Args a{};
    a.B = B; a.D = D; a.H = H; a.W = W;
    return static_cast<long long>(plan(a)) * NW16 * sizeof(float);
}

int mvd_conv3d_c16c16_wgrad(const float* gy, const float* x, float* gw, void* workspace, long long workspace_bytes, int B,
                            int D, int H, int W, void* stream) {
    using namespace mvd::c16;
    MVD_REQUIRE(gy && x && gw && workspace, "null pointer argument");
    if (int rc = check_shape(B, D, H, W)) return rc;
    MVD_REQUIRE(mvd::aligned16(x) && mvd::aligned16(gy), "activation pointers must be 16-byte aligned");
    cudaStream_t st = mvd::as_stream(stream);
    Args pa{};
    pa.B = B; pa.D = D; pa.H = H; pa.W = W;
    const int items = plan(pa);
    MVD_REQUIRE(workspace_bytes >= static_cast<long long>(items) * NW16 * 4, "workspace too small: %lld < %lld", workspace_bytes,
                static_cast<long long>(items) * NW16 * 4);
    WArgs a{};
    a.part = static_cast<float*>(workspace); a.B = B; a.D = D; a.H = H; a.W = W;
    a.tiles_h = pa.tiles_h; a.tiles_w = pa.tiles_w; a.dsplit = pa.dsplit; a.dlen = pa.dlen;
    CUtensorMap map_x, map_gy;
    if (int rc = make_x_map(&map_x, x, pa)) return rc;
    {
        const uint64_t Wd = W, Hd = H, Dd = D;
        const uint64_t dims[5] = {C, Wd, Hd, Dd, static_cast<uint64_t>(B)};
        const uint64_t str[4] = {C * 4, Wd * C * 4, Wd * Hd * C * 4, Wd * Hd * Dd * C * 4};
        const uint32_t box[5] = {C, TW, TH, 1, 1};
        if (int rc = mvd::make_f32_tensor_map(&map_gy, gy, 5, dims, str, box, 64)) return rc;
    }
    bool attr_done = false;       // set on every call (a few hundred ns): the attribute is per device, a process-wide latch is not
    if (!attr_done) {
        cudaFuncSetAttribute(conv3d_c16c16_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
        attr_done = true;
    }
    conv3d_c16c16_wgrad_kernel<<<items, WG_THREADS, WG_SMEM, st>>>(map_x, map_gy, a);
    if (int rc = mvd::check_launch("conv3d_c16c16_wgrad")) return rc;
    conv3d_c16c16_wgrad_reduce_kernel<<<(NW16 + 3) / 4, 128, 0, st>>>(a.part, gw, items);
    return mvd::check_launch("conv3d_c16c16_wgrad_reduce");
}


int mvd_conv3d_c16c16_tc(const float* in, const float* w, float* out, double* bn_sums, int B, int D, int H, int W, int mode,
                         int passes, int flags, void* stream) {
    using namespace mvd::c16;
    MVD_REQUIRE(in && w && out, "null pointer argument");
    if (int rc = check_shape(B, D, H, W)) return rc;
    MVD_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (forward) or 1 (data gradient), got %d", mode);
    MVD_REQUIRE(passes == 1 || passes == 3, "passes must be 1 (TF32) or 3 (3xTF32), got %d", passes);
    MVD_REQUIRE(mvd::aligned16(in) && mvd::aligned16(out), "activation pointers must be 16-byte aligned");
    cudaStream_t st = mvd::as_stream(stream);
    tc::TArgs a{};
    a.w = w; a.out = out; a.bn_sums = bn_sums; a.mode = mode; a.B = B; a.D = D; a.H = H; a.W = W;
    a.tiles_h = (H + tc::TH - 1) / tc::TH;
    a.tiles_w = (W + tc::TW - 1) / tc::TW;
    a.dbg = flags;
    // split the depth axis so that the work items fill whole waves of resident CTAs (each chunk re-reads 2 halo slices)
    const int base = B * a.tiles_h * a.tiles_w, slots = mvd::sm_count() * (passes == 3 ? 1 : 2);
    int best = 1;
    double best_eff = 0.0;
    for (int ds = 1; ds <= 16 && D / ds >= 8; ++ds) {
        const int dlen = (D + ds - 1) / ds, items = base * ((D + dlen - 1) / dlen);
        const double waves = static_cast<double>(items) / slots;
        const double eff = waves / static_cast<double>((items + slots - 1) / slots) * dlen / (dlen + 2.0);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = ds;
        }
    }
    a.dlen = (D + best - 1) / best;
    a.dsplit = (D + a.dlen - 1) / a.dlen;
    const int items = base * a.dsplit;
    CUtensorMap map;
    {
        const uint64_t Wd = W, Hd = H, Dd = D;
        const uint64_t dims[5] = {C, Wd, Hd, Dd, static_cast<uint64_t>(B)};
        const uint64_t str[4] = {C * 4, Wd * C * 4, Wd * Hd * C * 4, Wd * Hd * Dd * C * 4};
        const uint32_t box[5] = {C, tc::HW, tc::HH, 1, 1};
        if (int rc = mvd::make_f32_tensor_map(&map, in, 5, dims, str, box, 64)) return rc;
    }
    bool attr_done = false;       // set on every call (a few hundred ns): the attribute is per device, a process-wide latch is not
    if (!attr_done) {
        cudaFuncSetAttribute(tc::conv3d_c16c16_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<1>::ALLOC);
        cudaFuncSetAttribute(tc::conv3d_c16c16_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<3>::ALLOC);
        attr_done = true;
    }
    if (passes == 3) tc::conv3d_c16c16_tc_kernel<3><<<items, tc::THREADS, tc::Cfg<3>::ALLOC, st>>>(map, a);
    else tc::conv3d_c16c16_tc_kernel<1><<<items, tc::THREADS, tc::Cfg<1>::ALLOC, st>>>(map, a);
    return mvd::check_launch("conv3d_c16c16_tc");
}


static int wgrad_tc_plan(mvd::c16::tc::WTArgs& a) {
    using namespace mvd::c16;
    a.tiles_h = (a.H + tc::TH - 1) / tc::TH;
    a.tiles_w = (a.W + tc::TW - 1) / tc::TW;
    const int base = a.B * a.tiles_h * a.tiles_w, slots = mvd::sm_count();
    int best = 1;
    double best_eff = 0.0;
    for (int ds = 1; ds <= 16 && a.D / ds >= 8; ++ds) {
        const int dlen = (a.D + ds - 1) / ds, items = base * ((a.D + dlen - 1) / dlen);
        const double waves = static_cast<double>(items) / slots;
        const double eff = waves / static_cast<double>((items + slots - 1) / slots) * dlen / (dlen + 2.0);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = ds;
        }
    }
    a.dlen = (a.D + best - 1) / best;
    a.dsplit = (a.D + a.dlen - 1) / a.dlen;
    return base * a.dsplit;
}

long long mvd_conv3d_c16c16_wgrad_tc_workspace_bytes(int B, int D, int H, int W) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    mvd::c16::tc::WTArgs a{};
    a.B = B; a.D = D; a.H = H; a.W = W;
    return 2ll * wgrad_tc_plan(a) * mvd::c16::NW16 * sizeof(float);
}

int mvd_conv3d_c16c16_wgrad_tc(const float* gy, const float* x, float* gw, void* workspace, long long workspace_bytes, int B,
                               int D, int H, int W, void* stream) {
    using namespace mvd::c16;
    MVD_REQUIRE(gy && x && gw && workspace, "null pointer argument");
    if (int rc = check_shape(B, D, H, W)) return rc;
    MVD_REQUIRE(mvd::aligned16(x) && mvd::aligned16(gy), "activation pointers must be 16-byte aligned");
    cudaStream_t st = mvd::as_stream(stream);
    tc::WTArgs a{};
    a.part = static_cast<float*>(workspace); a.B = B; a.D = D; a.H = H; a.W = W;
    const int items = wgrad_tc_plan(a);
    MVD_REQUIRE(workspace_bytes >= 2ll * items * NW16 * 4, "workspace too small: %lld < %lld", workspace_bytes, 2ll * items * NW16 * 4);
    MVD_REQUIRE(W % 2 == 0, "the tcgen05 weight gradient pairs columns: W must be even (got %d)", W);
    CUtensorMap map_x, map_gy;
    const uint64_t Wd = W, Hd = H, Dd = D;
    const uint64_t dims[5] = {2 * C, Wd / 2, Hd, Dd, static_cast<uint64_t>(B)};             // paired view: 2 columns x 16 channels = 128 B
    const uint64_t str[4] = {2 * C * 4, Wd * C * 4, Wd * Hd * C * 4, Wd * Hd * Dd * C * 4};
    const uint32_t boxx[5] = {2 * C, tc::WP / 2, tc::HH, 1, 1}, boxg[5] = {2 * C, tc::WP / 2, tc::TH, 1, 1};
    if (int rc = mvd::make_f32_tensor_map(&map_x, x, 5, dims, str, boxx, 12832)) return rc;
    if (int rc = mvd::make_f32_tensor_map(&map_gy, gy, 5, dims, str, boxg, 12832)) return rc;
    bool attr_done = false;       // set on every call (a few hundred ns): the attribute is per device, a process-wide latch is not
    if (!attr_done) {
        cudaFuncSetAttribute(tc::conv3d_c16c16_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::WgCfg::ALLOC);
        attr_done = true;
    }
    tc::conv3d_c16c16_wgrad_tc_kernel<<<items, tc::THREADS, tc::WgCfg::ALLOC, st>>>(map_x, map_gy, a);
    if (int rc = mvd::check_launch("conv3d_c16c16_wgrad_tc")) return rc;
    conv3d_c16c16_wgrad_reduce_kernel<<<(NW16 + 3) / 4, 128, 0, st>>>(a.part, gw, 2 * items);
    return mvd::check_launch("conv3d_c16c16_wgrad_reduce");
}

}
