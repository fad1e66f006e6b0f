// Training-mode BatchNorm for channels-last activations, fused with what surrounds it in the reference's conv
// blocks: ReLU (ConvBnReLU3D movedepth/networks/resnet_encoder.py:175-182, Conv2d 453-475, ResNet blocks) and the
// residual add of the ResNet basic block.  The activation is a row-major [M, C] matrix (M = N*D*H*W, C contiguous).
//
//   forward : stats   (1 read)          per-channel sum / sum of squares, fp32 per thread, fp64 across the grid
//             [multi-GPU: the 2C fp64 sums are all-reduced between the two kernels == SyncBatchNorm]
//             finalize                  mean, invstd, running statistics, scale/shift
//             apply   (1 read, 1 write) y = relu(x*scale + shift (+ residual))   or   relu(x*scale + shift) + residual (relu flag bit 1)
//   backward: reduce  (2-3 reads)       sum(g), sum(g*xhat)  with g = gy * (y > 0)
//             apply   (2-3 reads, 1-2 writes)  gx = scale*(g - mean(g) - xhat*mean(g*xhat)), gres = g; gw, gb
// cuDNN's NHWC BatchNorm kernels move the 283 MB full-resolution reg3d activations at ~1.4 TB/s and need separate
// ReLU / add passes; these are plain HBM-bound float4 streams.
#include "peer.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {
namespace bn {

constexpr int THREADS = 256;

__device__ __forceinline__ void add4(float (&s)[4], const float4& v) {
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
}

// blockDim = 256 threads, q = C/4 channel quads; thread t handles quad t % q of rows t / q + k * (256 / q).
struct Map {
    int q, rows_per_iter, quad, row0;
};
template <int T = THREADS>
__device__ __forceinline__ Map make_map(int C) {
    Map m;
    m.q = C >> 2;
    m.rows_per_iter = T / m.q;
    m.quad = threadIdx.x % m.q;
    m.row0 = threadIdx.x / m.q;
    return m;
}

// Per-thread partial sums are kept in fp64: with fp32 partials the variance E[x^2] - mean^2 of a layer whose mean dominates its
// spread loses digits that tiny-batch layers then amplify (ResNet50 at batch 1: the gradient norm of the pose encoder moved by
// 2.7e-3 with the row-to-thread mapping alone).  The kernels are HBM-bound; eight fp64 operations per 16 B load are free.
__device__ __forceinline__ void add4(double (&s)[4], const float4& v) {
    s[0] += static_cast<double>(v.x); s[1] += static_cast<double>(v.y); s[2] += static_cast<double>(v.z); s[3] += static_cast<double>(v.w);
}
__device__ __forceinline__ void addsq4(double (&s)[4], const float4& v) {
    s[0] = fma(static_cast<double>(v.x), static_cast<double>(v.x), s[0]);
    s[1] = fma(static_cast<double>(v.y), static_cast<double>(v.y), s[1]);
    s[2] = fma(static_cast<double>(v.z), static_cast<double>(v.z), s[2]);
    s[3] = fma(static_cast<double>(v.w), static_cast<double>(v.w), s[3]);
}

// reduce 8 per-thread partials over the threads that share a channel quad, then one fp64 atomic per channel per CTA
template <int T = THREADS, typename P = float>
__device__ __forceinline__ void block_reduce_to_global(const P (&a)[4], const P (&b)[4], const Map& m, double* out, int C) {
    __shared__ P red[T][8];
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        red[t][k] = a[k];
        red[t][4 + k] = b[k];
    }
    __syncthreads();
    // thread (quad, j) for j < 8 sums column j of its quad over the rows_per_iter threads
    for (int o = t; o < m.q * 8; o += T) {
        const int quad = o >> 3, j = o & 7;
        double s = 0.0;
        for (int r = 0; r < m.rows_per_iter; ++r) s += static_cast<double>(red[r * m.q + quad][j]);
        const int ch = quad * 4 + (j & 3);
        atomicAdd(out + (j < 4 ? ch : C + ch), s);
    }
}

// Data-parallel training (SyncBatchNorm): the LAST block to finish the reduction exchanges the 2C fp64 sums with the other
// ranks over NVLink peer memory (csrc/peer.cuh) and leaves the global sums in place -- no separate exchange launch between
// the reduction and the kernel that consumes it.  sums[2C] is the arrival counter (zeroed with the sums).
__device__ __forceinline__ void exchange_if_last(double* sums, int C, const PeerArgs pa, double* local_copy = nullptr) {
    if (pa.peers == nullptr) return;
    __shared__ int s_last;
    __threadfence();                                   // this thread's atomics are visible device-wide
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(reinterpret_cast<unsigned int*>(sums + 2 * C), 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (local_copy != nullptr)                         // this rank's own sums (the BatchNorm parameter gradients)
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) local_copy[i] = ld_volatile_f64(sums + i);
    peer_allreduce_block(sums, sums, 2 * C, pa);
}

// ---- forward
__global__ void __launch_bounds__(THREADS) bn_stats_kernel(const float* __restrict__ x, long long M, int C, double* __restrict__ sums,
                                                           const PeerArgs pa) {
    const Map m = make_map(C);
    double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
    const float4* xp = reinterpret_cast<const float4*>(x);
    for (long long r = static_cast<long long>(blockIdx.x) * m.rows_per_iter + m.row0; r < M;
         r += static_cast<long long>(gridDim.x) * m.rows_per_iter) {
        const float4 v = __ldg(xp + r * m.q + m.quad);
        add4(s, v);
        addsq4(ss, v);
    }
    block_reduce_to_global<THREADS, double>(s, ss, m, sums, C);
    exchange_if_last(sums, C, pa);
}

// one thread per channel.  stats: [mean C][invstd C][scale C][shift C]
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ weight,
                                   const float* __restrict__ bias, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps, float* __restrict__ stats, int C,
                                   long long* __restrict__ num_batches_tracked) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked != nullptr) num_batches_tracked[0] += 1;      // the module's step counter rides along
    if (c >= C) return;
    const double mean = sums[c] / count;
    double var = sums[C + c] / count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
    const float scale = w * invstd;
    stats[c] = static_cast<float>(mean);
    stats[C + c] = invstd;
    stats[2 * C + c] = scale;
    stats[3 * C + c] = b - static_cast<float>(mean) * scale;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
}

__global__ void __launch_bounds__(THREADS) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                           const float* __restrict__ stats, float* __restrict__ y, long long M,
                                                           int C, int relu) {
    const Map m = make_map(C);
    const float4 sc = *reinterpret_cast<const float4*>(stats + 2 * C + 4 * m.quad);
    const float4 sh = *reinterpret_cast<const float4*>(stats + 3 * C + 4 * m.quad);
    const float4* xp = reinterpret_cast<const float4*>(x);
    const float4* rp = reinterpret_cast<const float4*>(res);
    float4* yp = reinterpret_cast<float4*>(y);
    for (long long r = static_cast<long long>(blockIdx.x) * m.rows_per_iter + m.row0; r < M;
         r += static_cast<long long>(gridDim.x) * m.rows_per_iter) {
        const long long i = r * m.q + m.quad;
        const float4 v = __ldg(xp + i);
        float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
        if (rp && !(relu & 2)) {                         // ResNet block: add, then ReLU
            const float4 rv = __ldg(rp + i);
            o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
        }
        if (relu & 1) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
        if (rp && (relu & 2)) {                          // U-Net skip: ReLU, then add (resnet_encoder.py:272-276)
            const float4 rv = __ldg(rp + i);
            o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
        }
        yp[i] = o;
    }
}

// ---- backward
__device__ __forceinline__ float4 masked(const float4& g, const float4& y, int relu) {
    if (!relu) return g;
    return make_float4(y.x > 0.f ? g.x : 0.f, y.y > 0.f ? g.y : 0.f, y.z > 0.f ? g.z : 0.f, y.w > 0.f ? g.w : 0.f);
}

// the forward's pre-activation x*scale + shift, bit-identical to what bn_apply_kernel clamped (no residual): its sign is
// the ReLU mask, so the backward does not have to read y
__device__ __forceinline__ float4 affine(const float4& v, const float4& sc, const float4& sh) {
    return make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
}

// sums2: [sum g (C)][sum g*xhat (C)]
__global__ void __launch_bounds__(THREADS) bn_bwd_reduce_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                                const float* __restrict__ y, const float* __restrict__ stats,
                                                                double* __restrict__ sums2, double* __restrict__ local_sums2,
                                                                long long M, int C, int relu, const PeerArgs pa) {
    const Map m = make_map(C);
    const float4 mean = *reinterpret_cast<const float4*>(stats + 4 * m.quad);
    const float4 istd = *reinterpret_cast<const float4*>(stats + C + 4 * m.quad);
    const float4 sc = *reinterpret_cast<const float4*>(stats + 2 * C + 4 * m.quad);
    const float4 sh = *reinterpret_cast<const float4*>(stats + 3 * C + 4 * m.quad);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* gp = reinterpret_cast<const float4*>(gy);
    const float4* xp = reinterpret_cast<const float4*>(x);
    const float4* yp = reinterpret_cast<const float4*>(y);
    for (long long r = static_cast<long long>(blockIdx.x) * m.rows_per_iter + m.row0; r < M;
         r += static_cast<long long>(gridDim.x) * m.rows_per_iter) {
        const long long i = r * m.q + m.quad;
        float4 g = __ldg(gp + i);
        const float4 v = __ldg(xp + i);
        if (relu) g = masked(g, yp ? __ldg(yp + i) : affine(v, sc, sh), 1);
        add4(s, g);
        sx[0] = fmaf(g.x, (v.x - mean.x) * istd.x, sx[0]);
        sx[1] = fmaf(g.y, (v.y - mean.y) * istd.y, sx[1]);
        sx[2] = fmaf(g.z, (v.z - mean.z) * istd.z, sx[2]);
        sx[3] = fmaf(g.w, (v.w - mean.w) * istd.w, sx[3]);
    }
    block_reduce_to_global(s, sx, m, sums2, C);
    exchange_if_last(sums2, C, pa, local_sums2);
}

__global__ void __launch_bounds__(THREADS) bn_bwd_apply_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                               const float* __restrict__ y, const float* __restrict__ stats,
                                                               const float* __restrict__ weight, const double* __restrict__ sums2,
                                                               double count, float* __restrict__ gx, float* __restrict__ gres,
                                                               float* __restrict__ gw, float* __restrict__ gb, long long M, int C,
                                                               int relu) {
    const Map m = make_map(C);
    const float4 mean = *reinterpret_cast<const float4*>(stats + 4 * m.quad);
    const float4 istd = *reinterpret_cast<const float4*>(stats + C + 4 * m.quad);
    const float4 sc = *reinterpret_cast<const float4*>(stats + 2 * C + 4 * m.quad);
    const float4 sh = *reinterpret_cast<const float4*>(stats + 3 * C + 4 * m.quad);
    float k[4], mg[4], mgx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = 4 * m.quad + j;
        const float w = weight ? weight[c] : 1.f;
        k[j] = w * reinterpret_cast<const float*>(&istd)[j];
        mg[j] = static_cast<float>(sums2[c] / count);
        mgx[j] = static_cast<float>(sums2[C + c] / count);
    }
    if (blockIdx.x == 0) {                                 // parameter gradients (NULL when sums2 was all-reduced: the
        for (int c = threadIdx.x; c < C; c += THREADS) {   // caller then takes them from its local sums)
            if (gw) gw[c] = static_cast<float>(sums2[C + c]);
            if (gb) gb[c] = static_cast<float>(sums2[c]);
        }
    }
    const float4* gp = reinterpret_cast<const float4*>(gy);
    const float4* xp = reinterpret_cast<const float4*>(x);
    const float4* yp = reinterpret_cast<const float4*>(y);
    float4* gxp = reinterpret_cast<float4*>(gx);
    float4* grp = reinterpret_cast<float4*>(gres);
    for (long long r = static_cast<long long>(blockIdx.x) * m.rows_per_iter + m.row0; r < M;
         r += static_cast<long long>(gridDim.x) * m.rows_per_iter) {
        const long long i = r * m.q + m.quad;
        float4 g = __ldg(gp + i);
        const float4 v = __ldg(xp + i);
        if (relu) g = masked(g, yp ? __ldg(yp + i) : affine(v, sc, sh), 1);
        float4 o;
        o.x = k[0] * (g.x - mg[0] - (v.x - mean.x) * istd.x * mgx[0]);
        o.y = k[1] * (g.y - mg[1] - (v.y - mean.y) * istd.y * mgx[1]);
        o.z = k[2] * (g.z - mg[2] - (v.z - mean.z) * istd.z * mgx[2]);
        o.w = k[3] * (g.w - mg[3] - (v.w - mean.w) * istd.w * mgx[3]);
        gxp[i] = o;
        if (grp) grp[i] = g;
    }
}

// ---- one-kernel forward / backward for activations that fit the L2 ---------------------------------------------------
// Most BatchNorm layers of the step are small (ResNet / FPN4 / coarse reg3d levels: 1.5 - 50 MB against a 126 MB L2): the
// three-kernel form costs them four graph nodes (memset, stats, finalize, apply) of mostly launch latency, and reads x from
// DRAM twice.  The fused form is ONE launch of <= one CTA per SM: phase 1 reduces, a grid barrier (arrival counter; the last
// block to arrive performs the SyncBatchNorm exchange over NVLink peer memory before it opens the barrier), phase 2 applies
// with the second read of x served by the L2.  The workspace (2*1024 fp64 sums + 4 control words) starts zeroed and is handed
// back zeroed: the last block to have consumed the sums clears them, so no memset node precedes the kernel and every layer
// on one stream can share one workspace.  Co-residency: the grid is at most 2/3 of the SM count, a block is 512 threads of
// <= 64 registers with 24 KB of shared memory (two fit one SM), so the two such kernels the trainer's two streams can have in
// flight need at most 1/3 of the machine's block slots: they are always fully resident together even while an NCCL kernel
// holds tens of SMs -- the barrier cannot deadlock -- and the spin is bounded (control word 3 records a timeout instead of
// hanging the GPU).
constexpr int FT = 512;
constexpr int FUSED_CMAX = 1024;
constexpr long long BARRIER_SPIN_LIMIT = 8000000000ll;

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ctl: [0] arrivals, [1] barrier open, [2] blocks that have consumed the sums, [3] timeout flag (sticky)
__device__ __forceinline__ void grid_barrier_with_exchange(double* sums, unsigned int* ctl, int C, const PeerArgs pa, double* local_copy) {
    __shared__ int s_last;
    __threadfence();                                   // this block's atomics are visible device-wide
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ctl, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (pa.peers != nullptr) {
            if (local_copy != nullptr)
                for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) local_copy[i] = ld_volatile_f64(sums + i);
            peer_allreduce_block(sums, sums, 2 * C, pa);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu_u32(ctl + 1, 1u);
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(ctl + 1) == 0u) {
                if (clock64() - t0 > BARRIER_SPIN_LIMIT) {
                    ctl[3] = 1u;
                    break;
                }
                __nanosleep(64);
            }
        }
        __syncthreads();
    }
}

// every thread of the block has taken what it needs from the sums: the last block to say so hands the workspace back zeroed
__device__ __forceinline__ void release_workspace(double* sums, unsigned int* ctl, int C) {
    __shared__ int s_last_out;
    __syncthreads();
    if (threadIdx.x == 0) s_last_out = (atomicAdd(ctl + 2, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last_out) {
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sums[i] = 0.0;
        if (threadIdx.x < 3) ctl[threadIdx.x] = 0u;
    }
}

__global__ void __launch_bounds__(FT, 2) bn_fwd_fused_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                          const float* __restrict__ weight, const float* __restrict__ bias,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var,
                                                          long long* __restrict__ num_batches_tracked, float momentum, float eps,
                                                          double count, float* __restrict__ stats, float* __restrict__ y,
                                                          long long M, int C, int relu, double* __restrict__ ws, const PeerArgs pa) {
    __shared__ __align__(16) float s_scale[FUSED_CMAX], s_shift[FUSED_CMAX];
    const Map m = make_map<FT>(C);
    unsigned int* ctl = reinterpret_cast<unsigned int*>(ws + 2 * FUSED_CMAX);
    const float4* xp = reinterpret_cast<const float4*>(x);
    const long long step = static_cast<long long>(gridDim.x) * m.rows_per_iter;
    const long long r0 = static_cast<long long>(blockIdx.x) * m.rows_per_iter + m.row0;
    {
        double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
        long long r = r0;
        for (; r + 3 * step < M; r += 4 * step) {                                  // four independent loads in flight
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(xp + (r + u * step) * m.q + m.quad);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                add4(s, v[u]);
                addsq4(ss, v[u]);
            }
        }
        for (; r < M; r += step) {
            const float4 v = __ldg(xp + r * m.q + m.quad);
            add4(s, v);
            addsq4(ss, v);
        }
        block_reduce_to_global<FT, double>(s, ss, m, ws, C);
    }
    grid_barrier_with_exchange(ws, ctl, C, pa, nullptr);
    // the arithmetic of bn_finalize_kernel, once per channel per block; block 0 also publishes it
    for (int c = threadIdx.x; c < C; c += FT) {
        const double mean = ld_volatile_f64(ws + c) / count;
        double var = ld_volatile_f64(ws + C + c) / count - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        const float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
        const float scale = w * invstd;
        const float shift = b - static_cast<float>(mean) * scale;
        s_scale[c] = scale;
        s_shift[c] = shift;
        if (blockIdx.x == 0) {
            stats[c] = static_cast<float>(mean);
            stats[C + c] = invstd;
            stats[2 * C + c] = scale;
            stats[3 * C + c] = shift;
            if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
            if (running_var) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) num_batches_tracked[0] += 1;
    release_workspace(ws, ctl, C);                     // begins with __syncthreads(): s_scale / s_shift are complete
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + 4 * m.quad);
    const float4 sh = *reinterpret_cast<const float4*>(s_shift + 4 * m.quad);
    const float4* rp = reinterpret_cast<const float4*>(res);
    float4* yp = reinterpret_cast<float4*>(y);
    auto finish = [&](float4 o, long long i) {
        if (rp && !(relu & 2)) {
            const float4 rv = __ldg(rp + i);
            o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
        }
        if (relu & 1) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
        if (rp && (relu & 2)) {
            const float4 rv = __ldg(rp + i);
            o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
        }
        yp[i] = o;
    };
    long long r = r0;
    for (; r + 3 * step < M; r += 4 * step) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(xp + (r + u * step) * m.q + m.quad);
#pragma unroll
        for (int u = 0; u < 4; ++u) finish(affine(v[u], sc, sh), (r + u * step) * m.q + m.quad);
    }
    for (; r < M; r += step) finish(affine(__ldg(xp + r * m.q + m.quad), sc, sh), r * m.q + m.quad);
}

__global__ void __launch_bounds__(FT, 2) bn_bwd_fused_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                          const float* __restrict__ y, const float* __restrict__ stats,
                                                          const float* __restrict__ weight, double count, float* __restrict__ gx,
                                                          float* __restrict__ gres, float* __restrict__ gw, float* __restrict__ gb,
                                                          double* __restrict__ local_sums2, long long M, int C, int relu,
                                                          double* __restrict__ ws, const PeerArgs pa) {
    const Map m = make_map<FT>(C);
    unsigned int* ctl = reinterpret_cast<unsigned int*>(ws + 2 * FUSED_CMAX);
    const float4 mean = *reinterpret_cast<const float4*>(stats + 4 * m.quad);
    const float4 istd = *reinterpret_cast<const float4*>(stats + C + 4 * m.quad);
    const float4 sc = *reinterpret_cast<const float4*>(stats + 2 * C + 4 * m.quad);
    const float4 sh = *reinterpret_cast<const float4*>(stats + 3 * C + 4 * m.quad);
    const float4* gp = reinterpret_cast<const float4*>(gy);
    const float4* xp = reinterpret_cast<const float4*>(x);
    const float4* yp = reinterpret_cast<const float4*>(y);
    const long long step = static_cast<long long>(gridDim.x) * m.rows_per_iter;
    const long long r0 = static_cast<long long>(blockIdx.x) * m.rows_per_iter + m.row0;
    auto load_g = [&](long long i, const float4& v) {
        float4 g = __ldg(gp + i);
        if (relu) g = masked(g, yp ? __ldg(yp + i) : affine(v, sc, sh), 1);
        return g;
    };
    {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f};
        auto acc = [&](const float4& g, const float4& v) {
            add4(s, g);
            sx[0] = fmaf(g.x, (v.x - mean.x) * istd.x, sx[0]);
            sx[1] = fmaf(g.y, (v.y - mean.y) * istd.y, sx[1]);
            sx[2] = fmaf(g.z, (v.z - mean.z) * istd.z, sx[2]);
            sx[3] = fmaf(g.w, (v.w - mean.w) * istd.w, sx[3]);
        };
        long long r = r0;
        for (; r + step < M; r += 2 * step) {
            const long long i0 = r * m.q + m.quad, i1 = (r + step) * m.q + m.quad;
            const float4 v0 = __ldg(xp + i0), v1 = __ldg(xp + i1);
            const float4 g0 = load_g(i0, v0), g1 = load_g(i1, v1);
            acc(g0, v0);
            acc(g1, v1);
        }
        for (; r < M; r += step) {
            const long long i = r * m.q + m.quad;
            const float4 v = __ldg(xp + i);
            acc(load_g(i, v), v);
        }
        block_reduce_to_global<FT>(s, sx, m, ws, C);
    }
    grid_barrier_with_exchange(ws, ctl, C, pa, local_sums2);
    float k[4], mg[4], mgx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = 4 * m.quad + j;
        const float w = weight ? weight[c] : 1.f;
        k[j] = w * reinterpret_cast<const float*>(&istd)[j];
        mg[j] = static_cast<float>(ld_volatile_f64(ws + c) / count);
        mgx[j] = static_cast<float>(ld_volatile_f64(ws + C + c) / count);
    }
    if (blockIdx.x == 0) {                                 // parameter gradients (NULL under SyncBatchNorm: local_sums2 has them)
        for (int c = threadIdx.x; c < C; c += FT) {
            if (gw) gw[c] = static_cast<float>(ld_volatile_f64(ws + C + c));
            if (gb) gb[c] = static_cast<float>(ld_volatile_f64(ws + c));
        }
    }
    release_workspace(ws, ctl, C);
    float4* gxp = reinterpret_cast<float4*>(gx);
    float4* grp = reinterpret_cast<float4*>(gres);
    auto emit = [&](long long i, const float4& g, const float4& v) {
        float4 o;
        o.x = k[0] * (g.x - mg[0] - (v.x - mean.x) * istd.x * mgx[0]);
        o.y = k[1] * (g.y - mg[1] - (v.y - mean.y) * istd.y * mgx[1]);
        o.z = k[2] * (g.z - mg[2] - (v.z - mean.z) * istd.z * mgx[2]);
        o.w = k[3] * (g.w - mg[3] - (v.w - mean.w) * istd.w * mgx[3]);
        gxp[i] = o;
        if (grp) grp[i] = g;
    };
    long long r = r0;
    for (; r + step < M; r += 2 * step) {
        const long long i0 = r * m.q + m.quad, i1 = (r + step) * m.q + m.quad;
        const float4 v0 = __ldg(xp + i0), v1 = __ldg(xp + i1);
        const float4 g0 = load_g(i0, v0), g1 = load_g(i1, v1);
        emit(i0, g0, v0);
        emit(i1, g1, v1);
    }
    for (; r < M; r += step) {
        const long long i = r * m.q + m.quad;
        const float4 v = __ldg(xp + i);
        emit(i, load_g(i, v), v);
    }
}

// at most 2/3 of the SM count (see above); small tensors get as many blocks as give every thread >= 2 rows
static int fused_grid_for(long long M, int C) {
    const int rows_per_iter = FT / (C >> 2);
    const long long blocks = (M + 2LL * rows_per_iter - 1) / (2LL * rows_per_iter);
    const long long cap = (2LL * sm_count()) / 3;
    return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

static int grid_for(long long M, int C) {
    const int rows_per_iter = THREADS / (C >> 2);
    const long long blocks = (M + rows_per_iter - 1) / rows_per_iter;
    const long long cap = static_cast<long long>(sm_count()) * 8;
    return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// Reductions end with 2C fp64 atomics per block on the SAME 2C addresses: with one block per 256/q rows a mid-sized tensor
// (e.g. [6,64,48,160]) launched 1184 blocks of ~2 iterations each and spent its time in 150 k contended atomics.  Give every
// thread >= 16 rows (the full-resolution volumes still fill 8 blocks per SM).
static int reduce_grid_for(long long M, int C) {
    const int rows_per_iter = THREADS / (C >> 2);
    const long long blocks = (M + 16LL * rows_per_iter - 1) / (16LL * rows_per_iter);
    const long long cap = static_cast<long long>(sm_count()) * 8;
    return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

static int check(long long M, int C) {
    MVD_REQUIRE(M > 0 && C >= 4 && C <= 1024 && (C & (C - 1)) == 0, "BatchNorm kernels need C a power of two in [4,1024] (got M=%lld C=%d)", M, C);
    return 0;
}

}  // namespace bn
}  // namespace mvd

extern "C" {

static int check_peer(const unsigned long long* peers, int rank, int world, int nmax, int C) {
    if (peers == nullptr) return 0;
    MVD_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    MVD_REQUIRE(2 * C <= nmax, "2C = %d statistics exceed the exchange buffer's %d", 2 * C, nmax);
    return 0;
}

int mvd_bn_stats(const float* x, long long M, int C, double* sums, const unsigned long long* peers, int rank, int world, int nmax,
                 void* stream) {
    using namespace mvd::bn;
    MVD_REQUIRE(x && sums, "null pointer argument");
    if (int rc = check(M, C)) return rc;
    if (int rc = check_peer(peers, rank, world, nmax, C)) return rc;
    cudaStream_t st = mvd::as_stream(stream);
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * (2 * C + 1), st);
    if (e != cudaSuccess) return mvd::fail(static_cast<int>(e), "bn_stats memset: %s", cudaGetErrorString(e));
    bn_stats_kernel<<<reduce_grid_for(M, C), THREADS, 0, st>>>(x, M, C, sums, mvd::PeerArgs{peers, rank, world, nmax});
    return mvd::check_launch("bn_stats");
}

int mvd_bn_finalize(const double* sums, double count, const float* weight, const float* bias, float* running_mean,
                    float* running_var, float momentum, float eps, float* stats, int C, long long* num_batches_tracked, void* stream) {
    using namespace mvd::bn;
    MVD_REQUIRE(sums && stats && count > 0, "bad argument");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, mvd::as_stream(stream)>>>(sums, count, weight, bias, running_mean, running_var,
                                                                            momentum, eps, stats, C, num_batches_tracked);
    return mvd::check_launch("bn_finalize");
}

int mvd_bn_apply(const float* x, const float* residual, const float* stats, float* y, long long M, int C, int relu, void* stream) {
    using namespace mvd::bn;
    MVD_REQUIRE(x && stats && y, "null pointer argument");
    if (int rc = check(M, C)) return rc;
    bn_apply_kernel<<<grid_for(M, C), THREADS, 0, mvd::as_stream(stream)>>>(x, residual, stats, y, M, C, relu);
    return mvd::check_launch("bn_apply");
}

int mvd_bn_bwd_reduce(const float* gy, const float* x, const float* y, const float* stats, double* sums2, double* local_sums2,
                      long long M, int C, int relu, const unsigned long long* peers, int rank, int world, int nmax, void* stream) {
    using namespace mvd::bn;
    MVD_REQUIRE(gy && x && stats && sums2, "null pointer argument");
    MVD_REQUIRE(peers == nullptr || local_sums2 != nullptr, "data-parallel reduce needs local_sums2 (this rank's own sums)");
    if (int rc = check(M, C)) return rc;
    if (int rc = check_peer(peers, rank, world, nmax, C)) return rc;
    cudaStream_t st = mvd::as_stream(stream);
    cudaError_t e = cudaMemsetAsync(sums2, 0, sizeof(double) * (2 * C + 1), st);
    if (e != cudaSuccess) return mvd::fail(static_cast<int>(e), "bn_bwd_reduce memset: %s", cudaGetErrorString(e));
    bn_bwd_reduce_kernel<<<reduce_grid_for(M, C), THREADS, 0, st>>>(gy, x, y, stats, sums2, local_sums2, M, C, relu,
                                                             mvd::PeerArgs{peers, rank, world, nmax});
    return mvd::check_launch("bn_bwd_reduce");
}

int mvd_bn_bwd_apply(const float* gy, const float* x, const float* y, const float* stats, const float* weight,
                     const double* sums2, double count, float* gx, float* gres, float* gw, float* gb, long long M, int C,
                     int relu, void* stream) {
    using namespace mvd::bn;
    MVD_REQUIRE(gy && x && stats && sums2 && gx && count > 0, "bad argument");
    if (int rc = check(M, C)) return rc;
    bn_bwd_apply_kernel<<<grid_for(M, C), THREADS, 0, mvd::as_stream(stream)>>>(gy, x, y, stats, weight, sums2, count, gx, gres, gw,
                                                                                 gb, M, C, relu);
    return mvd::check_launch("bn_bwd_apply");
}

int mvd_bn_workspace_doubles(void) { return 2 * mvd::bn::FUSED_CMAX + 2; }

int mvd_bn_fwd_fused(const float* x, const float* residual, const float* weight, const float* bias, float* running_mean,
                     float* running_var, long long* num_batches_tracked, float momentum, float eps, double count, float* stats,
                     float* y, long long M, int C, int relu, double* workspace, const unsigned long long* peers, int rank,
                     int world, int nmax, void* stream) {
    using namespace mvd::bn;
    MVD_REQUIRE(x && stats && y && workspace && count > 0, "bad argument");
    if (int rc = check(M, C)) return rc;
    if (int rc = check_peer(peers, rank, world, nmax, C)) return rc;
    bn_fwd_fused_kernel<<<fused_grid_for(M, C), FT, 0, mvd::as_stream(stream)>>>(
        x, residual, weight, bias, running_mean, running_var, num_batches_tracked, momentum, eps, count, stats, y, M, C, relu,
        workspace, mvd::PeerArgs{peers, rank, world, nmax});
    return mvd::check_launch("bn_fwd_fused");
}

int mvd_bn_bwd_fused(const float* gy, const float* x, const float* y, const float* stats, const float* weight, double count,
                     float* gx, float* gres, float* gw, float* gb, double* local_sums2, long long M, int C, int relu,
                     double* workspace, const unsigned long long* peers, int rank, int world, int nmax, void* stream) {
    using namespace mvd::bn;
    MVD_REQUIRE(gy && x && stats && gx && workspace && count > 0, "bad argument");
    MVD_REQUIRE(peers == nullptr || local_sums2 != nullptr, "data-parallel backward needs local_sums2 (this rank's own sums)");
    if (int rc = check(M, C)) return rc;
    if (int rc = check_peer(peers, rank, world, nmax, C)) return rc;
    bn_bwd_fused_kernel<<<fused_grid_for(M, C), FT, 0, mvd::as_stream(stream)>>>(
        gy, x, y, stats, weight, count, gx, gres, gw, gb, local_sums2, M, C, relu, workspace, mvd::PeerArgs{peers, rank, world, nmax});
    return mvd::check_launch("bn_bwd_fused");
}

}
