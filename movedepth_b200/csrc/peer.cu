// Small all-reduce over NVLink peer memory for the SyncBatchNorm statistics: the 2C fp64 sums of one BatchNorm pass
// are exchanged by ONE single-CTA kernel per rank -- every rank stores its vector into its slot of every peer's
// symmetric buffer (remote st.global over NVLink/NVSwitch), raises a per-peer epoch flag (st.release.sys), spins on the
// flags the peers raise in its own buffer (ld.acquire.sys) and adds the slots in rank order (bitwise identical result on
// every rank).  190 of these per training step replace 190 NCCL calls (launch + protocol latency ~30 us each at 8
// GPUs); being plain kernels they live inside the step's CUDA graph.
//
// Symmetric buffer layout (same on every rank; allocated by the caller, zero-initialised):
//   [0, 8)                       epoch counter of this rank (local)
//   [8, 16)                      time-out marker: set to the epoch at which a peer's flag did not arrive within ~4 s
//                                (the wait is bounded so that a dead or never-launched peer cannot hang the GPU)
//   [64, 64 + 8*world)           flags: flag[p] = last epoch rank p has published into this buffer
//   [4096, ...)                  data [2 slots][world][nmax] doubles      (slot = epoch & 1)
#include "common.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {

constexpr int PEER_HEADER = 4096;
constexpr long long PEER_SPIN_LIMIT = 8000000000ll;       // SM clocks (~4 s at 1.97 GHz)

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_f64_kernel(const double* __restrict__ local, double* __restrict__ out, int n,
                                                                 const unsigned long long* __restrict__ peers, int rank, int world,
                                                                 int nmax) {
    __shared__ unsigned long long s_epoch;
    const int tid = threadIdx.x;
    unsigned char* mine = reinterpret_cast<unsigned char*>(peers[rank]);
    if (tid == 0) {
        unsigned long long* ctr = reinterpret_cast<unsigned long long*>(mine);
        s_epoch = *ctr + 1ull;
        *ctr = s_epoch;
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const size_t slot = static_cast<size_t>(epoch & 1ull);
    // 1. publish: my vector into slot [slot][rank] of every rank's buffer (the local copy included)
    for (int p = 0; p < world; ++p) {
        double* dst = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(peers[p]) + PEER_HEADER) + (slot * world + rank) * nmax;
        for (int i = tid; i < n; i += blockDim.x) dst[i] = local[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag on peer `tid`, then wait for peer `tid`'s flag in my buffer
    if (tid < world) {
        st_release_sys(reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(peers[tid]) + 64) + rank, epoch);
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + 64) + tid;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < epoch) {
            if (clock64() - t0 > PEER_SPIN_LIMIT) {
                reinterpret_cast<unsigned long long*>(mine)[1] = epoch;
                break;
            }
        }
    }
    __syncthreads();
    // 3. reduce in rank order
    const double* src = reinterpret_cast<const double*>(mine + PEER_HEADER) + slot * world * nmax;
    for (int i = tid; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int p = 0; p < world; ++p) s += ld_volatile_f64(src + static_cast<size_t>(p) * nmax + i);
        out[i] = s;
    }
}

}  // namespace mvd

extern "C" {

long long mvd_peer_allreduce_buffer_bytes(int world, int nmax) {
    if (world <= 0 || nmax <= 0) return 0;
    return static_cast<long long>(mvd::PEER_HEADER) + 2ll * world * nmax * static_cast<long long>(sizeof(double));
}

int mvd_peer_allreduce_f64(const double* local, double* out, int n, const unsigned long long* peers, int rank, int world, int nmax,
                           void* stream) {
    MVD_REQUIRE(local && out && peers, "null pointer argument");
    MVD_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    MVD_REQUIRE(n > 0 && n <= nmax, "vector length %d exceeds the exchange buffer's %d", n, nmax);
    mvd::peer_allreduce_f64_kernel<<<1, 256, 0, mvd::as_stream(stream)>>>(local, out, n, peers, rank, world, nmax);
    return mvd::check_launch("peer_allreduce_f64");
}

}
