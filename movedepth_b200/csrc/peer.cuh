// Device side of the NVLink peer-memory all-reduce of small fp64 vectors (SyncBatchNorm statistics); see peer.cu for the
// protocol and the symmetric buffer layout.  `peer_allreduce_block` is executed by ONE thread block: either the stand-alone
// kernel of peer.cu, or the LAST block of a BatchNorm reduction kernel (csrc/bn.cu), which saves the separate launch.
#pragma once
#include "common.cuh"

namespace mvd {

constexpr int PEER_HEADER = 4096;
constexpr long long PEER_SPIN_LIMIT = 8000000000ll;       // SM clocks (~4 s at 1.97 GHz)

struct PeerArgs {
    const unsigned long long* peers;   // device table of the ranks' symmetric buffer addresses; nullptr: single process
    int rank, world, nmax;
};

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

// out[i] = sum over ranks of local[i], i < n (out may alias local).  Called by every thread of one block.
__device__ __forceinline__ void peer_allreduce_block(const double* local, double* out, int n, const PeerArgs pa) {
    __shared__ unsigned long long s_epoch;
    const int tid = threadIdx.x, nthr = blockDim.x;
    unsigned char* mine = reinterpret_cast<unsigned char*>(pa.peers[pa.rank]);
    __syncthreads();
    if (tid == 0) {
        unsigned long long* ctr = reinterpret_cast<unsigned long long*>(mine);
        s_epoch = *ctr + 1ull;
        *ctr = s_epoch;
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const size_t slot = static_cast<size_t>(epoch & 1ull);
    // 1. publish: my vector into slot [slot][rank] of every rank's buffer (the local copy included)
    for (int p = 0; p < pa.world; ++p) {
        double* dst = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(pa.peers[p]) + PEER_HEADER) +
                      (slot * pa.world + pa.rank) * pa.nmax;
        for (int i = tid; i < n; i += nthr) dst[i] = ld_volatile_f64(local + i);
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag on peer `tid`, then wait for peer `tid`'s flag in my buffer
    if (tid < pa.world) {
        st_release_sys(reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(pa.peers[tid]) + 64) + pa.rank, epoch);
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
    // 3. reduce in rank order (bitwise identical on every rank)
    const double* src = reinterpret_cast<const double*>(mine + PEER_HEADER) + slot * pa.world * pa.nmax;
    for (int i = tid; i < n; i += nthr) {
        double s = 0.0;
        for (int p = 0; p < pa.world; ++p) s += ld_volatile_f64(src + static_cast<size_t>(p) * pa.nmax + i);
        out[i] = s;
    }
}

}  // namespace mvd
