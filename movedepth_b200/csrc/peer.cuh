// Device side of the NVLink peer-memory all-reduce of small fp64 vectors (SyncBatchNorm statistics); see peer.cu for the
// protocol and the symmetric buffer layout.  `peer_allreduce_block` is executed by ONE thread block: either the stand-alone
// kernel of peer.cu, or the LAST block of a BatchNorm reduction kernel (csrc/bn.cu), which saves the separate launch.
#pragma once
#include "common.cuh"

namespace mvd {

constexpr int PEER_HEADER = 4096;
constexpr int PEER_ENTRY = 16;                            // bytes per double in flight: {low, tag, high, tag}
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

__device__ __forceinline__ void st_volatile_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_volatile_v4(const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
}

// out[i] = sum over ranks of local[i], i < n (out may alias local).  Called by every thread of one block.
// Low-latency protocol (flag-in-data, as NCCL's LL): every double travels as two 8-byte words {low half, epoch},
// {high half, epoch}; an aligned 8-byte store is one NVLink transaction, so the receiver simply polls each word until its
// epoch tag matches -- no __threadfence_system() and no separate flag round trip (they cost ~8 of the former 13 us).
// Entry layout per rank buffer: [2 slots (epoch parity)][world][nmax] x 16 B.
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
    const uint32_t tag = static_cast<uint32_t>(s_epoch);             // never 0: the buffers start zeroed
    const size_t slot = static_cast<size_t>(s_epoch & 1ull);
    // 1. publish my vector into entry [slot][rank] of every OTHER rank's buffer; my own values stay in registers / local
    for (int i = tid; i < n; i += nthr) {
        const double v = ld_volatile_f64(local + i);
        const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
        const uint32_t lo = static_cast<uint32_t>(bits), hi = static_cast<uint32_t>(bits >> 32);
        for (int p = 0; p < pa.world; ++p) {
            if (p == pa.rank) continue;
            unsigned char* dst = reinterpret_cast<unsigned char*>(pa.peers[p]) + PEER_HEADER +
                                 ((slot * pa.world + pa.rank) * pa.nmax + i) * 16;
            st_volatile_v4(dst, lo, tag, hi, tag);
        }
    }
    // 2. gather: poll the other ranks' entries in my buffer until their tags match, add in rank order (bitwise identical
    //    on every rank: the own contribution enters at its rank position)
    const long long t0 = clock64();
    bool timed_out = false;
    for (int i = tid; i < n; i += nthr) {
        const double own = ld_volatile_f64(local + i);
        double s = 0.0;
        for (int p = 0; p < pa.world; ++p) {
            if (p == pa.rank) {
                s += own;
                continue;
            }
            const unsigned char* src = mine + PEER_HEADER + ((slot * pa.world + p) * pa.nmax + i) * 16;
            uint32_t lo, t1, hi, t2;
            ld_volatile_v4(src, lo, t1, hi, t2);
            while ((t1 != tag || t2 != tag) && !timed_out) {
                if (clock64() - t0 > PEER_SPIN_LIMIT) {
                    timed_out = true;
                    reinterpret_cast<unsigned long long*>(mine)[1] = s_epoch;
                }
                ld_volatile_v4(src, lo, t1, hi, t2);
            }
            s += __longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(hi) << 32) | lo));
        }
        out[i] = s;
    }
}

}  // namespace mvd
