// Small all-reduce over NVLink peer memory for the SyncBatchNorm statistics: the 2C fp64 sums of one BatchNorm pass are
// exchanged by ONE thread block per rank -- every rank stores its vector into its entries of every peer's symmetric buffer
// (remote st.volatile over NVLink/NVSwitch), each double as two 8-byte words tagged with the call's epoch; the receiver
// polls the words of its own buffer until the tags match and adds the contributions in rank order (bitwise identical result
// on every rank).  No fence, no separate flag (NCCL's LL idea).  190 of these per training step replace 190 NCCL calls
// (launch + protocol latency ~25-30 us each); inside the BatchNorm reduction kernels (csrc/bn.cu) they are not even a
// launch of their own, and being plain kernels they live inside the step's CUDA graph.
//
// Symmetric buffer layout (same on every rank; allocated by the caller, zero-initialised):
//   [0, 8)       epoch counter of this rank (local; the 32-bit tag of a call is its low word, never 0)
//   [8, 16)      time-out marker: set to the epoch at which a peer's data did not arrive within ~4 s (the wait is bounded
//                so that a dead or never-launched peer cannot hang the GPU)
//   [4096, ...)  entries [2 slots = epoch parity][world][nmax] x 16 B = {low half, tag, high half, tag}
#include "peer.cuh"
#include "../../include/movedepth_b200.h"

namespace mvd {

__global__ void __launch_bounds__(256) peer_allreduce_f64_kernel(const double* __restrict__ local, double* __restrict__ out, int n,
                                                                 const PeerArgs pa) {
    peer_allreduce_block(local, out, n, pa);
}

}  // namespace mvd

extern "C" {

long long mvd_peer_allreduce_buffer_bytes(int world, int nmax) {
    if (world <= 0 || nmax <= 0) return 0;
    return static_cast<long long>(mvd::PEER_HEADER) + 2ll * world * nmax * mvd::PEER_ENTRY;
}

int mvd_peer_allreduce_f64(const double* local, double* out, int n, const unsigned long long* peers, int rank, int world, int nmax,
                           void* stream) {
    MVD_REQUIRE(local && out && peers, "null pointer argument");
    MVD_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    MVD_REQUIRE(n > 0 && n <= nmax, "vector length %d exceeds the exchange buffer's %d", n, nmax);
    mvd::peer_allreduce_f64_kernel<<<1, 256, 0, mvd::as_stream(stream)>>>(local, out, n, mvd::PeerArgs{peers, rank, world, nmax});
    return mvd::check_launch("peer_allreduce_f64");
}

}
