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
    return static_cast<long long>(mvd::PEER_HEADER) + 2ll * world * nmax * static_cast<long long>(sizeof(double));
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
