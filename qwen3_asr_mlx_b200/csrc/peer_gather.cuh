// Final gather of the data-parallel launcher over NVLink peer memory ("only a final gather of embeddings").
// Every rank writes the embedding rows of ITS utterances straight into EVERY rank's output buffer (P2P-mapped symmetric
// memory, NVSwitch gives each GPU full bandwidth to every peer) at their FINAL row positions, i.e. the transfer and the
// restore of the original utterance order are one kernel: no max-padded staging buffer, no second pass over the data.
// HBM / NVLink-bound byte work: each local row is read once and written n_peers times with 16-byte accesses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qasr {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  void* p[kMaxPeers];
};

// One CTA per local row.  dst_rows[r] = row index of local row r in the gathered (original-order) matrix.
__global__ void __launch_bounds__(128)
scatter_rows_to_peers_kernel(const uint4* __restrict__ local, const long long* __restrict__ dst_rows, int vec_per_row,
                             PeerPtrs peers, int n_peers) {
  const long long r = blockIdx.x;
  const long long d = __ldg(dst_rows + r);
  const uint4* __restrict__ src = local + r * vec_per_row;
  for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x) {
    const uint4 v = __ldg(src + i);
#pragma unroll 4
    for (int p = 0; p < n_peers; ++p) static_cast<uint4*>(peers.p[p])[d * vec_per_row + i] = v;
  }
}

}  // namespace qasr
