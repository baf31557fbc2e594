// Result exchange of the vocabulary-sharded scan over peer memory (NVLink / NVSwitch) instead of
// NCCL: after the local scan every rank owns, for all Q query rows, its shard's (top-k values,
// indices, statistics).  Rank j merges query rows [j*Q/N, (j+1)*Q/N):
//   push 1   ONE kernel stores, for every peer p, p's row range of the local record straight
//            into p's receive area (16-byte stores over NVLink) and, when its last CTA is done,
//            adds 1 to p's arrival counter (system-scope fence + atomic);
//   wait     the consumer's STREAM waits for its counter with a stream memory operation
//            (cuStreamWaitValue32, GEQ on a monotone counter): no SM spins, nothing to deadlock;
//   merge    merge_ranks_kernel over the N received lists (merge.cu);
//   push 2   the same kernel stores the merged rows into every rank's result area; wait; one
//            device copy hands the result to the caller's arrays.
// Per step: 2 small kernels + 2 stream waits where the NCCL path runs a grouped send/recv and a
// grouped all-gather (~170 us at N = 8 for 5 MB of records: launch and protocol latency, not bytes).
// The blocks are plain cudaMalloc memory opened with CUDA IPC (mcl_peer_alloc / mcl_peer_open).
//
// Reuse of the areas needs no extra handshake: a rank enters step e+1 only after it has seen
// every peer's merged rows of step e, and a peer sends those only after its own merge of step e has
// read its receive area; a peer's merged rows of step e+1 need this rank's push of step e+1, which
// follows this rank's copy-out of step e in stream order.  When push 2 is skipped (the caller takes
// only the rows it merged) nothing after the merge holds a fast rank back, but its push 1 of step
// e+2 still needs every peer's push 1 of step e+1, which follows that peer's merge of step e: a
// rank runs at most ONE step ahead of any peer's wait.  The receive areas AND the arrival counters
// of push 1 therefore alternate with the step's parity (api.cu): counter [e & 1] holds arrivals of
// steps e, e-2, ... only, so reaching its target means that all pieces of step e have landed.
#include "kernels.h"

namespace mcl {

// PushParams (kernels.h): seg[i] = one contiguous piece, src -> offset dst_off in the block of peer
// seg_peer[i] (sizes and offsets are multiples of 16); peer[] = block bases, own block included;
// counter_off = the arrival counter bumped in every OTHER rank's block; done = local word that
// counts finished CTAs and returns to 0.
__global__ void __launch_bounds__(256)
p2p_push_kernel(const __grid_constant__ PushParams p) {
  for (int s = 0; s < p.nseg; ++s) {
    const uint4* src = reinterpret_cast<const uint4*>(p.seg[s].src);
    uint4* dst = reinterpret_cast<uint4*>(p.peer[p.seg_peer[s]] + p.seg[s].dst_off);
    const size_t n = p.seg[s].bytes >> 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      dst[i] = __ldcg(src + i);
  }
  __threadfence_system();                // this thread's stores are visible system-wide ...
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(p.done, 1u) == gridDim.x - 1) {
      __threadfence_system();            // ... and so are those of every CTA counted before this one
      *p.done = 0u;
      for (int r = 0; r < p.npeer; ++r)
        if (r != p.self)
          atomicAdd_system(reinterpret_cast<unsigned*>(p.peer[r] + p.counter_off), 1u);
    }
  }
}

cudaError_t launch_p2p_push(const PushParams& p, int sm_count, cudaStream_t s) {
  size_t bytes = 0;
  for (int i = 0; i < p.nseg; ++i) bytes += p.seg[i].bytes;
  int grid = (int)((bytes / 16 + 256 * 8 - 1) / (256 * 8));       // ~8 vectors per thread
  grid = grid < 1 ? 1 : (grid > sm_count ? sm_count : grid);
  p2p_push_kernel<<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace mcl
