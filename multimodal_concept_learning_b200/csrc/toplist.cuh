// Warp-level sorted top-64 list shared by the merge and selection kernels.  Keys are 64 bit:
// high word = order-preserving float key, low word = ~row index, so a plain descending sort
// realises "value desc, index asc" -- the documented lowest-index-wins tie rule.
#pragma once
#include "common.cuh"

namespace mcl {

__device__ __forceinline__ uint64_t pack_key(float v, uint32_t idx) {
  return ((uint64_t)f2key(v) << 32) | (uint64_t)(uint32_t)(~idx);
}
__device__ __forceinline__ uint64_t shfl_xor64(uint64_t v, int j) {
  return __shfl_xor_sync(0xffffffffu, v, j);
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ uint64_t max64(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t min64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// compare-exchange stage j (< 32) for element position p; desc = block sorted descending
__device__ __forceinline__ uint64_t cx(uint64_t mine, int j, bool lower, bool desc) {
  const uint64_t other = shfl_xor64(mine, j);
  return (lower == desc) ? max64(mine, other) : min64(mine, other);
}

// Element p = i*32 + lane lives in register k[i] of `lane`.
// Sorts all 64 descending starting from bitonic-network size `first_size`
// (2 = full sort, 64 = bitonic merge of an already bitonic sequence).
__device__ __forceinline__ void bitonic64_desc(uint64_t& k0, uint64_t& k1, int lane,
                                               int first_size) {
  for (int size = first_size; size <= 64; size <<= 1) {
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j == 32) {  // partner is the other register of this lane; size == 64: descending
        const uint64_t hi = max64(k0, k1), lo = min64(k0, k1);
        k0 = hi; k1 = lo;
      } else {
        const bool lower = (lane & j) == 0;
        const bool desc0 = (size == 64) ? true : (size == 32 ? true : ((lane & size) == 0));
        const bool desc1 = (size == 64) ? true : (size == 32 ? false : ((lane & size) == 0));
        k0 = cx(k0, j, lower, desc0);
        k1 = cx(k1, j, lower, desc1);
      }
    }
  }
}

struct TopList {
  uint64_t r0, r1;  // running top-64, descending over p = i*32 + lane
  __device__ __forceinline__ void init() { r0 = 0ull; r1 = 0ull; }
  // merge a batch of 64 unsorted keys (b0 = positions 0..31, b1 = 32..63)
  // same, for a batch that is ALREADY sorted descending (another list, a rank's top-k):
  // reverse + element-wise max + one 6-stage bitonic merge instead of a 21-stage sort
  __device__ __forceinline__ void push_sorted(uint64_t b0, uint64_t b1, int lane) {
    const uint64_t rb0 = shfl64(b1, 31 - lane);
    const uint64_t rb1 = shfl64(b0, 31 - lane);
    r0 = max64(r0, rb0);
    r1 = max64(r1, rb1);
    bitonic64_desc(r0, r1, lane, 64);
  }
  __device__ __forceinline__ void push(uint64_t b0, uint64_t b1, int lane) {
    bitonic64_desc(b0, b1, lane, 2);
    // reversed batch: position p takes batch element 63 - p
    const uint64_t rb0 = shfl64(b1, 31 - lane);
    const uint64_t rb1 = shfl64(b0, 31 - lane);
    r0 = max64(r0, rb0);
    r1 = max64(r1, rb1);
    bitonic64_desc(r0, r1, lane, 64);
  }
};

}  // namespace mcl
