// Per-query-row streaming state of the scan epilogue: online log-sum-exp, running sum,
// label pick-up and the lazy-threshold top-k candidate buffer.
//
// One THREAD owns one query row (a TMEM lane in the tcgen05 kernel).  Scores arrive in
// chunks of 32 consecutive table rows, already multiplied by the per-table-row scale;
// they are kept in "y space": z = y * rs with rs = inv_norm_q * scale > 0 a per-row
// constant, so ordering, max and top-k do not depend on rs and it is folded into the
// exp2 argument (a = rs * log2 e) and into the final outputs only.
//
// Top-k: a candidate is appended to the row's buffer (global memory, L2 resident) when
// y > tau, tau being the k-th best value at the last compaction.  When fewer than 32 free
// entries remain the WARP compacts that row cooperatively: exact k-th value by radix
// select on order-preserving keys, stable keep of everything above it plus the earliest
// ties.  Because table rows are visited in increasing order inside a slot, "earliest"
// is "lowest index", which is the documented tie rule.
#pragma once
#include "common.cuh"

namespace mcl {

struct RowState {
  float m;        // running max of y
  float s;        // sum exp2((y - m) * a)
  float sum_y;    // sum of y
  float tau;      // append threshold
  float y_label;  // y at the label column (0 if not seen)
  int cnt;        // entries in the candidate buffer
  uint2* buf;     // this row's candidate buffer inside the active slot

  __device__ __forceinline__ void reset(uint2* b) {
    m = -INFINITY; s = 0.f; sum_y = 0.f; tau = -INFINITY; y_label = 0.f; cnt = 0; buf = b;
  }
};

// One chunk of 32 scores for this thread's row.  col0 = local table row of y[0];
// n_valid < 32 only in the ragged last chunk (TAIL), whose out-of-range columns are masked.
template <bool TAIL>
__device__ __forceinline__ void row_process_chunk(RowState& st, float (&y)[kChunk], int col0,
                                                  int n_valid, float a, int lab_local) {
  if (TAIL) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i)
      if (i >= n_valid) y[i] = -INFINITY;
  }
  float cm = y[0];
#pragma unroll
  for (int i = 1; i < kChunk; ++i) cm = fmaxf(cm, y[i]);

  // label column (at most once per row and slot)
  if (lab_local >= col0 && lab_local < col0 + kChunk) {
    const int off = lab_local - col0;
#pragma unroll
    for (int i = 0; i < kChunk; ++i)
      if (i == off) st.y_label = y[i];
  }

  // online log-sum-exp in base 2
  const float m_new = fmaxf(st.m, cm);
  const float corr = exp2f((st.m - m_new) * a);  // first chunk: exp2(-inf) = 0, s is 0 anyway
  const float mb = m_new * a;
  float acc = 0.f, sy = 0.f;
#pragma unroll
  for (int i = 0; i < kChunk; ++i) {
    acc += exp2f(fmaf(y[i], a, -mb));            // masked columns: exp2(-inf) = 0
    if (TAIL) sy += (i < n_valid) ? y[i] : 0.f; else sy += y[i];
  }
  st.s = fmaf(st.s, corr, acc);
  st.m = m_new;
  st.sum_y += sy;

  // lazy-threshold candidate append
  if (cm > st.tau) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      if (y[i] > st.tau) {
        st.buf[st.cnt] = make_uint2(__float_as_uint(y[i]), (uint32_t)(col0 + i));
        ++st.cnt;
      }
    }
  }
}

// Warp-cooperative compaction of every row of this warp whose buffer could overflow on
// the next chunk.  warp_buf = buffer of the warp's lane-0 row; rows are kCandCap apart.
// Must be called by all 32 lanes (converged).
__device__ __forceinline__ void warp_compact_rows(RowState& st, int k, uint2* warp_buf, int lane) {
  unsigned need = __ballot_sync(0xffffffffu, st.cnt + kChunk > kCandCap);
  const unsigned lt = (1u << lane) - 1u;
  while (need) {
    const int r = __ffs(need) - 1;
    need &= need - 1;
    const int n = __shfl_sync(0xffffffffu, st.cnt, r);
    uint2* b = warp_buf + (size_t)r * kCandCap;
    __threadfence_block();
    __syncwarp();  // lane r's appends are visible to the warp
    uint32_t key[kCandCap / 32], idx[kCandCap / 32];
#pragma unroll
    for (int i = 0; i < kCandCap / 32; ++i) {
      const int j = lane + 32 * i;
      key[i] = 0u; idx[i] = 0u;                       // key 0 is below every real score
      if (j < n) {
        const uint2 e = __ldcg(b + j);
        key[i] = f2key(__uint_as_float(e.x));
        idx[i] = e.y;
      }
    }
    // radix select: key of the k-th largest entry; kr = how many of its ties to keep
    uint32_t prefix = 0u;
    int kr = k;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t want = (prefix >> bit) | 1u;
      int c = 0;
#pragma unroll
      for (int i = 0; i < kCandCap / 32; ++i) c += ((key[i] >> bit) == want) ? 1 : 0;
      const int tot = __reduce_add_sync(0xffffffffu, c);
      if (tot >= kr) prefix |= (1u << bit); else kr -= tot;
    }
    __syncwarp();  // every lane holds its entries in registers before any overwrite
    int base = 0, ties_seen = 0;
#pragma unroll
    for (int i = 0; i < kCandCap / 32; ++i) {
      const bool gt = key[i] > prefix;
      const bool eq = key[i] == prefix;
      const unsigned eqm = __ballot_sync(0xffffffffu, eq);
      const bool keep = gt || (eq && (ties_seen + __popc(eqm & lt)) < kr);
      ties_seen += __popc(eqm);
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) b[base + __popc(km & lt)] = make_uint2(__float_as_uint(key2f(key[i])), idx[i]);
      base += __popc(km);
    }
    __syncwarp();
    if (lane == r) { st.cnt = base; st.tau = key2f(prefix); }
  }
}

// Close a slot: candidate count and (m, s, sum_z, z_label) in z space for this row.
__device__ __forceinline__ void row_flush(const RowState& st, float rs, int* cnt_out,
                                          float4* stats_out) {
  *cnt_out = st.cnt;
  *stats_out = make_float4(st.m * rs, st.s, st.sum_y * rs, st.y_label * rs);
}

}  // namespace mcl
