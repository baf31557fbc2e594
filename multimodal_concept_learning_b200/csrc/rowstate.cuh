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
// y > tau.  tau is a LOWER BOUND of the row's final k-th best score, from two sources:
//   - the row's own buffer: when fewer than 32 free entries remain the WARP compacts it
//     cooperatively -- a threshold with k..k+kSlack entries at or above it is found by
//     bisection on the order-preserving keys (warp-wide counts), everything below is dropped;
//   - the other CTAs scanning the same query row over other table chunks, which publish
//     their thresholds through one word per row in global memory (atomicMax); any chunk's
//     k-th best is a lower bound of the global k-th best.
// Neither source can drop a member of the true top-k: an element is discarded only when k
// others with a larger value -- or an equal value and a lower table row, because a chunk is
// visited in increasing row order -- are known to exist.  The exact ordering, including the
// lowest-index-wins tie rule, is established by merge.cu.  When bisection cannot separate
// the entries (many exactly equal scores, e.g. duplicated table rows) the compaction falls
// back to an exact radix select that keeps the earliest ties.
#pragma once
#include "common.cuh"
#include "plan.h"

namespace mcl {

constexpr int kSlack = 14;   // a compaction leaves k .. k+kSlack entries
#ifndef MCL_APPEND_GROUP
#define MCL_APPEND_GROUP 8   // columns per append group (4 and 8 measured: profiles/README.md)
#endif

// 2^x on the SFU (MUFU.EX2), one instruction: ~2 ulp, flushes results below 2^-126 to 0,
// which is far inside the rtol of a sum of >= 1 terms of magnitude 1 (the row max).
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct RowState {
  float m;        // running max of y
  float s;        // sum exp2((y - m) * a)
  float sum_y;    // sum of y
  float tau;      // append threshold (strict: append when y > tau)
  float y_label;  // y at the label column (0 if not seen)
  int cnt;        // entries in the candidate buffer
  uint2* buf;     // this row's candidate buffer inside the active slot

  __device__ __forceinline__ void reset(uint2* b) {
    m = -INFINITY; s = 0.f; sum_y = 0.f; tau = -INFINITY; y_label = 0.f; cnt = 0; buf = b;
  }
};

// One chunk of 32 scores for this thread's row; called by all 32 lanes of a warp, converged.  col0 = local table row of y[0];
// n_valid < 32 only in the ragged last chunk (TAIL), whose out-of-range columns are masked.
// CAP: tanh soft-capped logits z' = c*tanh(z/c) (Gemma-2 style heads, modeling_gemma3.py:653-656):
// rc = rs/c, `a` = c*log2(e); max / sum-exp / sum run on t = tanh(y*rc), ranking stays on y
// (tanh is monotone), and row_flush scales by c.
template <bool TAIL, bool CAP>
__device__ __forceinline__ void row_process_chunk(RowState& st, float (&y)[kChunk], int col0,
                                                  int n_valid, float a, int lab_local, float rc) {
  if (TAIL) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i)
      if (i >= n_valid) y[i] = -INFINITY;
  }
  // maxima of the groups of kGroup columns (they also steer the candidate append below)
  constexpr int kGroup = MCL_APPEND_GROUP, kGroups = kChunk / kGroup;
  float gm[kGroups];
#pragma unroll
  for (int g = 0; g < kGroups; ++g) {
    float m4[kGroup / 4];
#pragma unroll
    for (int h = 0; h < kGroup / 4; ++h)
      m4[h] = fmaxf(fmaxf(y[kGroup * g + 4 * h], y[kGroup * g + 4 * h + 1]),
                    fmaxf(y[kGroup * g + 4 * h + 2], y[kGroup * g + 4 * h + 3]));
    gm[g] = m4[0];
#pragma unroll
    for (int h = 1; h < kGroup / 4; ++h) gm[g] = fmaxf(gm[g], m4[h]);
  }
  float cm = gm[0];
#pragma unroll
  for (int g = 1; g < kGroups; ++g) cm = fmaxf(cm, gm[g]);

  // label column (at most once per row and slot)
  if (lab_local >= col0 && lab_local < col0 + kChunk) {
    const int off = lab_local - col0;
#pragma unroll
    for (int i = 0; i < kChunk; ++i)
      if (i == off) st.y_label = y[i];
  }

  // online log-sum-exp in base 2; four partial sums keep the FADD chains short (two resident
  // warps per scheduler cannot hide a 32-deep dependent chain)
  const float m_new = fmaxf(st.m, cm);
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, sy[4] = {0.f, 0.f, 0.f, 0.f};
  if (!CAP) {
    const float corr = ex2_fast((st.m - m_new) * a);  // first chunk: exp2(-inf) = 0, s is 0 anyway
    const float mb = m_new * a;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      acc[i & 3] += ex2_fast(fmaf(y[i], a, -mb));     // masked columns: exp2(-inf) = 0
      if (TAIL) sy[i & 3] += (i < n_valid) ? y[i] : 0.f; else sy[i & 3] += y[i];
    }
    st.s = fmaf(st.s, corr, (acc[0] + acc[1]) + (acc[2] + acc[3]));
  } else {
    const float mt_new = tanhf(m_new * rc);
    const float corr = ex2_fast((tanhf(st.m * rc) - mt_new) * a);   // first chunk: s is 0
    const float mb = mt_new * a;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const float t = tanhf(y[i] * rc);
      const float e = ex2_fast(fmaf(t, a, -mb));
      if (TAIL) { acc[i & 3] += (i < n_valid) ? e : 0.f; sy[i & 3] += (i < n_valid) ? t : 0.f; }
      else { acc[i & 3] += e; sy[i & 3] += t; }
    }
    st.s = fmaf(st.s, corr, (acc[0] + acc[1]) + (acc[2] + acc[3]));
  }
  st.m = m_new;
  st.sum_y += (sy[0] + sy[1]) + (sy[2] + sy[3]);

  // lazy-threshold candidate append.  In steady state a warp's 32 rows hold one or two
  // candidates per chunk between them: only the groups of kGroup columns that contain one run
  // the store sequence (column order is kept: the tie rule relies on it).
  // The group decisions are made warp-wide (one REDUX of the group mask): uniform branches need
  // no reconvergence barriers, which cost more than the predicated-off stores they would skip.
  unsigned gmask = 0u;
#pragma unroll
  for (int g = 0; g < kGroups; ++g) gmask |= (gm[g] > st.tau) ? (1u << g) : 0u;
  gmask = __reduce_or_sync(0xffffffffu, gmask);
  if (gmask) {
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
      if (gmask & (1u << g)) {
#pragma unroll
        for (int i = kGroup * g; i < kGroup * g + kGroup; ++i) {
          if (y[i] > st.tau) {
            st.buf[st.cnt] = make_uint2(__float_as_uint(y[i]), (uint32_t)(col0 + i));
            ++st.cnt;
          }
        }
      }
    }
  }
}

// k = 1 specialisation (every training / evaluation call site of the reference is an argmax:
// multimodal_training.py:276, vision_training.py:132): the row's running maximum IS its top-1,
// so no candidate buffer, no threshold, no compaction -- st.m holds the best y and st.cnt the
// table row it came from.  A chunk updates the argmax only when its maximum is strictly larger
// (columns are visited in increasing order inside a slot: the earliest maximum wins the tie).
template <bool TAIL, bool CAP>
__device__ __forceinline__ void row_process_chunk_top1(RowState& st, float (&y)[kChunk], int col0,
                                                       int n_valid, float a, int lab_local, float rc) {
  if (TAIL) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i)
      if (i >= n_valid) y[i] = -INFINITY;
  }
  float m8[kChunk / 4];
#pragma unroll
  for (int h = 0; h < kChunk / 4; ++h)
    m8[h] = fmaxf(fmaxf(y[4 * h], y[4 * h + 1]), fmaxf(y[4 * h + 2], y[4 * h + 3]));
  const float cm = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                         fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
  if (lab_local >= col0 && lab_local < col0 + kChunk) {
    const int off = lab_local - col0;
#pragma unroll
    for (int i = 0; i < kChunk; ++i)
      if (i == off) st.y_label = y[i];
  }
  if (cm > st.m) {                 // rare after the first tiles: ~ln(columns) times per row
    int bi = 0;
#pragma unroll
    for (int i = kChunk - 1; i >= 0; --i)
      if (y[i] == cm) bi = i;
    st.cnt = col0 + bi;
  }
  const float m_new = fmaxf(st.m, cm);
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, sy[4] = {0.f, 0.f, 0.f, 0.f};
  if (!CAP) {
    const float corr = ex2_fast((st.m - m_new) * a);
    const float mb = m_new * a;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      acc[i & 3] += ex2_fast(fmaf(y[i], a, -mb));
      if (TAIL) sy[i & 3] += (i < n_valid) ? y[i] : 0.f; else sy[i & 3] += y[i];
    }
    st.s = fmaf(st.s, corr, (acc[0] + acc[1]) + (acc[2] + acc[3]));
  } else {
    const float mt_new = tanhf(m_new * rc);
    const float corr = ex2_fast((tanhf(st.m * rc) - mt_new) * a);
    const float mb = mt_new * a;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const float t = tanhf(y[i] * rc);
      const float e = ex2_fast(fmaf(t, a, -mb));
      if (TAIL) { acc[i & 3] += (i < n_valid) ? e : 0.f; sy[i & 3] += (i < n_valid) ? t : 0.f; }
      else { acc[i & 3] += e; sy[i & 3] += t; }
    }
    st.s = fmaf(st.s, corr, (acc[0] + acc[1]) + (acc[2] + acc[3]));
  }
  st.m = m_new;
  st.sum_y += (sy[0] + sy[1]) + (sy[2] + sy[3]);
}

// Close a k = 1 slot: the argmax is the slot's single candidate, its value the slot's threshold.
__device__ __forceinline__ void row_flush_top1(RowState& st) {
  if (st.m > -INFINITY) {
    st.buf[0] = make_uint2(__float_as_uint(st.m), (uint32_t)st.cnt);
    st.cnt = 1;
  } else {
    st.cnt = 0;
  }
  st.tau = st.m;
}

__device__ __forceinline__ int warp_count_ge(const uint32_t (&key)[kCandCap / 32], uint32_t x) {
  int c = 0;
#pragma unroll
  for (int i = 0; i < kCandCap / 32; ++i) c += (key[i] >= x) ? 1 : 0;
  return __reduce_add_sync(0xffffffffu, c);
}

// Warp-cooperative compaction of every row of this warp whose buffer could overflow on
// the next chunk.  warp_buf = buffer of the warp's lane-0 row; rows are kCandCap apart.
// tau_pub = this lane's word of the shared threshold array (nullable).  Must be called by
// all 32 lanes (converged).
//
// Joint threshold (joint_warp != nullptr): a slot sees only 1/n of a row's columns, so its own
// k-th best -- and the maximum of the slots' k-th bests, which tau_pub carries -- passes ~n times
// more candidates than the row's true k-th best would.  Each of the first ng2 slot halves of a
// wave therefore also publishes x_i, a key that at least m = ceil(k / ng2) of ITS entries reach
// (the m-th largest of the 32 lane maxima of the buffer: m max-reductions, no extra loads).
// Once all ng2 words of a row are set, min_i x_i is reached by >= ng2 * m >= k scores of the row
// at distinct table rows, i.e. it is a lower bound of the row's k-th best -- about as tight as
// the k-th best of ALL columns seen so far -- and every slot of the row filters with it
// (scan_tc.cu).  joint_warp = the word of the warp's lane-0 row; rows are kJointWords apart.
// Measured on B200 (profiles/README.md): ~3x fewer appends, but the per-tile reads and the stale
// bound between compactions cost as much as they save except on C2; library option 12 turns it on.
__device__ __forceinline__ void warp_compact_rows(RowState& st, int k, uint2* warp_buf, int lane,
                                                  uint32_t* tau_pub, uint32_t* joint_warp = nullptr,
                                                  int joint_m = 0) {
  unsigned need = __ballot_sync(0xffffffffu, st.cnt + kChunk > kCandCap);
  if (need == 0) return;
  const unsigned lt = (1u << lane) - 1u;
  __syncwarp();  // orders the owners' appends before the warp's reads (no fence needed)
  // The reload of a row's entries is an L2 round trip (~2/3 of a compaction).  At the start of
  // a table chunk all 32 rows of a warp fill up together, so the entries of the next TWO rows
  // are requested before the current row is processed.
  struct Pending { uint2 raw[kCandCap / 32]; int r, n; uint32_t tau; };
  auto fetch = [&](Pending& pd) {                      // pop a row and issue its loads (no use yet)
    pd.r = -1; pd.n = 0; pd.tau = 0u;
#pragma unroll
    for (int i = 0; i < kCandCap / 32; ++i) pd.raw[i] = make_uint2(0u, 0u);
    if (need) {
      pd.r = __ffs(need) - 1;
      need &= need - 1;
      pd.n = __shfl_sync(0xffffffffu, st.cnt, pd.r);
      pd.tau = f2key(__shfl_sync(0xffffffffu, st.tau, pd.r));
      const uint2* bp = warp_buf + (size_t)pd.r * kCandCap;
#pragma unroll
      for (int i = 0; i < kCandCap / 32; ++i)
        if (lane + 32 * i < pd.n) pd.raw[i] = __ldcg(bp + lane + 32 * i);
    }
  };
  Pending nx1, nx2;
  fetch(nx1);
  fetch(nx2);
  while (nx1.r >= 0) {
    const int r = nx1.r;
    const int n = nx1.n;
    const uint32_t tau_key = nx1.tau;
    uint2* b = warp_buf + (size_t)r * kCandCap;
    uint32_t key[kCandCap / 32], idx[kCandCap / 32];
#pragma unroll
    for (int i = 0; i < kCandCap / 32; ++i) {
      const int j = lane + 32 * i;
      key[i] = (j < n) ? f2key(__uint_as_float(nx1.raw[i].x)) : 0u;   // 0 is below every real score
      idx[i] = nx1.raw[i].y;
    }
    nx1 = nx2;
    fetch(nx2);
    // ---- bisection: x with k <= #{key >= x} <= k + kSlack ------------------------------
    uint32_t mx = key[0];
#pragma unroll
    for (int i = 1; i < kCandCap / 32; ++i) mx = max(mx, key[i]);
    if (joint_warp) {                                // m-th largest lane maximum of the buffer
      uint32_t v = mx, xj = 0u;
      for (int t = 0; t < joint_m; ++t) {
        xj = __reduce_max_sync(0xffffffffu, v);
        const unsigned who = __ballot_sync(0xffffffffu, v == xj);
        if (lane == __ffs(who) - 1) v = 0u;          // retire ONE lane holding it
      }
      // (lane maxima move when a buffer is repacked: keep the best bound ever established)
      if (lane == 0 && xj > 1u) atomicMax(joint_warp + (size_t)r * kJointWords, xj);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    // entries AT the threshold count: an entry equal to the row's own tau was appended before tau
    // rose to its value (earlier table row: it wins the tie) and may be one of the k best
    uint32_t lo = tau_key;
    uint32_t hi = (mx == 0xffffffffu) ? mx : mx + 1u;  // #{key >= hi} = 0 < k
    uint32_t x = lo;
    int cx = warp_count_ge(key, lo);
    // (a threshold raised by another CTA can leave fewer than k live entries: keep just those)
    bool found = (cx <= k + kSlack);
#pragma unroll 1
    for (int it = 0; it < 34 && !found && hi - lo > 1u; ++it) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      const int c = warp_count_ge(key, mid);
      if (c < k) hi = mid;
      else if (c > k + kSlack) lo = mid;
      else { x = mid; cx = c; found = true; }
    }
    uint32_t cls_lo, cls_hi;
    int kr;
    if (found) {                     // keep everything at or above x
      cls_lo = x; cls_hi = 0xffffffffu; kr = cx;
    } else {
      // ---- exact radix select (ties): key of the k-th largest, kr = ties of it to keep ---
      uint32_t prefix = 0u;
      kr = k;
#pragma unroll 1
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t want = (prefix >> bit) | 1u;
        int c = 0;
#pragma unroll
        for (int i = 0; i < kCandCap / 32; ++i) c += ((key[i] >> bit) == want) ? 1 : 0;
        const int tot = __reduce_add_sync(0xffffffffu, c);
        if (tot >= kr) prefix |= (1u << bit); else kr -= tot;
      }
      cls_lo = prefix; cls_hi = prefix;
    }
    // entries above the class are kept; inside the class the first kr in buffer order
    __syncwarp();  // every lane holds its entries in registers before any overwrite
    int base = 0, ties_seen = 0;
#pragma unroll
    for (int i = 0; i < kCandCap / 32; ++i) {
      const bool gt = key[i] > cls_hi;
      const bool eq = key[i] >= cls_lo && key[i] <= cls_hi;
      const unsigned eqm = __ballot_sync(0xffffffffu, eq);
      const bool keep = gt || (eq && (ties_seen + __popc(eqm & lt)) < kr);
      ties_seen += __popc(eqm);
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) b[base + __popc(km & lt)] = make_uint2(__float_as_uint(key2f(key[i])), idx[i]);
      base += __popc(km);
    }
    if (lane == r) {
      st.cnt = base;
      if (base >= k) {                           // k entries at or above cls_lo are known
        st.tau = fmaxf(st.tau, key2f(cls_lo));
        if (tau_pub) atomicMax(tau_pub, cls_lo); // a lower bound of this row's final k-th best
      }
    }
  }
  __syncwarp();
}

// Fold in the threshold other CTAs published for this row: keep y >= shared, i.e. y > the
// float just below it (an equal score in another chunk may still lose the index tie-break).
__device__ __forceinline__ void row_apply_shared_tau(RowState& st, uint32_t shared_key) {
  if (shared_key > 1u) st.tau = fmaxf(st.tau, key2f(shared_key - 1u));
}

// Close a slot: candidate count and (m, s, sum_z, z_label) in z space for this row.
// softcap c > 0: z' = c*tanh(z/c); sum_y already holds the sum of tanh values.
// The row's final threshold travels with the count: it is a lower bound of the row's global
// k-th best score, which lets the merge drop most candidates of the OTHER slots unsorted.
__device__ __forceinline__ void row_flush(const RowState& st, float rs, float softcap, int2* cnt_out,
                                          float4* stats_out) {
  *cnt_out = make_int2(st.cnt, (int)f2key(st.tau));
  if (softcap > 0.f) {
    const float rc = rs / softcap;
    *stats_out = make_float4(softcap * tanhf(st.m * rc), st.s, softcap * st.sum_y,
                             softcap * tanhf(st.y_label * rc));
  } else {
    *stats_out = make_float4(st.m * rs, st.s, st.sum_y * rs, st.y_label * rs);
  }
}

}  // namespace mcl
