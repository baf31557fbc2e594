// Merge kernels: (1) the partial-result slots of one scan launch -> final [Q,k] / [Q,4];
// (2) R dense per-shard results -> final (the step after the NCCL all-gather).
// One warp per query row; a running sorted list of 64 (key) entries lives in registers,
// two per lane, and is updated by warp-shuffle bitonic networks.  Keys are 64 bit:
// high word = order-preserving float key, low word = ~row index, so a plain descending
// sort realises "value desc, index asc" -- the documented lowest-index-wins tie rule.
#include <atomic>
#include "kernels.h"
#include "toplist.cuh"

namespace mcl {

std::atomic<int> g_merge_variant{0};   // library option 19: 1 = the streaming-fold merge kernels only (A/B)

// kWPR warps per query row (1 for the usual handful of slots, 16 when a small Q was split over
// all SMs).  Phase A: the largest threshold any slot recorded for the row is a lower bound of
// its global k-th best, so only candidates at or above it can matter -- typically ~2k of the
// up to 128 per slot -- and everything else is dropped without being sorted.  Phase B: each
// warp folds its slots' survivors, 64 at a time through a shared-memory queue, into a running
// sorted top-64.  Phase C (kWPR > 1): warp 0 folds the other warps' lists and writes the row.
constexpr int kPendCap = 96;

template <int kWPR>
__global__ void __launch_bounds__(128 > 32 * kWPR ? 128 : 32 * kWPR)
merge_slots_kernel(SlotView sv, const SlotMap map, int Q, int k, const float* __restrict__ inv_q,
                   float scale, float softcap, long long index_base, float* __restrict__ topk_val,
                   long long* __restrict__ topk_idx, float4* __restrict__ row_stats) {
  constexpr int kWarps = (128 > 32 * kWPR ? 128 : 32 * kWPR) / 32;
  constexpr int kRows = kWarps / kWPR;                  // rows per CTA
  __shared__ unsigned long long pend[kWarps][kPendCap];
  __shared__ unsigned long long lists[kWarps][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp % kWPR;                           // this warp's rank within its row
  const int row = blockIdx.x * kRows + warp / kWPR;
  const bool live = row < Q;
  const int rb = live ? row / kBlockM : 0, r_in = live ? row % kBlockM : 0;
  const int slot0 = rb * map.stride;
  const int nsplit = live ? slotmap_count(map, rb) : 0;   // slots this row block really owns
  const unsigned lt = (1u << lane) - 1u;

  // ---- phase A: bound, and (warp 0 of the row) the merged statistics ---------------------
  // Rows with at most 32 slots (every multi-row-block plan): lane i holds slot i's count, threshold
  // and statistics -- ONE round trip -- and phase B below fetches the candidate lists of two slots
  // at a time with all their loads in flight.  (A warp used to walk count -> 32 entries -> next 32
  // entries ... slot after slot: ~18 dependent L2 round trips per row, which made this kernel
  // 100 us at 8192 rows x 10 slots -- 12 % of an eight-way shard's step.)
  const bool few = nsplit <= 32;
  int2 c_l = make_int2(0, 0);
  float4 st_l = make_float4(-INFINITY, 0.f, 0.f, 0.f);
  if (live && few && lane < nsplit) {
    c_l = __ldcg(&sv.cnt[(size_t)(slot0 + lane) * kBlockM + r_in]);
    if (wr == 0) st_l = __ldcg(&sv.stats[(size_t)(slot0 + lane) * kBlockM + r_in]);
  }
  uint32_t bound = (uint32_t)c_l.y;
  if (live && !few)
    for (int i = lane; i < nsplit; i += 32)
      bound = max(bound, (uint32_t)sv.cnt[(size_t)(slot0 + i) * kBlockM + r_in].y);
  bound = __reduce_max_sync(0xffffffffu, bound);
  if (live && wr == 0) {
    float m = st_l.x;
    if (!few)
      for (int i = lane; i < nsplit; i += 32)
        m = fmaxf(m, sv.stats[(size_t)(slot0 + i) * kBlockM + r_in].x);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f, sum_z = 0.f, z_label = 0.f;
    if (few) {
      s = (st_l.y > 0.f) ? st_l.y * expf(st_l.x - m) : 0.f;
      sum_z = st_l.z;
      z_label = st_l.w;
    } else {
      for (int i = lane; i < nsplit; i += 32) {
        const float4 st = sv.stats[(size_t)(slot0 + i) * kBlockM + r_in];
        s += (st.y > 0.f) ? st.y * expf(st.x - m) : 0.f;
        sum_z += st.z;
        z_label += st.w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      sum_z += __shfl_xor_sync(0xffffffffu, sum_z, o);
      z_label += __shfl_xor_sync(0xffffffffu, z_label, o);
    }
    if (lane == 0) row_stats[row] = make_float4(m, s, sum_z, z_label);
  }

  // ---- phase B ----------------------------------------------------------------------------
  const float rs = live ? (inv_q ? inv_q[row] : 1.f) * scale : 1.f;
  TopList top; top.init();
  unsigned long long kth = 0ull;          // key of the warp's current k-th best (0 = none yet)
  int npend = 0;
  unsigned long long* q = pend[warp];
  auto flush64 = [&]() {                  // fold the first 64 queued keys, keep the rest
    const unsigned long long b0 = q[lane], b1 = q[32 + lane];
    const unsigned long long rest = (64 + lane < npend) ? q[64 + lane] : 0ull;
    __syncwarp();
    top.push(b0, b1, lane);
    npend -= 64;
    if (lane < npend) q[lane] = rest;
    __syncwarp();
    const int p = k - 1;
    kth = shfl64(p < 32 ? top.r0 : top.r1, p & 31);
  };
  auto consume = [&](const uint2 e, bool valid) {        // (all 32 lanes call)
    unsigned long long key = 0ull;
    if (valid && f2key(__uint_as_float(e.x)) >= bound) {  // can still be among the row's k best
      // rank by the OUTPUT value z (what callers and the rank merge see), so that scores
      // whose z round to the same float tie-break by table row everywhere
      float z = __uint_as_float(e.x) * rs;
      if (softcap > 0.f) z = softcap * tanhf(z / softcap);
      key = pack_key(z, e.y);
    }
    const bool keep = key > kth;          // keys are unique, so > loses nothing
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    if (keep) q[npend + __popc(km & lt)] = key;
    npend += __popc(km);
    __syncwarp();
    if (npend >= 64) flush64();
  };
  if (live && few) {
    constexpr int kT = kCandCap / 32;     // loads per lane and slot
    for (int i = wr; i < nsplit; i += 2 * kWPR) {         // (warp-uniform)
      const int i1 = i + kWPR;
      const int n0 = __shfl_sync(0xffffffffu, c_l.x, i);
      const int n1 = __shfl_sync(0xffffffffu, c_l.x, i1 & 31);
      const bool two = i1 < nsplit;
      const uint2* b0 = sv.cand + ((size_t)(slot0 + i) * kBlockM + r_in) * kCandCap;
      const uint2* b1 = sv.cand + ((size_t)(slot0 + (two ? i1 : i)) * kBlockM + r_in) * kCandCap;
      uint2 e0[kT], e1[kT];
#pragma unroll
      for (int t = 0; t < kT; ++t) {
        const int j = 32 * t + lane;
        e0[t] = (j < n0) ? __ldcg(b0 + j) : make_uint2(0u, 0u);
        e1[t] = (two && j < n1) ? __ldcg(b1 + j) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int t = 0; t < kT; ++t)
        if (32 * t < n0) consume(e0[t], 32 * t + lane < n0);
      if (two) {
#pragma unroll
        for (int t = 0; t < kT; ++t)
          if (32 * t < n1) consume(e1[t], 32 * t + lane < n1);
      }
    }
  } else if (live) {
    for (int i = wr; i < nsplit; i += kWPR) {
      const int slot = slot0 + i;
      const int n = sv.cnt[(size_t)slot * kBlockM + r_in].x;
      const uint2* b = sv.cand + ((size_t)slot * kBlockM + r_in) * kCandCap;
      for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        consume(j < n ? b[j] : make_uint2(0u, 0u), j < n);
      }
    }
  }
  if (live && npend > 0) {                // tail: pad the queue to 64 with empty keys
    if (lane + npend < 64) q[npend + lane] = 0ull;
    if (lane + npend + 32 < 64) q[npend + 32 + lane] = 0ull;
    __syncwarp();
    npend = 64;
    flush64();
  }
  // ---- phase C ----------------------------------------------------------------------------
  if (kWPR > 1) {
    lists[warp][lane] = top.r0;
    lists[warp][32 + lane] = top.r1;
    __syncthreads();
    if (wr != 0) return;
    for (int w = 1; w < kWPR; ++w) top.push_sorted(lists[warp + w][lane], lists[warp + w][32 + lane], lane);
  }
  if (!live) return;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = i * 32 + lane;
    const uint64_t key = i ? top.r1 : top.r0;
    if (p < k) {
      const bool empty = (key == 0ull);
      topk_val[(size_t)row * k + p] = empty ? -INFINITY : key2f((uint32_t)(key >> 32));
      topk_idx[(size_t)row * k + p] =
          empty ? -1ll : index_base + (long long)(uint32_t)(~(uint32_t)key);
    }
  }
}

// One warp per row, selection before sorting (rows with at most 32 slots, launches with enough
// rows to fill the chip).  ncu on the kernel above at 8192 rows x 10 slots
// (profiles/r02_merge_slots_c3_8_ncu_raw.csv): 61 % issue utilisation, 7 k warp instructions per
// row, 70 % of them ISETP / SEL / SHFL of the bitonic networks -- the slots' own thresholds pass
// ~4k candidates per row, which cost four or five 64-key sorts.  Here:
//   1. lane i reads slot i's count, threshold and statistics (one round trip); the candidate lists
//      are fetched two slots at a time with all loads in flight; entries at or above the largest
//      slot threshold (a lower bound of the row's k-th best) become 64-bit keys in shared memory;
//   2. with more than 64 of them a pivot p with k <= #{key >= p} <= 64 is searched with warp-wide
//      counts (16 keys per lane in registers, ~50 instructions per count): the first pivots
//      interpolate on the logarithm of the counts -- score tails are close to exponential --, then
//      plain bisection of the 64-bit key interval, which always ends because keys are unique;
//   3. ONE 64-key sort of the keys >= p.
// More than 512 keys above the bound (adversarial thresholds): the streaming fold of the kernel
// above.  Same keys, same order: bit-identical outputs (tests/test_gpu_parity.py compare the
// kernels through library option 19).
constexpr int kSurvCap = 512;

__global__ void __launch_bounds__(128)
merge_rows_kernel(SlotView sv, const SlotMap map, int Q, int k, const float* __restrict__ inv_q,
                  float scale, float softcap, long long index_base, float* __restrict__ topk_val,
                  long long* __restrict__ topk_idx, float4* __restrict__ row_stats) {
  __shared__ unsigned long long surv[4][kSurvCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 4 + warp;
  if (row >= Q) return;                                   // (no block-wide barrier below)
  const int rb = row / kBlockM, r_in = row % kBlockM;
  const int slot0 = rb * map.stride;
  const int nsplit = slotmap_count(map, rb);              // <= 32 (the launcher checks the stride)
  const unsigned lt = (1u << lane) - 1u;
  unsigned long long* sq = surv[warp];

  int2 c_l = make_int2(0, 0);
  float4 st_l = make_float4(-INFINITY, 0.f, 0.f, 0.f);
  if (lane < nsplit) {
    c_l = __ldcg(&sv.cnt[(size_t)(slot0 + lane) * kBlockM + r_in]);
    st_l = __ldcg(&sv.stats[(size_t)(slot0 + lane) * kBlockM + r_in]);
  }
  const float rs = (inv_q ? inv_q[row] : 1.f) * scale;
  const uint32_t bound = __reduce_max_sync(0xffffffffu, (uint32_t)c_l.y);
  {
    float m = st_l.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = (st_l.y > 0.f) ? st_l.y * expf(st_l.x - m) : 0.f, sum_z = st_l.z, z_label = st_l.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      sum_z += __shfl_xor_sync(0xffffffffu, sum_z, o);
      z_label += __shfl_xor_sync(0xffffffffu, z_label, o);
    }
    if (lane == 0) row_stats[row] = make_float4(m, s, sum_z, z_label);
  }

  // ---- 1. candidates at or above the bound -> keys in shared memory --------------------------
  int ns = 0;                                             // (warp-uniform)
  auto keep_key = [&](const uint2 e, bool valid) {
    unsigned long long key = 0ull;
    if (valid && f2key(__uint_as_float(e.x)) >= bound) {
      float z = __uint_as_float(e.x) * rs;                // rank by the OUTPUT value (see above)
      if (softcap > 0.f) z = softcap * tanhf(z / softcap);
      key = pack_key(z, e.y);
    }
    const unsigned km = __ballot_sync(0xffffffffu, key != 0ull);
    const int pos = ns + __popc(km & lt);
    if (key != 0ull && pos < kSurvCap) sq[pos] = key;
    ns += __popc(km);
  };
  constexpr int kT = kCandCap / 32;                       // loads per lane and slot
  for (int i = 0; i < nsplit; i += 2) {
    const int n0 = __shfl_sync(0xffffffffu, c_l.x, i);
    const int n1 = __shfl_sync(0xffffffffu, c_l.x, (i + 1) & 31);
    const bool two = i + 1 < nsplit;
    const uint2* b0 = sv.cand + ((size_t)(slot0 + i) * kBlockM + r_in) * kCandCap;
    const uint2* b1 = b0 + (two ? (size_t)kBlockM * kCandCap : 0);
    uint2 e0[kT], e1[kT];
#pragma unroll
    for (int t = 0; t < kT; ++t) {
      const int j = 32 * t + lane;
      e0[t] = (j < n0) ? __ldcg(b0 + j) : make_uint2(0u, 0u);
      e1[t] = (two && j < n1) ? __ldcg(b1 + j) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int t = 0; t < kT; ++t)
      if (32 * t < n0) keep_key(e0[t], 32 * t + lane < n0);
    if (two) {
#pragma unroll
      for (int t = 0; t < kT; ++t)
        if (32 * t < n1) keep_key(e1[t], 32 * t + lane < n1);
    }
  }
  __syncwarp();

  TopList top; top.init();
  if (ns > kSurvCap) {
    // ---- too many keys above the bound: streaming fold (the algorithm of merge_slots_kernel)
    unsigned long long kth = 0ull;
    int npend = 0;
    auto flush64 = [&]() {
      const unsigned long long b0 = sq[lane], b1 = sq[32 + lane];
      const unsigned long long rest = (64 + lane < npend) ? sq[64 + lane] : 0ull;
      __syncwarp();
      top.push(b0, b1, lane);
      npend -= 64;
      if (lane < npend) sq[lane] = rest;
      __syncwarp();
      const int p = k - 1;
      kth = shfl64(p < 32 ? top.r0 : top.r1, p & 31);
    };
    for (int i = 0; i < nsplit; ++i) {
      const int n = __shfl_sync(0xffffffffu, c_l.x, i);
      const uint2* b = sv.cand + ((size_t)(slot0 + i) * kBlockM + r_in) * kCandCap;
      for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        unsigned long long key = 0ull;
        if (j < n) {
          const uint2 e = __ldcg(b + j);
          if (f2key(__uint_as_float(e.x)) >= bound) {
            float z = __uint_as_float(e.x) * rs;
            if (softcap > 0.f) z = softcap * tanhf(z / softcap);
            key = pack_key(z, e.y);
          }
        }
        const bool keep = key > kth;
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) sq[npend + __popc(km & lt)] = key;
        npend += __popc(km);
        __syncwarp();
        if (npend >= 64) flush64();
      }
    }
    if (npend > 0) {
      if (lane + npend < 64) sq[npend + lane] = 0ull;
      if (lane + npend + 32 < 64) sq[npend + 32 + lane] = 0ull;
      __syncwarp();
      npend = 64;
      flush64();
    }
  } else {
    // ---- 2. a pivot with k <= #{key >= pivot} <= 64 ---------------------------------------------
    unsigned long long kr[kSurvCap / 32];
#pragma unroll
    for (int i = 0; i < kSurvCap / 32; ++i) kr[i] = (32 * i + lane < ns) ? sq[32 * i + lane] : 0ull;
    unsigned long long pivot = 1ull;                       // every key (keys are never 0)
    if (ns > 64) {
      unsigned long long mx = 0ull, mn = ~0ull;
#pragma unroll
      for (int i = 0; i < kSurvCap / 32; ++i) { mx = max64(mx, kr[i]); if (kr[i]) mn = min64(mn, kr[i]); }
      const uint32_t hi_w = __reduce_max_sync(0xffffffffu, (uint32_t)(mx >> 32));
      const uint32_t lo_w = __reduce_min_sync(0xffffffffu, (uint32_t)(mn >> 32));
      unsigned long long lo = (unsigned long long)lo_w << 32;          // #{>= lo} = ns   (> 64)
      unsigned long long hi = hi_w == 0xffffffffu ? ~0ull : ((unsigned long long)hi_w + 1ull) << 32;   // #{>= hi} = 0 (< k)
      float c_lo = (float)ns, c_hi = 0.5f;
      for (int it = 0;; ++it) {
        unsigned long long p;
        if (it < 6) {                                      // log-linear interpolation, aimed at (k + 64) / 2
          const float f = (__log2f(c_lo) - __log2f(0.5f * (float)(k + 64))) / (__log2f(c_lo) - __log2f(c_hi));
          p = lo + (unsigned long long)((double)(hi - lo) * (double)fminf(fmaxf(f, 0.02f), 0.98f));
        } else {
          p = lo + ((hi - lo) >> 1);
        }
        if (p <= lo) p = lo + 1ull;                        // (hi - lo >= 2 while the search runs: see below)
        int c = 0;
#pragma unroll
        for (int i = 0; i < kSurvCap / 32; ++i) c += (kr[i] >= p) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c > 64) { lo = p; c_lo = (float)c; }
        else if (c < k) { hi = p; c_hi = fmaxf((float)c, 0.5f); }
        else { pivot = p; break; }
        // keys are unique: #{>= lo} > 64 and #{>= hi} < k put more than 64 - k + 1 >= 1 keys inside
        // [lo, hi), so the interval cannot shrink below two values before a pivot is found
      }
    }
    // ---- 3. one sort of the keys at or above the pivot ------------------------------------------
    int nk = 0;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kSurvCap / 32; ++i) {
      if (32 * i < ns) {                                   // (uniform)
        const bool keep = kr[i] >= pivot && kr[i] != 0ull;
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) sq[nk + __popc(km & lt)] = kr[i];        // nk <= 64: positions below the sources read above
        nk += __popc(km);
      }
    }
    __syncwarp();
    const unsigned long long b0 = lane < nk ? sq[lane] : 0ull, b1 = 32 + lane < nk ? sq[32 + lane] : 0ull;
    top.push(b0, b1, lane);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = i * 32 + lane;
    const uint64_t key = i ? top.r1 : top.r0;
    if (p < k) {
      const bool empty = (key == 0ull);
      topk_val[(size_t)row * k + p] = empty ? -INFINITY : key2f((uint32_t)(key >> 32));
      topk_idx[(size_t)row * k + p] = empty ? -1ll : index_base + (long long)(uint32_t)(~(uint32_t)key);
    }
  }
}

__global__ void __launch_bounds__(128)
merge_ranks_kernel(const char* __restrict__ val_b, const char* __restrict__ idx_b,
                   const char* __restrict__ stats_b, size_t val_stride, size_t idx_stride,
                   size_t stats_stride, int R, int Q, int k, float* __restrict__ out_val,
                   long long* __restrict__ out_idx, float4* __restrict__ out_stats) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= Q) return;
  float m = -INFINITY;
  auto stats_of = [&](int r) {
    return reinterpret_cast<const float4*>(stats_b + (size_t)r * stats_stride)[row];
  };
  for (int r = lane; r < R; r += 32) m = fmaxf(m, stats_of(r).x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f, sum_z = 0.f, z_label = 0.f;
  for (int r = lane; r < R; r += 32) {
    const float4 st = stats_of(r);
    s += (st.y > 0.f) ? st.y * expf(st.x - m) : 0.f;
    sum_z += st.z;
    z_label += st.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sum_z += __shfl_xor_sync(0xffffffffu, sum_z, o);
    z_label += __shfl_xor_sync(0xffffffffu, z_label, o);
  }
  if (lane == 0) out_stats[row] = make_float4(m, s, sum_z, z_label);

  TopList top; top.init();
  for (int r = 0; r < R; ++r) {
    const float* v = reinterpret_cast<const float*>(val_b + (size_t)r * val_stride) + (size_t)row * k;
    const long long* ix =
        reinterpret_cast<const long long*>(idx_b + (size_t)r * idx_stride) + (size_t)row * k;
    uint64_t b0 = 0ull, b1 = 0ull;
    if (lane < k && ix[lane] >= 0) b0 = pack_key(v[lane], (uint32_t)ix[lane]);
    if (lane + 32 < k && ix[lane + 32] >= 0) b1 = pack_key(v[lane + 32], (uint32_t)ix[lane + 32]);
    // a rank's list is sorted (value desc, row asc) by contract; verify cheaply, sort if not
    const uint64_t nxt0 = shfl64(b0, (lane + 1) & 31), nxt1 = shfl64(b1, (lane + 1) & 31);
    const uint64_t first1 = shfl64(b1, 0);
    const bool ok = (lane == 31) ? (b0 >= first1) : (b0 >= nxt0 && b1 >= nxt1);
    if (__all_sync(0xffffffffu, ok)) top.push_sorted(b0, b1, lane); else top.push(b0, b1, lane);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = i * 32 + lane;
    const uint64_t key = i ? top.r1 : top.r0;
    if (p < k) {
      const bool empty = (key == 0ull);
      out_val[(size_t)row * k + p] = empty ? -INFINITY : key2f((uint32_t)(key >> 32));
      out_idx[(size_t)row * k + p] = empty ? -1ll : (long long)(uint32_t)(~(uint32_t)key);
    }
  }
}

// Clears the shared thresholds and drift counters before a scan.  A kernel of our own instead
// of cudaMemsetAsync so that it can ask for the scan kernel's shared-memory carve-out: the
// driver's memset kernel runs with the default one, and the SMs drain and reconfigure twice
// around it (~10-30 us per scan, the bulk of a 16-query scan's time).
__global__ void __launch_bounds__(256) zero_words_kernel(uint4* __restrict__ p, size_t n16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// Threshold seeding (scan_tc_kernel.cuh, kModeSeed): one warp per query row finds the k-th
// largest of the row's n <= 128 chunk maxima -- k distinct table rows reach it, so it is a lower
// bound of the row's k-th best score -- by an MSB-first radix descent on the order-preserving
// keys, and writes it into the row's shared threshold word.  The same launch clears the drift
// counters and joint words that follow the thresholds in the workspace.
constexpr int kMaxSeedChunks = 128;
__global__ void __launch_bounds__(256)
seed_select_kernel(const uint32_t* __restrict__ seed_max, int n_chunks, long long ld, long long rows,
                   int k, uint32_t* __restrict__ tau_shared, uint4* __restrict__ zero_ptr, size_t zero_n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_n16; i += (size_t)gridDim.x * blockDim.x)
    zero_ptr[i] = make_uint4(0u, 0u, 0u, 0u);
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  uint32_t key[kMaxSeedChunks / 32];
#pragma unroll
  for (int i = 0; i < kMaxSeedChunks / 32; ++i) {
    const int c = lane + 32 * i;
    key[i] = c < n_chunks ? __ldcg(seed_max + (size_t)c * ld + row) : 0u;   // 0 is below every real score
  }
  uint32_t prefix = 0u;
  int kr = k;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t want = (prefix >> bit) | 1u;
    int c = 0;
#pragma unroll
    for (int i = 0; i < kMaxSeedChunks / 32; ++i) c += ((key[i] >> bit) == want) ? 1 : 0;
    const int tot = __reduce_add_sync(0xffffffffu, c);
    if (tot >= kr) prefix |= (1u << bit); else kr -= tot;
  }
  if (lane == 0) tau_shared[row] = (n_chunks >= k) ? prefix : 0u;
}

cudaError_t launch_seed_select(const uint32_t* seed_max, int n_chunks, int64_t ld, int64_t rows, int k,
                               uint32_t* tau_shared, void* zero_ptr, size_t zero_bytes, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if (n_chunks > kMaxSeedChunks) return cudaErrorInvalidValue;
  static std::atomic<bool> pref_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !pref_set[dev].load()) {
    cudaFuncSetAttribute(seed_select_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    pref_set[dev].store(true);
  }
  seed_select_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(seed_max, n_chunks, (long long)ld, (long long)rows, k,
                                                               tau_shared, (uint4*)zero_ptr, zero_bytes / 16);
  return cudaGetLastError();
}

cudaError_t launch_zero(void* ptr, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return cudaSuccess;
  static std::atomic<bool> pref_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !pref_set[dev].load()) {
    cudaFuncSetAttribute(zero_words_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    pref_set[dev].store(true);
  }
  const size_t n16 = bytes / 16;   // the workspace carve-up keeps both ends 256-byte aligned
  zero_words_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, s>>>((uint4*)ptr, n16);
  return cudaGetLastError();
}

cudaError_t launch_merge_slots(const SlotView& sv, const SlotMap& map, int64_t Q, int k,
                               const float* inv_q, float scale, float softcap, int64_t index_base,
                               float* topk_val, int64_t* topk_idx, float* row_stats,
                               cudaStream_t s) {
  if (Q == 0) return cudaSuccess;
  // Same shared-memory carve-out as the scan kernel that precedes this one in every step:
  // alternating carve-outs makes the SMs drain and reconfigure before each launch.
  static std::atomic<bool> pref_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !pref_set[dev].load()) {
    cudaFuncSetAttribute(merge_slots_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(merge_slots_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(merge_slots_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(merge_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(merge_ranks_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    pref_set[dev].store(true);
  }
  // enough rows to fill the chip with one warp each, at most 32 slots per row: select, then sort once
  if (map.stride <= 32 && Q >= 2048 && !g_merge_variant.load())
    merge_rows_kernel<<<(unsigned)((Q + 3) / 4), 128, 0, s>>>(
        sv, map, (int)Q, k, inv_q, scale, softcap, (long long)index_base, topk_val, (long long*)topk_idx,
        (float4*)row_stats);
  // few rows with many slots (small Q split over all SMs): more warps per row
  else if (map.stride > 32 && Q <= 4096)
    merge_slots_kernel<16><<<(unsigned)Q, 32 * 16, 0, s>>>(
        sv, map, (int)Q, k, inv_q, scale, softcap, (long long)index_base, topk_val, (long long*)topk_idx,
        (float4*)row_stats);
  // A warp walks its slots one after the other, so rows with 8+ slots are split over four warps
  // while the launch still fits the chip in a wave or two
  else if (map.stride >= 8 && Q <= 16384)
    merge_slots_kernel<4><<<(unsigned)Q, 128, 0, s>>>(
        sv, map, (int)Q, k, inv_q, scale, softcap, (long long)index_base, topk_val, (long long*)topk_idx,
        (float4*)row_stats);
  else
    merge_slots_kernel<1><<<(unsigned)((Q + 3) / 4), 128, 0, s>>>(
        sv, map, (int)Q, k, inv_q, scale, softcap, (long long)index_base, topk_val, (long long*)topk_idx,
        (float4*)row_stats);
  return cudaGetLastError();
}

cudaError_t launch_merge_ranks(const float* val, const int64_t* idx, const float* stats,
                               size_t val_stride, size_t idx_stride, size_t stats_stride, int R,
                               int64_t Q, int k, float* out_val, int64_t* out_idx,
                               float* out_stats, cudaStream_t s) {
  if (Q == 0) return cudaSuccess;
  merge_ranks_kernel<<<(unsigned)((Q + 3) / 4), 128, 0, s>>>(
      (const char*)val, (const char*)idx, (const char*)stats, val_stride, idx_stride, stats_stride,
      R, (int)Q, k, out_val, (long long*)out_idx, (float4*)out_stats);
  return cudaGetLastError();
}

}  // namespace mcl
