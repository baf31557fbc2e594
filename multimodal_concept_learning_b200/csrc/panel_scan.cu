// Small query batches (Q <= 128: one row block -- the reference's own workloads score 6-96
// concept tokens against the vocabulary, token_embedding_analysis.py:183-260) in ONE launch.
//
// With a handful of query rows the table read bounds the scan (AI ~ Q flop/B), so the kernel is
// laid out for the table stream, not for the tensor pipe:
//   phase 1  every CTA owns a contiguous range of 128-row table panels.  The PANEL is the UMMA
//            M operand and the (zero-padded) query batch the N operand: D[128 table rows x NQ
//            queries], NQ = 16 .. 128.  TMA stages 16 KB table slices (+ the matching NQ x 64
//            query slice, L2 resident) into a ring of up to ten stages, so ~150 KB of table
//            bytes are in flight per SM; the grid is every SM, two or three panels each.
//            Epilogue: thread = table row = TMEM lane; one per-row scale (1/||t||), NQ per-query
//            scales from shared memory.  Per panel and query an epilogue warp leaves
//              - its 32 scores in the workspace, [Q][ld] fp32 (4 Q/(2 D) of the table bytes, L2
//                resident; the warp's rows are consecutive: 128 B stores),
//              - the maximum of the 32 as an order-preserving key (REDUX), [Q][V/32],
//              - and folds (max, sum exp(z - max), sum z) of the 32 -- a transposing shuffle
//                reduction, 16 queries in 16 shuffles -- into per-query running values that the
//                CTA writes once at the end, [grid][Q].
//   barrier  CTAs take a ticket when their panels are done; the last min(Q, grid) arrivals wait
//            for the others (grid <= SM count, one CTA per SM: all are resident) and each
//            selects for one query row.  The two counters are reset by the last selector, so a
//            launch leaves them as it found them: zero.
//   phase 2  one CTA per query row, 512 threads, no pass over the row: the statistics are the
//            fold of the grid's partial values; the ceil(k/16)-th largest group maximum of every
//            warp, minimised over the 16 warps, is reached by >= k scores at distinct table
//            rows, i.e. it is a lower bound of the k-th best; only the ~k groups of 32 scores
//            whose maximum reaches it are read back, and their scores at or above the bound
//            (~2k on random data) go to shared memory as 64-bit keys (value key, ~row).  Warps
//            sort 64 survivors each in registers (toplist.cuh), warp 0 folds the lists.  More
//            than 1024 such groups or 2048 survivors (adversarial layouts, long runs of equal
//            scores): exact MSB-first radix select over the row's 64-bit keys, eight passes.
// Same values (z = (acc * inv_t) * (inv_q * scale), soft-cap included) and tie rule (value desc,
// table row asc) as the streaming path: tests/test_gpu_parity.py asserts bit-identical top-k.
#include <cuda.h>
#include <stdio.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include "kernels.h"
#include "rownorm.cuh"
#include "rowstate.cuh"
#include "toplist.cuh"

namespace mcl {

constexpr int kPnThreads = 512;
constexpr int kPnWarps = kPnThreads / 32;
constexpr int kPnM = 128;                                  // table rows per panel (UMMA M, TMEM lanes)
constexpr uint32_t kPnTBytes = kPnM * kBlockK * 2;         // 16 KB: one K slice of a panel
constexpr uint32_t kPnRingBytes = 192 * 1024;
constexpr int kPnAcc = 4;                                  // accumulator stages in TMEM
constexpr int kPnLiveCap = 2048;                           // survivors of the bound kept in shared memory
constexpr uint32_t kPnBarBytes = 512;
constexpr int kPnBatch = 8;                                // 16-byte loads in flight per thread (phase 2)

constexpr int kPnGroupCap = 1024;                          // groups whose maximum reaches the bound
struct PnScratch {
  unsigned long long live[kPnLiveCap];
  union {
    unsigned long long lists[kPnWarps][64];                // phase 2: the warps' sorted lists
    float4 part[4][128];                                   // phase 1: the epilogue warps' (max, sum exp, sum z)
  };
  union {
    int glist[kPnGroupCap];
    int hist[256];
  };
  float red[kPnWarps][4];
  uint32_t wkey[kPnWarps];
  unsigned long long prefix;
  int kk, cnt, gcnt, ticket;
};
constexpr uint32_t kPnSmemBytes = kPnRingBytes + kPnBarBytes + 2 * 128 * 4 + sizeof(PnScratch) + 1024;
static_assert(kPnSmemBytes <= 227 * 1024, "shared memory of the panel scan");

struct PnParams {
  int Q, V, D, k, num_kb, num_tiles;
  const float* inv_q;              // nullable
  const float* inv_t;              // nullable
  float scale, softcap;            // softcap 0 = off
  const __nv_bfloat16* q_rows;     // non-null: 1/||q_row|| is computed here (rownorm.cuh)
  long long ldq;
  long long index_base;
  const long long* labels;         // nullable
  float* scores;                   // [Q][ld]
  long long ld;
  uint32_t* gmax;                  // [Q][gld] keys of the maxima of 32 consecutive scores
  long long gld;
  float4* part;                    // [grid][Q] (max, sum exp(z - max), sum z, -) over a CTA's panels
  float* topk_val;
  long long* topk_idx;
  float4* row_stats;
  uint32_t* sync;                  // [0] CTAs done with phase 1, [1] CTAs done selecting; zero between launches
  unsigned long long* fault;       // mapped host word: barrier waits that gave up (nullable)
  unsigned long long* timing;      // nullable: [grid][4] globaltimer at CTA start / panels done / barrier passed / end
};

__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// f(z, col, valid) for every score of the row, kPnBatch 16-byte L2 loads in flight per thread;
// called by all 32 lanes of a warp together (valid = false past the row's end), so that f may use
// warp collectives.  Column c lives in thread (c / 4) % kPnThreads; the n % 4 tail columns in
// threads 0 .. 2.
template <typename F>
__device__ __forceinline__ void for_each_score(const float* __restrict__ src, int n, int tid, F&& f) {
  const int nvec = n >> 2;
  const float4* src4 = reinterpret_cast<const float4*>(src);
  for (int base = 0; base < nvec; base += kPnThreads * kPnBatch) {   // (block-uniform trip count)
    float4 x[kPnBatch];
#pragma unroll
    for (int u = 0; u < kPnBatch; ++u) {
      const int i = base + u * kPnThreads + tid;
      x[u] = i < nvec ? __ldcg(src4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kPnBatch; ++u) {
      const int i = base + u * kPnThreads + tid;
      const bool ok = i < nvec;
      f(x[u].x, 4 * i, ok); f(x[u].y, 4 * i + 1, ok); f(x[u].z, 4 * i + 2, ok); f(x[u].w, 4 * i + 3, ok);
    }
  }
  if (tid < 32) {
    const int c = 4 * nvec + tid;
    f(c < n ? __ldcg(src + c) : 0.f, c, c < n);
  }
}

// (max, sum exp(z - max)) pairs fold like this; exp2(-inf) = 0 covers the empty side
__device__ __forceinline__ void lse_fold(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  s = (m > -INFINITY ? s * ex2_fast((m - mn) * kLog2e) : 0.f) + (m2 > -INFINITY ? s2 * ex2_fast((m2 - mn) * kLog2e) : 0.f);
  m = mn;
}

// one append per warp and call: `keep` lanes get consecutive positions from *counter
__device__ __forceinline__ int warp_append(bool keep, int* counter, int lane) {
  const unsigned kept = __ballot_sync(0xffffffffu, keep);
  if (!kept) return 0;
  const int leader = __ffs(kept) - 1;
  int pos = 0;
  if (lane == leader) pos = atomicAdd(counter, __popc(kept));
  return __shfl_sync(0xffffffffu, pos, leader) + __popc(kept & ((1u << lane) - 1u));
}

// ---- phase 2: exact top-k + statistics of one query row by the whole CTA ----------------------
__device__ __forceinline__ void pn_select_row(const PnParams& p, PnScratch& sc, int row, int tid, int G) {
  const int lane = tid & 31, warp = tid >> 5;
  const int n = p.V, k = p.k;                     // k <= V (checked by the host)
  const int ng = (n + 31) >> 5;                   // groups of 32 consecutive scores
  const float* src = p.scores + (size_t)row * p.ld;
  const uint32_t* gk = p.gmax + (size_t)row * p.gld;

  // (all loads of the first round trip are issued before any is used: the CTAs' partial statistics
  // and the group maxima, which stay in registers for the second look when the row has <= 2048 groups)
  const bool few = ng <= 4 * kPnThreads;
  float4 pv = make_float4(-INFINITY, 0.f, 0.f, 0.f);
  if (tid < G) pv = __ldcg(p.part + (size_t)tid * p.Q + row);
  float zl = 0.f;                                 // the label's score (thread 0)
  if (tid == 0 && p.labels) {
    const long long lg = p.labels[row], l = lg - p.index_base;
    if (lg != -100 && l >= 0 && l < n) zl = __ldcg(src + l);
  }
  uint32_t gkey[4] = {0u, 0u, 0u, 0u};            // keys of finite scores are never 0
  uint32_t mx = 0u;
  if (few) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = tid + i * kPnThreads;
      if (g < ng) gkey[i] = __ldcg(gk + g);
    }
    mx = max(max(gkey[0], gkey[1]), max(gkey[2], gkey[3]));
  } else {
    for (int g = tid; g < ng; g += kPnThreads) mx = max(mx, __ldcg(gk + g));
  }
  // statistics: fold the CTAs' partial values
  {
    float m = pv.x, s = pv.y, sz = pv.z;
    for (int b = tid + kPnThreads; b < G; b += kPnThreads) {
      const float4 v = __ldcg(p.part + (size_t)b * p.Q + row);
      lse_fold(m, s, v.x, v.y);
      sz += v.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      lse_fold(m, s, m2, s2);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) { sc.red[warp][0] = m; sc.red[warp][1] = s; sc.red[warp][2] = sz; }
  }
  // a lower bound of the k-th best score from the group maxima: warps whose 32 lanes all hold a
  // group each name the J-th largest of their lane maxima, nfull * J >= k
  const int nfull = min(kPnWarps, ng >> 5);
  const int J = nfull > 0 ? (k + nfull - 1) / nfull : 33;
  {
    uint32_t x = mx, wb = 0u;
    if (J <= 32)
      for (int j = 0; j < J; ++j) {
        wb = __reduce_max_sync(0xffffffffu, x);
        const unsigned eq = __ballot_sync(0xffffffffu, x == wb);
        if (lane == __ffs(eq) - 1) x = 0u;        // take out ONE lane that holds it
      }
    if (lane == 0) sc.wkey[warp] = (J <= 32 && warp < nfull) ? wb : 0xffffffffu;
    if (tid == 0) { sc.cnt = 0; sc.gcnt = 0; }
  }
  __syncthreads();
  uint32_t bound = 0u;                            // 0: every score is a candidate (small tables)
  if (J <= 32) {
    bound = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < kPnWarps; ++w) bound = min(bound, sc.wkey[w]);
  }
  if (tid == 0) {
    float m = -INFINITY, s = 0.f, sz = 0.f;
#pragma unroll
    for (int w = 0; w < kPnWarps; ++w) { lse_fold(m, s, sc.red[w][0], sc.red[w][1]); sz += sc.red[w][2]; }
    p.row_stats[row] = make_float4(m, s, sz, zl);
  }
  // groups whose maximum reaches the bound (block-uniform trip counts: warp collectives inside)
  if (few) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool hit = gkey[i] != 0u && gkey[i] >= bound;
      const int pos = warp_append(hit, &sc.gcnt, lane);
      if (hit && pos < kPnGroupCap) sc.glist[pos] = tid + i * kPnThreads;
    }
  } else {
    for (int g0 = 0; g0 < ng; g0 += kPnThreads) {
      const int g = g0 + tid;
      const bool hit = g < ng && __ldcg(gk + g) >= bound;
      const int pos = warp_append(hit, &sc.gcnt, lane);
      if (hit && pos < kPnGroupCap) sc.glist[pos] = g;
    }
  }
  __syncthreads();
  const int ngq = sc.gcnt;
  bool exact = ngq > kPnGroupCap;
  if (!exact) {
    // their scores at or above the bound -> shared memory, four groups per warp in flight
    for (int i0 = warp * 4; i0 < ngq; i0 += kPnWarps * 4) {
      float z[4];
      int c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = (i0 + u < ngq) ? sc.glist[i0 + u] * 32 + lane : n;
        z[u] = c[u] < n ? __ldcg(src + c[u]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t key = f2key(z[u]);
        const bool keep = c[u] < n && key >= bound;
        const int pos = warp_append(keep, &sc.cnt, lane);
        if (keep && pos < kPnLiveCap) sc.live[pos] = ((unsigned long long)key << 32) | (unsigned long long)(uint32_t)(~(uint32_t)c[u]);
      }
    }
    __syncthreads();
    exact = sc.cnt > kPnLiveCap;
  }
  if (exact) {
    // ---- the bound kept too many keys: exact MSB-first radix select over (value key, ~row),
    // one byte per pass (the keys are unique, so exactly k of them are >= the k-th largest)
    unsigned long long prefix = 0ull;
    int kk = k;
    __syncthreads();
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      if (tid < 256) sc.hist[tid] = 0;
      __syncthreads();
      for_each_score(src, n, tid, [&](float z, int c, bool ok) {
        const unsigned long long key = ((unsigned long long)f2key(z) << 32) | (unsigned long long)(uint32_t)(~(uint32_t)c);
        const bool lv = ok && (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8)));
        const int bin = lv ? (int)((key >> shift) & 255ull) : 256;
        const unsigned peers = __match_any_sync(0xffffffffu, bin);   // warp-aggregated: equal scores share a bin
        if (lv && lane == __ffs(peers) - 1) atomicAdd(&sc.hist[bin], __popc(peers));
      });
      __syncthreads();
      if (warp == 0) {
        int mine = 0;                              // lane l owns bins 255-8l .. 248-8l
#pragma unroll
        for (int j = 0; j < 8; ++j) mine += sc.hist[255 - 8 * lane - j];
        int above = mine;                          // inclusive scan over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, above, o);
          if (lane >= o) above += t;
        }
        const int before = above - mine;
        if (before < kk && kk <= above) {
          int cum = before, bin = 255 - 8 * lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int h = sc.hist[255 - 8 * lane - j];
            if (cum < kk && kk <= cum + h) { bin = 255 - 8 * lane - j; break; }
            cum += h;
          }
          sc.prefix = prefix | ((unsigned long long)bin << shift);
          sc.kk = kk - cum;
        }
      }
      __syncthreads();
      prefix = sc.prefix;
      kk = sc.kk;
    }
    if (tid == 0) sc.cnt = 0;
    __syncthreads();
    for_each_score(src, n, tid, [&](float z, int c, bool ok) {
      const unsigned long long key = ((unsigned long long)f2key(z) << 32) | (unsigned long long)(uint32_t)(~(uint32_t)c);
      const bool keep = ok && key >= prefix;
      const int pos = warp_append(keep, &sc.cnt, lane);
      if (keep && pos < kPnLiveCap) sc.live[pos] = key;
    });
    __syncthreads();
  }
  const int nlive = min(sc.cnt, kPnLiveCap);
  // warps sort 64 survivors at a time into a running top-64 of their own
  const int nchunk = (nlive + 63) >> 6;
  if (warp < nchunk) {
    TopList top; top.init();
    for (int c = warp; c < nchunk; c += kPnWarps) {
      const int i0 = c * 64 + lane, i1 = i0 + 32;
      top.push(i0 < nlive ? sc.live[i0] : 0ull, i1 < nlive ? sc.live[i1] : 0ull, lane);
    }
    sc.lists[warp][lane] = top.r0;
    sc.lists[warp][lane + 32] = top.r1;
  }
  __syncthreads();
  if (warp == 0) {
    TopList top;
    top.r0 = sc.lists[0][lane];
    top.r1 = sc.lists[0][lane + 32];
    const int nl = min(nchunk, kPnWarps);
    for (int w = 1; w < nl; ++w) top.push_sorted(sc.lists[w][lane], sc.lists[w][lane + 32], lane);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pos = i * 32 + lane;
      const unsigned long long key = i ? top.r1 : top.r0;
      if (pos < k) {
        const bool empty = (key == 0ull);
        p.topk_val[(size_t)row * k + pos] = empty ? -INFINITY : key2f((uint32_t)(key >> 32));
        p.topk_idx[(size_t)row * k + pos] = empty ? -1ll : p.index_base + (long long)(uint32_t)(~(uint32_t)key);
      }
    }
  }
  __syncthreads();                                 // the scratch is reused by the CTA's next row
}

template <int NQ>
__global__ void __launch_bounds__(kPnThreads, 1)
panel_scan_kernel(const __grid_constant__ CUtensorMap tm_t, const __grid_constant__ CUtensorMap tm_t32,
                  const __grid_constant__ CUtensorMap tm_q, const PnParams p) {
  constexpr uint32_t kQBytes = NQ * kBlockK * 2;                 // 2 .. 16 KB
  constexpr uint32_t kStageBytes = kPnTBytes + kQBytes;
  constexpr uint32_t kStages = kPnRingBytes / kStageBytes;       // 10 (NQ = 16) .. 6 (NQ = 128)
  constexpr uint32_t kTmemCols = kPnAcc * NQ;                    // 64 .. 512: a power of two >= 32
  // epilogue warps: kEG per TMEM lane quarter, which take the chunks of 16 queries in turn (with
  // 128 queries the statistics of a panel are ~3 k warp instructions: one warp per quarter would
  // bound the scan)
  constexpr int kEG = NQ >= 48 ? 3 : NQ / 16;
  static_assert(kStages >= 4 && (2 * kStages + 2 * kPnAcc) * 8 + 8 <= kPnBarBytes, "barrier area");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B atoms
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kPnRingBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](uint32_t a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](uint32_t a) { return bar_base + 8u * (2 * kStages + kPnAcc + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2 * kPnAcc);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kPnRingBytes + 8u * (2 * kStages + 2 * kPnAcc));
  float* c_rs = reinterpret_cast<float*>(smem_gen + kPnRingBytes + kPnBarBytes);   // [128] inv_q * scale
  float* c_rc = c_rs + 128;                                                        // [128] rs / softcap
  PnScratch& sc = *reinterpret_cast<PnScratch*>(smem_gen + kPnRingBytes + kPnBarBytes + 2 * 128 * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto stamp = [&](int i) {
    if (p.timing && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.timing[4 * blockIdx.x + i] = t;
    }
  };
  stamp(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_t);
    tma_prefetch_desc(&tm_t32);
    tma_prefetch_desc(&tm_q);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (uint32_t s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (uint32_t a = 0; a < kPnAcc; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * kEG); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int G = (int)gridDim.x;
  // This CTA's table rows: a contiguous range of whole groups of 32 rows (the unit of the group
  // maxima), walked in panels of 128; the last panel may be short -- only its rows are fetched
  // (32-row TMA boxes), the MMA runs over the stale rest of the stage and nobody reads those lanes.
  const int ngroups = (p.V + 31) >> 5;
  const int r0 = (int)((long long)blockIdx.x * ngroups / G) * 32;
  const int r1 = min(p.V, (int)((long long)(blockIdx.x + 1) * ngroups / G) * 32);
  const int npanel = (r1 - r0 + kPnM - 1) / kPnM;
  const int num_kb = p.num_kb;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < npanel; ++t) {
        const int row = r0 + t * kPnM;
        const int nbox = min(4, (r1 - row + 31) >> 5);          // 32-row boxes of this panel
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_backoff(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * kStageBytes;
          if (nbox == 4) {
            mbar_expect_tx(full_bar(stage), kStageBytes);
            tma_load_2d(sa, &tm_t, full_bar(stage), kb * kBlockK, row, kL2EvictNormal);
          } else {
            mbar_expect_tx(full_bar(stage), kQBytes + nbox * 4096u);
            for (int b = 0; b < nbox; ++b)                       // 32 rows x 128 B = four swizzle atoms
              tma_load_2d(sa + b * 4096u, &tm_t32, full_bar(stage), kb * kBlockK, row + 32 * b, kL2EvictNormal);
          }
          tma_load_2d(sa + kPnTBytes, &tm_q, full_bar(stage), kb * kBlockK, 0, kL2EvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kPnM, NQ);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = 0; t < npanel; ++t) {
        mbar_wait_backoff(tempty_bar(acc), acc_phase ^ 1u);      // the epilogue has drained this stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * NQ;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint64_t adesc = umma_desc_sw128(sa);              // table panel: M
          const uint64_t bdesc = umma_desc_sw128(sa + kPnTBytes);  // queries: N
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
          umma_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == kPnAcc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ==================== per-query scales, then the epilogue (warps 4..7) ====================
    if (p.q_rows) {
      // two rows per memory round trip and warp; rownorm.cuh: the bits of row_inv_norm_kernel
      for (int r0 = (warp - 2) * 2; r0 < p.Q; r0 += (kPnWarps - 2) * 2) {
        float v[2];
        const int nr = min(2, p.Q - r0);
        warp_rows_inv_norm<__nv_bfloat16, 2>(p.q_rows + (long long)r0 * p.ldq, p.ldq, nr, p.D, lane, v);
        if (lane < nr) {
          const float rs = (lane ? v[1] : v[0]) * p.scale;
          c_rs[r0 + lane] = rs;
          c_rc[r0 + lane] = p.softcap > 0.f ? rs / p.softcap : 0.f;
        }
      }
    } else {
      for (int r = tid - 64; r < p.Q; r += kPnThreads - 64) {
        const float rs = (p.inv_q ? p.inv_q[r] : 1.f) * p.scale;
        c_rs[r] = rs;
        c_rc[r] = p.softcap > 0.f ? rs / p.softcap : 0.f;
      }
    }
    named_bar_sync(1, kPnThreads - 64);
    if (warp >= 4 && warp < 4 + 4 * kEG) {
      const int quarter = warp & 3;                      // TMEM lanes this warp may read
      const int eg = (warp - 4) >> 2;                    // chunk j of 16 queries is this warp's if j % kEG == eg
      const int last_j = ((NQ / 16 - 1 - eg) / kEG) * kEG + eg;
      const bool cap = p.softcap > 0.f;
      const int qown = (lane >> 1) & 15;                 // the query (of 16) whose sums end in this lane
      // ... and its running (max, sum exp, sum z) live in shared memory, sc.part[quarter][query]: the
      // chunk body below is ONE copy of ~300 instructions that every chunk and panel reuses.  (Unrolled
      // over the 8 chunks of 128 queries with the running values in registers it was ~2 k instructions
      // executed once per panel: ncu, 96 queries: 6.1 "no instruction" stalls per issue, 0.22 IPC.)
      float4* run = sc.part[quarter];
      if ((lane & 1) == 0)
        for (int j = eg; j < NQ / 16; j += kEG) run[j * 16 + qown] = make_float4(-INFINITY, 0.f, 0.f, 0.f);
      __syncwarp();
      uint32_t acc = 0, acc_phase = 0;
      for (int t = 0; t < npanel; ++t) {
        const int wrow = r0 + t * kPnM + quarter * 32;   // first table row of this warp (a whole group)
        const int row = wrow + lane;                     // table row of this thread
        const bool rv = row < r1;
        const bool warp_rows = wrow < r1;                // (uniform) the warp holds table rows
        const float it = (p.inv_t && rv) ? __ldg(p.inv_t + row) : 1.f;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * NQ;
        float* dst = p.scores + row;
        uint32_t* gdst = p.gmax + (size_t)(wrow >> 5);
#pragma unroll 1
        for (int j = eg; j < NQ / 16; j += kEG) {        // this warp's chunks of 16 queries
          float a[16];
          tmem_ld_32x32_x16(taddr + j * 16, a);
          if (j == last_j) {                             // every tcgen05.ld of this warp and panel has landed
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          if (!warp_rows || j * 16 >= p.Q) continue;     // (uniform)
          uint32_t gkey[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int qi = j * 16 + i;
            const float y = a[i] * it;
            a[i] = cap ? p.softcap * tanhf(y * c_rc[qi]) : y * c_rs[qi];
            if (rv && qi < p.Q) __stcg(dst + (size_t)qi * p.ld, a[i]);   // 32 consecutive table rows per warp: 128 B
            gkey[i] = __reduce_max_sync(0xffffffffu, rv ? f2key(a[i]) : 0u);
          }
          uint32_t kown = gkey[0];
#pragma unroll
          for (int i = 1; i < 16; ++i) kown = (qown == i) ? gkey[i] : kown;
          if ((lane & 1) == 0 && j * 16 + qown < p.Q) gdst[(size_t)(j * 16 + qown) * p.gld] = kown;
          // exp(z - group max) and z, summed over the warp's 32 rows for 16 queries at once: every
          // step halves the values a lane carries and doubles the rows each value covers
          float e[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            e[i] = rv ? ex2_fast((a[i] - key2f(gkey[i])) * kLog2e) : 0.f;
            a[i] = rv ? a[i] : 0.f;
          }
#pragma unroll
          for (int w = 8; w >= 1; w >>= 1) {             // lane bit 2w picks the upper w values
            const bool up = (lane & (2 * w)) != 0;
#pragma unroll
            for (int h = 0; h < w; ++h) {
              const float se = up ? e[h] : e[h + w], ke = up ? e[h + w] : e[h];
              const float sa = up ? a[h] : a[h + w], ka = up ? a[h + w] : a[h];
              e[h] = ke + __shfl_xor_sync(0xffffffffu, se, 2 * w);
              a[h] = ka + __shfl_xor_sync(0xffffffffu, sa, 2 * w);
            }
          }
          e[0] += __shfl_xor_sync(0xffffffffu, e[0], 1);
          a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
          if ((lane & 1) == 0) {                         // (only this lane ever touches the entry)
            float4 r = run[j * 16 + qown];
            lse_fold(r.x, r.y, key2f(kown), e[0]);
            r.z += a[0];
            run[j * 16 + qown] = r;
          }
        }
        if (++acc == kPnAcc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (tid < p.Q) {                                       // the CTA's (max, sum exp, sum z) per query
    float m = -INFINITY, s = 0.f, sz = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float4 v = sc.part[w][tid];
      lse_fold(m, s, v.x, v.y);
      sz += v.z;
    }
    __stcg(p.part + (size_t)blockIdx.x * p.Q + tid, make_float4(m, s, sz, 0.f));
  }
  __threadfence();                                       // scores, group maxima, partial sums: visible before the ticket
  __syncthreads();
  stamp(1);
  // ---- barrier: the last min(Q, G) CTAs to get here select, one query row at a time each ----
  if (tid == 0) sc.ticket = (int)atomicAdd(p.sync, 1u);
  __syncthreads();
  const int nsel = min(p.Q, G);
  const int my = sc.ticket - (G - nsel);
  if (my < 0) { stamp(2); stamp(3); return; }
  if (tid == 0) {
    // grid <= SM count and one CTA per SM: every CTA is resident (or done) on a GPU that is not
    // shared; bounded (~2 s) so that foreign work holding SMs cannot hang the stream for good
    long long spin = 0;
    while (ld_acquire_u32(p.sync) < (uint32_t)G && spin < (1ll << 24)) { __nanosleep(100); ++spin; }
    sc.cnt = (spin == (1ll << 24)) ? -1 : 0;
    if (spin == (1ll << 24) && p.fault) atomicAdd_system(p.fault, 1ull);
  }
  __syncthreads();
  const bool gave_up = sc.cnt < 0;
  __syncthreads();
  __threadfence();
  stamp(2);
  for (int row = my; row < p.Q; row += nsel) {
    if (gave_up) {                                       // fail loudly: NaN scores, no rows
      for (int i = tid; i < p.k; i += kPnThreads) {
        p.topk_val[(size_t)row * p.k + i] = __int_as_float(0x7fc00000);
        p.topk_idx[(size_t)row * p.k + i] = -1ll;
      }
      if (tid == 0) p.row_stats[row] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
    } else {
      pn_select_row(p, sc, row, tid, G);
    }
  }
  __syncthreads();
  stamp(3);
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(p.sync + 1, 1u) == (uint32_t)(nsel - 1)) {   // every selector is past its wait
      p.sync[0] = 0u;
      p.sync[1] = 0u;
      __threadfence();
    }
  }
}

// ---- host side -----------------------------------------------------------------------------
namespace {
// zeroed words owned by the library: 64 pairs per device, handed out round robin so that scans
// in flight on different streams do not share a pair
uint32_t* g_pn_sync[64];
unsigned long long* g_pn_fault[64];
std::atomic<unsigned> g_pn_next{0};
std::mutex g_pn_mu;

template <int NQ>
cudaError_t pn_launch(int grid, cudaStream_t s, const CUtensorMap& tm_t, const CUtensorMap& tm_t32, const CUtensorMap& tm_q,
                      const PnParams& p) {
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev].load()) {
    cudaError_t e = cudaFuncSetAttribute(panel_scan_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPnSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev].store(true);
  }
  panel_scan_kernel<NQ><<<grid, kPnThreads, kPnSmemBytes, s>>>(tm_t, tm_t32, tm_q, p);
  return cudaGetLastError();
}
}  // namespace

long long panel_barrier_faults() {
  long long n = 0;
  for (auto* w : g_pn_fault)
    if (w) n += (long long)*(volatile unsigned long long*)w;
  return n;
}

cudaError_t launch_panel_scan(const ScanArgs& a, int sm_count, float* scores, int64_t ld, uint32_t* gmax,
                              int64_t gld, float* part, float* topk_val,
                              int64_t* topk_idx, float* row_stats, cudaStream_t s, char* err, size_t errlen) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  {
    std::lock_guard<std::mutex> lk(g_pn_mu);
    if (!g_pn_sync[dev]) {
      // first scan on this device (never inside a stream capture: graphed.py warms up first)
      uint32_t* w = nullptr;
      cudaError_t e = cudaMalloc(&w, 64 * 64);
      if (e != cudaSuccess) { snprintf(err, errlen, "cudaMalloc(panel sync words)"); return e; }
      e = cudaMemset(w, 0, 64 * 64);
      if (e != cudaSuccess) { cudaFree(w); snprintf(err, errlen, "cudaMemset(panel sync words)"); return e; }
      void* h = nullptr;
      if (cudaHostAlloc(&h, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
        *(volatile unsigned long long*)h = 0ull;
        g_pn_fault[dev] = (unsigned long long*)h;
      } else {
        cudaGetLastError();
      }
      g_pn_sync[dev] = w;
    }
  }
  const int NQ = a.Q <= 16 ? 16 : (a.Q <= 32 ? 32 : (a.Q <= 64 ? 64 : 128));
  CUtensorMap tm_t, tm_t32, tm_q;
  if (!make_tmap_bf16(&tm_t, a.table, a.V, a.D, a.ldt, kPnM) || !make_tmap_bf16(&tm_t32, a.table, a.V, a.D, a.ldt, 32) ||
      !make_tmap_bf16(&tm_q, a.q, a.Q, a.D, a.ldq, NQ)) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (Q=%lld V=%lld D=%lld ldq=%lld ldt=%lld)",
             (long long)a.Q, (long long)a.V, (long long)a.D, (long long)a.ldq, (long long)a.ldt);
    return cudaErrorInvalidValue;
  }
  PnParams p{};
  p.Q = (int)a.Q; p.V = (int)a.V; p.D = (int)a.D; p.k = a.k;
  p.num_kb = (int)((a.D + kBlockK - 1) / kBlockK);
  p.num_tiles = (int)((a.V + kPnM - 1) / kPnM);
  p.inv_q = a.inv_q; p.inv_t = a.inv_t; p.scale = a.scale; p.softcap = a.softcap;
  p.q_rows = a.qnorm_in_kernel ? (const __nv_bfloat16*)a.q : nullptr;
  p.ldq = a.ldq;
  p.index_base = a.index_base; p.labels = (const long long*)a.labels;
  p.scores = scores; p.ld = ld;
  p.gmax = gmax; p.gld = gld; p.part = (float4*)part;
  p.topk_val = topk_val; p.topk_idx = (long long*)topk_idx; p.row_stats = (float4*)row_stats;
  // pairs 0..31 rotate over direct launches, 32..63 over launches recorded into CUDA graphs (a graph
  // keeps its pair for life: it must not meet a direct launch on another stream in the same pair)
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  static std::atomic<unsigned> next_captured{0};
  const unsigned pair = cap == cudaStreamCaptureStatusActive ? 32u + next_captured.fetch_add(1) % 32u
                                                             : g_pn_next.fetch_add(1) % 32u;
  p.sync = g_pn_sync[dev] + 16 * pair;
  p.fault = g_pn_fault[dev];
  p.timing = (unsigned long long*)a.timing;
  const int grid = std::min(sm_count, (int)((a.V + 31) / 32));   // every CTA owns at least one group of 32 rows
  switch (NQ) {
    case 16: return pn_launch<16>(grid, s, tm_t, tm_t32, tm_q, p);
    case 32: return pn_launch<32>(grid, s, tm_t, tm_t32, tm_q, p);
    case 64: return pn_launch<64>(grid, s, tm_t, tm_t32, tm_q, p);
    default: return pn_launch<128>(grid, s, tm_t, tm_t32, tm_q, p);
  }
}

}  // namespace mcl
