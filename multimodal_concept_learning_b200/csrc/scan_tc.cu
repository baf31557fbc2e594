// The product path: fused similarity scan on Blackwell tensor cores.
//
//   scores tile [128 queries x 256 table rows] = Q_tile (bf16, K-major) * T_tile^T
//   - operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) into a 4-stage ring,
//   - multiplied by tcgen05.mma (kind::f16, M128 N256 K16) issued by ONE thread,
//   - accumulated in TMEM (2 x 256 fp32 columns, double buffered),
//   - drained by four epilogue warps with tcgen05.ld: thread = query row = TMEM lane, so the
//     online log-sum-exp, running sum, label pick-up and top-k filter are thread-private
//     (rowstate.cuh) and the [Q x V] score matrix never leaves the SM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
//
// Scheduling: persistent CTAs.  The job list is (row-group, table tile) with a row-group =
// g consecutive 128-query row blocks; CTA c belongs to group c / g as member c % g, and a
// group owns a contiguous range of jobs ("stream-K" over the list), in which member r scans
// row block rg*g + r.  The g members of a group walk the same table tiles at the same time,
// so a table tile is fetched from HBM once per group and hit in L2 by the other members,
// while the contiguous split keeps all 148 SMs busy for any Q.  Every (CTA, row block)
// range ends in a partial-result slot; merge.cu combines the slots.
#include <cuda.h>
#include <stdio.h>
#include <algorithm>
#include "rowstate.cuh"
#include "kernels.h"

namespace mcl {

constexpr int kStages = 4;
constexpr int kTcThreads = 192;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr uint32_t kBBytes = kBlockN * kBlockK * 2;   // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;   // 48 KB
constexpr uint32_t kTmemCols = 512;                   // 2 accumulator stages x 256 columns
constexpr uint32_t kBarBytes = 256;
constexpr uint32_t kTcSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + align slack

struct TcParams {
  int Q, V, D, k;
  int num_rb, num_vt, num_kb;
  int g, jpg, max_seg;
  long long total_jobs;
  const float* inv_q;
  const float* inv_t;
  float scale;
  long long index_base;
  const long long* labels;
  SlotView sv;
  float* dbg_scores;
};

__global__ void __launch_bounds__(kTcThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_t,
               const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](uint32_t a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](uint32_t a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * kStageBytes + 8u * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_t);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (uint32_t s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (uint32_t a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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

  // this CTA's contiguous job range
  const int grp = blockIdx.x / p.g, member = blockIdx.x % p.g;
  const long long j0 = (long long)grp * p.jpg;
  const long long j1 = (j0 + p.jpg < p.total_jobs) ? j0 + p.jpg : p.total_jobs;
  const int rg_first = (int)(j0 / p.num_vt);
  const int vt_first = (int)(j0 % p.num_vt);

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int rg = rg_first, vt = vt_first;
      for (long long j = j0; j < j1; ++j) {
        const int rb = rg * p.g + member;
        if (rb < p.num_rb) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), kStageBytes);
            const uint32_t sa = smem_base + stage * kStageBytes;
            tma_load_2d(sa, &tm_q, full_bar(stage), kb * kBlockK, rb * kBlockM);
            tma_load_2d(sa + kABytes, &tm_t, full_bar(stage), kb * kBlockK, vt * kBlockN);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
        if (++vt == p.num_vt) { vt = 0; ++rg; }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kBlockN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      int rg = rg_first, vt = vt_first;
      for (long long j = j0; j < j1; ++j) {
        const int rb = rg * p.g + member;
        if (rb < p.num_rb) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // epilogue has drained this stage
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kBlockN;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * kStageBytes;
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = umma_desc_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)   // +32 B per K=16 step inside the 128 B row
              umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
            umma_commit(empty_bar(stage));            // smem slot reusable once the MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(tfull_bar(acc));                // accumulator complete
          acc ^= 1u; if (acc == 0) acc_phase ^= 1u;
        }
        if (++vt == p.num_vt) { vt = 0; ++rg; }
      }
    }
  } else {
    // ============================ epilogue ================================
    const int quarter = warp & 3;                      // TMEM lanes this warp may read
    const int row_in_tile = quarter * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    int rg = rg_first, vt = vt_first;
    RowState st;
    bool open = false;
    float rs = 1.f, a = kLog2e;
    int lab_local = -1;
    long long row = 0;
    int slot = 0;
    uint2* warp_buf = nullptr;
    for (long long j = j0; j < j1; ++j) {
      const int rb = rg * p.g + member;
      const bool last_of_rg = (vt == p.num_vt - 1) || (j == j1 - 1);
      if (rb < p.num_rb) {
        if (!open) {
          slot = blockIdx.x * p.max_seg + (rg - rg_first);
          uint2* slot_buf = p.sv.cand + (size_t)slot * kBlockM * kCandCap;
          st.reset(slot_buf + (size_t)row_in_tile * kCandCap);
          warp_buf = slot_buf + (size_t)(quarter * 32) * kCandCap;
          row = (long long)rb * kBlockM + row_in_tile;
          rs = ((row < p.Q && p.inv_q) ? p.inv_q[row] : 1.f) * p.scale;
          a = rs * kLog2e;
          lab_local = -1;
          if (p.labels && row < p.Q) {
            const long long lg = p.labels[row];
            const long long l = lg - p.index_base;
            if (lg != -100 && l >= 0 && l < p.V) lab_local = (int)l;
          }
          open = true;
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kBlockN;
        const int tile_col0 = vt * kBlockN;
#pragma unroll 1
        for (int c = 0; c < kBlockN / kChunk; ++c) {
          const int col0 = tile_col0 + c * kChunk;
          if (col0 >= p.V) break;                       // warp-uniform
          float y[kChunk];
          __syncwarp();
          tmem_ld_32x32(taddr + c * kChunk, y);
          const int n_valid = min(kChunk, p.V - col0);
          if (p.inv_t) {
            if (n_valid == kChunk) {
              const float4* cs4 = reinterpret_cast<const float4*>(p.inv_t + col0);
#pragma unroll
              for (int i = 0; i < kChunk / 4; ++i) {
                const float4 cs = __ldg(cs4 + i);
                y[4 * i] *= cs.x; y[4 * i + 1] *= cs.y; y[4 * i + 2] *= cs.z; y[4 * i + 3] *= cs.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < kChunk; ++i) y[i] *= (i < n_valid) ? __ldg(p.inv_t + col0 + i) : 1.f;
            }
          }
          if (p.dbg_scores && row < p.Q) {
#pragma unroll
            for (int i = 0; i < kChunk; ++i)
              if (i < n_valid) p.dbg_scores[(size_t)row * p.V + col0 + i] = y[i] * rs;
          }
          if (n_valid == kChunk) row_process_chunk<false>(st, y, col0, kChunk, a, lab_local);
          else row_process_chunk<true>(st, y, col0, n_valid, a, lab_local);
          __syncwarp();
          warp_compact_rows(st, p.k, warp_buf, lane);
        }
        // hand the accumulator stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        acc ^= 1u; if (acc == 0) acc_phase ^= 1u;
        if (last_of_rg) {
          row_flush(st, rs, p.sv.cnt + (size_t)slot * kBlockM + row_in_tile,
                    p.sv.stats + (size_t)slot * kBlockM + row_in_tile);
          open = false;
        }
      }
      if (++vt == p.num_vt) { vt = 0; ++rg; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

TcSchedule make_tc_schedule(int64_t Q, int64_t V, int64_t D, int sm_count, int force_ctas,
                            int force_g) {
  TcSchedule s{};
  s.num_rb = (int)((Q + kBlockM - 1) / kBlockM);
  s.num_vt = (int)((V + kBlockN - 1) / kBlockN);
  s.num_kb = (int)((D + kBlockK - 1) / kBlockK);
  const int ctas = (force_ctas > 0) ? std::min(force_ctas, sm_count) : sm_count;
  // Cost model per candidate g: tensor time ~ jobs per group; HBM time ~ one table pass per
  // row-group (members of a group share tiles through L2; groups do not).
  const double t_tile = 2.0 * kBlockM * kBlockN * (double)s.num_kb * kBlockK / 9.0e12;  // s, per-SM ~9 TF/s
  const double table_bytes = (double)V * (double)D * 2.0;
  double best = 1e300;
  int best_g = 1;
  const int gmax = std::max(1, std::min(s.num_rb, ctas));
  for (int g = 1; g <= gmax; ++g) {
    const int ng = ctas / g;
    if (ng < 1) break;
    const long long num_rg = (s.num_rb + g - 1) / g;
    const long long total = num_rg * s.num_vt;
    const long long jpg = (total + ng - 1) / ng;
    const double t_mma = (double)jpg * t_tile;
    const double t_hbm = (double)num_rg * table_bytes / 5.5e12;
    const double cost = std::max(t_mma, t_hbm) + 0.15 * t_hbm;
    if (cost < best * 0.999 || (cost <= best * 1.001 && g > best_g)) { best = std::min(best, cost); best_g = g; }
  }
  s.g = (force_g > 0) ? std::min(force_g, gmax) : best_g;
  s.num_groups = std::max(1, ctas / s.g);
  s.num_rg = (s.num_rb + s.g - 1) / s.g;
  s.total_jobs = (long long)s.num_rg * s.num_vt;
  s.jpg = (int)((s.total_jobs + s.num_groups - 1) / s.num_groups);
  if (s.jpg < 1) s.jpg = 1;
  s.max_seg = (s.jpg + s.num_vt - 1) / s.num_vt + 1;
  s.grid = s.num_groups * s.g;
  return s;
}

Workspace carve_workspace(void* base, int nslots) {
  Workspace w{};
  w.nslots = nslots;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_cand = take((size_t)nslots * kBlockM * kCandCap * sizeof(uint2));
  const size_t o_cnt = take((size_t)nslots * kBlockM * sizeof(int));
  const size_t o_stats = take((size_t)nslots * kBlockM * sizeof(float4));
  w.bytes = off;
  if (base) {
    uint8_t* b = (uint8_t*)base;
    w.sv.cand = (uint2*)(b + o_cand);
    w.sv.cnt = (int*)(b + o_cnt);
    w.sv.stats = (float4*)(b + o_stats);
  }
  return w;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] (pitch ld elements) -> tiles of box_rows x 64, SWIZZLE_128B
static bool make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_scan_tc(const ScanArgs& a, const TcSchedule& sch, const SlotView& sv,
                           cudaStream_t s, char* err, size_t errlen) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kTcSmemBytes);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute(smem=%u)", kTcSmemBytes); return e; }
    attr_set = true;
  }
  CUtensorMap tm_q, tm_t;
  if (!make_tmap(&tm_q, a.q, a.Q, a.D, a.ldq, kBlockM) ||
      !make_tmap(&tm_t, a.table, a.V, a.D, a.ldt, kBlockN)) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (Q=%lld V=%lld D=%lld ldq=%lld ldt=%lld)",
             (long long)a.Q, (long long)a.V, (long long)a.D, (long long)a.ldq, (long long)a.ldt);
    return cudaErrorInvalidValue;
  }
  TcParams p{};
  p.Q = (int)a.Q; p.V = (int)a.V; p.D = (int)a.D; p.k = a.k;
  p.num_rb = sch.num_rb; p.num_vt = sch.num_vt; p.num_kb = sch.num_kb;
  p.g = sch.g; p.jpg = sch.jpg; p.max_seg = sch.max_seg; p.total_jobs = sch.total_jobs;
  p.inv_q = a.inv_q; p.inv_t = a.inv_t; p.scale = a.scale;
  p.index_base = a.index_base; p.labels = (const long long*)a.labels;
  p.sv = sv; p.dbg_scores = a.dbg_scores;
  scan_tc_kernel<<<sch.grid, kTcThreads, kTcSmemBytes, s>>>(tm_q, tm_t, p);
  return cudaGetLastError();
}

}  // namespace mcl
