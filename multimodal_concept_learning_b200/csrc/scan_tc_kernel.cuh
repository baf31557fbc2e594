// The product path: fused similarity scan on Blackwell tensor cores.
//
//   scores tile [128 queries x 256 table rows] = Q_tile (bf16, K-major) * T_tile^T
//   - operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) into an mbarrier ring of
//     192 KB (6 stages per CTA of a pair, 4 for a single CTA),
//   - multiplied by tcgen05.mma (kind::f16, K16; M256 N256 issued by ONE thread for a CTA pair,
//     M128 N256 for a single CTA),
//   - accumulated in TMEM (2 x 256 fp32 columns, double buffered),
//   - drained by eight epilogue warps with tcgen05.ld: thread = query row = TMEM lane, so the
//     online log-sum-exp, running sum, label pick-up and top-k filter are thread-private
//     (rowstate.cuh) and the [Q x V] score matrix never leaves the SM.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue.  A warp may only read the TMEM lane quarter warp_id % 4, so two warps
// share each quarter and split the tile's 256 columns in halves (own row state, own slot):
// two resident warps per scheduler hide each other's latencies.  (Separate top-k and LSE warps,
// 16 in all, were measured and are no faster: profiles/README.md.)
//
// Scheduling (plan.h): persistent workers (a worker = one CTA, or a CTA pair sharing one
// cta_group::2 MMA).  Row units are taken in waves of gu; inside a wave the workers form nfull
// groups of one worker per row unit that walk the same table tiles side by side -- so at any
// time the chip works on gu row units (their query tiles stay L2 resident) and streams a few
// table positions, each tile fetched from HBM once per group -- and the workers that do not
// fill another group share the tail tiles of all row units, so that no SM idles when the row
// blocks do not divide the SM count.  Every segment (worker, row unit, tile range) ends in a
// partial-result slot; merge.cu combines the slots of a row block.
#pragma once
#include <cuda.h>
#include "rowstate.cuh"
#include "rownorm.cuh"
#include "kernels.h"

namespace mcl {

constexpr int kEpiWarps = 8;                          // two per TMEM lane quarter: column halves
constexpr int kTcThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr uint32_t kBBytes = kBlockN * kBlockK * 2;   // 32 KB
constexpr uint32_t kRingBytes = 192 * 1024;           // operand ring: 4 x 48 KB, or 6 x 32 KB per CTA of a pair
constexpr uint32_t kTmemCols = 512;                   // 2 accumulator stages x 256 columns
constexpr uint32_t kBarBytes = 256;
constexpr uint32_t kCsBytes = kEpiWarps * 2 * (kBlockN / 2) * 4;  // per epilogue warp: 2 x 128 table-row scales
constexpr uint32_t kTcSmemBytes = kRingBytes + kBarBytes + kCsBytes + 1024;  // + align slack


struct TcParams {
  int Q, V, D, k;
  TcPlan plan;
  const float* inv_q;
  const float* inv_t;
  float scale;
  long long index_base;
  const long long* labels;
  SlotView sv;
  float* dbg_scores;
  unsigned long long* timing;   // nullable: [grid][2] globaltimer at CTA start / end
  uint32_t* tau_shared;         // [padded rows] order-preserving keys, zeroed before the launch
  int* sync_ctr;                // [plan_nctr] CTAs that started a window, zeroed
  uint32_t* joint;              // [padded rows][kJointWords] joint-threshold words, zeroed (nullable)
  float softcap;                // 0 = off; c > 0: logits are c*tanh(z/c) (kCap instantiation)
  unsigned long long pol_q, pol_t;   // L2 eviction priority of the query / table tile loads
  float* small_scores;          // small-batch path (select.cu): [Q][small_ld] scores, top-k filter off
  int small_ld;
  int tile_stride;              // seed mode: the plan's tile t is table tile t * tile_stride
  uint32_t* seed_max;           // seed mode: [chunks of the sample][seed_ld] keys of the chunk maxima
  int seed_ld;                  // padded query rows
  unsigned long long* drift_timeouts;   // mapped host word: drift waits that gave up (nullable)
  // small query batches (one row block, Q <= 64): the kernel computes 1/||q_row|| itself instead of
  // reading inv_q (one launch less where launches bound the step) and CTA 0 publishes the values
  const __nv_bfloat16* q_rows;  // nullable: off
  long long ldq;
  float* inv_q_out;             // [padded rows] (CTA 0 writes; the merge reads)
  uint32_t* clear_words;        // words CTA 0 zeroes for the kernel that FOLLOWS this one (nullable)
  int n_clear;
  // grad mode (backward of the fused cross-entropy): dL/dz tiles, bf16, [p_rows][ldp]
  __nv_bfloat16* p_out;
  long long ldp;
  int p_rows;
  const float* lse;             // [Q] log-sum-exp saved by the forward (natural log)
  const float* grad_loss;       // device scalar: dL/d(loss)
  float grad_coef;              // 1 / rows with a label (the mean's denominator)
  float eps_over_v, one_minus_eps;
};

// Epilogue modes (one template instantiation each):
//   kModeTopK  streaming top-k filter + statistics (the general path)
//   kModeTop1  k = 1: running (max, argmax) in registers, no candidate buffers (rowstate.cuh)
//   kModeSeed  threshold seeding pre-pass: scores of a SAMPLE of table tiles, of which only the
//              maximum of every 32-column chunk is kept, as an order-preserving key; the k-th
//              largest chunk maximum of a row is reached by k distinct table rows, i.e. it is a
//              lower bound of the row's k-th best score (seed_select_kernel writes it into the
//              shared threshold words the main scan starts from).  Epilogue-bound shapes
//              (D <= 1536) lose more time to the cold start of every slot's top-k filter
//              (threshold -inf: every score is a candidate) than this sample costs.
//   kModeGrad  backward of the fused cross-entropy: the scores are recomputed tile by tile and turned
//              into dL/dz = (exp(z - lse) - (1-eps) [col = label] - eps/V) * dL/dloss / n_valid with the
//              log-sum-exp the forward saved, rounded to bf16 and written as the operand of the two
//              gradient GEMMs (gemm_tc.cu).  With a soft-cap the factor dz'/dz = 1 - tanh^2 rides along.
constexpr int kModeTopK = 0, kModeTop1 = 1, kModeSeed = 2, kModeGrad = 3;

// kCS = CTAs per cluster.  kCS = 1: every CTA multiplies its own 128 x 256 tile
// (cta_group::1, 48 KB of operands per K slice, 4 stages).  kCS = 2: two consecutive members of
// a group (same table tiles, different query row blocks) form a CTA pair and ONE
// tcgen05.mma.cta_group::2 (M = 256) issued by the even CTA multiplies both row blocks: each
// CTA stages only its own A tile and HALF of the B tile (32 KB per K slice, 6 stages) and the
// tensor core reads the B halves out of both CTAs' shared memory.  That cuts shared-memory
// traffic (TMA writes + MMA reads) by a third -- the single-CTA tile is bound by it -- and each
// CTA still finds its own 128 rows x 256 columns of accumulators in its own TMEM, so the
// epilogue is identical.
template <int kCS, bool kCap, int kMode>
__global__ void __launch_bounds__(kTcThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_t,
               const TcParams p) {
  constexpr uint32_t kStageBytes = kABytes + kBBytes / kCS;   // per CTA
  constexpr uint32_t kStages = kRingBytes / kStageBytes;      // 4 or 6
  constexpr uint32_t kSliceRows = kBlockN / kCS;              // table rows of a tile this CTA stages
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms
  const uint32_t bar_base = smem_base + kRingBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](uint32_t a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](uint32_t a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kRingBytes + 8u * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long t_start = 0;
  if (p.timing && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_t);
  }
  if (p.clear_words && blockIdx.x == 0)     // e.g. the selection kernel's per-row arrival counters
    for (int i = threadIdx.x; i < p.n_clear; i += kTcThreads) p.clear_words[i] = 0u;
  if (warp == 1) {
    if (lane == 0) {
      // full: the (leader's) producer arms it; empty / tfull: one tcgen05.commit arrival;
      // tempty: every epilogue warp of every CTA whose accumulators the MMA overwrites
      for (uint32_t s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (uint32_t a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiWarps * kCS); }
      fence_barrier_init();
    }
    __syncwarp();
    if (kCS == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    else { tmem_alloc_2sm(tmem_slot, kTmemCols); tmem_relinquish_2sm(); }
  }
  tc_fence_before();
  __syncthreads();
  if (kCS > 1) cluster_sync_all();      // the peer's barriers and TMEM exist before they are used
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t crank = (kCS > 1) ? cluster_ctarank() : 0u;
  const bool leader = (crank == 0);

  // every role walks this worker's segments (plan.h) with its own iterator
  const int num_kb = p.plan.num_kb;
  const int tstride = (kMode == kModeSeed) ? p.tile_stride : 1;
  SegIter it;
  seg_iter_init(it, (int)blockIdx.x / kCS);
  Seg sg;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      while (seg_iter_next(p.plan, it, sg)) {
        const int rb = sg.unit * kCS + (int)crank;
        const int win = p.plan.win;
        // (a worker alone on its tiles has nobody to drift from; the seed pass is a few tiles long)
        int* ctr = (kMode != kModeSeed && kMode != kModeGrad && sg.sync >= 0 && sg.members > kCS) ? p.sync_ctr + sg.sync : nullptr;
        for (int vt = sg.vt0; vt < sg.vt1; ++vt) {
          // Drift bound: the members of a full group read the same table tiles and rely on L2
          // to fetch each from HBM once; nothing else keeps them together, and SMs differ in
          // speed by a few percent.  A member announces every window of `win` tiles it starts
          // and may not start window w before all members have started window w-2.
          if (ctr && (vt - sg.vt0) % win == 0) {
            const int w = (vt - sg.vt0) / win;
            atomicAdd(ctr + w, 1);
            if (w >= 2) {
              // bounded (~1 s): if a co-resident peer never shows up (SMs held by foreign work)
              // give up the L2 locality rather than hang
              const volatile int* c = ctr + (w - 2);
              int spin = 0;
              for (; *c < sg.members && spin < (1 << 22); ++spin) __nanosleep(256);
              // (counted where the host can see it without a copy: mcl_set_option(103, 0))
              if (spin == (1 << 22) && p.drift_timeouts) atomicAdd_system(p.drift_timeouts, 1ull);
            }
          }
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait_backoff(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * kStageBytes;
            if (kCS == 1) {
              mbar_expect_tx(full_bar(stage), kStageBytes);
              tma_load_2d(sa, &tm_q, full_bar(stage), kb * kBlockK, rb * kBlockM, p.pol_q);
              tma_load_2d(sa + kABytes, &tm_t, full_bar(stage), kb * kBlockK, vt * tstride * kBlockN, p.pol_t);
            } else {
              // both CTAs' bytes are counted on the leader's barrier, which its producer arms
              if (leader) mbar_expect_tx(full_bar(stage), 2 * kStageBytes);
              tma_load_2d_2sm(sa, &tm_q, full_bar(stage), kb * kBlockK, rb * kBlockM, p.pol_q);
              tma_load_2d_2sm(sa + kABytes, &tm_t, full_bar(stage), kb * kBlockK,
                              vt * tstride * kBlockN + (int)(crank * kSliceRows), p.pol_t);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * kCS, kBlockN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      while (seg_iter_next(p.plan, it, sg)) {
        const int ntile = sg.vt1 - sg.vt0;
        for (int t = 0; t < ntile; ++t) {
          mbar_wait_backoff(tempty_bar(acc), acc_phase ^ 1u);   // epilogue has drained this stage
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kBlockN;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * kStageBytes;
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = umma_desc_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) { // +32 B per K=16 step inside the 128 B row
              if (kCS == 1) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
              else umma_bf16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
            }
            // smem slot reusable once the MMAs retire (in both CTAs of a pair)
            if (kCS == 1) umma_commit(empty_bar(stage)); else umma_commit_2sm(empty_bar(stage));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          // accumulator complete (each CTA's epilogue waits on its own barrier)
          if (kCS == 1) umma_commit(tfull_bar(acc)); else umma_commit_2sm(tfull_bar(acc));
          acc ^= 1u; if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ============================ epilogue ================================
    const int quarter = warp & 3;                      // TMEM lanes this warp may read
    const int half = (warp - 2) >> 2;                  // column half of every tile
    const int row_in_tile = quarter * 32 + lane;
    uint32_t acc = 0, acc_phase = 0, tb = 0;
    RowState st;
    float rs = 1.f, a = kLog2e;
    int lab_local = -1;
    long long row = 0;
    int slot = 0;
    uint2* warp_buf = nullptr;
    // Per-table-row scales of a tile: global -> registers one tile ahead -> a warp-private
    // shared-memory copy (broadcast reads).  Warp-private so that the four epilogue warps
    // never wait for one another: a warp busy compacting must not stall the other three.
    constexpr int kHalfN = kBlockN / 2;
    float* cs_warp = reinterpret_cast<float*>(smem_gen + kRingBytes + kBarBytes) +
                     (warp - 2) * 2 * kHalfN;
    float4 cs_ra = make_float4(1.f, 1.f, 1.f, 1.f);
    int cs_vt = -1;
    auto load_cs = [&](int vt_) {                      // lane l: this half's columns 4l..4l+3
      const int c0 = vt_ * tstride * kBlockN + half * kHalfN + 4 * lane;
      if (c0 + 3 < p.V) cs_ra = __ldg(reinterpret_cast<const float4*>(p.inv_t + c0));
      else cs_ra = make_float4(c0 < p.V ? p.inv_t[c0] : 1.f, c0 + 1 < p.V ? p.inv_t[c0 + 1] : 1.f,
                               c0 + 2 < p.V ? p.inv_t[c0 + 2] : 1.f, 1.f);
      cs_vt = vt_;
    };
    Seg nx;
    bool have = seg_iter_next(p.plan, it, sg);
    while (have) {
      const bool have_next = seg_iter_next(p.plan, it, nx);
      const int vt0 = sg.vt0, vt1 = sg.vt1;
      const int rb = sg.unit * kCS + (int)crank;    // may be a padding row block inside a cluster
      slot = (rb * p.plan.S + sg.j) * 2 + half;
      uint2* slot_buf = p.sv.cand + (size_t)slot * kBlockM * kCandCap;
      st.reset(slot_buf + (size_t)row_in_tile * kCandCap);
      warp_buf = slot_buf + (size_t)(quarter * 32) * kCandCap;
      row = (long long)rb * kBlockM + row_in_tile;
      if (kMode == kModeTopK && (row >= p.Q || p.small_scores)) st.tau = INFINITY;   // padding rows (and the small-batch path) never append
      float invq = 1.f;
      if (p.q_rows) {
        // the warp's 32 rows one after the other, 32 lanes per row (rownorm.cuh: the same function,
        // hence the same bits, as row_inv_norm_kernel); hidden behind the first tile's loads
        const long long r0 = (long long)rb * kBlockM + quarter * 32;
        for (int i = 0; i < 32 && r0 + i < p.Q; i += 4) {      // four rows per memory round trip
          float v[4];
          const int nr = (int)(p.Q - (r0 + i) < 4 ? p.Q - (r0 + i) : 4);
          warp_rows_inv_norm<__nv_bfloat16, 4>(p.q_rows + (r0 + i) * p.ldq, p.ldq, nr, p.D, lane, v);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (lane == i + r) invq = v[r];
        }
        if (blockIdx.x == 0 && half == 0 && row < p.Q) p.inv_q_out[row] = invq;
      } else if (row < p.Q && p.inv_q) {
        invq = p.inv_q[row];
      }
      rs = invq * p.scale;
      float g_row = 0.f, lse2 = 0.f;                   // grad mode: dL/dloss / n_valid, lse * log2(e)
      if (kMode == kModeGrad && row < p.Q && (!p.labels || p.labels[row] != -100)) {
        g_row = __ldg(p.grad_loss) * p.grad_coef;
        lse2 = p.lse[row] * kLog2e;
      }
      a = (kCap ? p.softcap : rs) * kLog2e;
      const float rc = kCap ? rs / p.softcap : 0.f;
      lab_local = -1;
      if (p.labels && row < p.Q) {
        const long long lg = p.labels[row];
        const long long l = lg - p.index_base;
        if (lg != -100 && l >= 0 && l < p.V) lab_local = (int)l;
      }
      // threshold word shared by all workers that scan this query row (other table tiles)
      uint32_t* tau_pub = (kMode == kModeTopK && p.tau_shared) ? p.tau_shared + row : nullptr;
      uint32_t tau_seen = tau_pub ? __ldcg(tau_pub) : 0u;   // later segments start below a tight bound
      // joint threshold of the row's slots (rowstate.cuh): read by every segment, published by
      // the column halves of the wave's first groups
      const int ng2 = (kMode == kModeTopK && p.joint) ? sg.ng2 : 0;
      const uint32_t* joint_row = ng2 ? p.joint + (size_t)row * kJointWords : nullptr;
      uint32_t* joint_warp = (ng2 && sg.jw >= 0)
          ? p.joint + ((size_t)rb * kBlockM + quarter * 32) * kJointWords + sg.jw + half : nullptr;
      const int joint_m = ng2 ? (p.k + ng2 - 1) / ng2 : 0;
      uint4 ja = make_uint4(0u, 0u, 0u, 0u), jb = ja;
      if (joint_row) { ja = ld_cg_v4_pinned(joint_row); if (ng2 > 4) jb = ld_cg_v4_pinned(joint_row + 4); }
      const int next_vt0 = have_next ? nx.vt0 : -1;
      for (int vt = vt0; vt < vt1; ++vt) {
        // (both consume values requested one tile ago and request the next ones; the threshold
        // first, so that its use does not wait on the scoreboard of the loads issued just before)
        if (tau_pub) {
          row_apply_shared_tau(st, tau_seen);
          tau_seen = ld_cg_u32_pinned(tau_pub);
        }
        if (joint_row) {
          const uint32_t w[8] = {ja.x, ja.y, ja.z, ja.w, jb.x, jb.y, jb.z, jb.w};
          uint32_t jmin = 0xffffffffu;
          bool all_set = true;
#pragma unroll
          for (int i = 0; i < kJointWords; ++i)
            if (i < ng2) { all_set = all_set && w[i] != 0u; jmin = min(jmin, w[i]); }
          if (all_set) row_apply_shared_tau(st, jmin);
          ja = ld_cg_v4_pinned(joint_row);
          if (ng2 > 4) jb = ld_cg_v4_pinned(joint_row + 4);
        }
        if (p.inv_t) {
          if (cs_vt != vt) load_cs(vt);                // first tile of a run: exposed once
          reinterpret_cast<float4*>(cs_warp + tb * kHalfN)[lane] = cs_ra;
          __syncwarp();
          if (vt + 1 < vt1) load_cs(vt + 1); else if (next_vt0 >= 0) load_cs(next_vt0);
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kBlockN + half * kHalfN;
        const int tile_col0 = vt * tstride * kBlockN + half * kHalfN;   // first column of this warp's half
        const int nch = max(0, min(kHalfN / kChunk, (p.V - tile_col0 + kChunk - 1) / kChunk));
        const float* cs_tile = cs_warp + tb * kHalfN;

        auto release_acc = [&]() {                     // every tcgen05.ld of this tile has landed
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kCS == 1 || leader) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_remote(tempty_bar(acc), 0);   // the MMA issuer lives in the leader
          }
        };
        auto consume = [&](float (&y)[kChunk], int c) {
          const int col0 = tile_col0 + c * kChunk;
          const int n_valid = min(kChunk, p.V - col0);
          if (p.inv_t) {
            const float4* cs4 = reinterpret_cast<const float4*>(cs_tile + c * kChunk);
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) {
              const float4 cs = cs4[i];
              y[4 * i] *= cs.x; y[4 * i + 1] *= cs.y; y[4 * i + 2] *= cs.z; y[4 * i + 3] *= cs.w;
            }
          }
          if (p.dbg_scores && row < p.Q) {
#pragma unroll
            for (int i = 0; i < kChunk; ++i)
              if (i < n_valid)
                p.dbg_scores[(size_t)row * p.V + col0 + i] =
                    kCap ? p.softcap * tanhf(y[i] * rc) : y[i] * rs;
          }
          if (p.small_scores && row < p.Q) {             // 128 B per row and chunk, -inf past the table's end
            float4* dst = reinterpret_cast<float4*>(p.small_scores + (size_t)row * p.small_ld + col0);
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) {
              float z[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float v = kCap ? p.softcap * tanhf(y[4 * i + j] * rc) : y[4 * i + j] * rs;
                z[j] = (4 * i + j < n_valid) ? v : -INFINITY;
              }
              dst[i] = make_float4(z[0], z[1], z[2], z[3]);
            }
          }
          if (kMode == kModeGrad) {
            uint32_t packed[kChunk / 2];
#pragma unroll
            for (int i = 0; i < kChunk; i += 2) {
              float v[2];
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                float pr, d = 1.f;
                if (kCap) {
                  const float t = tanhf(y[i + j] * rc);
                  pr = ex2_fast(fmaf(t, a, -lse2));      // a = softcap * log2(e)
                  d = 1.f - t * t;
                } else {
                  pr = ex2_fast(fmaf(y[i + j], a, -lse2));
                }
                float x = pr - p.eps_over_v;
                if (col0 + i + j == lab_local) x -= p.one_minus_eps;
                v[j] = (i + j < n_valid) ? x * g_row * d : 0.f;
              }
              const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
              packed[i / 2] = *reinterpret_cast<const uint32_t*>(&h);
            }
            if (row < (long long)p.p_rows) {
              uint4* dst = reinterpret_cast<uint4*>(p.p_out + row * p.ldp + col0);
#pragma unroll
              for (int i = 0; i < kChunk / 8; ++i)
                dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
            }
          } else if (kMode == kModeSeed) {
            // sample tiles are whole tiles: only the chunk maximum is kept (as a key, coalesced
            // over the warp's 32 rows)
            float m8[kChunk / 4];
#pragma unroll
            for (int h = 0; h < kChunk / 4; ++h)
              m8[h] = fmaxf(fmaxf(y[4 * h], y[4 * h + 1]), fmaxf(y[4 * h + 2], y[4 * h + 3]));
            const float cm = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                                   fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
            const int chunk_id = vt * (kBlockN / kChunk) + half * (kHalfN / kChunk) + c;
            p.seed_max[(size_t)chunk_id * p.seed_ld + row] = f2key(cm);
          } else if (kMode == kModeTop1) {
            if (n_valid == kChunk) row_process_chunk_top1<false, kCap>(st, y, col0, kChunk, a, lab_local, rc);
            else row_process_chunk_top1<true, kCap>(st, y, col0, n_valid, a, lab_local, rc);
          } else {
            if (n_valid == kChunk) row_process_chunk<false, kCap>(st, y, col0, kChunk, a, lab_local, rc);
            else row_process_chunk<true, kCap>(st, y, col0, n_valid, a, lab_local, rc);
            __syncwarp();
            // (nothing follows the last chunk of the slot: leave its buffer to the merge)
            if (!(vt + 1 == vt1 && c + 1 == nch))
              warp_compact_rows(st, p.k, warp_buf, lane, tau_pub, joint_warp, joint_m);
          }
        };

        // TMEM -> registers one chunk at a time; the other epilogue warp on this scheduler
        // covers the load latency, and a single copy of the chunk code spares the I-cache
        float y[kChunk];
        if (nch == 0) release_acc();                   // this half lies beyond the table's end
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          __syncwarp();
          tmem_ld_issue(taddr + c * kChunk, y);
          tmem_ld_wait(y);
          if (c + 1 == nch) release_acc();
          consume(y, c);
        }
        acc ^= 1u; if (acc == 0) acc_phase ^= 1u;
        tb ^= 1u;
      }
      if (kMode == kModeTop1) row_flush_top1(st);
      if (kMode != kModeSeed && kMode != kModeGrad)
        row_flush(st, rs, kCap ? p.softcap : 0.f, p.sv.cnt + (size_t)slot * kBlockM + row_in_tile,
                  p.sv.stats + (size_t)slot * kBlockM + row_in_tile);
      sg = nx;
      have = have_next;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.timing && threadIdx.x == 0) {   // after the barrier: the epilogue warps are done too
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    p.timing[2 * blockIdx.x] = t_start;
    p.timing[2 * blockIdx.x + 1] = t1;
  }
  if (kCS > 1) cluster_sync_all();      // no CTA leaves while a peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (kCS == 1) tmem_dealloc(tmem_base, kTmemCols); else tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}


// Per-mode launchers: each epilogue mode is instantiated in a translation unit of its own
// (scan_tc_m{0,1,2}.cu) so that the variants compile in parallel.
cudaError_t tc_set_smem_attr_mode0();
cudaError_t tc_set_smem_attr_mode1();
cudaError_t tc_set_smem_attr_mode2();
cudaError_t tc_set_smem_attr_mode3();
cudaError_t tc_launch_mode0(const cudaLaunchConfig_t* cfg, int cs, bool cap, const CUtensorMap& tm_q,
                            const CUtensorMap& tm_t, const TcParams& p);
cudaError_t tc_launch_mode1(const cudaLaunchConfig_t* cfg, int cs, bool cap, const CUtensorMap& tm_q,
                            const CUtensorMap& tm_t, const TcParams& p);
cudaError_t tc_launch_mode2(const cudaLaunchConfig_t* cfg, int cs, bool cap, const CUtensorMap& tm_q,
                            const CUtensorMap& tm_t, const TcParams& p);
cudaError_t tc_launch_mode3(const cudaLaunchConfig_t* cfg, int cs, bool cap, const CUtensorMap& tm_q,
                            const CUtensorMap& tm_t, const TcParams& p);

}  // namespace mcl
