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
#include <cuda.h>
#include <stdio.h>
#include <algorithm>
#include <atomic>
#include "rowstate.cuh"
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
};

// kCS = CTAs per cluster.  kCS = 1: every CTA multiplies its own 128 x 256 tile
// (cta_group::1, 48 KB of operands per K slice, 4 stages).  kCS = 2: two consecutive members of
// a group (same table tiles, different query row blocks) form a CTA pair and ONE
// tcgen05.mma.cta_group::2 (M = 256) issued by the even CTA multiplies both row blocks: each
// CTA stages only its own A tile and HALF of the B tile (32 KB per K slice, 6 stages) and the
// tensor core reads the B halves out of both CTAs' shared memory.  That cuts shared-memory
// traffic (TMA writes + MMA reads) by a third -- the single-CTA tile is bound by it -- and each
// CTA still finds its own 128 rows x 256 columns of accumulators in its own TMEM, so the
// epilogue is identical.
template <int kCS, bool kCap>
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
        int* ctr = sg.sync >= 0 ? p.sync_ctr + sg.sync : nullptr;
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
              for (int spin = 0; *c < sg.members && spin < (1 << 22); ++spin) __nanosleep(256);
            }
          }
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait_backoff(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * kStageBytes;
            if (kCS == 1) {
              mbar_expect_tx(full_bar(stage), kStageBytes);
              tma_load_2d(sa, &tm_q, full_bar(stage), kb * kBlockK, rb * kBlockM, p.pol_q);
              tma_load_2d(sa + kABytes, &tm_t, full_bar(stage), kb * kBlockK, vt * kBlockN, p.pol_t);
            } else {
              // both CTAs' bytes are counted on the leader's barrier, which its producer arms
              if (leader) mbar_expect_tx(full_bar(stage), 2 * kStageBytes);
              tma_load_2d_2sm(sa, &tm_q, full_bar(stage), kb * kBlockK, rb * kBlockM, p.pol_q);
              tma_load_2d_2sm(sa + kABytes, &tm_t, full_bar(stage), kb * kBlockK,
                              vt * kBlockN + (int)(crank * kSliceRows), p.pol_t);
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
      const int c0 = vt_ * kBlockN + half * kHalfN + 4 * lane;
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
      if (row >= p.Q || p.small_scores) st.tau = INFINITY;   // padding rows (and the small-batch path) never append
      rs = ((row < p.Q && p.inv_q) ? p.inv_q[row] : 1.f) * p.scale;
      a = (kCap ? p.softcap : rs) * kLog2e;
      const float rc = kCap ? rs / p.softcap : 0.f;
      lab_local = -1;
      if (p.labels && row < p.Q) {
        const long long lg = p.labels[row];
        const long long l = lg - p.index_base;
        if (lg != -100 && l >= 0 && l < p.V) lab_local = (int)l;
      }
      // threshold word shared by all workers that scan this query row (other table tiles)
      uint32_t* tau_pub = p.tau_shared ? p.tau_shared + row : nullptr;
      uint32_t tau_seen = tau_pub ? __ldcg(tau_pub) : 0u;   // later segments start below a tight bound
      // joint threshold of the row's slots (rowstate.cuh): read by every segment, published by
      // the column halves of the wave's first groups
      const int ng2 = p.joint ? sg.ng2 : 0;
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
        const int tile_col0 = vt * kBlockN + half * kHalfN;   // first column of this warp's half
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
          if (n_valid == kChunk) row_process_chunk<false, kCap>(st, y, col0, kChunk, a, lab_local, rc);
          else row_process_chunk<true, kCap>(st, y, col0, n_valid, a, lab_local, rc);
          __syncwarp();
          // (nothing follows the last chunk of the slot: leave its buffer to the merge)
          if (!(vt + 1 == vt1 && c + 1 == nch))
            warp_compact_rows(st, p.k, warp_buf, lane, tau_pub, joint_warp, joint_m);
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

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

// One node of a wave's plan (plan.h): R row units x tiles [a, T) on W >= R workers, and,
// recursively, the nodes below it.  Returns the tiles on the critical path.  Every segment
// restarts the top-k filter of its rows (threshold -inf -> bursts of buffer compactions while
// it tightens): `cold` tiles' worth of time at the start of a wave, `warm` once the row's other
// slots have published a threshold.  Below the top node a row unit gets at most kMaxSubGroups
// more slots per node: every slot costs workspace, a restart and merge work.
constexpr int kMaxSubGroups = 4;
struct NodePlan { long crit; long tiles_hbm; int n; Node nd[kMaxNodes]; };
static NodePlan plan_node(int R, int a, int T, int W, int depth, int r0, int w0, bool tails, int cold,
                          int warm) {
  const int len = T - a;
  const int start = depth == 0 ? cold : warm;
  NodePlan best{};
  Node nd{};
  nd.r0 = r0; nd.R = R; nd.a = a; nd.w0 = w0;
  int nfull = std::max(1, std::min(W / R, len));
  if (depth > 0) nfull = std::min(nfull, kMaxSubGroups);
  const int wr = W - nfull * R;
  // even split, the other workers idle
  nd.tpc = (len + nfull - 1) / nfull;
  nd.nfull = (len + nd.tpc - 1) / nd.tpc;       // no group without tiles
  nd.t0 = T; nd.wr = 0; nd.passes = 0;
  best.crit = nd.tpc + start; best.tiles_hbm = len; best.n = 1; best.nd[0] = nd;
  if (!tails || wr <= 0 || nfull >= len || depth + 1 >= kMaxNodes || W / R > nfull) return best;
  // groups take tpc tiles each, the wr tail workers the rest of the tiles for all R row units.
  // Perfect balance is tpc = R*len/W.
  const int hi = len / nfull;
  const int centre = (int)((long long)R * len / W);
  for (int tpc = std::max(1, centre - 2); tpc <= std::min(hi, centre + 2); ++tpc) {
    const int t0 = a + nfull * tpc, tl = T - t0;
    if (tl < 1) continue;
    const int passes = R / wr, rest = R - passes * wr;
    long tail = (long)passes * (tl + warm);
    NodePlan sub{};
    if (rest > 0) {
      sub = plan_node(rest, t0, T, wr, depth + 1, r0 + passes * wr, w0 + nfull * R, tails, cold, warm);
      tail += sub.crit;
    }
    const long crit = std::max<long>(tpc + start, tail);
    if (crit < best.crit) {
      nd.nfull = nfull; nd.tpc = tpc; nd.t0 = t0; nd.wr = wr; nd.passes = passes;
      best.crit = crit;
      best.tiles_hbm = (long)nfull * tpc + (long)passes * tl + sub.tiles_hbm;
      best.n = 1 + sub.n;
      best.nd[0] = nd;
      for (int i = 0; i < sub.n; ++i) best.nd[1 + i] = sub.nd[i];
    }
  }
  return best;
}

static void finish_chain(const NodePlan& np, int win, Chain* ch, int* nsync, int* workers, int T) {
  ch->n = np.n;
  int jbase = 0, sync = 0;
  for (int i = 0; i < np.n; ++i) {
    Node nd = np.nd[i];
    nd.jbase = jbase;
    nd.sync0 = sync;
    nd.nwin_g = (nd.tpc + win - 1) / win;
    nd.nwin_t = nd.wr ? (T - nd.t0 + win - 1) / win : 0;
    jbase += nd.nfull;
    sync += nd.nfull * nd.nwin_g + nd.passes * nd.nwin_t;
    *workers = std::max(*workers, nd.w0 + nd.nfull * nd.R + nd.wr);
    ch->nd[i] = nd;
  }
  *nsync = std::max(*nsync, sync);
}

static TcPlan make_tc_plan_uncached(int64_t Q, int64_t V, int64_t D, int sm_count, const PlanKnobs& kn) {
  TcPlan p{};
  p.num_rb = (int)((Q + kBlockM - 1) / kBlockM);
  p.num_vt = (int)((V + kBlockN - 1) / kBlockN);
  p.num_kb = (int)((D + kBlockK - 1) / kBlockK);
  const int ctas = (kn.ctas > 0) ? std::min(kn.ctas, sm_count) : sm_count;
  // CTA pairs (one cta_group::2 MMA over two row blocks) whenever there are two row blocks
  p.cs = (p.num_rb >= 2 && ctas >= 2 && kn.cluster != 1) ? 2 : 1;
  const int W = std::max(1, ctas / p.cs);
  const int T = p.num_vt;
  p.ru = (p.num_rb + p.cs - 1) / p.cs;
  // Cost model (fitted to B200 measurements, profiles/): time = tensor time of the critical
  // path + HBM traffic / bandwidth.  Traffic = the table tiles each group and tail pass
  // streams, plus the share of the L2-level operand traffic that misses once the working set
  // (gu query units + a few table streams of ~4 tiles) outgrows the usable L2.  (The kernel
  // runs at the board's power cap: HBM and L2 traffic cost clock, not only time.)
  const double a_unit = (double)p.cs * kBlockM * p.num_kb * kBlockK * 2.0;
  const double b_tile = (double)kBlockN * p.num_kb * kBlockK * 2.0;
  const double t_tile = 2.0 * kBlockM * kBlockN * (double)p.num_kb * kBlockK / 9.0e12;
  const double l2_level = (double)p.num_rb * T * (a_unit / p.cs + b_tile);
  // Usable L2 and the price of a DRAM byte, fitted to the ncu captures and option sweeps in
  // profiles/ (C3 gu=32: 6.4-7.0 GB read; C4 gu=37 / 24: 65.8 / 41.9 GB, 54.5 / 52.1 ms).  The
  // kernel is tensor bound, so DRAM bytes cost board power (clock), not bandwidth.
  const double l2_cap = 110.0e6, miss_weight = 0.3, dram_price = 1.0 / 1.2e13;
  // candidate buffers are hot too: ~1 KB per query row, slot and column half
  const double buf_slot = (double)p.cs * 2.0 * kBlockM * 1024.0;
  // restart of the top-k filter per segment, in tiles (see plan_node); option 8 scales it
  const int cold = std::max(1, (int)(1.5e-4 * kn.seg_penalty / t_tile + 0.5));
  const int warm = std::max(1, (int)(0.4e-4 * kn.seg_penalty / t_tile + 0.5));
  const bool tails = kn.leftover != 0;
  double best = 1e300;
  int best_gu = 1;
  const int gmax = std::max(1, std::min(p.ru, W));
  for (int gu = 1; gu <= gmax; ++gu) {
    const int waves = (p.ru + gu - 1) / gu;
    const NodePlan full = plan_node(gu, 0, T, W, 0, 0, 0, tails, cold, warm);
    const NodePlan last = plan_node(p.ru - (waves - 1) * gu, 0, T, W, 0, 0, 0, tails, cold, warm);
    const double tiles = (double)(waves - 1) * full.crit + last.crit;
    const NodePlan& wide = waves > 1 ? full : last;
    const int streams = wide.nd[0].nfull;
    const int slots = wide.nd[0].nfull + (wide.nd[0].wr ? 1 : 0);
    const double ws = gu * (a_unit + slots * buf_slot) + streams * 4.0 * b_tile;
    const double miss = ws > l2_cap ? 1.0 - l2_cap / ws : 0.0;
    const double dram = ((double)(waves - 1) * full.tiles_hbm + last.tiles_hbm) * b_tile +
                        (double)Q * D * 2.0 + miss_weight * miss * l2_level;
    double cost = tiles * t_tile + dram * dram_price;
    if (p.cs == 1) cost *= 1.08;         // no CTA pairing: 48 instead of 32 KB per K slice from L2
    if (cost < best) { best = cost; best_gu = gu; }
  }
  p.gu = (kn.gu > 0) ? std::min(kn.gu, gmax) : best_gu;
  p.waves = (p.ru + p.gu - 1) / p.gu;
  const NodePlan full = plan_node(p.gu, 0, T, W, 0, 0, 0, tails, cold, warm);
  const NodePlan last = plan_node(p.ru - (p.waves - 1) * p.gu, 0, T, W, 0, 0, 0, tails, cold, warm);
  // window of the drift bound: ~3 windows of every stream must fit in the L2 share left
  // after the resident query units
  const int streams = std::max(last.nd[0].nfull, p.waves > 1 ? full.nd[0].nfull : 0);
  const double l2_stream = std::max(8.0e6, 90.0e6 - p.gu * a_unit);
  p.win = (int)std::max(1.0, std::min(16.0, l2_stream / (3.0 * b_tile * streams)));
  if (kn.win > 0) p.win = kn.win;
  p.nsync = 0;
  p.workers = 0;
  finish_chain(last, p.win, &p.last, &p.nsync, &p.workers, T);
  if (p.waves > 1) finish_chain(full, p.win, &p.full, &p.nsync, &p.workers, T);
  else p.full = p.last;
  p.S = 1;
  for (int w = (p.waves > 1 ? 0 : 1); w < 2; ++w) {   // slot stride: the widest row unit
    const int u0 = w ? (p.waves - 1) * p.gu : 0;
    const int rw = w ? p.ru - u0 : p.gu;
    for (int u = 0; u < rw; ++u) p.S = std::max(p.S, plan_unit_slots(p, u0 + u));
  }
  return p;
}

// Planning tries every wave size with a small recursive search: cached, because the same few
// shapes are scanned over and over.
TcPlan make_tc_plan(int64_t Q, int64_t V, int64_t D, int sm_count, const PlanKnobs& kn) {
  struct Key { int64_t Q, V, D; int sm; PlanKnobs kn; };
  struct Entry { Key k; TcPlan p; bool ok; };
  thread_local Entry cache[16] = {};
  thread_local int next = 0;
  auto same = [&](const Key& k) {
    return k.Q == Q && k.V == V && k.D == D && k.sm == sm_count && k.kn.ctas == kn.ctas && k.kn.gu == kn.gu &&
           k.kn.cluster == kn.cluster && k.kn.leftover == kn.leftover && k.kn.seg_penalty == kn.seg_penalty &&
           k.kn.win == kn.win;
  };
  for (const Entry& e : cache)
    if (e.ok && same(e.k)) return e.p;
  Entry e{Key{Q, V, D, sm_count, kn}, make_tc_plan_uncached(Q, V, D, sm_count, kn), true};
  cache[next] = e;
  next = (next + 1) % 16;
  return e.p;
}

Workspace carve_workspace(void* base, int nslots, int num_rb, int nctr, size_t extra_bytes) {
  Workspace w{};
  w.nslots = nslots;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_time = take(2 * 1024 * sizeof(unsigned long long));   // always at offset 0
  const size_t o_tau = take((size_t)num_rb * kBlockM * sizeof(uint32_t));
  const size_t o_ctr = take((size_t)nctr * sizeof(int));
  const size_t o_joint = take((size_t)num_rb * kBlockM * kJointWords * sizeof(uint32_t));
  const size_t o_cand = take((size_t)nslots * kBlockM * kCandCap * sizeof(uint2));
  const size_t o_cnt = take((size_t)nslots * kBlockM * sizeof(int2));
  const size_t o_stats = take((size_t)nslots * kBlockM * sizeof(float4));
  const size_t o_extra = take(extra_bytes);
  w.bytes = off;
  w.zero_bytes = o_joint + (((size_t)num_rb * kBlockM * kJointWords * sizeof(uint32_t) + 255) & ~(size_t)255) - o_tau;
  if (base) {
    uint8_t* b = (uint8_t*)base;
    w.timing = (void*)(b + o_time);
    w.tau_shared = (void*)(b + o_tau);
    w.sync_ctr = (void*)(b + o_ctr);
    w.joint = (void*)(b + o_joint);
    w.sv.cand = (uint2*)(b + o_cand);
    w.sv.cnt = (int2*)(b + o_cnt);
    w.sv.stats = (float4*)(b + o_stats);
    w.extra = (void*)(b + o_extra);
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

// 2-D bf16 row-major [rows, cols] (pitch ld elements) -> tiles of box_rows x 64, SWIZZLE_128B.
// Descriptors depend only on (base, shape, pitch, box); the last few are cached per thread
// because the same table (and query buffer) is scanned over and over.
static bool make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows) {
  struct Entry { const void* base; int64_t rows, cols, ld; int box; CUtensorMap tm; bool ok; };
  thread_local Entry cache[8] = {};
  thread_local int next = 0;
  for (const Entry& e : cache)
    if (e.ok && e.base == base && e.rows == rows && e.cols == cols && e.ld == ld && e.box == box_rows) {
      *tm = e.tm;
      return true;
    }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  cache[next] = Entry{base, rows, cols, ld, box_rows, *tm, true};
  next = (next + 1) % 8;
  return true;
}

cudaError_t launch_scan_tc(const ScanArgs& a, const TcPlan& plan, const SlotView& sv,
                           cudaStream_t s, char* err, size_t errlen) {
  // the opt-in to > 48 KB of dynamic shared memory is per device
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev].load()) {
    cudaError_t e = cudaSuccess;
    const void* kernels[4] = {(const void*)scan_tc_kernel<1, false>, (const void*)scan_tc_kernel<2, false>,
                              (const void*)scan_tc_kernel<1, true>, (const void*)scan_tc_kernel<2, true>};
    for (const void* kfn : kernels)
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute(smem=%u)", kTcSmemBytes); return e; }
    attr_set[dev].store(true);
  }
  const int cs = plan.cs;
  CUtensorMap tm_q, tm_t;
  if (!make_tmap(&tm_q, a.q, a.Q, a.D, a.ldq, kBlockM) ||
      !make_tmap(&tm_t, a.table, a.V, a.D, a.ldt, kBlockN / cs)) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (Q=%lld V=%lld D=%lld ldq=%lld ldt=%lld)",
             (long long)a.Q, (long long)a.V, (long long)a.D, (long long)a.ldq, (long long)a.ldt);
    return cudaErrorInvalidValue;
  }
  TcParams p{};
  p.Q = (int)a.Q; p.V = (int)a.V; p.D = (int)a.D; p.k = a.k;
  p.plan = plan;
  p.inv_q = a.inv_q; p.inv_t = a.inv_t; p.scale = a.scale;
  p.index_base = a.index_base; p.labels = (const long long*)a.labels;
  p.sv = sv; p.dbg_scores = a.dbg_scores; p.timing = (unsigned long long*)a.timing;
  p.tau_shared = (uint32_t*)a.tau_shared;
  p.sync_ctr = (int*)a.sync_ctr;
  p.joint = (uint32_t*)a.joint;
  p.softcap = a.softcap;
  p.small_scores = a.small_scores; p.small_ld = (int)a.small_ld;
  p.pol_q = (a.l2_mode & 1) ? kL2EvictLast : kL2EvictNormal;
  p.pol_t = (a.l2_mode & 2) ? kL2EvictFirst : ((a.l2_mode & 4) ? kL2EvictLast : kL2EvictNormal);
  const bool cap = a.softcap > 0.f;
  // The drift bound makes CTAs wait for one another, so all of them should be resident at
  // once: grid <= SM count with one CTA per SM (192 KB of shared memory) gives that on an
  // otherwise idle GPU, and the wait is bounded in case foreign work holds SMs.  (A cooperative
  // launch would enforce it, but costs ~40 us of launch latency and cannot be combined with
  // clusters.)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan_grid(plan));
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = kTcSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  if (cs == 2) {
    cfg.attrs = attr + 1;
    cfg.numAttrs = 1;
    return cap ? cudaLaunchKernelEx(&cfg, scan_tc_kernel<2, true>, tm_q, tm_t, p)
               : cudaLaunchKernelEx(&cfg, scan_tc_kernel<2, false>, tm_q, tm_t, p);
  }
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  return cap ? cudaLaunchKernelEx(&cfg, scan_tc_kernel<1, true>, tm_q, tm_t, p)
             : cudaLaunchKernelEx(&cfg, scan_tc_kernel<1, false>, tm_q, tm_t, p);
}

}  // namespace mcl
