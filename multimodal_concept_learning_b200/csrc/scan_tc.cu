// Host side of the tcgen05 scan: tile planner, workspace carving, tensor maps, launch.
// The kernel template lives in scan_tc_kernel.cuh; its epilogue modes are instantiated in
// scan_tc_m{0,1,2}.cu.
#include <stdio.h>
#include <algorithm>
#include <atomic>
#include "scan_tc_kernel.cuh"

namespace mcl {

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
  // (kn.filter: 0 = the top-k filter starts cold in every slot; 1 = thresholds seeded by the
  // pre-pass: a segment restarts almost warm; 2 = no filter at all (k = 1 epilogue, seed pass))
  const double cold_s = kn.filter == 0 ? 1.5e-4 : (kn.filter == 1 ? 0.2e-4 : 0.0);
  const double warm_s = kn.filter == 0 ? 0.4e-4 : (kn.filter == 1 ? 0.1e-4 : 0.0);
  const int cold = std::max(1, (int)(cold_s * kn.seg_penalty / t_tile + 0.5));
  const int warm = std::max(1, (int)(warm_s * kn.seg_penalty / t_tile + 0.5));
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
           k.kn.cluster == kn.cluster && k.kn.leftover == kn.leftover && k.kn.seg_penalty == kn.seg_penalty && k.kn.filter == kn.filter &&
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
  const size_t o_qn = take((size_t)num_rb * kBlockM * sizeof(float));
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
    w.inv_q = (float*)(b + o_qn);
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
bool make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld,
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

// One mapped (zero-copy) host word per device counts the drift waits that timed out: written by
// the kernel with a system-scope atomic on the rare time-out only, read by the host without a
// copy or a synchronisation.
static unsigned long long* g_drift_words[64];
static unsigned long long* drift_word(int dev) {
  static std::atomic<bool> tried[64];
  if (dev < 0 || dev >= 64) return nullptr;
  if (!tried[dev].exchange(true)) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      *(volatile unsigned long long*)h = 0ull;
      g_drift_words[dev] = (unsigned long long*)h;
    } else {
      cudaGetLastError();
    }
  }
  return g_drift_words[dev];
}
long long drift_timeouts_total() {
  long long n = 0;
  for (auto* w : g_drift_words)
    if (w) n += (long long)*(volatile unsigned long long*)w;
  return n;
}

cudaError_t launch_scan_tc(const ScanArgs& a, const TcPlan& plan, const SlotView& sv,
                           cudaStream_t s, char* err, size_t errlen) {
  // the opt-in to > 48 KB of dynamic shared memory is per device
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev].load()) {
    cudaError_t e = tc_set_smem_attr_mode0();
    if (e == cudaSuccess) e = tc_set_smem_attr_mode1();
    if (e == cudaSuccess) e = tc_set_smem_attr_mode2();
    if (e == cudaSuccess) e = tc_set_smem_attr_mode3();
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute(smem=%u)", kTcSmemBytes); return e; }
    attr_set[dev].store(true);
  }
  const int cs = plan.cs;
  CUtensorMap tm_q, tm_t;
  if (!make_tmap_bf16(&tm_q, a.q, a.Q, a.D, a.ldq, kBlockM) ||
      !make_tmap_bf16(&tm_t, a.table, a.V, a.D, a.ldt, kBlockN / cs)) {
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
  p.tile_stride = a.tile_stride > 0 ? a.tile_stride : 1;
  p.seed_max = (uint32_t*)a.seed_max; p.seed_ld = (int)a.seed_ld;
  p.drift_timeouts = drift_word(dev);
  p.q_rows = a.qnorm_in_kernel ? (const __nv_bfloat16*)a.q : nullptr;
  p.ldq = a.ldq;
  p.inv_q_out = a.inv_q_out;
  p.clear_words = (uint32_t*)a.clear_words; p.n_clear = a.n_clear;
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
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = cs == 2 ? attr : nullptr;
  cfg.numAttrs = cs == 2 ? 1 : 0;
  p.p_out = (__nv_bfloat16*)a.p_out; p.ldp = a.ldp; p.p_rows = (int)a.p_rows;
  p.lse = a.lse; p.grad_loss = a.grad_loss; p.grad_coef = a.grad_coef;
  p.eps_over_v = a.eps_over_v; p.one_minus_eps = a.one_minus_eps;
  if (a.mode == kModeGrad) return tc_launch_mode3(&cfg, cs, cap, tm_q, tm_t, p);
  if (a.mode == kModeSeed) return tc_launch_mode2(&cfg, cs, false, tm_q, tm_t, p);
  if (a.mode == kModeTop1) return tc_launch_mode1(&cfg, cs, cap, tm_q, tm_t, p);
  return tc_launch_mode0(&cfg, cs, cap, tm_q, tm_t, p);
}

}  // namespace mcl
