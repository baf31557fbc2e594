// extern "C" surface of libmcl_sm100.so (include/mcl.h).  Argument checking, workspace
// carving, kernel sequencing on the caller's stream, error strings, and the NCCL
// communicator (resolved with dlopen so the library loads on machines without NCCL).
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include "../../include/mcl.h"
#include "kernels.h"

using namespace mcl;
namespace mcl { extern std::atomic<int> g_gather_variant, g_merge_variant; }

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<long long> g_opt_ctas{0}, g_opt_g{0}, g_opt_simt{0}, g_opt_timing{0}, g_opt_cluster{0}, g_opt_allgather{0}, g_opt_phase{0};
std::atomic<long long> g_opt_leftover{1}, g_opt_segpen{1}, g_opt_l2{0}, g_opt_win{0}, g_opt_nosmall{0}, g_opt_joint{0};
std::atomic<long long> g_opt_noseed{0}, g_opt_notop1{0}, g_opt_filter{0}, g_opt_noqnorm{0}, g_opt_nopanel{0};

PlanKnobs knobs() {
  PlanKnobs k;
  k.ctas = (int)g_opt_ctas.load(); k.gu = (int)g_opt_g.load(); k.cluster = (int)g_opt_cluster.load();
  k.leftover = (int)g_opt_leftover.load(); k.seg_penalty = (int)g_opt_segpen.load(); k.win = (int)g_opt_win.load();
  k.filter = (int)g_opt_filter.load();   // plan introspection only: scans pick it per call (scan_layout)
  return k;
}
float g_phase_ms[3] = {0.f, 0.f, 0.f};   // last scan: memset, scan kernel, merge kernel (option 6)

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(MCL_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DevInfo { int sm = 0, major = 0, minor = 0; bool ok = false; };
bool dev_info(DevInfo* out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  static std::mutex mu;
  static DevInfo cache[64];
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 0 || dev >= 64) return false;
  if (!cache[dev].ok) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return false;
    cache[dev].sm = p.multiProcessorCount;
    cache[dev].major = p.major;
    cache[dev].minor = p.minor;
    cache[dev].ok = true;
  }
  *out = cache[dev];
  return true;
}
int require_sm100(DevInfo* d) {
  if (!dev_info(d)) {
    cudaGetLastError();
    return fail(MCL_ERR_CUDA, "no usable CUDA device (libmcl_sm100 has no CPU fallback)");
  }
  if (d->major != 10)
    return fail(MCL_ERR_UNSUPPORTED_ARCH, "device is sm_%d%d; libmcl_sm100 is built for sm_100a only",
                d->major, d->minor);
  return MCL_OK;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }
size_t elt_size(int dtype) { return dtype == MCL_DTYPE_BF16 ? 2 : 4; }

int simt_nsplit(int64_t Q, int64_t V, int sm) {
  const int num_rb = (int)((Q + kBlockM - 1) / kBlockM);
  const int num_chunks = (int)((V + kChunk - 1) / kChunk);
  int ns = (4 * sm + num_rb - 1) / num_rb;
  if (ns > num_chunks) ns = num_chunks;
  if (ns < 1) ns = 1;
  // no empty splits: chunks_per_split = ceil(chunks / ns), then ns = ceil(chunks / cps)
  const int cps = (num_chunks + ns - 1) / ns;
  return (num_chunks + cps - 1) / cps;
}

bool use_tc(int dtype) { return dtype == MCL_DTYPE_BF16 && !g_opt_simt.load(); }

// slot layout of one scan: the tcgen05 plan, or `nsplit` uniform slots per row block
struct ScanLayout {
  bool tc; TcPlan plan; int nsplit, nslots, rows_padded, nctr;
  // small-batch path (select.cu): one row block, scores dumped [Q][small_ld] + per-range lists
  bool small; int64_t small_ld, gld; int splits; size_t scores_bytes, gmax_bytes, extra_bytes;
  int mode;            // epilogue mode of the tcgen05 scan: top-k / top-1 (k == 1)
  // threshold seeding pre-pass (scan_tc_kernel.cuh, kModeSeed): seed_tiles sample tiles, every
  // seed_stride-th tile of the table, scanned with a plan of their own
  bool seed; TcPlan seed_plan; int seed_tiles, seed_stride; int64_t seed_ld; size_t seed_bytes;
};
constexpr size_t kSmallScoreBytesMax = 64u << 20;   // the dump must stay L2 resident
ScanLayout scan_layout(int64_t Q, int64_t V, int64_t D, int k, int dtype, int sm) {
  ScanLayout L{};
  L.tc = use_tc(dtype);
  const int num_rb = (int)((Q + kBlockM - 1) / kBlockM);
  if (L.tc) {
    L.mode = (k == 1 && !g_opt_notop1.load()) ? 1 : 0;
    const int num_vt = (int)((V + kBlockN - 1) / kBlockN), num_kb = (int)((D + kBlockK - 1) / kBlockK);
    L.small_ld = (V + kChunk - 1) / kChunk * kChunk;
    L.scores_bytes = (((size_t)Q * L.small_ld * sizeof(float)) + 255) & ~(size_t)255;
    // (the two-kernel path is built on the top-k epilogue mode; k = 1 reaches the one-launch path only)
    L.small = num_rb == 1 && (L.mode == 0 || !g_opt_nopanel.load()) && L.scores_bytes <= kSmallScoreBytesMax &&
              !g_opt_nosmall.load();
    // Seeding pays where the epilogue, not the tensor pipe, bounds the scan (D <= 1536) and the
    // sample -- ~2.5 k chunk maxima per row, at most 16 tiles -- is under a tenth of the table.
    const int nt = std::min(16, (5 * k / 2 + 7) / 8);
    L.seed = !L.small && L.mode == 0 && !g_opt_noseed.load() && num_kb <= 24 && nt >= 1 && nt * 8 >= k &&
             num_vt >= 10 * nt;
    PlanKnobs kn = knobs();
    kn.filter = L.mode == 1 ? 2 : (L.seed ? 1 : 0);     // what a segment's restart costs (scan_tc.cu)
    L.plan = make_tc_plan(Q, V, D, sm, kn);
    L.nslots = plan_nslots(L.plan);
    L.rows_padded = L.plan.ru * L.plan.cs;
    L.nctr = plan_nctr(L.plan);
    if (L.seed) {
      L.seed_tiles = nt;
      L.seed_stride = (L.plan.num_vt - 1) / nt;          // never the last (ragged) tile
      kn.filter = 2;
      L.seed_plan = make_tc_plan(Q, (int64_t)nt * kBlockN, D, sm, kn);
      L.seed_ld = (int64_t)std::max(L.rows_padded, L.seed_plan.ru * L.seed_plan.cs) * kBlockM;
      L.seed_bytes = (((size_t)nt * (kBlockN / kChunk) * L.seed_ld * sizeof(uint32_t)) + 255) & ~(size_t)255;
      L.extra_bytes = L.seed_bytes;
    }
    if (L.small) {
      L.splits = select_splits(V);
      // lists sized for k = MCL_MAX_K: the workspace query does not depend on k
      const size_t lists = ((((size_t)L.splits * Q * MCL_MAX_K * 4) + 255) & ~(size_t)255) +
                           (size_t)L.splits * Q * MCL_MAX_K * 8;
      // one-launch path (panel_scan.cu): keys of the maxima of 32 consecutive scores, [Q][gld], and
      // the CTAs' partial statistics, [sm][Q] float4 -- in the place of the two-kernel path's lists
      L.gld = (((V + 127) / 128) * 4 + 31) / 32 * 32;
      L.gmax_bytes = (((size_t)Q * L.gld * sizeof(uint32_t)) + 255) & ~(size_t)255;
      const size_t panel = L.gmax_bytes + (size_t)sm * Q * 16;
      L.extra_bytes = L.scores_bytes + std::max(lists, panel);
    }
  } else {
    L.nsplit = simt_nsplit(Q, V, sm);
    L.nslots = num_rb * L.nsplit;
    L.rows_padded = num_rb;
    L.nctr = 0;
  }
  return L;
}

int check_scan_args(const void* q, const void* table, int dtype, int64_t Q, int64_t V, int64_t D,
                    int64_t ldq, int64_t ldt, const float* inv_q, const float* inv_t, float scale,
                    int k, const float* topk_val, const int64_t* topk_idx, const float* row_stats) {
  if (dtype != MCL_DTYPE_BF16 && dtype != MCL_DTYPE_F32) return fail(MCL_ERR_BAD_ARG, "dtype %d", dtype);
  if (Q < 0 || V < 1 || D < 1) return fail(MCL_ERR_BAD_ARG, "bad shape Q=%lld V=%lld D=%lld", (long long)Q, (long long)V, (long long)D);
  if (Q >= (1ll << 31) || V >= (1ll << 31) || D >= (1ll << 31)) return fail(MCL_ERR_BAD_ARG, "shape exceeds 2^31");
  if (k < 1 || k > MCL_MAX_K || k > V) return fail(MCL_ERR_BAD_ARG, "k=%d must be in [1, min(V=%lld, %d)]", k, (long long)V, MCL_MAX_K);
  if (!(scale > 0.f) || !(scale < INFINITY)) return fail(MCL_ERR_BAD_ARG, "scale must be finite and > 0");
  if (ldq < D || ldt < D) return fail(MCL_ERR_BAD_ARG, "ld < D");
  if (Q > 0 && (!q || !topk_val || !topk_idx || !row_stats)) return fail(MCL_ERR_BAD_ARG, "null pointer");
  if (!table) return fail(MCL_ERR_BAD_ARG, "null table");
  const size_t es = elt_size(dtype);
  if (!aligned16(q) || !aligned16(table) || (ldq * es) % 16 || (ldt * es) % 16)
    return fail(MCL_ERR_UNALIGNED, "q/table base and row pitch must be multiples of 16 bytes");
  if ((inv_t && !aligned16(inv_t)) || !aligned16(row_stats))
    return fail(MCL_ERR_UNALIGNED, "inv_norm_t / row_stats must be 16-byte aligned");
  (void)inv_q;
  return MCL_OK;
}

int scan_impl(const void* q, const void* table, int dtype, int64_t Q, int64_t V, int64_t D,
              int64_t ldq, int64_t ldt, const float* inv_q, const float* inv_t, float scale, int k,
              int64_t index_base, const int64_t* labels, float* topk_val, int64_t* topk_idx,
              float* row_stats, void* workspace, size_t workspace_bytes, float* dbg,
              cudaStream_t stream, float softcap = 0.f, int flags = 0) {
  int rc = check_scan_args(q, table, dtype, Q, V, D, ldq, ldt, inv_q, inv_t, scale, k, topk_val,
                           topk_idx, row_stats);
  if (rc) return rc;
  if (!(softcap >= 0.f) || !(softcap < INFINITY)) return fail(MCL_ERR_BAD_ARG, "softcap must be finite and >= 0");
  DevInfo di;
  if ((rc = require_sm100(&di))) return rc;
  if (Q == 0) return MCL_OK;
  const ScanLayout L = scan_layout(Q, V, D, k, dtype, di.sm);
  Workspace ws = carve_workspace(workspace, L.nslots, L.rows_padded, L.nctr, L.extra_bytes);
  if (!workspace || workspace_bytes < ws.bytes || !aligned16(workspace))
    return fail(MCL_ERR_WORKSPACE_TOO_SMALL, "workspace %zu B < required %zu B (or null/unaligned)",
                workspace_bytes, ws.bytes);
  // MCL_SCAN_NORMALIZE_Q: the library forms 1/||q_row|| itself -- inside the scan kernel for small
  // batches (no extra launch), else with the row kernel into a workspace scratch
  const bool want_qnorm = (flags & MCL_SCAN_NORMALIZE_Q) && !inv_q;
  // One-row-block batches: ONE launch (panel_scan.cu: table panels on the M side, scores dumped to
  // the workspace, grid barrier, exact selection); option 18 restores the two-kernel path
  // (scan_tc_kernel with the filter off + row_select_kernel) for A/B measurements.
  const bool panel = L.small && !dbg && !g_opt_nopanel.load();
  const bool qnorm_in_kernel = want_qnorm && L.tc && dtype == MCL_DTYPE_BF16 && (panel || Q <= 64) && !g_opt_noqnorm.load();
  if (want_qnorm && !qnorm_in_kernel) {
    cudaError_t e0 = launch_row_inv_norm(q, dtype, Q, D, ldq, ws.inv_q, stream);
    if (e0 != cudaSuccess) return cuda_fail(e0, "row_inv_norm launch (queries)");
    g_launches++;
  }
  if (want_qnorm) inv_q = ws.inv_q;                 // (the merge reads what the scan kernel published)
  ScanArgs a{q, table, dtype, Q, V, D, ldq, ldt, inv_q, inv_t, scale, k, index_base, labels, dbg,
             g_opt_timing.load() ? ws.timing : nullptr, ws.tau_shared, ws.sync_ctr,
             g_opt_joint.load() ? ws.joint : nullptr, softcap,
             (int)g_opt_l2.load(), nullptr, 0, L.mode, 1, nullptr, 0,
             qnorm_in_kernel ? 1 : 0, ws.inv_q, nullptr, 0};
  // Small batches (one row block) without a caller-side score dump: the scan keeps only the
  // statistics and drops the scores into the workspace; select.cu picks the top-k from them.
  const bool small = L.small && !dbg;
  if (small) {
    a.small_scores = (float*)ws.extra; a.small_ld = L.small_ld;
    // the selection kernel counts range arrivals per row in the (otherwise unused) threshold words:
    // the scan kernel zeroes them, so this path needs no clear launch
    a.tau_shared = nullptr;
    a.clear_words = ws.tau_shared; a.n_clear = (int)(L.rows_padded * kBlockM);
  }
  SlotMap map{};
  cudaError_t e;
  if (panel) {
    char msg[256] = "";
    a.small_scores = nullptr;
    char* x = (char*)ws.extra + L.scores_bytes;
    e = launch_panel_scan(a, di.sm, (float*)ws.extra, L.small_ld, (uint32_t*)x, L.gld, (float*)(x + L.gmax_bytes),
                          topk_val, topk_idx, row_stats, stream, msg, sizeof(msg));
    if (e != cudaSuccess) return fail(MCL_ERR_CUDA, "panel scan launch: %s %s", cudaGetErrorString(e), msg);
    g_launches++;
    return MCL_OK;
  }
  const bool phases = g_opt_phase.load() != 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (phases) {
    for (auto& x : ev) cudaEventCreate(&x);
    cudaEventRecord(ev[0], stream);
  }
  if (L.tc) {
    char msg[256] = "";
    if (L.seed && !dbg) {
      // threshold seeding: sample scan -> chunk maxima -> k-th largest per row into the shared
      // threshold words (the same kernel clears the drift counters and joint words)
      ScanArgs sa = a;
      sa.mode = 2; sa.tile_stride = L.seed_stride; sa.seed_max = ws.extra; sa.seed_ld = L.seed_ld;
      sa.labels = nullptr; sa.timing = nullptr; sa.dbg_scores = nullptr; sa.small_scores = nullptr;
      e = launch_scan_tc(sa, L.seed_plan, ws.sv, stream, msg, sizeof(msg));
      if (e != cudaSuccess) return fail(MCL_ERR_CUDA, "seed scan launch: %s %s", cudaGetErrorString(e), msg);
      g_launches++;
      e = launch_seed_select((const uint32_t*)ws.extra, L.seed_tiles * (kBlockN / kChunk), L.seed_ld,
                             (int64_t)L.rows_padded * kBlockM, k, (uint32_t*)ws.tau_shared, ws.sync_ctr,
                             ws.zero_bytes - ((char*)ws.sync_ctr - (char*)ws.tau_shared), stream);
      if (e != cudaSuccess) return cuda_fail(e, "seed select launch");
      g_launches++;
    } else if (!small && (L.mode == 0 || L.plan.ru > 1)) {
      // (a k = 1 scan of a single row unit reads neither thresholds nor drift counters; neither
      // does the small-batch path)
      e = launch_zero(ws.tau_shared, ws.zero_bytes, stream);
      if (e != cudaSuccess) return cuda_fail(e, "clearing the shared thresholds");
      g_launches++;
    }
    if (phases) cudaEventRecord(ev[1], stream);
    e = launch_scan_tc(a, L.plan, ws.sv, stream, msg, sizeof(msg));
    if (e != cudaSuccess) return fail(MCL_ERR_CUDA, "scan_tc launch: %s %s", cudaGetErrorString(e), msg);
    map.stride = L.plan.S * 2; map.uniform = 0; map.plan = L.plan;
  } else {
    e = launch_scan_simt(a, ws.sv, L.nsplit, stream);
    if (e != cudaSuccess) return cuda_fail(e, "scan_simt launch");
    map.stride = L.nsplit; map.uniform = L.nsplit;
  }
  g_launches++;
  if (phases) cudaEventRecord(ev[2], stream);
  if (small) {
    char* lists = (char*)ws.extra + L.scores_bytes;
    float* list_val = (float*)lists;
    int64_t* list_idx = (int64_t*)(lists + ((((size_t)L.splits * Q * MCL_MAX_K * 4) + 255) & ~(size_t)255));
    e = launch_select_small(a.small_scores, L.small_ld, V, Q, k, ws.sv, map, list_val, list_idx, ws.tau_shared, index_base,
                            topk_val, topk_idx, row_stats, stream);
    if (e != cudaSuccess) return cuda_fail(e, "select launch");
    g_launches++;
  } else {
    e = launch_merge_slots(ws.sv, map, Q, k, inv_q, scale, softcap, index_base, topk_val, topk_idx, row_stats, stream);
    if (e != cudaSuccess) return cuda_fail(e, "merge launch");
    g_launches++;
  }
  if (phases) {   // debug only: synchronises
    cudaEventRecord(ev[3], stream);
    cudaEventSynchronize(ev[3]);
    if (L.tc)
      for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&g_phase_ms[i], ev[i], ev[i + 1]);
    for (auto& x : ev) cudaEventDestroy(x);
  }
  return MCL_OK;
}

// ---- NCCL through dlopen ---------------------------------------------------------------
struct NcclUid { char internal[128]; };
typedef int (*nccl_get_uid_t)(NcclUid*);
typedef int (*nccl_init_rank_t)(void**, int, NcclUid, int);
typedef int (*nccl_destroy_t)(void*);
typedef int (*nccl_allgather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_t)(int);
typedef int (*nccl_sendrecv_t)(void*, size_t, int, int, void*, cudaStream_t);   // send (const void*) / recv
typedef int (*nccl_group_t)(void);
struct NcclApi {
  void* h = nullptr;
  nccl_get_uid_t get_uid = nullptr;
  nccl_init_rank_t init_rank = nullptr;
  nccl_destroy_t destroy = nullptr;
  nccl_allgather_t allgather = nullptr;
  nccl_errstr_t errstr = nullptr;
  nccl_sendrecv_t send = nullptr, recv = nullptr;
  nccl_group_t group_start = nullptr, group_end = nullptr;
};
NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) return;
    api.get_uid = (nccl_get_uid_t)dlsym(api.h, "ncclGetUniqueId");
    api.init_rank = (nccl_init_rank_t)dlsym(api.h, "ncclCommInitRank");
    api.destroy = (nccl_destroy_t)dlsym(api.h, "ncclCommDestroy");
    api.allgather = (nccl_allgather_t)dlsym(api.h, "ncclAllGather");
    api.errstr = (nccl_errstr_t)dlsym(api.h, "ncclGetErrorString");
    api.send = (nccl_sendrecv_t)dlsym(api.h, "ncclSend");
    api.recv = (nccl_sendrecv_t)dlsym(api.h, "ncclRecv");
    api.group_start = (nccl_group_t)dlsym(api.h, "ncclGroupStart");
    api.group_end = (nccl_group_t)dlsym(api.h, "ncclGroupEnd");
  });
  if (!api.h || !api.get_uid || !api.init_rank || !api.destroy || !api.allgather) return nullptr;
  return &api;
}
int nccl_fail(NcclApi* n, int rc, const char* what) {
  return fail(MCL_ERR_NCCL, "%s: %s", what, (n && n->errstr) ? n->errstr(rc) : "nccl error");
}

struct Record { size_t val_off, idx_off, stats_off, bytes; };
Record record_layout(int64_t Q, int k) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  Record r;
  r.val_off = 0;
  r.idx_off = up((size_t)Q * k * 4);
  r.stats_off = r.idx_off + up((size_t)Q * k * 8);
  r.bytes = r.stats_off + up((size_t)Q * 16);
  return r;
}

}  // namespace

extern "C" {

int mcl_version(void) { return MCL_VERSION; }
const char* mcl_last_error(void) { return g_err; }

int mcl_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  DevInfo d;
  if (!dev_info(&d)) {
    cudaGetLastError();
    return fail(MCL_ERR_CUDA, "no usable CUDA device");
  }
  if (sm_count) *sm_count = d.sm;
  if (cc_major) *cc_major = d.major;
  if (cc_minor) *cc_minor = d.minor;
  return MCL_OK;
}

int mcl_row_inv_norm(const void* x, int dtype, int64_t rows, int64_t dim, int64_t ld,
                     float* inv_norm_out, mcl_stream_t stream) {
  if (dtype != MCL_DTYPE_BF16 && dtype != MCL_DTYPE_F32) return fail(MCL_ERR_BAD_ARG, "dtype %d", dtype);
  if (rows < 0 || dim < 1 || ld < dim) return fail(MCL_ERR_BAD_ARG, "bad shape rows=%lld dim=%lld ld=%lld", (long long)rows, (long long)dim, (long long)ld);
  if (rows > 0 && (!x || !inv_norm_out)) return fail(MCL_ERR_BAD_ARG, "null pointer");
  if (!aligned16(x) || (ld * elt_size(dtype)) % 16) return fail(MCL_ERR_UNALIGNED, "x base and row pitch must be multiples of 16 bytes");
  DevInfo di;
  int rc = require_sm100(&di);
  if (rc) return rc;
  cudaError_t e = launch_row_inv_norm(x, dtype, rows, dim, ld, inv_norm_out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "row_inv_norm launch");
  if (rows) g_launches++;
  return MCL_OK;
}

int mcl_gather_mean(const void* table, int dtype, int64_t V, int64_t D, int64_t ld,
                    const int64_t* offsets, const int64_t* ids, int64_t Q, int normalize, void* out,
                    int64_t ld_out, int* bad_id_flag, mcl_stream_t stream) {
  if (dtype != MCL_DTYPE_BF16 && dtype != MCL_DTYPE_F32) return fail(MCL_ERR_BAD_ARG, "dtype %d", dtype);
  if (V < 1 || D < 1 || ld < D || ld_out < D || Q < 0) return fail(MCL_ERR_BAD_ARG, "bad shape");
  if (!table || (Q > 0 && (!offsets || !out))) return fail(MCL_ERR_BAD_ARG, "null pointer");
  const size_t es = elt_size(dtype);
  if (!aligned16(table) || !aligned16(out) || (ld * es) % 16 || (ld_out * es) % 16)
    return fail(MCL_ERR_UNALIGNED, "table/out base and row pitch must be multiples of 16 bytes");
  DevInfo di;
  int rc = require_sm100(&di);
  if (rc) return rc;
  cudaError_t e = launch_gather_mean(table, dtype, V, D, ld, offsets, ids, Q, normalize, out, ld_out,
                                     bad_id_flag, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "gather_mean launch");
  if (Q) g_launches++;
  return MCL_OK;
}

size_t mcl_scan_workspace_bytes(int64_t Q, int64_t V_local, int64_t D, int k, int dtype) {
  DevInfo di;
  if (!dev_info(&di)) { cudaGetLastError(); di.sm = 148; }
  if (Q <= 0 || V_local <= 0 || D <= 0) return 256;
  const ScanLayout L = scan_layout(Q, V_local, D, k < 1 ? 1 : k, dtype, di.sm);
  return carve_workspace(nullptr, L.nslots, L.rows_padded, L.nctr, L.extra_bytes).bytes;
}

int mcl_concept_scan(const void* q, const void* table, int dtype, int64_t Q, int64_t V_local,
                     int64_t D, int64_t ldq, int64_t ldt, const float* inv_norm_q,
                     const float* inv_norm_t, float scale, int k, int64_t index_base,
                     const int64_t* labels, float* topk_val, int64_t* topk_idx, float* row_stats,
                     void* workspace, size_t workspace_bytes, mcl_stream_t stream) {
  return scan_impl(q, table, dtype, Q, V_local, D, ldq, ldt, inv_norm_q, inv_norm_t, scale, k,
                   index_base, labels, topk_val, topk_idx, row_stats, workspace, workspace_bytes,
                   nullptr, (cudaStream_t)stream);
}

int mcl_concept_scan_ex(const void* q, const void* table, int dtype, int64_t Q, int64_t V_local,
                        int64_t D, int64_t ldq, int64_t ldt, const float* inv_norm_q,
                        const float* inv_norm_t, float scale, float softcap, int k, int64_t index_base,
                        const int64_t* labels, float* topk_val, int64_t* topk_idx, float* row_stats,
                        void* workspace, size_t workspace_bytes, int flags, mcl_stream_t stream) {
  return scan_impl(q, table, dtype, Q, V_local, D, ldq, ldt, inv_norm_q, inv_norm_t, scale, k,
                   index_base, labels, topk_val, topk_idx, row_stats, workspace, workspace_bytes,
                   nullptr, (cudaStream_t)stream, softcap, flags);
}

int mcl_concept_scan_softcap(const void* q, const void* table, int dtype, int64_t Q, int64_t V_local,
                             int64_t D, int64_t ldq, int64_t ldt, const float* inv_norm_q,
                             const float* inv_norm_t, float scale, float softcap, int k,
                             int64_t index_base, const int64_t* labels, float* topk_val,
                             int64_t* topk_idx, float* row_stats, void* workspace,
                             size_t workspace_bytes, float* scores_out, mcl_stream_t stream) {
  return scan_impl(q, table, dtype, Q, V_local, D, ldq, ldt, inv_norm_q, inv_norm_t, scale, k,
                   index_base, labels, topk_val, topk_idx, row_stats, workspace, workspace_bytes,
                   scores_out, (cudaStream_t)stream, softcap);
}

int mcl_concept_scan_debug(const void* q, const void* table, int dtype, int64_t Q, int64_t V_local,
                           int64_t D, int64_t ldq, int64_t ldt, const float* inv_norm_q,
                           const float* inv_norm_t, float scale, int k, int64_t index_base,
                           const int64_t* labels, float* topk_val, int64_t* topk_idx,
                           float* row_stats, void* workspace, size_t workspace_bytes,
                           float* scores_out, mcl_stream_t stream) {
  return scan_impl(q, table, dtype, Q, V_local, D, ldq, ldt, inv_norm_q, inv_norm_t, scale, k,
                   index_base, labels, topk_val, topk_idx, row_stats, workspace, workspace_bytes,
                   scores_out, (cudaStream_t)stream);
}

int mcl_similarity_matrix(const void* q, const void* table, int dtype, int64_t Q, int64_t V,
                          int64_t D, int64_t ldq, int64_t ldt, const float* inv_norm_q,
                          const float* inv_norm_t, float scale, float* scores_out, void* workspace,
                          size_t workspace_bytes, mcl_stream_t stream) {
  // The same scan kernels with the score dump on; the k=1 outputs go to the workspace tail.
  if (!scores_out && Q > 0) return fail(MCL_ERR_BAD_ARG, "null scores_out");
  const size_t need = mcl_scan_workspace_bytes(Q, V, D, 1, dtype);
  const size_t extra = ((size_t)(Q > 0 ? Q : 0) * 32 + 255) & ~(size_t)255;
  if (!workspace || workspace_bytes < need + extra)
    return fail(MCL_ERR_WORKSPACE_TOO_SMALL, "workspace %zu B < required %zu B", workspace_bytes, need + extra);
  char* tail = (char*)workspace + need;
  return scan_impl(q, table, dtype, Q, V, D, ldq, ldt, inv_norm_q, inv_norm_t, scale, 1, 0, nullptr,
                   (float*)(tail + (size_t)Q * 16), (int64_t*)(tail + (size_t)Q * 24), (float*)tail,
                   workspace, need, scores_out, (cudaStream_t)stream);
}

size_t mcl_similarity_workspace_bytes(int64_t Q, int64_t V, int64_t D, int dtype) {
  return mcl_scan_workspace_bytes(Q, V, D, 1, dtype) + (((size_t)(Q > 0 ? Q : 0) * 32 + 255) & ~(size_t)255);
}

int mcl_merge(const float* val, const int64_t* idx, const float* stats, int R, int64_t Q, int k,
              float* out_val, int64_t* out_idx, float* out_stats, mcl_stream_t stream) {
  if (R < 1 || Q < 0 || k < 1 || k > MCL_MAX_K) return fail(MCL_ERR_BAD_ARG, "bad R/Q/k");
  if (Q > 0 && (!val || !idx || !stats || !out_val || !out_idx || !out_stats)) return fail(MCL_ERR_BAD_ARG, "null pointer");
  if (!aligned16(stats) || !aligned16(out_stats)) return fail(MCL_ERR_UNALIGNED, "stats must be 16-byte aligned");
  DevInfo di;
  int rc = require_sm100(&di);
  if (rc) return rc;
  cudaError_t e = launch_merge_ranks(val, idx, stats, (size_t)Q * k * 4, (size_t)Q * k * 8,
                                     (size_t)Q * 16, R, Q, k, out_val, out_idx, out_stats,
                                     (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "merge launch");
  if (Q) g_launches++;
  return MCL_OK;
}

int mcl_ce_from_stats(const float* row_stats, const int64_t* labels, int64_t Q, float label_smoothing,
                      int64_t vocab, float* loss_rows, float* loss_mean, mcl_stream_t stream) {
  if (Q < 0 || vocab < 1 || !(label_smoothing >= 0.f) || !(label_smoothing <= 1.f))
    return fail(MCL_ERR_BAD_ARG, "bad Q / vocab / label_smoothing");
  if (!loss_mean || (Q > 0 && (!row_stats || !labels))) return fail(MCL_ERR_BAD_ARG, "null pointer");
  if (!aligned16(row_stats)) return fail(MCL_ERR_UNALIGNED, "row_stats must be 16-byte aligned");
  DevInfo di;
  int rc = require_sm100(&di);
  if (rc) return rc;
  cudaError_t e = launch_ce_from_stats(row_stats, labels, Q, label_smoothing, vocab, loss_rows, loss_mean,
                                       (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "ce_from_stats launch");
  g_launches++;
  return MCL_OK;
}

// ---- backward of the fused cross-entropy ------------------------------------------------------
namespace {
struct BwdBlocking { int64_t nb, vc, p_bytes; };
constexpr size_t kBwdPBytes = 64u << 20;          // dL/dz block: written once, read twice, mostly out of L2
BwdBlocking bwd_blocking(int64_t n, int64_t V, int dtype) {
  BwdBlocking b{};
  if (dtype == MCL_DTYPE_BF16) {
    const int64_t n_pad = (n + kBlockM - 1) / kBlockM * kBlockM;
    b.nb = n_pad < 4096 ? n_pad : 4096;
    int64_t vc = (int64_t)(kBwdPBytes / (size_t)(b.nb * 2)) / kBlockN * kBlockN;
    const int64_t v_pad = (V + kBlockN - 1) / kBlockN * kBlockN;
    b.vc = vc < kBlockN ? kBlockN : (vc > v_pad ? v_pad : vc);
    b.p_bytes = b.nb * b.vc * 2;
  } else {
    b.nb = n;
    int64_t vc = (int64_t)(kBwdPBytes / (size_t)((n > 0 ? n : 1) * 4)) / 64 * 64;
    const int64_t v_pad = (V + 63) / 64 * 64;
    b.vc = vc < 64 ? 64 : (vc > v_pad ? v_pad : vc);
    b.p_bytes = b.nb * b.vc * 4;
  }
  return b;
}
}  // namespace

size_t mcl_ce_backward_workspace_bytes(int64_t n, int64_t V, int64_t D, int dtype) {
  (void)D;
  if (n <= 0 || V <= 0) return 256;
  return (size_t)((bwd_blocking(n, V, dtype).p_bytes + 255) & ~(int64_t)255);
}

int64_t mcl_ce_backward_block_rows(int64_t n, int64_t V, int dtype) {
  if (n <= 0 || V <= 0) return 0;
  return bwd_blocking(n, V, dtype).nb;
}

int mcl_ce_backward(const void* q, const void* table, int dtype, int64_t n, int64_t V, int64_t D, int64_t ldq,
                    int64_t ldt, const float* lse, const int64_t* labels, float scale, float softcap,
                    float label_smoothing, int64_t vocab_total, const float* grad_loss, int64_t n_valid,
                    float* grad_q, float* grad_table, void* workspace, size_t workspace_bytes,
                    mcl_stream_t stream_) {
  return mcl_ce_backward_ex(q, table, dtype, n, V, D, ldq, ldt, lse, labels, scale, softcap, label_smoothing,
                            vocab_total, grad_loss, n_valid, grad_q, grad_table, MCL_DTYPE_F32, workspace,
                            workspace_bytes, stream_);
}

int mcl_ce_backward_ex(const void* q, const void* table, int dtype, int64_t n, int64_t V, int64_t D, int64_t ldq,
                       int64_t ldt, const float* lse, const int64_t* labels, float scale, float softcap,
                       float label_smoothing, int64_t vocab_total, const float* grad_loss, int64_t n_valid,
                       float* grad_q, void* grad_table_, int grad_table_dtype, void* workspace,
                       size_t workspace_bytes, mcl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  float* grad_table = (float*)grad_table_;
  if (grad_table_dtype != MCL_DTYPE_F32 && grad_table_dtype != MCL_DTYPE_BF16)
    return fail(MCL_ERR_BAD_ARG, "grad_table_dtype %d", grad_table_dtype);
  const bool gt_bf16 = grad_table_ && grad_table_dtype == MCL_DTYPE_BF16;
  if (gt_bf16 && (dtype != MCL_DTYPE_BF16 || n > mcl_ce_backward_block_rows(n, V, dtype) || (D % 8)))
    return fail(MCL_ERR_BAD_ARG, "a bf16 table gradient needs bf16 inputs, D %% 8 == 0 and all %lld rows in one block "
                "of mcl_ce_backward_block_rows (no accumulation across row blocks)", (long long)n);
  if (dtype != MCL_DTYPE_BF16 && dtype != MCL_DTYPE_F32) return fail(MCL_ERR_BAD_ARG, "dtype %d", dtype);
  if (n < 0 || V < 1 || D < 1 || ldq < D || ldt < D || vocab_total < 1 || n_valid < 1)
    return fail(MCL_ERR_BAD_ARG, "bad shape n=%lld V=%lld D=%lld", (long long)n, (long long)V, (long long)D);
  if (n >= (1ll << 31) || V >= (1ll << 31)) return fail(MCL_ERR_BAD_ARG, "shape exceeds 2^31");
  if (!(scale > 0.f) || !(softcap >= 0.f) || !(label_smoothing >= 0.f) || !(label_smoothing <= 1.f))
    return fail(MCL_ERR_BAD_ARG, "bad scale / softcap / label_smoothing");
  if (n > 0 && (!q || !lse || !labels || !grad_loss)) return fail(MCL_ERR_BAD_ARG, "null pointer");
  if (!table) return fail(MCL_ERR_BAD_ARG, "null table");
  const size_t es = elt_size(dtype);
  if (!aligned16(q) || !aligned16(table) || (ldq * es) % 16 || (ldt * es) % 16 || !aligned16(grad_q) ||
      !aligned16(grad_table) || (D % 4))
    return fail(MCL_ERR_UNALIGNED, "q / table / gradient rows must be 16-byte aligned");
  DevInfo di;
  int rc = require_sm100(&di);
  if (rc) return rc;
  if (n == 0 || (!grad_q && !grad_table)) return MCL_OK;
  const BwdBlocking bl = bwd_blocking(n, V, dtype);
  if (!workspace || workspace_bytes < (size_t)bl.p_bytes || !aligned16(workspace))
    return fail(MCL_ERR_WORKSPACE_TOO_SMALL, "workspace %zu B < required %lld B", workspace_bytes, (long long)bl.p_bytes);
  const float eps_over_v = label_smoothing / (float)vocab_total, one_minus_eps = 1.f - label_smoothing;
  const float coef = 1.f / (float)n_valid;
  cudaError_t e;
  if (dtype == MCL_DTYPE_F32) {
    // check path: Z = q Tc^T -> dL/dz in place -> the two gradient products, chunk by chunk over the table
    float* P = (float*)workspace;
    const float* qf = (const float*)q;
    const float* tf = (const float*)table;
    for (int64_t v0 = 0; v0 < V; v0 += bl.vc) {
      const int64_t vc = V - v0 < bl.vc ? V - v0 : bl.vc;
      e = launch_gemm_simt(qf, ldq, 1, tf + v0 * ldt, 1, ldt, P, bl.vc, n, vc, D, 0, stream);
      if (e == cudaSuccess)
        e = launch_dz_simt(P, bl.vc, n, vc, v0, lse, labels, scale, softcap, eps_over_v, one_minus_eps, grad_loss,
                           coef * scale, stream);
      if (e == cudaSuccess && grad_q)
        e = launch_gemm_simt(P, bl.vc, 1, tf + v0 * ldt, ldt, 1, grad_q, D, n, D, vc, v0 > 0, stream);
      if (e == cudaSuccess && grad_table)
        e = launch_gemm_simt(P, 1, bl.vc, qf, ldq, 1, grad_table + v0 * D, D, vc, D, n, 0, stream);
      if (e != cudaSuccess) return cuda_fail(e, "backward (fp32 check path) launch");
      g_launches += 2 + (grad_q ? 1 : 0) + (grad_table ? 1 : 0);
    }
    return MCL_OK;
  }
  // tcgen05 path: for every block of rows, walk the table in chunks: dL/dz block (scan kernel, grad
  // epilogue) -> dL/dq += P T (accumulates over the chunks) and dL/dT chunk (+)= P^T q (accumulates
  // over the row blocks)
  const __nv_bfloat16* qb = (const __nv_bfloat16*)q;
  const __nv_bfloat16* tb = (const __nv_bfloat16*)table;
  PlanKnobs kn = knobs();
  kn.filter = 2;
  char msg[256] = "";
  for (int64_t r0 = 0; r0 < n; r0 += bl.nb) {
    const int64_t rows = n - r0 < bl.nb ? n - r0 : bl.nb;
    for (int64_t v0 = 0; v0 < V; v0 += bl.vc) {
      const int64_t vc = V - v0 < bl.vc ? V - v0 : bl.vc;
      const TcPlan plan = make_tc_plan(rows, vc, D, di.sm, kn);
      ScanArgs a{};
      a.q = qb + r0 * ldq; a.table = tb + v0 * ldt; a.dtype = dtype; a.Q = rows; a.V = vc; a.D = D;
      a.ldq = ldq; a.ldt = ldt; a.scale = scale; a.k = 1; a.index_base = v0; a.labels = labels + r0;
      a.softcap = softcap; a.mode = 3; a.tile_stride = 1;
      a.p_out = workspace; a.ldp = bl.vc; a.p_rows = bl.nb;
      a.lse = lse + r0; a.grad_loss = grad_loss; a.grad_coef = coef * scale;
      a.eps_over_v = eps_over_v; a.one_minus_eps = one_minus_eps;
      e = launch_scan_tc(a, plan, SlotView{}, stream, msg, sizeof(msg));
      if (e != cudaSuccess) return fail(MCL_ERR_CUDA, "backward dL/dz launch: %s %s", cudaGetErrorString(e), msg);
      g_launches++;
      if (grad_q) {
        e = launch_gemm_tc(workspace, 0, bl.vc, tb + v0 * ldt, 1, ldt, grad_q + r0 * D, D, rows, D, vc, v0 > 0,
                           di.sm, stream);
        if (e != cudaSuccess) return cuda_fail(e, "backward dL/dq GEMM launch");
        g_launches++;
      }
      if (grad_table) {
        // (one row block: the product is final -- with a bf16 gradient the accumulators are rounded once
        // on the way out, half the bytes of the largest write of the backward and no cast pass after it)
        float* gt_chunk = gt_bf16 ? (float*)((__nv_bfloat16*)grad_table_ + v0 * D) : grad_table + v0 * D;
        e = launch_gemm_tc(workspace, 1, bl.vc, qb + r0 * ldq, 1, ldq, gt_chunk, D, vc, D, rows, r0 > 0,
                           di.sm, stream, gt_bf16 ? 1 : 0);
        if (e != cudaSuccess) return cuda_fail(e, "backward dL/dT GEMM launch");
        g_launches++;
      }
    }
  }
  return MCL_OK;
}

int mcl_gemm_bf16(const void* a, int a_mn, int64_t lda, const void* b, int b_mn, int64_t ldb, float* c,
                  int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate, mcl_stream_t stream) {
  if (M < 0 || N < 0 || K < 1 || !a || !b || !c) return fail(MCL_ERR_BAD_ARG, "bad GEMM arguments");
  if (!aligned16(a) || !aligned16(b) || !aligned16(c) || (lda * 2) % 16 || (ldb * 2) % 16)
    return fail(MCL_ERR_UNALIGNED, "operand bases and pitches must be multiples of 16 bytes");
  DevInfo di;
  int rc = require_sm100(&di);
  if (rc) return rc;
  cudaError_t e = launch_gemm_tc(a, a_mn, lda, b, b_mn, ldb, c, ldc, M, N, K, accumulate, di.sm, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "gemm launch");
  g_launches++;
  return MCL_OK;
}

int mcl_comm_unique_id(void* unique_id_out) {
  NcclApi* n = nccl();
  if (!n) {
    const char* why = dlerror();                     // (one call: dlerror() clears what it returns)
    return fail(MCL_ERR_NCCL, "libnccl.so.2 not loadable: %s", why ? why : "");
  }
  if (!unique_id_out) return fail(MCL_ERR_BAD_ARG, "null pointer");
  int rc = n->get_uid((NcclUid*)unique_id_out);
  if (rc) return nccl_fail(n, rc, "ncclGetUniqueId");
  return MCL_OK;
}

int mcl_comm_init(const void* unique_id, int world, int rank, void** comm_out) {
  NcclApi* n = nccl();
  if (!n) return fail(MCL_ERR_NCCL, "libnccl.so.2 not loadable");
  if (!unique_id || !comm_out || world < 1 || rank < 0 || rank >= world) return fail(MCL_ERR_BAD_ARG, "bad comm args");
  NcclUid uid;
  memcpy(&uid, unique_id, sizeof(uid));
  void* comm = nullptr;
  int rc = n->init_rank(&comm, world, uid, rank);
  if (rc) return nccl_fail(n, rc, "ncclCommInitRank");
  *comm_out = comm;
  return MCL_OK;
}

int mcl_comm_destroy(void* comm) {
  NcclApi* n = nccl();
  if (!n) return fail(MCL_ERR_NCCL, "libnccl.so.2 not loadable");
  if (!comm) return MCL_OK;
  int rc = n->destroy(comm);
  if (rc) return nccl_fail(n, rc, "ncclCommDestroy");
  return MCL_OK;
}

int mcl_comm_all_gather(void* comm, const void* send, void* recv, size_t bytes_per_rank,
                        mcl_stream_t stream) {
  NcclApi* n = nccl();
  if (!n) return fail(MCL_ERR_NCCL, "libnccl.so.2 not loadable");
  if (!comm || !send || !recv) return fail(MCL_ERR_BAD_ARG, "null pointer");
  if (bytes_per_rank == 0) return MCL_OK;
  const int rc = n->allgather(send, recv, bytes_per_rank, /*ncclInt8*/ 0, comm, (cudaStream_t)stream);
  if (rc) return nccl_fail(n, rc, "ncclAllGather");
  return MCL_OK;
}

size_t mcl_sharded_gather_bytes(int64_t Q, int k, int world) {
  if (Q < 0 || k < 1 || world < 1) return 0;
  // R full records (all-gather path) + room for R mini-records of Q/R rows (row-exchange path)
  const size_t rec = record_layout(Q, k).bytes;
  return rec * (size_t)world + rec + (size_t)world * 1024;
}

int mcl_concept_scan_sharded(const void* q, const void* table_shard, int dtype, int64_t Q,
                             int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                             const float* inv_norm_q, const float* inv_norm_t, float scale, int k,
                             int64_t index_base, const int64_t* labels, float* topk_val,
                             int64_t* topk_idx, float* row_stats, void* workspace,
                             size_t workspace_bytes, void* gather_buf, size_t gather_bytes,
                             void* comm, int world, int rank, mcl_stream_t stream) {
  return mcl_concept_scan_sharded_ex(q, table_shard, dtype, Q, V_local, D, ldq, ldt, inv_norm_q, inv_norm_t,
                                     scale, k, index_base, labels, topk_val, topk_idx, row_stats, workspace,
                                     workspace_bytes, gather_buf, gather_bytes, comm, world, rank, 0, stream);
}

int mcl_concept_scan_sharded_ex(const void* q, const void* table_shard, int dtype, int64_t Q,
                                int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                                const float* inv_norm_q, const float* inv_norm_t, float scale, int k,
                                int64_t index_base, const int64_t* labels, float* topk_val,
                                int64_t* topk_idx, float* row_stats, void* workspace,
                                size_t workspace_bytes, void* gather_buf, size_t gather_bytes,
                                void* comm, int world, int rank, int flags, mcl_stream_t stream) {
  if (world < 1 || rank < 0 || rank >= world) return fail(MCL_ERR_BAD_ARG, "bad world/rank");
  // the rank merge packs GLOBAL row ids into the low 32 bits of its sort keys (merge.cu)
  if (index_base < 0 || index_base + V_local > (1ll << 32))
    return fail(MCL_ERR_BAD_ARG, "global table rows must stay below 2^32 (index_base %lld + V_local %lld)",
                (long long)index_base, (long long)V_local);
  const Record rec = record_layout(Q, k);
  const size_t gather_need = mcl_sharded_gather_bytes(Q, k, world);
  if (!gather_buf || gather_bytes < gather_need || !aligned16(gather_buf))
    return fail(MCL_ERR_WORKSPACE_TOO_SMALL, "gather_buf %zu B < required %zu B", gather_bytes,
                gather_need);
  char* mine = (char*)gather_buf + rec.bytes * (size_t)rank;
  // 1. local scan straight into this rank's record of the gather buffer
  int rc = scan_impl(q, table_shard, dtype, Q, V_local, D, ldq, ldt, inv_norm_q, inv_norm_t, scale,
                     k, index_base, labels, (float*)(mine + rec.val_off),
                     (int64_t*)(mine + rec.idx_off), (float*)(mine + rec.stats_off), workspace,
                     workspace_bytes, nullptr, (cudaStream_t)stream, 0.f,
                     (flags & MCL_SHARDED_NORMALIZE_Q) ? MCL_SCAN_NORMALIZE_Q : 0);
  if (rc) return rc;
  NcclApi* n = nullptr;
  if (world > 1) {
    n = nccl();
    if (!n) return fail(MCL_ERR_NCCL, "libnccl.so.2 not loadable");
    if (!comm) return fail(MCL_ERR_BAD_ARG, "null communicator");
  }
  char* base = (char*)gather_buf;
  const bool exchange = world > 2 && Q % world == 0 && n->send && n->recv && n->group_start &&
                        n->group_end && !g_opt_allgather.load();
  if (!exchange) {
    // 2a. one all-gather of the packed records (in place), 3a. every rank merges all Q rows
    if (world > 1) {
      int nrc = n->allgather(mine, gather_buf, rec.bytes, /*ncclInt8*/ 0, comm, (cudaStream_t)stream);
      if (nrc) return nccl_fail(n, nrc, "ncclAllGather");
    }
    cudaError_t e = launch_merge_ranks((const float*)(base + rec.val_off),
                                       (const int64_t*)(base + rec.idx_off),
                                       (const float*)(base + rec.stats_off), rec.bytes, rec.bytes,
                                       rec.bytes, world, Q, k, topk_val, topk_idx, row_stats,
                                       (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "merge launch");
    if (Q) g_launches++;
    return MCL_OK;
  }
  // 2b. Larger worlds: rank j merges only query rows [j*Q/N, (j+1)*Q/N).  Every rank sends each
  // peer that peer's row range of its local lists (one grouped send/recv exchange, 1/N of the
  // all-gather's bytes), merges its own rows, and the merged rows are all-gathered straight into
  // the output arrays.  The local record sits in slot `rank` of gather_buf; the R mini-records
  // of Q/N rows are received behind the R full slots.
  const int64_t rows = Q / world;
  const Record mini = record_layout(rows, k);
  char* area = base + rec.bytes * (size_t)world;          // receive area behind the R records
  const size_t vb = (size_t)rows * k * 4, ib = (size_t)rows * k * 8, sb = (size_t)rows * 16;
  int nrc = n->group_start();
  if (nrc) return nccl_fail(n, nrc, "ncclGroupStart");
  for (int peer = 0; peer < world; ++peer) {
    char* dst = area + mini.bytes * (size_t)peer;          // rows of MINE as computed by `peer`
    const size_t r0 = (size_t)peer * rows;                  // rows of PEER as computed by me
    if (!nrc) nrc = n->send(mine + rec.val_off + r0 * k * 4, vb, 0, peer, comm, (cudaStream_t)stream);
    if (!nrc) nrc = n->send(mine + rec.idx_off + r0 * k * 8, ib, 0, peer, comm, (cudaStream_t)stream);
    if (!nrc) nrc = n->send(mine + rec.stats_off + r0 * 16, sb, 0, peer, comm, (cudaStream_t)stream);
    if (!nrc) nrc = n->recv(dst + mini.val_off, vb, 0, peer, comm, (cudaStream_t)stream);
    if (!nrc) nrc = n->recv(dst + mini.idx_off, ib, 0, peer, comm, (cudaStream_t)stream);
    if (!nrc) nrc = n->recv(dst + mini.stats_off, sb, 0, peer, comm, (cudaStream_t)stream);
  }
  const int erc = n->group_end();
  if (nrc || erc) return nccl_fail(n, nrc ? nrc : erc, "ncclSend/ncclRecv exchange");
  // 3b. merge my rows into their place in the outputs
  float* my_val = topk_val + (size_t)rank * rows * k;
  int64_t* my_idx = topk_idx + (size_t)rank * rows * k;
  float* my_stats = row_stats + (size_t)rank * rows * 4;
  cudaError_t e = launch_merge_ranks((const float*)(area + mini.val_off),
                                     (const int64_t*)(area + mini.idx_off),
                                     (const float*)(area + mini.stats_off), mini.bytes, mini.bytes,
                                     mini.bytes, world, rows, k, my_val, my_idx, my_stats,
                                     (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "merge launch");
  g_launches++;
  // 4b. all-gather the merged rows in place -- unless the caller only consumes its own row range
  // (host consumers: every rank copies its Q/N rows to the host)
  if (flags & MCL_SHARDED_LOCAL_ROWS) return MCL_OK;
  nrc = n->group_start();
  if (!nrc) nrc = n->allgather(my_val, topk_val, vb, 0, comm, (cudaStream_t)stream);
  if (!nrc) nrc = n->allgather(my_idx, topk_idx, ib, 0, comm, (cudaStream_t)stream);
  if (!nrc) nrc = n->allgather(my_stats, row_stats, sb, 0, comm, (cudaStream_t)stream);
  const int erc2 = n->group_end();
  if (nrc || erc2) return nccl_fail(n, nrc ? nrc : erc2, "ncclAllGather of the merged rows");
  return MCL_OK;
}

// ---- peer exchange of replicated query batches (copy engines over NVLink, no SMs) ----------
typedef int (*cu_wait32_t)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static cu_wait32_t get_wait32() {
  static cu_wait32_t fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (cu_wait32_t)p;
    else
      cudaGetLastError();
  });
  return fn;
}

// ---- sharded scan with the result exchange over peer memory (p2p_exchange.cu) -------------------
namespace {
struct P2PLayout { size_t xcount, fcount, done, mine, xin, fin, bytes; Record rec, mini; int64_t rows; };
P2PLayout p2p_layout(int64_t Q, int k, int world) {
  P2PLayout L{};
  L.rows = Q / world;
  L.rec = record_layout(Q, k);
  L.mini = record_layout(L.rows, k);
  L.xcount = 0; L.fcount = 64; L.done = 128;
  L.mine = 256;
  L.xin = L.mine + L.rec.bytes;                            // two receive areas, used in turn (see below)
  L.fin = L.xin + 2 * L.mini.bytes * (size_t)world;
  L.bytes = L.fin + L.rec.bytes;
  return L;
}
}  // namespace

size_t mcl_sharded_p2p_block_bytes(int64_t Q, int k, int world) {
  // 0: this (Q, k, world) cannot use the peer-memory exchange (pieces must be whole 16-byte vectors)
  if (world < 2 || world > kPushMaxPeers || Q < world || Q % world || k < 1 || ((Q / world) * k) % 4) return 0;
  return p2p_layout(Q, k, world).bytes;
}

int mcl_concept_scan_sharded_p2p(const void* q, const void* table_shard, int dtype, int64_t Q,
                                 int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                                 const float* inv_norm_q, const float* inv_norm_t, float scale, int k,
                                 int64_t index_base, const int64_t* labels, float* topk_val,
                                 int64_t* topk_idx, float* row_stats, void* workspace,
                                 size_t workspace_bytes, void* const* peer_blocks, size_t block_bytes,
                                 int world, int rank, uint32_t epoch, uint32_t full_epoch, int flags,
                                 mcl_stream_t stream) {
  if (world < 2 || world > kPushMaxPeers || rank < 0 || rank >= world) return fail(MCL_ERR_BAD_ARG, "bad world/rank");
  if (index_base < 0 || index_base + V_local > (1ll << 32))
    return fail(MCL_ERR_BAD_ARG, "global table rows must stay below 2^32 (index_base %lld + V_local %lld)",
                (long long)index_base, (long long)V_local);
  const size_t need = mcl_sharded_p2p_block_bytes(Q, k, world);
  if (!need) return fail(MCL_ERR_BAD_ARG, "peer-memory exchange needs Q %% world == 0 and (Q / world) * k %% 4 == 0");
  if (!peer_blocks || block_bytes < need) return fail(MCL_ERR_WORKSPACE_TOO_SMALL, "peer block %zu B < required %zu B", block_bytes, need);
  for (int r = 0; r < world; ++r)
    if (!peer_blocks[r] || !aligned16(peer_blocks[r])) return fail(MCL_ERR_BAD_ARG, "null / unaligned peer block %d", r);
  if (epoch == 0 || epoch > (1u << 31))
    return fail(MCL_ERR_BAD_ARG, "epoch counts the scans on these blocks from 1 (at most 2^31: take fresh blocks then)");
  if (!(flags & MCL_SHARDED_LOCAL_ROWS) && (full_epoch == 0 || full_epoch > epoch))
    return fail(MCL_ERR_BAD_ARG, "full_epoch counts the scans without MCL_SHARDED_LOCAL_ROWS on these blocks from 1 "
                "(this one included): the second arrival counter advances only in those");
  cu_wait32_t wait32 = get_wait32();
  if (!wait32) return fail(MCL_ERR_UNIMPLEMENTED, "cuStreamWaitValue32 is not available in this driver");
  DevInfo di;
  int rc = require_sm100(&di);
  if (rc) return rc;
  const P2PLayout L = p2p_layout(Q, k, world);
  const Record& rec = L.rec;
  const Record& mini = L.mini;
  const int64_t rows = L.rows;
  char* me = (char*)peer_blocks[rank];
  char* mine = me + L.mine;
  cudaStream_t s = (cudaStream_t)stream;
  // 1. local scan into this rank's record
  rc = scan_impl(q, table_shard, dtype, Q, V_local, D, ldq, ldt, inv_norm_q, inv_norm_t, scale, k, index_base,
                 labels, (float*)(mine + rec.val_off), (int64_t*)(mine + rec.idx_off), (float*)(mine + rec.stats_off),
                 workspace, workspace_bytes, nullptr, s, 0.f,
                 (flags & MCL_SHARDED_NORMALIZE_Q) ? MCL_SCAN_NORMALIZE_Q : 0);
  if (rc) return rc;
  const size_t vb = (size_t)rows * k * 4, ib = (size_t)rows * k * 8, sb = (size_t)rows * 16;
  // 2. every peer's row range of the record -> that peer's receive area, slot `rank`.  The receive
  // areas alternate with the epoch: with MCL_SHARDED_LOCAL_ROWS nothing after the merge makes a fast
  // rank wait for a slow one, but its push of step e+2 needs the slow rank's push of step e+1, which
  // follows that rank's merge of step e in stream order -- so the area of step e is free by then.
  const size_t xin_off = L.xin + (size_t)(epoch & 1u) * mini.bytes * (size_t)world;
  // The arrival counters alternate with the epoch like the areas: a peer can run at most ONE step
  // ahead of this rank's wait (same argument), so counter [e & 1] only ever holds arrivals of steps
  // e, e-2, ...: reaching its target proves that every peer's pieces of step e are here, never that
  // a fast peer's pieces of step e+1 made up for a slow peer's of step e.
  const size_t xcount_off = L.xcount + 4u * (size_t)(epoch & 1u);
  const uint32_t xcount_target = ((epoch + 1u) >> 1) * (uint32_t)(world - 1);   // steps of this parity so far
  PushParams pp{};
  pp.npeer = world; pp.self = rank; pp.counter_off = xcount_off; pp.done = (unsigned*)(me + L.done);
  for (int r = 0; r < world; ++r) pp.peer[r] = (char*)peer_blocks[r];
  for (int peer = 0; peer < world; ++peer) {
    const size_t r0 = (size_t)peer * rows, slot = xin_off + mini.bytes * (size_t)rank;
    const PushSeg segs[3] = {{mine + rec.val_off + r0 * k * 4, slot + mini.val_off, vb},
                             {mine + rec.idx_off + r0 * k * 8, slot + mini.idx_off, ib},
                             {mine + rec.stats_off + r0 * 16, slot + mini.stats_off, sb}};
    for (const PushSeg& sg : segs) { pp.seg[pp.nseg] = sg; pp.seg_peer[pp.nseg] = peer; ++pp.nseg; }
  }
  cudaError_t e = launch_p2p_push(pp, di.sm, s);
  if (e != cudaSuccess) return cuda_fail(e, "p2p push launch");
  g_launches++;
  // 3. wait for the other ranks' pieces (a stream memory operation: no SM spins)
  int wrc = wait32(s, (unsigned long long)(uintptr_t)(me + xcount_off), xcount_target, /*GEQ*/ 0u);
  if (wrc) return fail(MCL_ERR_CUDA, "cuStreamWaitValue32 failed with CUresult %d", wrc);
  // 4. merge my rows into their place in the result area
  char* fin = me + L.fin;
  const size_t o_val = rec.val_off + (size_t)rank * vb, o_idx = rec.idx_off + (size_t)rank * ib,
               o_st = rec.stats_off + (size_t)rank * sb;
  char* area = me + xin_off;
  e = launch_merge_ranks((const float*)(area + mini.val_off), (const int64_t*)(area + mini.idx_off),
                         (const float*)(area + mini.stats_off), mini.bytes, mini.bytes, mini.bytes, world, rows, k,
                         (float*)(fin + o_val), (int64_t*)(fin + o_idx), (float*)(fin + o_st), s);
  if (e != cudaSuccess) return cuda_fail(e, "merge launch");
  g_launches++;
  // (the copies out of the result area: ONE launch of the push kernel with the caller's arrays as the
  // only "peer" -- three cudaMemcpyAsync cost three launches)
  auto copy_out = [&](const char* s0, char* d0, size_t b0, const char* s1, char* d1, size_t b1, const char* s2,
                      char* d2, size_t b2) -> cudaError_t {
    if (!aligned16(d0) || !aligned16(d1) || !aligned16(d2)) {
      cudaError_t ce = cudaMemcpyAsync(d0, s0, b0, cudaMemcpyDeviceToDevice, s);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(d1, s1, b1, cudaMemcpyDeviceToDevice, s);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(d2, s2, b2, cudaMemcpyDeviceToDevice, s);
      return ce;
    }
    PushParams pc{};
    pc.npeer = 1; pc.self = 0; pc.counter_off = 0; pc.done = (unsigned*)(me + L.done);
    pc.peer[0] = nullptr;                                  // dst_off carries the whole address
    pc.nseg = 3;
    pc.seg[0] = PushSeg{s0, (size_t)(uintptr_t)d0, b0};
    pc.seg[1] = PushSeg{s1, (size_t)(uintptr_t)d1, b1};
    pc.seg[2] = PushSeg{s2, (size_t)(uintptr_t)d2, b2};
    g_launches++;
    return launch_p2p_push(pc, di.sm, s);
  };
  if (flags & MCL_SHARDED_LOCAL_ROWS) {   // the caller takes this rank's row range only
    e = copy_out(fin + o_val, (char*)(topk_val + (size_t)rank * rows * k), vb, fin + o_idx,
                 (char*)(topk_idx + (size_t)rank * rows * k), ib, fin + o_st, (char*)(row_stats + (size_t)rank * rows * 4), sb);
    if (e != cudaSuccess) return cuda_fail(e, "copying the merged rows out");
    return MCL_OK;
  }
  // 5. the merged rows -> every other rank's result area, same place
  PushParams pf{};
  pf.npeer = world; pf.self = rank; pf.counter_off = L.fcount; pf.done = (unsigned*)(me + L.done);
  for (int r = 0; r < world; ++r) pf.peer[r] = (char*)peer_blocks[r];
  for (int peer = 0; peer < world; ++peer) {
    if (peer == rank) continue;
    const PushSeg segs[3] = {{fin + o_val, L.fin + o_val, vb}, {fin + o_idx, L.fin + o_idx, ib}, {fin + o_st, L.fin + o_st, sb}};
    for (const PushSeg& sg : segs) { pf.seg[pf.nseg] = sg; pf.seg_peer[pf.nseg] = peer; ++pf.nseg; }
  }
  e = launch_p2p_push(pf, di.sm, s);
  if (e != cudaSuccess) return cuda_fail(e, "p2p push launch (merged rows)");
  g_launches++;
  wrc = wait32(s, (unsigned long long)(uintptr_t)(me + L.fcount), full_epoch * (uint32_t)(world - 1), /*GEQ*/ 0u);
  if (wrc) return fail(MCL_ERR_CUDA, "cuStreamWaitValue32 failed with CUresult %d", wrc);
  // 6. hand the result to the caller's arrays
  e = copy_out(fin + rec.val_off, (char*)topk_val, (size_t)Q * k * 4, fin + rec.idx_off, (char*)topk_idx,
               (size_t)Q * k * 8, fin + rec.stats_off, (char*)row_stats, (size_t)Q * 16);
  if (e != cudaSuccess) return cuda_fail(e, "copying the result out");
  return MCL_OK;
}

int mcl_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out) {
  if (!dev_ptr || !ipc_handle_out || bytes == 0) return fail(MCL_ERR_BAD_ARG, "bad peer_alloc args");
  static_assert(sizeof(cudaIpcMemHandle_t) == MCL_IPC_HANDLE_BYTES, "ipc handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc (peer buffer)");
  e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return MCL_OK;
}

int mcl_peer_free(void* dev_ptr) {
  if (!dev_ptr) return MCL_OK;
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFree (peer buffer)");
  return MCL_OK;
}

int mcl_peer_open(const void* ipc_handle, void** peer_ptr) {
  if (!ipc_handle || !peer_ptr) return fail(MCL_ERR_BAD_ARG, "bad peer_open args");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
  *peer_ptr = p;
  return MCL_OK;
}

int mcl_peer_close(void* peer_ptr) {
  if (!peer_ptr) return MCL_OK;
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
  return MCL_OK;
}

int mcl_memcpy_async(void* dst, const void* src, size_t bytes, mcl_stream_t stream) {
  if (bytes == 0) return MCL_OK;
  if (!dst || !src) return fail(MCL_ERR_BAD_ARG, "null pointer");
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync");
  return MCL_OK;
}

int mcl_stream_wait_value32(mcl_stream_t stream, const void* dev_addr, uint32_t value) {
  cu_wait32_t fn = get_wait32();
  if (!fn) return fail(MCL_ERR_UNIMPLEMENTED, "cuStreamWaitValue32 is not available in this driver");
  if (!dev_addr || ((uintptr_t)dev_addr & 3u)) return fail(MCL_ERR_BAD_ARG, "bad flag address");
  const int rc = fn((cudaStream_t)stream, (unsigned long long)(uintptr_t)dev_addr, value, /*GEQ*/ 0u);
  if (rc) return fail(MCL_ERR_CUDA, "cuStreamWaitValue32 failed with CUresult %d", rc);
  return MCL_OK;
}

int64_t mcl_set_option(int opt, int64_t value) {
  if (opt == 0) return g_opt_ctas.exchange(value);
  if (opt == 1) return g_opt_g.exchange(value);
  if (opt == 2) return g_opt_simt.exchange(value);
  if (opt == 3) return g_opt_timing.exchange(value);
  if (opt == 4) return g_opt_cluster.exchange(value);
  if (opt == 5) return g_opt_allgather.exchange(value);
  if (opt == 6) return g_opt_phase.exchange(value);
  if (opt == 7) return g_opt_leftover.exchange(value);
  if (opt == 8) return g_opt_segpen.exchange(value);
  if (opt == 9) return g_opt_l2.exchange(value);
  if (opt == 10) return g_opt_win.exchange(value);
  if (opt == 11) return g_opt_nosmall.exchange(value);
  if (opt == 12) return g_opt_joint.exchange(value);
  if (opt == 13) return g_opt_noseed.exchange(value);
  if (opt == 14) return g_opt_notop1.exchange(value);
  if (opt == 15) return g_opt_filter.exchange(value);
  if (opt == 16) return g_opt_noqnorm.exchange(value);
  if (opt == 17) return g_gather_variant.exchange((int)value);
  if (opt == 18) return g_opt_nopanel.exchange(value);
  if (opt == 19) return g_merge_variant.exchange((int)value);
  if (opt == 103) return drift_timeouts_total();
  if (opt == 104) return panel_barrier_faults();
  if (opt >= 100 && opt < 103) return (int64_t)(g_phase_ms[opt - 100] * 1.0e6f);   // read-back, ns
  return -1;
}

int64_t mcl_launch_count(void) { return g_launches.load(); }

int mcl_plan_scan(int64_t Q, int64_t V_local, int64_t D, int sm_count, int32_t* plan_out) {
  if (Q < 1 || V_local < 1 || D < 1 || sm_count < 1 || !plan_out) return fail(MCL_ERR_BAD_ARG, "bad plan args");
  const TcPlan s = make_tc_plan(Q, V_local, D, sm_count, knobs());
  int32_t v[MCL_PLAN_INTS] = {s.num_rb, s.num_vt, s.num_kb, s.cs, s.workers, s.gu, s.ru, s.waves, s.S,
                              plan_nslots(s), plan_grid(s), s.win, s.nsync, plan_nctr(s), s.full.n, s.last.n};
  int n = 16;
  for (const Chain* c : {&s.full, &s.last})
    for (int i = 0; i < kMaxNodes; ++i) {
      const Node z{};
      const Node& d = i < c->n ? c->nd[i] : z;
      v[n++] = d.r0; v[n++] = d.R; v[n++] = d.a; v[n++] = d.w0; v[n++] = d.nfull; v[n++] = d.tpc;
      v[n++] = d.t0; v[n++] = d.wr; v[n++] = d.passes;
    }
  for (int i = 0; i < MCL_PLAN_INTS; ++i) plan_out[i] = v[i];
  return MCL_OK;
}

int64_t mcl_plan_segments(int64_t Q, int64_t V_local, int64_t D, int sm_count, int32_t* segs_out,
                          int64_t cap) {
  if (Q < 1 || V_local < 1 || D < 1 || sm_count < 1 || cap < 0 || (cap > 0 && !segs_out))
    return fail(MCL_ERR_BAD_ARG, "bad plan args");
  const TcPlan s = make_tc_plan(Q, V_local, D, sm_count, knobs());
  int64_t n = 0;
  for (int w = 0; w < s.workers; ++w) {
    SegIter it;
    seg_iter_init(it, w);
    Seg g;
    for (; seg_iter_next(s, it, g); ++n)
      if (n < cap) {
        int32_t* o = segs_out + 6 * n;
        o[0] = w; o[1] = g.unit; o[2] = g.vt0; o[3] = g.vt1; o[4] = g.j; o[5] = g.sync;
      }
  }
  return n;
}

int mcl_plan_row_block_slots(int64_t Q, int64_t V_local, int64_t D, int sm_count, int64_t row_block,
                             int32_t* first_slot, int32_t* num_slots) {
  if (Q < 1 || V_local < 1 || D < 1 || sm_count < 1 || !first_slot || !num_slots)
    return fail(MCL_ERR_BAD_ARG, "bad plan args");
  const TcPlan s = make_tc_plan(Q, V_local, D, sm_count, knobs());
  if (row_block < 0 || row_block >= s.num_rb) return fail(MCL_ERR_BAD_ARG, "row block out of range");
  SlotMap map{};
  map.stride = s.S * 2; map.uniform = 0; map.plan = s;
  *first_slot = (int32_t)row_block * map.stride;
  *num_slots = slotmap_count(map, (int)row_block);
  return MCL_OK;
}

}  // extern "C"
