// Host-side launch interfaces shared by the translation units of libmcl_sm100.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"
#include "plan.h"

struct CUtensorMap_st;   // <cuda.h>: CUtensorMap

namespace mcl {

struct ScanArgs {
  const void* q;
  const void* table;
  int dtype;                 // 0 bf16, 1 f32
  int64_t Q, V, D, ldq, ldt;
  const float* inv_q;        // nullable
  const float* inv_t;        // nullable
  float scale;
  int k;
  int64_t index_base;
  const int64_t* labels;     // nullable
  float* dbg_scores;         // nullable, [Q,V]
  void* timing;              // nullable, [grid][2] uint64 (tcgen05 scan only)
  void* tau_shared;          // nullable, [num_rb*128] uint32 zeroed before launch (tcgen05 scan)
  void* sync_ctr;            // [plan_nctr] int zeroed before launch (tcgen05 scan)
  void* joint;               // [padded rows][kJointWords] uint32 zeroed before launch (nullable: off)
  float softcap;             // 0 = off; c > 0: logits are c*tanh(z/c)
  int l2_mode;               // bit 0: query tiles evict-last, bit 1: table tiles evict-first
  float* small_scores;       // small-batch path: [Q][small_ld] score dump, top-k filter off (nullable)
  int64_t small_ld;
  int mode;                  // epilogue mode of the tcgen05 scan (scan_tc_kernel.cuh): 0 top-k, 1 top-1, 2 seed
  int tile_stride;           // seed mode: plan tile t = table tile t * tile_stride
  void* seed_max;            // seed mode: [sample chunks][seed_ld] uint32 keys of the chunk maxima
  int64_t seed_ld;
  int qnorm_in_kernel;       // tcgen05 scan computes 1/||q_row|| itself (small batches) -> inv_q_out
  float* inv_q_out;
  void* clear_words;         // uint32 words the scan's CTA 0 zeroes for the next kernel (nullable)
  int n_clear;
  // grad mode (mode 3): dL/dz tiles for the backward GEMMs
  void* p_out; int64_t ldp; int64_t p_rows;
  const float* lse; const float* grad_loss; float grad_coef, eps_over_v, one_minus_eps;
};

// Tile plan of the tcgen05 scan (plan.h).  ctas / gu / cluster = 0: heuristic.  leftover = 0
// plans without tail workers (A/B measurements); seg_penalty = tiles charged per extra segment
// of a tail worker when balancing it against the groups.
struct PlanKnobs { int ctas = 0, gu = 0, cluster = 0, leftover = 1, seg_penalty = 1, win = 0, filter = 0; };
TcPlan make_tc_plan(int64_t Q, int64_t V, int64_t D, int sm_count, const PlanKnobs& knobs);
inline int plan_nslots(const TcPlan& p) { return p.ru * p.cs * p.S * 2; }   // incl. padding row blocks
inline int plan_nctr(const TcPlan& p) { return p.waves * p.nsync; }
inline int plan_grid(const TcPlan& p) { return p.workers * p.cs; }

struct Workspace {
  void* timing;   // [1024][2] uint64 at offset 0: per-CTA globaltimer start/end (debug option 3)
  void* tau_shared;   // one threshold word per (padded) query row, shared between CTAs
  void* sync_ctr;     // window-arrival counters of the tile scheduler (follow tau_shared)
  void* joint;        // joint-threshold words (follow sync_ctr)
  size_t zero_bytes;  // tau_shared + sync_ctr + joint: cleared before every scan
  SlotView sv;
  float* inv_q;       // [padded rows] scratch: query inverse norms the library computed itself
  void* extra;        // path-specific tail (small-batch path: score dump + per-range lists)
  int nslots;
  size_t bytes;
};
// carve `nslots` slots out of a caller buffer (base may be null to only size it)
Workspace carve_workspace(void* base, int nslots, int num_rb, int nctr, size_t extra_bytes);

// 2-D bf16 row-major [rows, cols] (pitch ld elements) -> TMA tiles of box_rows x 64, SWIZZLE_128B
bool make_tmap_bf16(::CUtensorMap_st* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
// C[M,N] fp32 (+)= A * B on tcgen05 (gemm_tc.cu); *_mn = 1: the operand is stored [K][M or N]
// (few output tiles and a long K: K is split over the SMs and the partial products are added with
// atomics -- the summation order of such a product is not fixed; out_bf16: c is a __nv_bfloat16*, pitch
// ldc elements, the fp32 accumulators are rounded once on the way out -- not with accumulate)
cudaError_t launch_gemm_tc(const void* a, int a_mn, int64_t lda, const void* b, int b_mn, int64_t ldb,
                           float* c, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate,
                           int sm_count, cudaStream_t s, int out_bf16 = 0);
// fp32 check path of the backward (bwd_simt.cu): C (+)= op(A) * op(B) with arbitrary strides, and Z -> dL/dz in place
cudaError_t launch_gemm_simt(const float* a, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_k, int64_t sb_n,
                             float* c, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate, cudaStream_t s);
cudaError_t launch_dz_simt(float* z, int64_t ldz, int64_t rows, int64_t cols, int64_t col_base, const float* lse,
                           const int64_t* labels, float scale, float softcap, float eps_over_v, float one_minus_eps,
                           const float* grad_loss, float grad_coef, cudaStream_t s);
cudaError_t launch_scan_tc(const ScanArgs& a, const TcPlan& plan, const SlotView& sv,
                           cudaStream_t s, char* err, size_t errlen);
long long drift_timeouts_total();   // drift waits of the tcgen05 scan that gave up (all devices)
cudaError_t launch_scan_simt(const ScanArgs& a, const SlotView& sv, int nsplit, cudaStream_t s);

// zero `bytes` (a multiple of 16, 16-byte aligned) on the stream
cudaError_t launch_zero(void* ptr, size_t bytes, cudaStream_t s);

// seed pre-pass: k-th largest chunk maximum per row -> tau_shared; clears zero_bytes at zero_ptr
cudaError_t launch_seed_select(const uint32_t* seed_max, int n_chunks, int64_t ld, int64_t rows, int k,
                               uint32_t* tau_shared, void* zero_ptr, size_t zero_bytes, cudaStream_t s);
// per-row CE and its mean over the rows with a label from the scan's row statistics
cudaError_t launch_ce_from_stats(const float* row_stats, const int64_t* labels, int64_t Q,
                                 float label_smoothing, int64_t vocab, float* loss_rows,
                                 float* loss_mean, cudaStream_t s);

// slots -> final [Q,k] / [Q,4]
cudaError_t launch_merge_slots(const SlotView& sv, const SlotMap& map, int64_t Q, int k,
                               const float* inv_q, float scale, float softcap, int64_t index_base,
                               float* topk_val, int64_t* topk_idx, float* row_stats,
                               cudaStream_t s);
// Small-batch path (select.cu): exact top-k of the dumped scores [Q][ld] + the slots' statistics.
// list_val / list_idx: [select_splits(V)][Q][k] scratch.
int select_splits(int64_t V);
cudaError_t launch_select_small(const float* scores, int64_t ld, int64_t V, int64_t Q, int k,
                                const SlotView& sv, const SlotMap& map, float* list_val,
                                int64_t* list_idx, void* row_ctr, int64_t index_base, float* topk_val,
                                int64_t* topk_idx, float* row_stats, cudaStream_t s);
// One-row-block batches in one launch (panel_scan.cu): scores [Q][ld] scratch -> final outputs.
cudaError_t launch_panel_scan(const ScanArgs& a, int sm_count, float* scores, int64_t ld, uint32_t* gmax,
                              int64_t gld, float* part, float* topk_val,
                              int64_t* topk_idx, float* row_stats, cudaStream_t s, char* err, size_t errlen);
long long panel_barrier_faults();   // grid-barrier waits of the panel scan that gave up (all devices)
// R per-shard results (rank r's arrays start r * stride bytes after the base) -> final
cudaError_t launch_merge_ranks(const float* val, const int64_t* idx, const float* stats,
                               size_t val_stride, size_t idx_stride, size_t stats_stride, int R,
                               int64_t Q, int k, float* out_val, int64_t* out_idx,
                               float* out_stats, cudaStream_t s);

// Peer-memory result exchange of the sharded scan (p2p_exchange.cu): stores `nseg` pieces into the
// peers' blocks and bumps their arrival counters once every CTA is done.
struct PushSeg { const char* src; size_t dst_off, bytes; };
constexpr int kPushMaxPeers = 16;
constexpr int kPushMaxSegs = 3 * kPushMaxPeers;
struct PushParams {
  int nseg, npeer;
  PushSeg seg[kPushMaxSegs];
  int seg_peer[kPushMaxSegs];
  char* peer[kPushMaxPeers];
  size_t counter_off;
  unsigned* done;
  int self;
};
cudaError_t launch_p2p_push(const PushParams& p, int sm_count, cudaStream_t s);

cudaError_t launch_row_inv_norm(const void* x, int dtype, int64_t rows, int64_t dim, int64_t ld,
                                float* out, cudaStream_t s);
cudaError_t launch_gather_mean(const void* table, int dtype, int64_t V, int64_t D, int64_t ld,
                               const int64_t* offsets, const int64_t* ids, int64_t Q,
                               int normalize, void* out, int64_t ld_out, int* bad_flag,
                               cudaStream_t s);

}  // namespace mcl
