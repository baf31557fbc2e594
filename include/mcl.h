/*
 * mcl.h -- C ABI of libmcl_sm100.so: the B200-native concept-embedding similarity scan.
 *
 * The reference (AskSid/multimodal_concept_learning) is pure Python and has no FFI of
 * its own; the path is reached through ordinary Python calls into torch / sklearn /
 * transformers.  Each entry point below therefore cites the *reference call site* whose
 * arithmetic it replaces (paths relative to the reference tree).  INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless marked
 *     "host"; sizes are in elements; matrices are row-major with the inner dim contiguous.
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never
 *     allocates or frees device memory (one explicit exception: mcl_peer_alloc / mcl_peer_free,
 *     below) and never synchronises the stream.
 *   - every call returns 0 on success or a negative MCL_ERR_* code; a message is kept in a
 *     thread-local string readable with mcl_last_error().  No C++ exception crosses the ABI.
 *   - there is no CPU fallback: on a device that is not compute capability 10.x the compute
 *     calls return MCL_ERR_UNSUPPORTED_ARCH.
 */
#ifndef MCL_H_
#define MCL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCL_VERSION 100

/* error codes */
#define MCL_OK 0
#define MCL_ERR_BAD_ARG (-1)
#define MCL_ERR_UNALIGNED (-2)          /* base % 16 != 0 or ld*sizeof(elt) % 16 != 0 */
#define MCL_ERR_UNSUPPORTED_ARCH (-3)   /* device is not sm_100 */
#define MCL_ERR_WORKSPACE_TOO_SMALL (-4)
#define MCL_ERR_CUDA (-5)
#define MCL_ERR_NCCL (-6)
#define MCL_ERR_UNIMPLEMENTED (-7)

/* element types of q / table */
#define MCL_DTYPE_BF16 0   /* product path: TMA + tcgen05 (UTCHMMA) + TMEM, fp32 accumulate */
#define MCL_DTYPE_F32 1    /* check path: fp32 inputs, fp32 FMA accumulate on CUDA cores     */

/* largest k the fused epilogue keeps per row */
#define MCL_MAX_K 64

/* row_stats layout: 4 floats per query row */
#define MCL_STAT_M 0       /* max_j z_j                       */
#define MCL_STAT_S 1       /* sum_j exp(z_j - m)              */
#define MCL_STAT_SUMZ 2    /* sum_j z_j  (label smoothing)    */
#define MCL_STAT_ZLABEL 3  /* z_label, 0 if ignored/not local */

typedef void* mcl_stream_t; /* a cudaStream_t */

int mcl_version(void);
const char* mcl_last_error(void); /* host string, thread-local, valid until the next call */

/* sm count and compute capability of the current device (host outputs). */
int mcl_device_info(int* sm_count, int* cc_major, int* cc_minor);

/*
 * inv_norm_out[i] = 1 / ||x[i,:]||_2 in fp32; rows with norm < 10*FLT_EPSILON give 1
 * (scikit-learn `normalize` semantics).
 * Replaces: the two sklearn `normalize` calls inside `cosine_similarity` at
 *   src/multimodal/token_embedding_analysis.py:244, and the numpy row-normalisation of
 *   random_experiments/multi_token_embedding/multi_token.ipynb cell 3 line 16.
 */
int mcl_row_inv_norm(const void* x, int dtype, int64_t rows, int64_t dim, int64_t ld,
                     float* inv_norm_out, mcl_stream_t stream);

/*
 * Multi-token concept embeddings: out[i,:] = mean(table[ids[offsets[i]:offsets[i+1]], :]),
 * empty segments give zeros; fp32 accumulation, one rounding to the table dtype; if
 * `normalize` != 0 the mean is L2-normalised (fp32) before the rounding.
 * Replaces: `average_embeddings_for_tokens`
 *   src/multimodal/token_embedding_analysis_imagenet.py:261-286 (hot line :281) and
 *   `get_averaged_embedding`, multi_token.ipynb cell 2 lines 1-12.
 * ids outside [0,V) -> MCL_ERR_BAD_ARG is NOT detectable without a sync; they are clamped
 * on the device and flagged in *bad_id_flag (device int, nullable).
 */
int mcl_gather_mean(const void* table, int dtype, int64_t V, int64_t D, int64_t ld,
                    const int64_t* offsets /*[Q+1]*/, const int64_t* ids /*[nnz]*/, int64_t Q,
                    int normalize, void* out /*[Q,D] table dtype*/, int64_t ld_out,
                    int* bad_id_flag, mcl_stream_t stream);

/*
 * The fused scan.  For every query row i (z_ij = <q_i, t_j> * rq_i * rt_j * scale, with
 * rq / rt the optional inverse row norms):
 *     topk_val[i,:], topk_idx[i,:]  the k largest z_ij, descending, exact ties broken by
 *                                   the lowest table row; idx = index_base + local row
 *     row_stats[i,:]                (m, s, sum_z, z_label) -- see MCL_STAT_*
 * For batches of more than one row block (Q > 128) the [Q x V] score matrix is never written
 * to memory: the top-k filter and the statistics run in the GEMM's epilogue.  Batches of one
 * row block (the reference's 6..96 concept tokens) are bound by the table read, not by the
 * scores: there ONE launch (csrc/panel_scan.cu) streams the table as 128-row UMMA panels, drops
 * the Q x V_local scores -- a few percent of the table bytes, L2 resident -- into the workspace
 * with the maxima of every 32 of them, and after an in-kernel grid barrier one CTA per query row
 * selects the top-k exactly (same values, same tie rule).  The barrier assumes the launch has the
 * GPU to itself (grid <= SM count); its wait is bounded, and a give-up yields NaN / -1 outputs and
 * counts in mcl_set_option(104, 0).  lse_i = m + log s;
 * CE_i = (1-eps)(lse_i - z_label) + eps (lse_i - sum_z / V) is formed by the caller.
 * Replaces, in one call:
 *   - sklearn `cosine_similarity` per pair   src/multimodal/token_embedding_analysis.py:237-246
 *   - HF lm_head GEMM + fp32 upcast + `F.cross_entropy(ignore_index=-100)`
 *                                            src/multimodal/mllm.py:115-120
 *   - `torch.argmax(logits, dim=-1)`         src/multimodal/multimodal_training.py:276
 *   - `CrossEntropyLoss(label_smoothing)` + `torch.max(logits, 1)`
 *                                            src/vision/vision_training.py:81-83,116,132
 * Constraints: 1 <= k <= min(V_local, MCL_MAX_K); scale > 0; D % 8 == 0 (bf16) so rows are
 * 16-byte aligned; bases 16-byte aligned.  labels are GLOBAL row ids (int64), -100 = ignore.
 */
size_t mcl_scan_workspace_bytes(int64_t Q, int64_t V_local, int64_t D, int k, int dtype);

int mcl_concept_scan(const void* q /*[Q,D]*/, const void* table /*[V_local,D]*/, int dtype,
                     int64_t Q, int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                     const float* inv_norm_q /*[Q] nullable*/,
                     const float* inv_norm_t /*[V_local] nullable*/, float scale, int k,
                     int64_t index_base, const int64_t* labels /*[Q] nullable*/,
                     float* topk_val /*[Q,k]*/, int64_t* topk_idx /*[Q,k]*/,
                     float* row_stats /*[Q,4]*/, void* workspace, size_t workspace_bytes,
                     mcl_stream_t stream);

/*
 * The same scan with every option in one call: softcap (0 = off) as in mcl_concept_scan_softcap,
 * and flags.  MCL_SCAN_NORMALIZE_Q: inv_norm_q is NULL and the library forms 1/||q_row|| itself
 * (sklearn zero-row rule) -- inside the scan kernel for batches of up to 64 rows, where launches
 * bound the step (the reference's 6..96 concept tokens), else with the row kernel into workspace
 * scratch.  The values are bit-identical to mcl_row_inv_norm's.
 */
#define MCL_SCAN_NORMALIZE_Q 1
int mcl_concept_scan_ex(const void* q, const void* table, int dtype, int64_t Q, int64_t V_local,
                        int64_t D, int64_t ldq, int64_t ldt, const float* inv_norm_q,
                        const float* inv_norm_t, float scale, float softcap, int k,
                        int64_t index_base, const int64_t* labels, float* topk_val,
                        int64_t* topk_idx, float* row_stats, void* workspace, size_t workspace_bytes,
                        int flags, mcl_stream_t stream);

/*
 * Soft-capped variant: every logit is z' = softcap * tanh(z / softcap) before the top-k values,
 * the log-sum-exp, sum_z and z_label are formed (ranking is by z: tanh is monotone).
 * softcap = 0 is mcl_concept_scan.  scores_out is nullable (a [Q,V_local] dump of z' for tests).
 * Replaces: the `final_logit_softcapping` branch of the HF LM head that mllm.py:115 reaches,
 *   site-packages/transformers/models/gemma3/modeling_gemma3.py:653-656 (None for Gemma-3-1B,
 *   30.0 for the Gemma-2-2B of random_experiments/multi_token_embedding).
 */
int mcl_concept_scan_softcap(const void* q, const void* table, int dtype, int64_t Q,
                             int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                             const float* inv_norm_q, const float* inv_norm_t, float scale,
                             float softcap, int k, int64_t index_base, const int64_t* labels,
                             float* topk_val, int64_t* topk_idx, float* row_stats, void* workspace,
                             size_t workspace_bytes, float* scores_out, mcl_stream_t stream);

/* Same call, additionally dumping z to scores_out [Q, V_local] fp32 (tests only). */
int mcl_concept_scan_debug(const void* q, const void* table, int dtype, int64_t Q,
                           int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                           const float* inv_norm_q, const float* inv_norm_t, float scale, int k,
                           int64_t index_base, const int64_t* labels, float* topk_val,
                           int64_t* topk_idx, float* row_stats, void* workspace,
                           size_t workspace_bytes, float* scores_out, mcl_stream_t stream);

/*
 * Dense similarity matrix scores_out[i,j] = z_ij (fp32, [Q,V] contiguous) for SMALL Q x V:
 * the all-pairs cosine matrix the token analysis turns into distances.
 * Replaces the O(n^2) Python loop of per-pair `cosine_similarity` calls,
 *   src/multimodal/token_embedding_analysis.py:237-246 (n = 6..1000 concept tokens).
 */
size_t mcl_similarity_workspace_bytes(int64_t Q, int64_t V, int64_t D, int dtype);
int mcl_similarity_matrix(const void* q, const void* table, int dtype, int64_t Q, int64_t V,
                          int64_t D, int64_t ldq, int64_t ldt, const float* inv_norm_q,
                          const float* inv_norm_t, float scale, float* scores_out, void* workspace,
                          size_t workspace_bytes, mcl_stream_t stream);

/*
 * Merge R partial results (one per vocabulary shard): candidates are re-ranked by
 * (value desc, index asc); m* = max m_r, s* = sum s_r exp(m_r - m*); sum_z and z_label add.
 * val/idx/stats are [R,Q,k] / [R,Q,k] / [R,Q,4]; idx < 0 marks an empty candidate; indices must be
 * below 2^32 (they are packed into the low word of the 64-bit sort keys).
 */
int mcl_merge(const float* val, const int64_t* idx, const float* stats, int R, int64_t Q, int k,
              float* out_val, int64_t* out_idx, float* out_stats, mcl_stream_t stream);

/*
 * Cross-entropy from row_stats (one launch, deterministic): loss_rows[i] =
 * (1-eps)(lse_i - z_label_i) + eps (lse_i - sum_z_i / vocab), 0 where labels[i] == -100 (nullable
 * output); loss_mean[0] = mean over the rows with a label (NaN if there is none, as torch),
 * loss_mean[1] = their number.
 * Replaces: `F.cross_entropy(logits.float(), shifted_labels, ignore_index=-100)` inside
 *   ForCausalLMLoss reached from src/multimodal/mllm.py:115-120, and
 *   `nn.CrossEntropyLoss(label_smoothing=eps)` at src/vision/vision_training.py:81-83,116.
 */
int mcl_ce_from_stats(const float* row_stats /*[Q,4]*/, const int64_t* labels /*[Q]*/, int64_t Q,
                      float label_smoothing, int64_t vocab, float* loss_rows /*[Q] nullable*/,
                      float* loss_mean /*[2]*/, mcl_stream_t stream);

/*
 * Backward of the fused cross-entropy: with loss = mean over the n rows
 * (all of which carry a label; the caller passes only those) of CE(scale * q T^T, labels) as the
 * forward scan computed it, and lse[i] the log-sum-exp that scan returned,
 *     dL/dz[i,j]  = (exp(z_ij - lse_i) - (1-eps) [j = label_i] - eps / vocab_total) * *grad_loss / n_valid
 *     grad_q      = scale * dL/dz  * T      [n, D]  fp32, pitch D
 *     grad_table  = scale * dL/dz^T * q     [V, D]  fp32, pitch D
 * (softcap > 0: z' = softcap * tanh(z / softcap) and the factor 1 - tanh^2 rides along.)  Either
 * gradient pointer may be NULL.  bf16 inputs: the scores are recomputed tile by tile by the tcgen05
 * scan kernel, whose grad epilogue writes dL/dz in bf16 for a block of <= 4096 rows x a chunk of
 * table rows (<= 64 MB, L2-resident) into the workspace, and two tcgen05 GEMMs (csrc/gemm_tc.cu,
 * MN-major operand descriptors: no transposed copy of anything) consume it; the [n x V] matrix is
 * never materialised.  fp32 inputs: CUDA-core check path (rtol 1e-4).
 * Replaces: `accelerator.backward(loss)` through lm_head + ForCausalLMLoss,
 *   src/multimodal/multimodal_training.py:140 (`language_embed_only` trains the table itself,
 *   src/multimodal/mllm.py:181-184), and through the classifier head, src/vision/vision_training.py:120.
 */
size_t mcl_ce_backward_workspace_bytes(int64_t n, int64_t V, int64_t D, int dtype);
int mcl_ce_backward(const void* q /*[n,D]*/, const void* table /*[V,D]*/, int dtype, int64_t n, int64_t V,
                    int64_t D, int64_t ldq, int64_t ldt, const float* lse /*[n]*/,
                    const int64_t* labels /*[n]*/, float scale, float softcap, float label_smoothing,
                    int64_t vocab_total, const float* grad_loss /*device scalar*/, int64_t n_valid,
                    float* grad_q /*[n,D] nullable*/, float* grad_table /*[V,D] nullable*/,
                    void* workspace, size_t workspace_bytes, mcl_stream_t stream);

/*
 * Same, with the table gradient's dtype: MCL_DTYPE_BF16 (bf16 inputs, D % 8 == 0, and all n rows in
 * one row block: n <= mcl_ce_backward_block_rows(n, V, dtype)) makes the last GEMM round its fp32
 * accumulators once and write [V, D] bf16 -- half the bytes of the backward's largest write and no
 * cast pass for a bf16 parameter.  With few rows (the reference's answer-only supervision: ~24 of
 * 1672 positions carry a label) dL/dq = dL/dz * T has a handful of output tiles and K = V: its K
 * range is split over the SMs and the partial products are added with atomics (the summation order
 * of grad_q is then not fixed from run to run; rtol 1e-6).
 */
int64_t mcl_ce_backward_block_rows(int64_t n, int64_t V, int dtype);
int mcl_ce_backward_ex(const void* q, const void* table, int dtype, int64_t n, int64_t V, int64_t D,
                       int64_t ldq, int64_t ldt, const float* lse, const int64_t* labels, float scale,
                       float softcap, float label_smoothing, int64_t vocab_total, const float* grad_loss,
                       int64_t n_valid, float* grad_q, void* grad_table, int grad_table_dtype,
                       void* workspace, size_t workspace_bytes, mcl_stream_t stream);

/*
 * The GEMM of that backward on its own: C[M,N] fp32 (+)= A * B, bf16 operands on tcgen05.
 * a_mn = 0: A is stored [M][K] (pitch lda); a_mn = 1: A is stored [K][M].  Same for B with N.
 */
int mcl_gemm_bf16(const void* a, int a_mn, int64_t lda, const void* b, int b_mn, int64_t ldb, float* c,
                  int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate, mcl_stream_t stream);

/*
 * Vocabulary-sharded scan over the GPUs of one NVSwitch box: local scan, ONE ncclAllGather
 * of the packed per-rank record, local merge -- all enqueued on `stream`.
 * The library owns only the communicator.  `unique_id` is the 128-byte ncclUniqueId (host),
 * produced on rank 0 by mcl_comm_unique_id and distributed by the caller
 * (torch.distributed's store in the Python host layer).
 */
int mcl_comm_unique_id(void* unique_id_out /*host, 128 B*/);
int mcl_comm_init(const void* unique_id /*host, 128 B*/, int world, int rank, void** comm_out);
int mcl_comm_destroy(void* comm);

/*
 * Plain all-gather over the library's communicator on `stream` (in place when
 * send == recv + rank * bytes_per_rank).  The host layer uses it to assemble a replicated query
 * batch from the 1/N slices each rank uploaded over its own PCIe link.
 */
int mcl_comm_all_gather(void* comm, const void* send, void* recv, size_t bytes_per_rank,
                        mcl_stream_t stream);

/* bytes of `gather_buf` needed by mcl_concept_scan_sharded */
size_t mcl_sharded_gather_bytes(int64_t Q, int k, int world);

int mcl_concept_scan_sharded(const void* q, const void* table_shard, int dtype, int64_t Q,
                             int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                             const float* inv_norm_q, const float* inv_norm_t, float scale,
                             int k, int64_t index_base, const int64_t* labels,
                             float* topk_val, int64_t* topk_idx, float* row_stats,
                             void* workspace, size_t workspace_bytes, void* gather_buf,
                             size_t gather_bytes, void* comm, int world, int rank,
                             mcl_stream_t stream);

/*
 * Same, with flags.  MCL_SHARDED_LOCAL_ROWS: with the row exchange (world > 2, Q % world == 0)
 * rank r merges query rows [r*Q/world, (r+1)*Q/world) only; the flag skips the final all-gather of
 * the merged rows, so only that row range of topk_val / topk_idx / row_stats is written -- for
 * consumers that take each rank's rows separately (host copies of 1/world of the result per rank).
 * Without the row exchange every rank merges every row and the flag changes nothing.
 */
#define MCL_SHARDED_LOCAL_ROWS 1
#define MCL_SHARDED_NORMALIZE_Q 2   /* inv_norm_q is NULL: the library forms the query norms (MCL_SCAN_NORMALIZE_Q) */
int mcl_concept_scan_sharded_ex(const void* q, const void* table_shard, int dtype, int64_t Q,
                                int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                                const float* inv_norm_q, const float* inv_norm_t, float scale,
                                int k, int64_t index_base, const int64_t* labels,
                                float* topk_val, int64_t* topk_idx, float* row_stats,
                                void* workspace, size_t workspace_bytes, void* gather_buf,
                                size_t gather_bytes, void* comm, int world, int rank, int flags,
                                mcl_stream_t stream);

/*
 * The sharded scan with the result exchange over PEER MEMORY instead of NCCL (csrc/p2p_exchange.cu):
 * every rank stores each peer's row range of its local lists straight into that peer's block
 * (one kernel of 16-byte stores over NVLink that ends in a system-scope add on the peer's arrival
 * counter), the consumer's stream waits for its counter with a stream memory operation, merges its
 * Q / world rows, and the merged rows travel the same way into every rank's result area.  Two small
 * kernels and two stream waits replace a grouped ncclSend/ncclRecv and a grouped ncclAllGather.
 *   peer_blocks[world]  HOST array of device pointers: block r is rank r's (own block at [rank]);
 *                       each is mcl_sharded_p2p_block_bytes(Q, k, world) bytes from mcl_peer_alloc
 *                       (zeroed), the others' opened with mcl_peer_open.  One set of blocks serves
 *                       one (Q, k) shape; `epoch` counts the scans issued on it: 1, 2, 3, ... up to
 *                       2^31 (the same on every rank -- the arrival counters are monotone and, like
 *                       the receive areas, alternate with the epoch's parity); `full_epoch`
 *                       counts those of them that ran WITHOUT MCL_SHARDED_LOCAL_ROWS, this one
 *                       included (the merged rows travel, and their counter advances, only then).
 *   mcl_sharded_p2p_block_bytes returns 0 when the shape cannot take this path (it needs
 *                       2 <= world <= 16, Q % world == 0 and (Q / world) * k % 4 == 0): use
 *                       mcl_concept_scan_sharded_ex then.
 * flags as mcl_concept_scan_sharded_ex.  Same outputs, bit for bit, as the NCCL paths.
 */
size_t mcl_sharded_p2p_block_bytes(int64_t Q, int k, int world);
int mcl_concept_scan_sharded_p2p(const void* q, const void* table_shard, int dtype, int64_t Q,
                                 int64_t V_local, int64_t D, int64_t ldq, int64_t ldt,
                                 const float* inv_norm_q, const float* inv_norm_t, float scale,
                                 int k, int64_t index_base, const int64_t* labels,
                                 float* topk_val, int64_t* topk_idx, float* row_stats,
                                 void* workspace, size_t workspace_bytes, void* const* peer_blocks,
                                 size_t block_bytes, int world, int rank, uint32_t epoch,
                                 uint32_t full_epoch, int flags, mcl_stream_t stream);

/*
 * Peer exchange of replicated query batches without SMs.  A sharded scan needs the whole query
 * batch on every GPU; when the batch arrives from the host, every rank uploads 1/N of it over
 * its own PCIe link and PUSHES that slice into the staging buffer of every peer with copy-engine
 * copies over NVLink (cudaMemcpyAsync on a side stream, overlapping the previous scan), followed
 * by a 4-byte step counter into the peer's flag word; the consumer's stream waits on its own flag
 * words with a stream memory operation.  No kernel, no NCCL call, no SM is involved.
 *   mcl_peer_alloc   cudaMalloc + zero + cudaIpcGetMemHandle: the ONE kind of device memory the
 *                    library allocates, because IPC handles need a plain cudaMalloc block (a
 *                    framework's caching / virtual-memory allocator does not give one);
 *                    freed by mcl_peer_free.  ipc_handle_out: MCL_IPC_HANDLE_BYTES host bytes,
 *                    distributed by the caller (torch.distributed in the Python host layer).
 *   mcl_peer_open    maps another rank's block into this process (cudaIpcOpenMemHandle).
 *   mcl_memcpy_async cudaMemcpyAsync(cudaMemcpyDefault) on `stream`: local, peer or pinned-host
 *                    pointers (UVA).
 *   mcl_stream_wait_value32   `stream` waits until (int32)(*dev_addr - value) >= 0
 *                    (cuStreamWaitValue32, GEQ): no host involvement, no spinning kernel.
 */
#define MCL_IPC_HANDLE_BYTES 64
int mcl_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out /*host, 64 B*/);
int mcl_peer_free(void* dev_ptr);
int mcl_peer_open(const void* ipc_handle /*host, 64 B*/, void** peer_ptr);
int mcl_peer_close(void* peer_ptr);
int mcl_memcpy_async(void* dst, const void* src, size_t bytes, mcl_stream_t stream);
int mcl_stream_wait_value32(mcl_stream_t stream, const void* dev_addr, uint32_t value);

/*
 * Tuning knobs (host-side, process-global).  opt: 0 = CTAs per launch (0 = all SMs),
 * 1 = row units per wave of the tile plan (0 = heuristic), 2 = route bf16 inputs
 * through the CUDA-core check kernel instead of tcgen05 (tests only), 3 = record per-CTA
 * start/end globaltimer stamps in the first 16 KB of the workspace, 4 = 1 disables CTA pairs
 * (cta_group::2 MMA), 5 = 1 forces the plain all-gather merge in the sharded scan (default:
 * row exchange for world > 2), 6 = 1 times the phases of every scan with CUDA events (debug,
 * synchronises), 7 = 0 plans without tail workers (default 1), 8 = tiles charged per extra
 * segment of a tail worker (default 1), 9 = L2 eviction priority of the TMA loads (bit 0: query
 * tiles evict-last, bit 1: table tiles evict-first), 10 = drift window in tiles (0 = heuristic),
 * 11 = 1 sends one-row-block batches through the streaming top-k path instead of the score-dump +
 * radix-select path (tests, A/B), 12 = 1 turns the joint threshold of a row's slots on (rowstate.cuh; measured: +3 % on C2,
 * -3..-6 % elsewhere, so off by default), 13 = 1 turns the threshold-seeding pre-pass off, 14 = 1 sends
 * k = 1 scans through the general top-k epilogue instead of the running-argmax one (tests, A/B),
 * 15 = what the planner charges a segment's restart in mcl_plan_* (0 = cold top-k filter, 1 = seeded
 * thresholds, 2 = no filter: k = 1 and the seed pass; the scans choose it themselves per call),
 * 16 = 1 keeps the query norms of MCL_SCAN_NORMALIZE_Q out of the scan kernel (tests, A/B),
 * 17 = 1 runs mcl_gather_mean on the register kernels instead of the bulk-copy rings (A/B),
 * 18 = 1 sends one-row-block batches through the two-kernel path (scan with the filter off +
 * selection kernel) instead of the one-launch panel scan (tests, A/B), 19 = 1 merges slots with the
 * streaming-fold kernels only (tests, A/B of merge_rows_kernel);
 * opt 100..102 read the last memset / scan / merge
 * time in ns; opt 103 reads how many drift waits of the scan kernel timed out (group members
 * that lost L2 locality because a peer CTA was not resident) since the process started; opt 104
 * reads how many grid-barrier waits of the panel scan gave up (its outputs are NaN / -1 then).
 * Returns the old value.
 */
int64_t mcl_set_option(int opt, int64_t value);

/* number of kernel launches the library has enqueued in this process (for bench.py) */
int64_t mcl_launch_count(void);

/*
 * Host-only introspection of the tcgen05 scan's tile plan (csrc/plan.h) for a device with
 * `sm_count` SMs; needs no GPU (used by the CPU tests).  plan_out[MCL_PLAN_INTS] =
 * {row blocks, table tiles, K slices, CTAs per worker, workers, row units per wave, row units,
 *  waves, slot stride per row block and column half, slots, grid, drift window, drift counters
 *  per wave, drift counters, nodes of a full wave, nodes of the last wave, then for each of the
 *  4 + 4 node positions (first row unit, row units, first tile, first worker, groups, tiles per
 *  group, first tail tile, tail workers, tail passes)}.
 */
#define MCL_PLAN_INTS 88
int mcl_plan_scan(int64_t Q, int64_t V_local, int64_t D, int sm_count, int32_t* plan_out);
/*
 * Every segment of that plan in execution order per worker, 6 ints each: {worker, row unit,
 * first tile, end tile, slot of the row unit, first drift counter or -1}.  Returns the number
 * of segments (writes at most `cap`), or a negative error code.
 */
int64_t mcl_plan_segments(int64_t Q, int64_t V_local, int64_t D, int sm_count, int32_t* segs_out,
                          int64_t cap);
/* The slots (column halves counted separately) the merge reads for one row block. */
int mcl_plan_row_block_slots(int64_t Q, int64_t V_local, int64_t D, int sm_count, int64_t row_block,
                             int32_t* first_slot, int32_t* num_slots);

#ifdef __cplusplus
}
#endif
#endif /* MCL_H_ */
