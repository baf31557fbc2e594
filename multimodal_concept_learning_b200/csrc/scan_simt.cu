// fp32-accumulate CHECK PATH of the concept scan (MCL_DTYPE_F32, and bf16 inputs when the
// caller forces it): plain CUDA-core FMAs, one thread per query row, the same epilogue
// (rowstate.cuh) and the same slot format as the tcgen05 kernel, so the two engines can be
// compared on the GPU at sizes the CPU oracle cannot reach.  Not a product path for speed:
// it exists because north_star asks for an fp32-accumulate check (rtol 1e-4).
#include "rowstate.cuh"
#include "kernels.h"

namespace mcl {

constexpr int kSimtK = 32;  // K slice

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}

template <typename T>
__global__ void __launch_bounds__(kBlockM, 2)
scan_simt_kernel(const T* __restrict__ q, const T* __restrict__ table, int Q, int V, int D,
                 long long ldq, long long ldt, const float* __restrict__ inv_q,
                 const float* __restrict__ inv_t, float scale, int k, long long index_base,
                 const long long* __restrict__ labels, SlotView sv, int nsplit,
                 int chunks_per_split, float* __restrict__ dbg_scores, float softcap) {
  __shared__ float qs[kSimtK][kBlockM + 1];           // transposed query slice
  __shared__ __align__(16) float ts[kChunk][kSimtK + 4];  // table slice, row = table row

  const int rb = blockIdx.x, split = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = rb * kBlockM + tid;
  const int slot = rb * nsplit + split;

  const int num_chunks = (V + kChunk - 1) / kChunk;
  const int c_begin = split * chunks_per_split;
  const int c_end = min(num_chunks, c_begin + chunks_per_split);

  RowState st;
  uint2* slot_buf = sv.cand + (size_t)slot * kBlockM * kCandCap;
  st.reset(slot_buf + (size_t)tid * kCandCap);
  uint2* warp_buf = slot_buf + (size_t)(warp * 32) * kCandCap;

  const float rs = (row < Q && inv_q ? inv_q[row] : 1.f) * scale;
  const float a = (softcap > 0.f ? softcap : rs) * kLog2e;
  const float rc = softcap > 0.f ? rs / softcap : 0.f;
  int lab_local = -1;
  if (labels && row < Q) {
    const long long l = labels[row] - index_base;
    if (labels[row] != -100 && l >= 0 && l < V) lab_local = (int)l;
  }

  for (int c = c_begin; c < c_end; ++c) {
    const int col0 = c * kChunk;
    float acc[kChunk];
#pragma unroll
    for (int i = 0; i < kChunk; ++i) acc[i] = 0.f;

    for (int d0 = 0; d0 < D; d0 += kSimtK) {
      __syncthreads();
      // query slice: warp w loads rows w*32 .. w*32+31, lane = d (coalesced along D)
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int r_local = warp * 32 + rr;
        const int r_glob = rb * kBlockM + r_local;
        const int d = d0 + lane;
        qs[lane][r_local] = (r_glob < Q && d < D) ? to_f32(q[(size_t)r_glob * ldq + d]) : 0.f;
      }
      // table slice: warp w loads table rows w*8 .. w*8+7 of the chunk
#pragma unroll
      for (int cc = 0; cc < kChunk / 4; ++cc) {
        const int c_local = warp * (kChunk / 4) + cc;
        const int t_row = col0 + c_local;
        const int d = d0 + lane;
        ts[c_local][lane] = (t_row < V && d < D) ? to_f32(table[(size_t)t_row * ldt + d]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int d4 = 0; d4 < kSimtK; d4 += 4) {
        const float q0 = qs[d4 + 0][tid], q1 = qs[d4 + 1][tid], q2 = qs[d4 + 2][tid],
                    q3 = qs[d4 + 3][tid];
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
          const float4 t4 = *reinterpret_cast<const float4*>(&ts[i][d4]);
          acc[i] = fmaf(q0, t4.x, acc[i]);
          acc[i] = fmaf(q1, t4.y, acc[i]);
          acc[i] = fmaf(q2, t4.z, acc[i]);
          acc[i] = fmaf(q3, t4.w, acc[i]);
        }
      }
    }
    // per-table-row scale, then the shared epilogue
    const int n_valid = min(kChunk, V - col0);
    if (inv_t) {
#pragma unroll
      for (int i = 0; i < kChunk; ++i) acc[i] *= (i < n_valid) ? __ldg(inv_t + col0 + i) : 1.f;
    }
    if (dbg_scores && row < Q) {
#pragma unroll
      for (int i = 0; i < kChunk; ++i)
        if (i < n_valid)
          dbg_scores[(size_t)row * V + col0 + i] =
              softcap > 0.f ? softcap * tanhf(acc[i] * rc) : acc[i] * rs;
    }
    if (softcap > 0.f) {
      if (n_valid == kChunk) row_process_chunk<false, true>(st, acc, col0, kChunk, a, lab_local, rc);
      else row_process_chunk<true, true>(st, acc, col0, n_valid, a, lab_local, rc);
    } else {
      if (n_valid == kChunk) row_process_chunk<false, false>(st, acc, col0, kChunk, a, lab_local, 0.f);
      else row_process_chunk<true, false>(st, acc, col0, n_valid, a, lab_local, 0.f);
    }
    __syncwarp();
    if (c + 1 < c_end) warp_compact_rows(st, k, warp_buf, lane, nullptr);   // not after the last chunk
  }
  row_flush(st, rs, softcap, sv.cnt + (size_t)slot * kBlockM + tid, sv.stats + (size_t)slot * kBlockM + tid);
}

cudaError_t launch_scan_simt(const ScanArgs& a, const SlotView& sv, int nsplit, cudaStream_t s) {
  const int num_rb = (int)((a.Q + kBlockM - 1) / kBlockM);
  const int num_chunks = (int)((a.V + kChunk - 1) / kChunk);
  const int cps = (num_chunks + nsplit - 1) / nsplit;
  dim3 grid(num_rb, nsplit);
  if (a.dtype == 1) {
    scan_simt_kernel<float><<<grid, kBlockM, 0, s>>>(
        (const float*)a.q, (const float*)a.table, (int)a.Q, (int)a.V, (int)a.D, a.ldq, a.ldt,
        a.inv_q, a.inv_t, a.scale, a.k, a.index_base, (const long long*)a.labels, sv, nsplit, cps,
        a.dbg_scores, a.softcap);
  } else {
    scan_simt_kernel<__nv_bfloat16><<<grid, kBlockM, 0, s>>>(
        (const __nv_bfloat16*)a.q, (const __nv_bfloat16*)a.table, (int)a.Q, (int)a.V, (int)a.D,
        a.ldq, a.ldt, a.inv_q, a.inv_t, a.scale, a.k, a.index_base, (const long long*)a.labels, sv,
        nsplit, cps, a.dbg_scores, a.softcap);
  }
  return cudaGetLastError();
}

}  // namespace mcl
