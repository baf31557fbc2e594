// HBM-bound row kernels (part (c) of the path): inverse L2 row norms and the multi-token
// gather-mean(-normalise).  One warp per row, 128-bit coalesced loads, fp32 arithmetic.
// Algorithmic bytes: inv-norm reads rows*dim*elt and writes rows*4; gather-mean reads
// nnz*D*elt (+ the CSR arrays) and writes Q*D*elt.
#include <algorithm>
#include <atomic>
#include "kernels.h"
#include "rownorm.cuh"

namespace mcl {

template <typename T>
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(const T* __restrict__ x, long long rows, int dim, long long ld,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float inv = warp_row_inv_norm<T>(x + row * ld, dim, lane);
  if (lane == 0) out[row] = inv;
}

// out[i,:] = mean of the gathered rows (fp32 sum in id order, true division, one rounding);
// with `normalize` the fp32 mean is scaled by 1/||mean|| (zero rows stay zero) first.
template <typename T>
__global__ void __launch_bounds__(256)
gather_mean_kernel(const T* __restrict__ table, long long V, int D, long long ld,
                   const long long* __restrict__ offsets, const long long* __restrict__ ids,
                   long long Q, int normalize, T* __restrict__ out, long long ld_out,
                   int* __restrict__ bad_flag) {
  constexpr int N = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= Q) return;
  const long long b = offsets[row], e = offsets[row + 1];
  const int n = (int)(e - b);
  const float fn = (float)(n > 0 ? n : 1);
  const int nvec = D / N;
  T* o = out + row * ld_out;

  float inv = 1.f;
  for (int pass = normalize ? 0 : 1; pass < 2; ++pass) {
    float ss = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float acc[N];
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = 0.f;
      for (long long j = b; j < e; ++j) {
        long long id = ids[j];
        if (id < 0 || id >= V) { if (bad_flag) *bad_flag = 1; id = id < 0 ? 0 : V - 1; }
        float a[N];
        Vec<T>::load(table + id * ld + (size_t)v * N, a);
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] += a[i];
      }
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = acc[i] / fn;
      if (pass == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) ss = fmaf(acc[i], acc[i], ss);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] *= inv;
        Vec<T>::store(o + (size_t)v * N, acc);
      }
    }
    for (int c = nvec * N + lane; c < D; c += 32) {  // ragged tail of D
      float acc = 0.f;
      for (long long j = b; j < e; ++j) {
        long long id = ids[j];
        if (id < 0 || id >= V) { if (bad_flag) *bad_flag = 1; id = id < 0 ? 0 : V - 1; }
        acc += Vec<T>::ld1(table + id * ld + c);
      }
      acc = acc / fn;
      if (pass == 0) ss = fmaf(acc, acc, ss); else Vec<T>::st1(o + c, acc * inv);
    }
    if (pass == 0) {
      const float nrm = sqrtf(warp_sum(ss));
      inv = (nrm < kTinyNorm) ? 1.0f : 1.0f / nrm;
    }
  }
}

// Register-resident variant for rows of at most 32 * NV 16-byte vectors (D <= 4096 bf16 with
// NV = 16): every table row is read ONCE -- the fp32 means stay in registers between the sum, the
// optional norm and the store, where the generic kernel above re-gathers the rows for its second
// pass -- and a lane has NV independent 16-byte loads in flight per gathered row (a gather is
// latency-bound: few warps fit at this register count, so the parallelism has to come from inside
// the thread).  Same arithmetic, same order: bit-identical outputs.
template <typename T, int NV>
__global__ void __launch_bounds__(128)
gather_mean_reg_kernel(const T* __restrict__ table, long long V, int D, long long ld,
                       const long long* __restrict__ offsets, const long long* __restrict__ ids,
                       long long Q, int normalize, T* __restrict__ out, long long ld_out,
                       int* __restrict__ bad_flag) {
  constexpr int N = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= Q) return;
  const long long b = offsets[row], e = offsets[row + 1];
  const int n = (int)(e - b);
  const float fn = (float)(n > 0 ? n : 1);
  const int nvec = D / N;
  const int ntail = D - nvec * N;                  // < N ragged columns: lane c owns column nvec*N + c
  T* o = out + row * ld_out;
  float acc[NV][N];
  float tail = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int c = 0; c < N; ++c) acc[i][c] = 0.f;
  for (long long j = b; j < e; ++j) {
    long long id = __ldg(ids + j);
    if (id < 0 || id >= V) { if (bad_flag) *bad_flag = 1; id = id < 0 ? 0 : V - 1; }
    const T* src = table + id * ld;
    uint4 raw[NV];                                 // all loads of the row first, widened on use
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      raw[i] = (v < nvec) ? __ldg(reinterpret_cast<const uint4*>(src + (size_t)v * N)) : make_uint4(0u, 0u, 0u, 0u);
    }
    if (lane < ntail) tail += Vec<T>::ld1(src + nvec * N + lane);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float a[N];
      Vec<T>::widen(raw[i], a);
#pragma unroll
      for (int c = 0; c < N; ++c) acc[i][c] += a[c];   // (+0 past nvec)
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
      acc[i][c] = acc[i][c] / fn;
      ss = fmaf(acc[i][c], acc[i][c], ss);         // (vectors past nvec hold zeros)
    }
  }
  tail = tail / fn;
  float inv = 1.f;
  if (normalize) {
    if (lane < ntail) ss = fmaf(tail, tail, ss);
    const float nrm = sqrtf(warp_sum(ss));
    inv = (nrm < kTinyNorm) ? 1.0f : 1.0f / nrm;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      if (normalize) {
#pragma unroll
        for (int c = 0; c < N; ++c) acc[i][c] *= inv;
      }
      Vec<T>::store(o + (size_t)v * N, acc[i]);
    }
  }
  if (lane < ntail) Vec<T>::st1(o + nvec * N + lane, normalize ? tail * inv : tail);
}

// Cross-entropy from the scan's row statistics (m, s, sum_z, z_label): per row
// (1-eps)(lse - z_label) + eps (lse - sum_z / V), 0 on rows whose label is -100, and the mean over
// the other rows (`F.cross_entropy(..., ignore_index=-100, label_smoothing=eps)`, 'mean').  One
// block, fixed summation order (deterministic); loss_mean[0] = mean (NaN without a valid row, as
// torch), loss_mean[1] = number of valid rows.
__global__ void __launch_bounds__(1024)
ce_from_stats_kernel(const float4* __restrict__ stats, const long long* __restrict__ labels, long long Q,
                     float eps, float vocab, float* __restrict__ loss_rows, float* __restrict__ loss_mean) {
  __shared__ double s_sum[32];
  __shared__ long long s_cnt[32];
  double sum = 0.0;
  long long cnt = 0;
  for (long long r = threadIdx.x; r < Q; r += blockDim.x) {
    const float4 st = stats[r];
    const bool valid = labels[r] != -100;
    const float lse = st.x + logf(st.y);
    const float l = valid ? (1.f - eps) * (lse - st.w) + eps * (lse - st.z / vocab) : 0.f;
    if (loss_rows) loss_rows[r] = l;
    sum += (double)l;
    cnt += valid ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; }
  __syncthreads();
  if (warp == 0) {
    sum = lane < (int)(blockDim.x >> 5) ? s_sum[lane] : 0.0;
    cnt = lane < (int)(blockDim.x >> 5) ? s_cnt[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
      loss_mean[0] = cnt > 0 ? (float)(sum / (double)cnt) : __int_as_float(0x7fc00000);
      loss_mean[1] = (float)cnt;
    }
  }
}

cudaError_t launch_ce_from_stats(const float* row_stats, const int64_t* labels, int64_t Q,
                                 float label_smoothing, int64_t vocab, float* loss_rows,
                                 float* loss_mean, cudaStream_t s) {
  ce_from_stats_kernel<<<1, 1024, 0, s>>>((const float4*)row_stats, (const long long*)labels,
                                          (long long)Q, label_smoothing, (float)vocab,
                                          loss_rows, loss_mean);
  return cudaGetLastError();
}

cudaError_t launch_row_inv_norm(const void* x, int dtype, int64_t rows, int64_t dim, int64_t ld,
                                float* out, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (dtype == 0)
    row_inv_norm_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, rows, (int)dim, ld, out);
  else
    row_inv_norm_kernel<float><<<grid, 256, 0, s>>>((const float*)x, rows, (int)dim, ld, out);
  return cudaGetLastError();
}

// Bulk-copy variant (the default for rows of at least 512 bytes): a gather is bound by the bytes
// it keeps in flight, and registers are the wrong place to park them (the register-resident kernel
// above reaches 16 % occupancy and 2.3 TB/s).  Here ONE elected thread per CTA streams every
// gathered table row into a ring of shared-memory slots with `cp.async.bulk` (TMA's 1-D form,
// completion on an mbarrier) -- up to ~190 KB in flight per SM -- and seven consumer warps, one
// output row each, add the rows up out of shared memory in id order, normalise and store.  The
// slot of source row j is (j - first j of the CTA) mod S: producer and consumers derive it from
// the CSR offsets alone.  Persistent CTAs over contiguous ranges of output rows.
constexpr int kGmThreads = 256;
constexpr int kGmConsumers = kGmThreads / 32 - 1;
constexpr uint32_t kGmRingBytes = 192 * 1024;

__device__ __forceinline__ void bulk_load_row(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <typename T, int NV>
__global__ void __launch_bounds__(kGmThreads, 1)
gather_mean_bulk_kernel(const T* __restrict__ table, long long V, int D, long long ld,
                        const long long* __restrict__ offsets, const long long* __restrict__ ids,
                        long long Q, int normalize, T* __restrict__ out, long long ld_out,
                        int* __restrict__ bad_flag, int rows_per_cta, int nslots, uint32_t slot_bytes) {
  constexpr int N = Vec<T>::N;
  extern __shared__ uint8_t gm_smem_raw[];
  const uint32_t ring = (smem_u32(gm_smem_raw) + 127u) & ~127u;
  uint8_t* ring_gen = gm_smem_raw + (ring - smem_u32(gm_smem_raw));
  const uint32_t bars = ring + (uint32_t)nslots * slot_bytes;        // full[nslots], empty[nslots]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const long long row1 = row0 + rows_per_cta < Q ? row0 + rows_per_cta : Q;
  if (row0 >= row1) return;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; ++s) { mbar_init(bars + 8u * s, 1); mbar_init(bars + 8u * (nslots + s), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long j0 = offsets[row0];                  // first source row of this CTA
  const uint32_t row_bytes = (uint32_t)D * sizeof(T);
  if (warp == 0) {
    // ---- producer: every source row of rows [row0, row1), in CSR order; the warp fetches 32 ids
    // at a time (one coalesced load), lane 0 issues the copies ----
    const long long j1 = offsets[row1];
    for (long long jb = j0; jb < j1; jb += 32) {
      long long my_id = (jb + lane < j1) ? ids[jb + lane] : 0;
      if (my_id < 0 || my_id >= V) { if (bad_flag) *bad_flag = 1; my_id = my_id < 0 ? 0 : V - 1; }
      const int cnt = (int)(j1 - jb < 32 ? j1 - jb : 32);
      for (int t = 0; t < cnt; ++t) {
        const long long id = __shfl_sync(0xffffffffu, my_id, t);
        if (lane == 0) {
          const long long c = jb + t - j0;
          const int slot = (int)(c % nslots);
          const uint32_t phase = (uint32_t)((c / nslots) & 1);
          mbar_wait(bars + 8u * (nslots + slot), phase ^ 1u);        // consumers have freed the slot
          mbar_expect_tx(bars + 8u * slot, row_bytes);
          bulk_load_row(ring + (uint32_t)slot * slot_bytes, table + id * ld, row_bytes, bars + 8u * slot);
        }
        __syncwarp();
      }
    }
    return;
  }
  // ---- consumers: warp w takes output rows row0 + (w-1), + kGmConsumers, ... ----
  const int nvec = D / N;                              // (D * sizeof(T) is a multiple of 16: no ragged tail)
  for (long long row = row0 + (warp - 1); row < row1; row += kGmConsumers) {
    const long long b = offsets[row], e = offsets[row + 1];
    const float fn = (float)(e > b ? e - b : 1);
    float acc[NV][N];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int c = 0; c < N; ++c) acc[i][c] = 0.f;
    for (long long j = b; j < e; ++j) {
      const long long c = j - j0;
      const int slot = (int)(c % nslots);
      mbar_wait(bars + 8u * slot, (uint32_t)((c / nslots) & 1));
      const uint4* src = reinterpret_cast<const uint4*>(ring_gen + (size_t)slot * slot_bytes);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
          float a[N];
          Vec<T>::widen(src[v], a);
#pragma unroll
          for (int cc = 0; cc < N; ++cc) acc[i][cc] += a[cc];
        }
      }
      __syncwarp();                                    // every lane has read the slot
      if (lane == 0) mbar_arrive(bars + 8u * (nslots + slot));
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int c = 0; c < N; ++c) {
        acc[i][c] = acc[i][c] / fn;
        ss = fmaf(acc[i][c], acc[i][c], ss);
      }
    float inv = 1.f;
    if (normalize) {
      const float nrm = sqrtf(warp_sum(ss));
      inv = (nrm < kTinyNorm) ? 1.0f : 1.0f / nrm;
    }
    T* o = out + row * ld_out;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        if (normalize) {
#pragma unroll
          for (int c = 0; c < N; ++c) acc[i][c] *= inv;
        }
        Vec<T>::store(o + (size_t)v * N, acc[i]);
      }
    }
  }
}

template <typename T, int NV>
static cudaError_t launch_gather_mean_bulk(const T* table, int64_t V, int64_t D, int64_t ld, const int64_t* offsets,
                                           const int64_t* ids, int64_t Q, int normalize, T* out, int64_t ld_out,
                                           int* bad_flag, cudaStream_t s) {
  const uint32_t slot_bytes = (uint32_t)((D * sizeof(T) + 127) & ~(size_t)127);
  int nslots = (int)(kGmRingBytes / slot_bytes);
  if (nslots > 48) nslots = 48;
  const size_t smem = (size_t)nslots * slot_bytes + 16u * nslots + 256;
  static std::atomic<bool> attr_set[64];
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  auto kfn = gather_mean_bulk_kernel<T, NV>;
  if (dev >= 0 && dev < 64 && !attr_set[dev].load()) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev].store(true);
  }
  int grid = (int)std::min<int64_t>(sm, (Q + kGmConsumers - 1) / kGmConsumers);
  const int rows_per_cta = (int)((Q + grid - 1) / grid);
  grid = (int)((Q + rows_per_cta - 1) / rows_per_cta);
  kfn<<<grid, kGmThreads, smem, s>>>(table, V, (int)D, ld, (const long long*)offsets, (const long long*)ids, Q,
                                     normalize, out, ld_out, bad_flag, rows_per_cta, nslots, slot_bytes);
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_gather_mean_t(const T* table, int64_t V, int64_t D, int64_t ld, const int64_t* offsets,
                                        const int64_t* ids, int64_t Q, int normalize, T* out, int64_t ld_out,
                                        int* bad_flag, cudaStream_t s) {
  const int64_t nvec = D / Vec<T>::N;
  // rows whose bytes are a multiple of 16 (no ragged tail), long enough for the staging to pay
  // and short enough for a lane's share to stay in registers: the bulk-copy kernel
  if ((D * sizeof(T)) % 16 == 0 && D * sizeof(T) >= 512 && nvec <= 32 * 16 && Q >= 64) {
    if (nvec <= 32 * 4) return launch_gather_mean_bulk<T, 4>(table, V, D, ld, offsets, ids, Q, normalize, out, ld_out, bad_flag, s);
    if (nvec <= 32 * 8) return launch_gather_mean_bulk<T, 8>(table, V, D, ld, offsets, ids, Q, normalize, out, ld_out, bad_flag, s);
    return launch_gather_mean_bulk<T, 16>(table, V, D, ld, offsets, ids, Q, normalize, out, ld_out, bad_flag, s);
  }
  const unsigned grid4 = (unsigned)((Q + 3) / 4);
#define MCL_GM_REG(NV)                                                                              \
  gather_mean_reg_kernel<T, NV><<<grid4, 128, 0, s>>>(table, V, (int)D, ld, (const long long*)offsets, \
                                                      (const long long*)ids, Q, normalize, out, ld_out, bad_flag)
  if (nvec <= 32 * 4) MCL_GM_REG(4);
  else if (nvec <= 32 * 8) MCL_GM_REG(8);
  else if (nvec <= 32 * 16) MCL_GM_REG(16);
  else
    gather_mean_kernel<T><<<(unsigned)((Q + 7) / 8), 256, 0, s>>>(table, V, (int)D, ld, (const long long*)offsets,
                                                                 (const long long*)ids, Q, normalize, out, ld_out,
                                                                 bad_flag);
#undef MCL_GM_REG
  return cudaGetLastError();
}

cudaError_t launch_gather_mean(const void* table, int dtype, int64_t V, int64_t D, int64_t ld,
                               const int64_t* offsets, const int64_t* ids, int64_t Q,
                               int normalize, void* out, int64_t ld_out, int* bad_flag,
                               cudaStream_t s) {
  if (Q == 0) return cudaSuccess;
  if (dtype == 0)
    return launch_gather_mean_t<__nv_bfloat16>((const __nv_bfloat16*)table, V, D, ld, offsets, ids, Q, normalize,
                                               (__nv_bfloat16*)out, ld_out, bad_flag, s);
  return launch_gather_mean_t<float>((const float*)table, V, D, ld, offsets, ids, Q, normalize, (float*)out,
                                     ld_out, bad_flag, s);
}

}  // namespace mcl
