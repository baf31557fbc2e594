// HBM-bound row kernels (part (c) of the path): inverse L2 row norms and the multi-token
// gather-mean(-normalise).  One warp per row, 128-bit coalesced loads, fp32 arithmetic.
// Algorithmic bytes: inv-norm reads rows*dim*elt and writes rows*4; gather-mean reads
// nnz*D*elt (+ the CSR arrays) and writes Q*D*elt.
#include <algorithm>
#include <atomic>
#include "kernels.h"
#include "rownorm.cuh"

namespace mcl {

std::atomic<int> g_gather_variant{0};   // library option 17: 1 = register kernels instead of the bulk-copy rings (A/B)

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

// (Measured on B200, profiles/r02_gather_mean_ncu_raw.csv: 2.3 TB/s = 0.36 of the copy peak for 1-4
// ids per row at D = 3584 -- 168 registers leave 12 warps per SM, and a row is a chain of dependent
// round trips (offsets -> id -> row), so the bytes in flight, not the DRAM, bound it; the kernel
// below therefore runs persistent warps that prefetch the next row's CSR entries.  A staging
// ring fed by cp.async.bulk with a single issuing thread was tried and measured SLOWER (1.2 TB/s,
// the consumers' per-row latency moved into the producer's slot waits) and was removed.)
// Register-resident variant for rows of at most 32 * NV 16-byte vectors (D <= 4096 bf16 with
// NV = 16): every table row is read ONCE -- the fp32 means stay in registers between the sum, the
// optional norm and the store, where the generic kernel above re-gathers the rows for its second
// pass -- and a lane has NV independent 16-byte loads in flight per gathered row (a gather is
// latency-bound: few warps fit at this register count, so the parallelism has to come from inside
// the thread).  Same arithmetic, same order: bit-identical outputs.
template <typename T, int NV, int MB>
__global__ void __launch_bounds__(128, MB)
gather_mean_reg_kernel(const T* __restrict__ table, long long V, int D, long long ld,
                       const long long* __restrict__ offsets, const long long* __restrict__ ids,
                       long long Q, int normalize, T* __restrict__ out, long long ld_out,
                       int* __restrict__ bad_flag) {
  constexpr int N = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const int nvec = D / N;
  const int ntail = D - nvec * N;                  // < N ragged columns: lane c owns column nvec*N + c
  // Persistent warps over rows w, w + W, ...  A row is a chain of dependent round trips
  // (offsets -> ids -> table rows): the NEXT row's offsets are requested before this row is
  // processed and its first 32 ids (one coalesced load) once this row's table loads are in flight,
  // so that only the table-row latency is exposed per row.
  const long long W = (long long)gridDim.x * 4;
  long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  long long b = 0, e = 0, my_id = 0;
  if (row < Q) {
    b = __ldg(offsets + row); e = __ldg(offsets + row + 1);
    if (lane < e - b) my_id = __ldg(ids + b + lane);
  }
  for (; row < Q; row += W) {
    const long long next = row + W;
    long long nb = 0, ne = 0, next_id = 0;
    if (next < Q) { nb = __ldg(offsets + next); ne = __ldg(offsets + next + 1); }
    const int n = (int)(e - b);
    const float fn = (float)(n > 0 ? n : 1);
    T* o = out + row * ld_out;
    float acc[NV][N];
    float tail = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int c = 0; c < N; ++c) acc[i][c] = 0.f;
    for (long long j = b; j < e; ++j) {
      long long id = (j - b < 32) ? __shfl_sync(0xffffffffu, my_id, (int)(j - b)) : __ldg(ids + j);
      if (id < 0 || id >= V) { if (bad_flag) *bad_flag = 1; id = id < 0 ? 0 : V - 1; }
      const T* src = table + id * ld;
      uint4 raw[NV];                               // all loads of the row first, widened on use
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        raw[i] = (v < nvec) ? __ldg(reinterpret_cast<const uint4*>(src + (size_t)v * N)) : make_uint4(0u, 0u, 0u, 0u);
      }
      if (lane < ntail) tail += Vec<T>::ld1(src + nvec * N + lane);
      if (j == b && lane < ne - nb) next_id = __ldg(ids + nb + lane);   // (nb / ne have landed by now)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float a[N];
        Vec<T>::widen(raw[i], a);
#pragma unroll
        for (int c = 0; c < N; ++c) acc[i][c] += a[c];   // (+0 past nvec)
      }
    }
    if (n == 0 && lane < ne - nb) next_id = __ldg(ids + nb + lane);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int c = 0; c < N; ++c) {
        acc[i][c] = acc[i][c] / fn;
        ss = fmaf(acc[i][c], acc[i][c], ss);       // (vectors past nvec hold zeros)
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
    b = nb; e = ne; my_id = next_id;
  }
}

// Bulk-copy variant (the default where a table row is a multiple of 16 bytes and fits the register
// budget): a gather is bound by the bytes in flight, and registers hold too few of them -- 12
// warps x one 7 KB row per SM at D = 3584.  Here every warp owns a ring of S shared-memory slots
// of one table row each and keeps S gathered rows in flight with cp.async.bulk (one copy of
// D * elt contiguous bytes per row, completion on the slot's mbarrier); the SAME warp consumes
// them in id order (conflict-free 16-byte shared loads, fp32 accumulators in registers) and
// refills the slot it has just drained, so there is no producer thread to wait for.  A warp takes
// a CONTIGUOUS range of output rows: its CSR entries are contiguous too and are read 32 at a
// time, one chunk ahead (ids and row ends live in registers, handed around by shuffles).
// ~200 KB per SM in flight instead of ~85.  Same arithmetic, same order: bit-identical outputs.
__device__ __forceinline__ void bulk_row_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// Exact x / n without the compiler's IEEE division (~40 dependent instructions per element): with
// r = RN(1 / n) computed once per row, q = RN(x r) and one Markstein correction
// q' = RN(q + RN(x - n q) r) is the correctly rounded quotient (the remainder is exact under FMA)
// whenever nothing under- or overflows; outside that range the true division is taken (a warp
// vote, practically never).  Bit-identical to x / n: tests compare with the register kernels,
// which divide, and oracle/check_div_identity.py samples the identity on the host for n = 1 .. 1000.
#ifndef MCL_GM_UNROLL
#define MCL_GM_UNROLL 1   // column-loop unroll of the ring path (1, 2 and 4 measured: profiles/README.md)
#endif
constexpr int kGmUnroll = MCL_GM_UNROLL;
template <typename T, int NV>
__global__ void __launch_bounds__(NV > 8 ? 256 : 512, 1)
gather_mean_bulk_kernel(const T* __restrict__ table, long long V, int D, long long ld,
                        const long long* __restrict__ offsets, const long long* __restrict__ ids,
                        long long Q, int normalize, T* __restrict__ out, long long ld_out,
                        int* __restrict__ bad_flag, int S, uint32_t slot_stride) {
  constexpr int N = Vec<T>::N;
  extern __shared__ __align__(128) uint8_t gm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nvec = D / N;                            // D % N == 0 on this path (rows are whole 16-byte vectors)
  const uint32_t rowbytes = (uint32_t)D * sizeof(T);
  const uint32_t base = smem_u32(gm_smem);
  const uint32_t bar0 = base + (uint32_t)(warp * S) * 8u;
  const uint32_t slot0 = base + (((uint32_t)(nwarps * S) * 8u + 127u) & ~127u) + (uint32_t)(warp * S) * slot_stride;
  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncwarp();

  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  const long long r_begin = gw * Q / GW, r_end = (gw + 1) * Q / GW;
  if (r_begin >= r_end) return;
  const long long g0 = __ldg(offsets + r_begin), g1 = __ldg(offsets + r_end);
  const long long total = g1 - g0;                   // gathered rows of this warp

  // CSR entries, 32 at a time and one chunk ahead: ids of gathered rows idbase .. idbase+31 (relative
  // to g0) in id_cur, the next 32 in id_nxt; ends of output rows rbase .. rbase+31 in end_cur / end_nxt
  long long idbase = 0, rbase = r_begin;
  auto load_ids = [&](long long rel) { return (rel + lane < total) ? __ldg(ids + g0 + rel + lane) : 0ll; };
  auto load_ends = [&](long long r) { return (r + lane < r_end) ? __ldg(offsets + r + lane + 1) : g1; };
  long long id_cur = load_ids(0), id_nxt = load_ids(32);
  long long end_cur = load_ends(r_begin), end_nxt = load_ends(r_begin + 32);

  long long issued = 0;                              // gathered rows requested so far
  int in_chunk = 0;                                  // = issued - idbase
  uint32_t is = 0;                                   // slot the next request goes to (= issued % S)
  uint32_t cs = 0, cphase = 0;                       // slot / mbarrier phase the next row is consumed from
  auto issue_one = [&]() {                           // (warp-uniform)
    if (in_chunk == 32) { id_cur = id_nxt; idbase += 32; in_chunk = 0; id_nxt = load_ids(idbase + 32); }
    long long id = __shfl_sync(0xffffffffu, id_cur, in_chunk);
    if (id < 0 || id >= V) { if (bad_flag && lane == 0) *bad_flag = 1; id = id < 0 ? 0 : V - 1; }
    if (lane == 0) {
      mbar_expect_tx(bar0 + 8u * is, rowbytes);
      bulk_row_g2s(slot0 + is * slot_stride, table + id * ld, rowbytes, bar0 + 8u * is);
    }
    ++issued; ++in_chunk;
    if (++is == (uint32_t)S) is = 0;
  };
  while (issued < total && issued < S) issue_one();

  long long b = g0;
  for (long long row = r_begin; row < r_end; ++row) {
    if (row - rbase == 32) { end_cur = end_nxt; rbase += 32; end_nxt = load_ends(rbase + 32); }
    const long long e = __shfl_sync(0xffffffffu, end_cur, (int)(row - rbase));   // (row end, absolute)
    const int n = (int)(e - b);
    const float fn = (float)(n > 0 ? n : 1);
    if (n <= S) {
      // ---- the row's gathered rows all fit the ring (the common case: 1-5 tokens per concept).
      // Column loops with a small body instead of NV x N unrolled accumulators: the unrolled form
      // is ~4 k instructions executed once per output row, and ncu showed it bound by instruction
      // fetch (stall "no instruction" 4.3 per issue, 0.23 IPC), not by memory.  With
      // normalisation the sums are formed twice (norm, then scale + store): shared-memory reads
      // are cheaper than keeping D fp32 means.
      uint32_t sw = cs, pw = cphase;
      for (int j = 0; j < n; ++j) {
        mbar_wait(bar0 + 8u * sw, pw);
        if (++sw == (uint32_t)S) { sw = 0; pw ^= 1u; }
      }
      const float rn = 1.0f / fn;
      const int nvl = (nvec + 31) >> 5;
      T* o = out + row * ld_out;
      float inv = 1.f;
      // With normalisation the fp32 means of rows of two or more tokens are parked in the slots they
      // came from between the two passes (a lane overwrites exactly the 16 bytes per slot it has just
      // read: 8 bf16 in, 2 x 4 fp32 out), so the second pass is a load, a scale and a store.
      const bool stash = normalize && n >= 2;
      const uint32_t s_a = slot0 + cs * slot_stride;
      const uint32_t s_b = slot0 + (cs + 1 == (uint32_t)S ? 0u : cs + 1) * slot_stride;
      for (int pass = normalize ? 0 : 1; pass < 2; ++pass) {
        float ss = 0.f;
#pragma unroll kGmUnroll
        for (int i = 0; i < nvl; ++i) {
          const int v = lane + 32 * i;
          float acc[N];
#pragma unroll
          for (int c = 0; c < N; ++c) acc[c] = 0.f;
          if (pass == 1 && stash) {
            if (v < nvec) {
              const uint4 lo = lds_v4(s_a + (uint32_t)v * 16u);
              acc[0] = __uint_as_float(lo.x); acc[1] = __uint_as_float(lo.y);
              acc[2] = __uint_as_float(lo.z); acc[3] = __uint_as_float(lo.w);
              if (N == 8) {
                const uint4 hi = lds_v4(s_b + (uint32_t)v * 16u);
                acc[N - 4] = __uint_as_float(hi.x); acc[N - 3] = __uint_as_float(hi.y);
                acc[N - 2] = __uint_as_float(hi.z); acc[N - 1] = __uint_as_float(hi.w);
              }
            }
          } else {
            if (v < nvec) {
              uint32_t sj = cs;
              for (int j = 0; j < n; ++j) {
                float x[N];
                Vec<T>::widen(lds_v4(slot0 + sj * slot_stride + (uint32_t)v * 16u), x);
#pragma unroll
                for (int c = 0; c < N; ++c) acc[c] += x[c];
                if (++sj == (uint32_t)S) sj = 0;
              }
            }
            if (n > 1) {                             // (uniform) exact x / n, see the note above the kernel
              float q[N];
              bool odd = false;
#pragma unroll
              for (int c = 0; c < N; ++c) {
                const float q0 = acc[c] * rn;
                q[c] = fmaf(fmaf(-fn, q0, acc[c]), rn, q0);
                const float ax = fabsf(acc[c]);
                odd = odd || (!(ax > 1e-30f && ax < 1e30f) && acc[c] != 0.f);
              }
              if (__any_sync(0xffffffffu, odd)) {
#pragma unroll
                for (int c = 0; c < N; ++c) q[c] = acc[c] / fn;
              }
#pragma unroll
              for (int c = 0; c < N; ++c) acc[c] = q[c];
            }
            if (pass == 0 && stash && v < nvec) {
              sts_v4(s_a + (uint32_t)v * 16u, make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]),
                                                         __float_as_uint(acc[2]), __float_as_uint(acc[3])));
              if (N == 8)
                sts_v4(s_b + (uint32_t)v * 16u, make_uint4(__float_as_uint(acc[N - 4]), __float_as_uint(acc[N - 3]),
                                                           __float_as_uint(acc[N - 2]), __float_as_uint(acc[N - 1])));
            }
          }
          if (pass == 0) {
#pragma unroll
            for (int c = 0; c < N; ++c) ss = fmaf(acc[c], acc[c], ss);
          } else if (v < nvec) {
            if (normalize) {
#pragma unroll
              for (int c = 0; c < N; ++c) acc[c] *= inv;
            }
            Vec<T>::store(o + (size_t)v * N, acc);
          }
        }
        if (pass == 0) {
          const float nrm = sqrtf(warp_sum(ss));
          inv = (nrm < kTinyNorm) ? 1.0f : 1.0f / nrm;
        }
      }
      cs = sw; cphase = pw;
      __syncwarp();                                  // every lane has read the slots: refill them
      for (int j = 0; j < n; ++j)
        if (issued < total) issue_one();
      b = e;
      continue;
    }
    // ---- long rows (more gathered rows than slots): fp32 accumulators in registers, one row at a time
    float acc[NV][N];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int c = 0; c < N; ++c) acc[i][c] = 0.f;
    for (long long j = b; j < e; ++j) {
      mbar_wait(bar0 + 8u * cs, cphase);
      const uint32_t src = slot0 + cs * slot_stride + (uint32_t)lane * 16u;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < nvec) {
          float a[N];
          Vec<T>::widen(lds_v4(src + 512u * i), a);
#pragma unroll
          for (int c = 0; c < N; ++c) acc[i][c] += a[c];
        }
      }
      if (++cs == (uint32_t)S) { cs = 0; cphase ^= 1u; }
      __syncwarp();                                  // every lane has read the slot: refill it
      if (issued < total) issue_one();
    }
    b = e;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int c = 0; c < N; ++c) {
        acc[i][c] = acc[i][c] / fn;
        ss = fmaf(acc[i][c], acc[i][c], ss);         // (vectors past nvec hold zeros)
      }
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

template <typename T>
static cudaError_t launch_gather_mean_t(const T* table, int64_t V, int64_t D, int64_t ld, const int64_t* offsets,
                                        const int64_t* ids, int64_t Q, int normalize, T* out, int64_t ld_out,
                                        int* bad_flag, cudaStream_t s) {
  const int64_t nvec = D / Vec<T>::N;
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  // bulk-copy rings (the default): one CTA per SM, 16 warps (rows up to 4 KB) or 8, S slots per warp
  // out of ~200 KB of shared memory; library option 17 = 1 selects the register kernels below (A/B)
  const size_t rowbytes = (size_t)D * sizeof(T);
  if (!g_gather_variant.load() && rowbytes % 16 == 0 && nvec <= 32 * 16) {
    const int NVr = nvec <= 32 * 4 ? 4 : (nvec <= 32 * 8 ? 8 : 16);
    const int threads = NVr > 8 ? 256 : 512, warps = threads / 32;
    const uint32_t slot_stride = (uint32_t)((rowbytes + 127) & ~(size_t)127);
    int S = (int)((226u * 1024u - 1024u) / ((size_t)warps * slot_stride));
    S = S > 8 ? 8 : S;
    if (S >= 2) {
      const size_t smem = (((size_t)warps * S * 8 + 127) & ~(size_t)127) + (size_t)warps * S * slot_stride;
      const long long want = (Q + warps - 1) / warps;
      const unsigned grid = (unsigned)(want < sm ? want : sm);
#define MCL_GM_BULK(NV)                                                                                        \
      do {                                                                                                     \
        static std::atomic<size_t> attr[64];                                                                   \
        if (dev >= 0 && dev < 64 && attr[dev].load() < smem) {                                                 \
          cudaError_t e = cudaFuncSetAttribute(gather_mean_bulk_kernel<T, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               227 * 1024);                                                    \
          if (e != cudaSuccess) return e;                                                                      \
          attr[dev].store(227 * 1024);                                                                         \
        }                                                                                                      \
        gather_mean_bulk_kernel<T, NV><<<grid, threads, smem, s>>>(table, V, (int)D, ld, (const long long*)offsets, \
            (const long long*)ids, Q, normalize, out, ld_out, bad_flag, S, slot_stride);                       \
      } while (0)
      if (NVr == 4) MCL_GM_BULK(4); else if (NVr == 8) MCL_GM_BULK(8); else MCL_GM_BULK(16);
#undef MCL_GM_BULK
      return cudaGetLastError();
    }
  }
  // persistent warps: as many blocks as fit the chip at this kernel's register count
#define MCL_GM_REG(NV, MB)                                                                          \
  do {                                                                                              \
    int occ = 1;                                                                                    \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gather_mean_reg_kernel<T, NV, MB>, 128, 0); \
    const long long want = (Q + 3) / 4, fit = (long long)sm * (occ > 0 ? occ : 1);                  \
    gather_mean_reg_kernel<T, NV, MB><<<(unsigned)(want < fit ? want : fit), 128, 0, s>>>(          \
        table, V, (int)D, ld, (const long long*)offsets, (const long long*)ids, Q, normalize, out, ld_out, bad_flag); \
  } while (0)
  // minimum blocks per SM = the register cap: 8 / 5 / 3 blocks of four warps keep the accumulators of
  // NV vectors per lane in registers (g_gather_variant = 1: two blocks, 255 registers, no spills -- A/B)
  if (nvec <= 32 * 4) MCL_GM_REG(4, 8);
  else if (nvec <= 32 * 8) MCL_GM_REG(8, 5);
  else if (nvec <= 32 * 16) MCL_GM_REG(16, 3);
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
