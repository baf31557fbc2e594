// Dense bf16 GEMM on tcgen05 for the BACKWARD of the fused cross-entropy:
//
//     C[M,N] (fp32) (+)= A[M,K] * B[K,N]      A, B bf16, fp32 accumulation in TMEM
//
// with either operand K-major (stored [rows = M or N][K], K contiguous -- the forward scan's
// layout) or MN-major (stored [K][M or N], the M / N index contiguous).  MN-major operands are what
// makes the backward transposition-free: with P = dL/dz [rows, table rows] (bf16, recomputed by the
// scan kernel's grad epilogue from the saved log-sum-exp),
//     dL/dq     = P * T        A = P K-major,            B = T  stored [K = table rows][N = D]: MN-major
//     dL/dtable = P^T * q      A = P stored [K = rows][M = table rows]: MN-major,  B = q MN-major
// so ONE copy of P and the table / query matrices as they lie in memory feed both products.
//
// Same machinery as the forward kernel (scan_tc_kernel.cuh): TMA (SWIZZLE_128B, OOB zero fill for
// ragged M / N / K) into a 4-stage mbarrier ring, one thread issuing tcgen05.mma (cta_group::1,
// M128 N256 K16), two TMEM accumulator stages so that the epilogue of tile t overlaps the MMAs of
// tile t+1, eight epilogue warps (lane quarter x column half) draining with tcgen05.ld.
// An MN-major operand is staged as boxes of 64 (M/N) x 64 (K) elements = 8 KB: rows of 128 bytes
// along M/N, 8-row swizzle atoms of 1024 bytes along K (the descriptor's stride byte offset),
// consecutive 64-wide M/N groups 8 KB apart (its leading byte offset); one K16 MMA step advances the
// start address by 16 rows = 2 KB.  (CUTLASS make_umma_desc<Major::MN>, LayoutType::B128.)
#include <cuda.h>
#include <stdio.h>
#include <atomic>
#include "kernels.h"

namespace mcl {

constexpr int kGemmEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kGemmEpiWarps;
constexpr uint32_t kGemmABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr uint32_t kGemmBBytes = kBlockN * kBlockK * 2;   // 32 KB
constexpr uint32_t kGemmStageBytes = kGemmABytes + kGemmBBytes;
constexpr uint32_t kGemmStages = 4;
constexpr uint32_t kGemmSmemBytes = kGemmStages * kGemmStageBytes + 256 + 1024;
constexpr uint32_t kMnGroupBytes = 64 * kBlockK * 2;      // one 64 x 64 box of an MN-major operand

struct GemmParams {
  int M, N, K;
  float* C;              // fp32 output; or bf16 (out_bf16) -- then the pointer is a __nv_bfloat16*
  long long ldc;
  int accumulate;        // 1: C += A*B, 0: C = A*B
  int m_tiles, n_tiles;
  // split K: work item w = (tile w % tiles, K blocks [ (w / tiles) * kb_per, ... + kb_per) ); with
  // ksplit > 1 the partial products are added with red.global.add.f32 (C zeroed by the launcher).
  // A product with few output tiles and a long K -- dL/dq = P T for a handful of labelled rows: 5
  // tiles, K = 262 k table rows -- otherwise runs on 5 of 148 SMs (measured: 1357 us of a 2.5 ms
  // backward at the reference's native shape).
  int ksplit, kb_per;
  int out_bf16;          // round the fp32 accumulators to bf16 on the way out (ksplit = 1, no accumulate)
};

// MN-major SWIZZLE_128B operand: LBO = 8 KB between 64-wide M/N groups, SBO = 1 KB between 8-row K groups
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(kMnGroupBytes >> 4) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <bool kAMN, bool kBMN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kGemmStages * kGemmStageBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kGemmStages + s); };
  auto tfull_bar = [&](uint32_t a) { return bar_base + 8u * (2 * kGemmStages + a); };
  auto tempty_bar = [&](uint32_t a) { return bar_base + 8u * (2 * kGemmStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kGemmStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kGemmStages * kGemmStageBytes + 8u * (2 * kGemmStages + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (uint32_t s = 0; s < kGemmStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (uint32_t a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kGemmEpiWarps); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int tiles = p.m_tiles * p.n_tiles;
  const int items = tiles * p.ksplit;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int t = w % tiles, ks = w / tiles;
        const int mb = t / p.n_tiles, nb = t - mb * p.n_tiles;     // tiles of one row block run side by side
        const int kb0 = ks * p.kb_per, kb1 = min(num_kb, kb0 + p.kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_backoff(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * kGemmStageBytes;
          const uint32_t sb = sa + kGemmABytes;
          mbar_expect_tx(full_bar(stage), kGemmStageBytes);
          if (!kAMN) {
            tma_load_2d(sa, &tm_a, full_bar(stage), kb * kBlockK, mb * kBlockM, kL2EvictNormal);
          } else {
#pragma unroll
            for (int g = 0; g < kBlockM / 64; ++g)
              tma_load_2d(sa + g * kMnGroupBytes, &tm_a, full_bar(stage), mb * kBlockM + g * 64, kb * kBlockK,
                          kL2EvictNormal);
          }
          if (!kBMN) {
            tma_load_2d(sb, &tm_b, full_bar(stage), kb * kBlockK, nb * kBlockN, kL2EvictNormal);
          } else {
#pragma unroll
            for (int g = 0; g < kBlockN / 64; ++g)
              tma_load_2d(sb + g * kMnGroupBytes, &tm_b, full_bar(stage), nb * kBlockN + g * 64, kb * kBlockK,
                          kL2EvictNormal);
          }
          if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kBlockN) | (kAMN ? (1u << 15) : 0u) | (kBMN ? (1u << 16) : 0u);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int kb0 = (w / tiles) * p.kb_per, kb1 = min(num_kb, kb0 + p.kb_per);
        mbar_wait_backoff(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlockN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kGemmStageBytes;
          const uint64_t adesc = kAMN ? umma_desc_mn_sw128(sa) : umma_desc_sw128(sa);
          const uint64_t bdesc = kBMN ? umma_desc_mn_sw128(sa + kGemmABytes) : umma_desc_sw128(sa + kGemmABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // K-major: +32 B inside the 128 B row; MN-major: +16 rows of 128 B (units of 16 B)
            const uint64_t ad = adesc + (kAMN ? 128u * k : 2u * k);
            const uint64_t bd = bdesc + (kBMN ? 128u * k : 2u * k);
            umma_bf16(d_tmem, ad, bd, idesc, (uint32_t)(kb != kb0 || k != 0));
          }
          umma_commit(empty_bar(stage));
          if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1u; if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    constexpr int kHalfN = kBlockN / 2;
    uint32_t acc = 0, acc_phase = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
      const int t = w % tiles;
      const int mb = t / p.n_tiles, nb = t - mb * p.n_tiles;
      const long long row = (long long)mb * kBlockM + quarter * 32 + lane;
      const int col_base = nb * kBlockN + half * kHalfN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kBlockN + half * kHalfN;
      float* crow = p.C + row * p.ldc;
#pragma unroll 1
      for (int c = 0; c < kHalfN / kChunk; ++c) {
        float y[kChunk];
        __syncwarp();
        tmem_ld_issue(taddr + c * kChunk, y);
        tmem_ld_wait(y);
        if (c + 1 == kHalfN / kChunk) {           // every tcgen05.ld of this tile has landed
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        const int col0 = col_base + c * kChunk;
        if (row < p.M && col0 < p.N && p.out_bf16) {
          __nv_bfloat16* brow = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc;
          if (col0 + kChunk <= p.N && (p.ldc & 7) == 0) {
            uint4* dst = reinterpret_cast<uint4*>(brow + col0);
#pragma unroll
            for (int i = 0; i < kChunk / 8; ++i) {
              uint4 v;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
              for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(y[8 * i + 2 * j], y[8 * i + 2 * j + 1]);
              dst[i] = v;
            }
          } else {
#pragma unroll
            for (int i = 0; i < kChunk; ++i)
              if (col0 + i < p.N) brow[col0 + i] = __float2bfloat16_rn(y[i]);
          }
        } else if (row < p.M && col0 < p.N && p.ksplit > 1) {
#pragma unroll
          for (int i = 0; i < kChunk; ++i)
            if (col0 + i < p.N) atomicAdd(crow + col0 + i, y[i]);   // (result unused: red.global.add.f32)
        } else if (row < p.M && col0 < p.N) {
          if (col0 + kChunk <= p.N && (p.ldc & 3) == 0) {
            float4* dst = reinterpret_cast<float4*>(crow + col0);
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) {
              float4 v = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
              if (p.accumulate) {
                const float4 o = dst[i];
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              dst[i] = v;
            }
          } else {
#pragma unroll
            for (int i = 0; i < kChunk; ++i)
              if (col0 + i < p.N) crow[col0 + i] = p.accumulate ? crow[col0 + i] + y[i] : y[i];
          }
        }
      }
      acc ^= 1u; if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// C[M,N] (+)= A * B.  a: K-major -> [M][K] with pitch lda, MN-major -> [K][M] with pitch lda; same for b.
cudaError_t launch_gemm_tc(const void* a, int a_mn, int64_t lda, const void* b, int b_mn, int64_t ldb,
                           float* c, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate,
                           int sm_count, cudaStream_t s, int out_bf16) {
  if (M == 0 || N == 0) return cudaSuccess;
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev].load()) {
    const void* kernels[4] = {(const void*)gemm_tc_kernel<false, false>, (const void*)gemm_tc_kernel<false, true>,
                              (const void*)gemm_tc_kernel<true, false>, (const void*)gemm_tc_kernel<true, true>};
    for (const void* kfn : kernels) {
      cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes);
      if (e != cudaSuccess) return e;
    }
    attr_set[dev].store(true);
  }
  CUtensorMap tm_a, tm_b;
  // K-major: rows = M (or N), cols = K, box 128 (256) rows x 64; MN-major: rows = K, cols = M (or N), box 64 x 64
  const bool ok_a = a_mn ? make_tmap_bf16(&tm_a, a, K, M, lda, 64) : make_tmap_bf16(&tm_a, a, M, K, lda, kBlockM);
  const bool ok_b = b_mn ? make_tmap_bf16(&tm_b, b, K, N, ldb, 64) : make_tmap_bf16(&tm_b, b, N, K, ldb, kBlockN);
  if (!ok_a || !ok_b) return cudaErrorInvalidValue;
  GemmParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.C = c; p.ldc = ldc; p.accumulate = accumulate;
  p.m_tiles = (int)((M + kBlockM - 1) / kBlockM);
  p.n_tiles = (int)((N + kBlockN - 1) / kBlockN);
  const int tiles = p.m_tiles * p.n_tiles;
  const int num_kb = (int)((K + kBlockK - 1) / kBlockK);
  p.out_bf16 = out_bf16;
  p.ksplit = 1; p.kb_per = num_kb;
  if (!out_bf16 && tiles * 2 <= sm_count && num_kb >= 16) {       // few output tiles, long K: split K over the idle SMs
    int ks = sm_count / tiles;
    if (ks > num_kb / 4) ks = num_kb / 4;
    p.kb_per = (num_kb + ks - 1) / ks;
    p.ksplit = (num_kb + p.kb_per - 1) / p.kb_per;                 // no empty split
    if (p.ksplit > 1 && !accumulate) {
      cudaError_t e = cudaMemset2DAsync(c, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, s);
      if (e != cudaSuccess) return e;
    }
  }
  if (out_bf16 && accumulate) return cudaErrorInvalidValue;
  const int items = tiles * p.ksplit;
  const int grid = items < sm_count ? items : sm_count;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = kGemmSmemBytes;
  cfg.stream = s;
  if (a_mn && b_mn) return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, true>, tm_a, tm_b, p);
  if (a_mn) return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, false>, tm_a, tm_b, p);
  if (b_mn) return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, true>, tm_a, tm_b, p);
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, false>, tm_a, tm_b, p);
}

}  // namespace mcl
