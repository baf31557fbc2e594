// fp32 CHECK PATH of the fused cross-entropy's backward (MCL_DTYPE_F32 inputs: the vision loop's
// fp16 / fp32 classifier head, and the rtol 1e-4 gradient checks): plain CUDA-core kernels, the
// same three products as the tcgen05 path -- Z = q T^T, dL/dq = P T, dL/dT = P^T q -- through ONE
// strided SIMT GEMM, plus the in-place Z -> P = dL/dz transform.  Not a product path for speed.
#include "kernels.h"

namespace mcl {

constexpr int kSgTile = 64, kSgK = 16;

// C[M,N] (+)= sum_k A(m,k) B(k,n), A(m,k) = a[m*sa_m + k*sa_k], B(k,n) = b[k*sb_k + n*sb_n].
// 16 x 16 threads, 4 x 4 outputs each, 64 x 64 x 16 shared-memory tiles.
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ a, long long sa_m, long long sa_k, const float* __restrict__ b,
                 long long sb_k, long long sb_n, float* __restrict__ c, long long ldc, int M, int N, int K,
                 int accumulate) {
  __shared__ float As[kSgK][kSgTile + 1], Bs[kSgK][kSgTile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * kSgTile, n0 = blockIdx.x * kSgTile;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kSgK) {
    for (int e = threadIdx.x; e < kSgTile * kSgK; e += 256) {
      const int kk = e / kSgTile, mm = e % kSgTile;
      const int m = m0 + mm, n = n0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? a[(long long)m * sa_m + (long long)k * sa_k] : 0.f;
      Bs[kk][mm] = (n < N && k < K) ? b[(long long)k * sb_k + (long long)n * sb_n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSgK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float* dst = c + (long long)m * ldc + n;
        *dst = accumulate ? *dst + acc[i][j] : acc[i][j];
      }
    }
}

// z[r, c] (raw dot products of columns col_base + c) -> dL/dz in place.
__global__ void __launch_bounds__(256)
dz_simt_kernel(float* __restrict__ z, long long ldz, long long rows, long long cols, long long col_base,
               const float* __restrict__ lse, const long long* __restrict__ labels, float scale, float softcap,
               float eps_over_v, float one_minus_eps, const float* __restrict__ grad_loss, float grad_coef) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = blockIdx.y;
  if (c >= cols || r >= rows) return;
  const long long lab = labels[r];
  float out = 0.f;
  if (lab != -100) {
    float v = z[r * ldz + c] * scale, d = 1.f;
    if (softcap > 0.f) {
      const float t = tanhf(v / softcap);
      v = softcap * t;
      d = 1.f - t * t;
    }
    float x = expf(v - lse[r]) - eps_over_v;
    if (col_base + c == lab) x -= one_minus_eps;
    out = x * (*grad_loss) * grad_coef * d;
  }
  z[r * ldz + c] = out;
}

cudaError_t launch_gemm_simt(const float* a, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_k, int64_t sb_n,
                             float* c, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate, cudaStream_t s) {
  if (M == 0 || N == 0) return cudaSuccess;
  dim3 grid((unsigned)((N + kSgTile - 1) / kSgTile), (unsigned)((M + kSgTile - 1) / kSgTile));
  gemm_simt_kernel<<<grid, 256, 0, s>>>(a, sa_m, sa_k, b, sb_k, sb_n, c, ldc, (int)M, (int)N, (int)K, accumulate);
  return cudaGetLastError();
}

cudaError_t launch_dz_simt(float* z, int64_t ldz, int64_t rows, int64_t cols, int64_t col_base, const float* lse,
                           const int64_t* labels, float scale, float softcap, float eps_over_v, float one_minus_eps,
                           const float* grad_loss, float grad_coef, cudaStream_t s) {
  if (rows == 0 || cols == 0) return cudaSuccess;
  dim3 grid((unsigned)((cols + 255) / 256), (unsigned)rows);
  dz_simt_kernel<<<grid, 256, 0, s>>>(z, ldz, rows, cols, col_base, lse, (const long long*)labels, scale, softcap,
                                      eps_over_v, one_minus_eps, grad_loss, grad_coef);
  return cudaGetLastError();
}

}  // namespace mcl
