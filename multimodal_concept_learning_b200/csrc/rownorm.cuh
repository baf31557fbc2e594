// Device helpers shared by the HBM-bound row kernels (rowops.cu) and the scan kernel, which
// computes the inverse norms of a SMALL query batch itself (one launch less where launches bound
// the step): 16-byte vector access of a row, warp reductions, and the inverse L2 norm of one row
// by one warp -- one function, so that both callers produce bit-identical values.
#pragma once
#include <float.h>
#include "common.cuh"

namespace mcl {

constexpr float kTinyNorm = 10.0f * FLT_EPSILON;  // sklearn: norms below this become 1

template <typename T> struct Vec;  // 16-byte vector of T
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = __bfloat1622float2(h[i]);
      f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
  }
  __device__ static __forceinline__ void widen(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = __bfloat1622float2(h[i]);
      f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
  __device__ static __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static __forceinline__ void widen(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ static __forceinline__ void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
  __device__ static __forceinline__ float ld1(const float* p) { return *p; }
  __device__ static __forceinline__ void st1(float* p, float v) { *p = v; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 1 / ||p[0 .. dim)|| by one warp (all 32 lanes call, every lane gets the result); rows with a
// norm below 10 * FLT_EPSILON give 1 (scikit-learn `normalize`).
template <typename T>
__device__ __forceinline__ float warp_row_inv_norm(const T* __restrict__ p, int dim, int lane) {
  constexpr int N = Vec<T>::N;
  const int nvec = dim / N;
  float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
  int v = lane;
  for (; v + 96 < nvec; v += 128) {  // four independent 16-byte loads in flight per lane
    float a[N], b[N], c[N], d[N];
    Vec<T>::load(p + (size_t)v * N, a);
    Vec<T>::load(p + (size_t)(v + 32) * N, b);
    Vec<T>::load(p + (size_t)(v + 64) * N, c);
    Vec<T>::load(p + (size_t)(v + 96) * N, d);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      ss0 = fmaf(a[i], a[i], ss0); ss1 = fmaf(b[i], b[i], ss1);
      ss2 = fmaf(c[i], c[i], ss2); ss3 = fmaf(d[i], d[i], ss3);
    }
  }
  for (; v < nvec; v += 32) {
    float a[N];
    Vec<T>::load(p + (size_t)v * N, a);
#pragma unroll
    for (int i = 0; i < N; ++i) ss0 = fmaf(a[i], a[i], ss0);
  }
  for (int e = nvec * N + lane; e < dim; e += 32) {
    const float a = Vec<T>::ld1(p + e);
    ss1 = fmaf(a, a, ss1);
  }
  const float n = sqrtf(warp_sum((ss0 + ss1) + (ss2 + ss3)));
  return (n < kTinyNorm) ? 1.0f : 1.0f / n;
}

// R rows at once (p, p + stride, ...): the loads of all R rows are issued before any is reduced, so
// a warp that needs the norms of several rows pays one memory round trip per R rows.  Per row the
// arithmetic and its order are those of warp_row_inv_norm: the results are bit-identical.
template <typename T, int R>
__device__ __forceinline__ void warp_rows_inv_norm(const T* __restrict__ p, long long stride, int nrows,
                                                   int dim, int lane, float (&out)[R]) {
  constexpr int N = Vec<T>::N;
  const int nvec = dim / N;
  float ss0[R], ss1[R], ss2[R], ss3[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { ss0[r] = 0.f; ss1[r] = 0.f; ss2[r] = 0.f; ss3[r] = 0.f; }
  int v = lane;
  for (; v + 96 < nvec; v += 128) {
    float a[R][N], b[R][N], c[R][N], d[R][N];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const T* q = p + (r < nrows ? r : 0) * stride;
      Vec<T>::load(q + (size_t)v * N, a[r]);
      Vec<T>::load(q + (size_t)(v + 32) * N, b[r]);
      Vec<T>::load(q + (size_t)(v + 64) * N, c[r]);
      Vec<T>::load(q + (size_t)(v + 96) * N, d[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        ss0[r] = fmaf(a[r][i], a[r][i], ss0[r]); ss1[r] = fmaf(b[r][i], b[r][i], ss1[r]);
        ss2[r] = fmaf(c[r][i], c[r][i], ss2[r]); ss3[r] = fmaf(d[r][i], d[r][i], ss3[r]);
      }
  }
  for (; v < nvec; v += 32) {
    float a[R][N];
#pragma unroll
    for (int r = 0; r < R; ++r) Vec<T>::load(p + (r < nrows ? r : 0) * stride + (size_t)v * N, a[r]);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < N; ++i) ss0[r] = fmaf(a[r][i], a[r][i], ss0[r]);
  }
  for (int e = nvec * N + lane; e < dim; e += 32) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float a = Vec<T>::ld1(p + (r < nrows ? r : 0) * stride + e);
      ss1[r] = fmaf(a, a, ss1[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float n = sqrtf(warp_sum((ss0[r] + ss1[r]) + (ss2[r] + ss3[r])));
    out[r] = (n < kTinyNorm) ? 1.0f : 1.0f / n;
  }
}

}  // namespace mcl
