// Shared pieces of the row-normalisation kernels (LayerNorm / RMSNorm fwd+bwd).
#pragma once
#include "common.cuh"
#include <type_traits>

namespace ab {

// Sum across the `tpr` threads that share a row. tpr is a power of two in [8, 1024]; rows never straddle a warp when
// tpr < 32. For tpr > 32 partial sums go through shared memory; `buf` alternates between two banks so one barrier per
// reduction is enough. EVERY thread of the CTA must call this (uniform control flow).
struct RowReducer {
  float* smem;  // [2 banks][64] floats (sum / maxv use the first 32 of a bank, sum2 all 64)
  int tpr, lane_r, rg, bank;
  __device__ __forceinline__ RowReducer(float* s, int tpr_) : smem(s), tpr(tpr_), bank(0) {
    lane_r = threadIdx.x % tpr_;
    rg = threadIdx.x / tpr_;
  }
  __device__ __forceinline__ float sum(float v) {
    if (tpr <= 32) {
      for (int o = tpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      return v;
    }
    v = warp_sum(v);
    const int wpr = tpr >> 5;            // warps per row
    const int w = threadIdx.x >> 5;      // warp in CTA; rows own consecutive warps
    float* b = smem + bank * 64;
    if ((threadIdx.x & 31) == 0) b[w] = v;
    __syncthreads();
    const int l = threadIdx.x & 31, w0 = rg * wpr;
    float t = l < wpr ? b[w0 + l] : 0.f;
    for (int o = wpr >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    t = __shfl_sync(0xffffffffu, t, 0);
    bank ^= 1;
    return t;
  }
  // two sums with one barrier (shared memory: [2 banks][2 values][32 warps])
  __device__ __forceinline__ void sum2(float& a, float& c) {
    if (tpr <= 32) {
      for (int o = tpr >> 1; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
      return;
    }
    a = warp_sum(a); c = warp_sum(c);
    const int wpr = tpr >> 5;
    const int w = threadIdx.x >> 5;
    float* b = smem + bank * 64;
    if ((threadIdx.x & 31) == 0) { b[w] = a; b[32 + w] = c; }
    __syncthreads();
    // lanes [0, wpr) of every warp pick one warp partial each, xor-shuffles finish (wpr is a power of two <= 32)
    const int l = threadIdx.x & 31, w0 = rg * wpr;
    float ta = l < wpr ? b[w0 + l] : 0.f, tc = l < wpr ? b[32 + w0 + l] : 0.f;
    for (int o = wpr >> 1; o > 0; o >>= 1) { ta += __shfl_xor_sync(0xffffffffu, ta, o); tc += __shfl_xor_sync(0xffffffffu, tc, o); }
    ta = __shfl_sync(0xffffffffu, ta, 0); tc = __shfl_sync(0xffffffffu, tc, 0);
    bank ^= 1;
    a = ta; c = tc;
  }
  __device__ __forceinline__ float maxv(float v) {
    if (tpr <= 32) {
      for (int o = tpr >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      return v;
    }
    v = warp_max(v);
    const int wpr = tpr >> 5;
    const int w = threadIdx.x >> 5;
    float* b = smem + bank * 64;
    if ((threadIdx.x & 31) == 0) b[w] = v;
    __syncthreads();
    float t = -INFINITY;
    const int w0 = rg * wpr;
    for (int i = 0; i < wpr; i++) t = fmaxf(t, b[w0 + i]);
    bank ^= 1;
    return t;
  }
};

template <typename T> __device__ __forceinline__ void unpack16(const uint4& raw, float (&r)[16 / sizeof(T)]) {
  words_to_float<T, (int)(16 / sizeof(T))>(reinterpret_cast<const uint32_t*>(&raw), r);
}

template <typename T, int N> __device__ __forceinline__ void decode_words(const uint32_t* w, float (&f)[N]) { words_to_float<T, N>(w, f); }
__device__ __forceinline__ float rsqrt_fast(float v) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }  // v = var + eps > 0, never denormal

struct NormCfg { int tpr, maxv, threads, rows_per_cta; bool ok; };

// Pick threads-per-row / vectors-per-thread for a row of `nvec` 16-byte vectors.
inline NormCfg norm_cfg(int nvec, int target_v = 4, int max_tpr = 1024) {
  NormCfg c{};
  int want = (nvec + target_v - 1) / target_v;
  int tpr = 8;
  while (tpr < want) tpr <<= 1;
  if (tpr > max_tpr) tpr = max_tpr;
  int mv = (nvec + tpr - 1) / tpr;
  int maxv = 1;
  while (maxv < mv) maxv <<= 1;
  c.tpr = tpr; c.maxv = maxv;
  c.threads = tpr > 256 ? tpr : 256;
  c.rows_per_cta = c.threads / tpr;
  c.ok = maxv <= 8;
  return c;
}

}  // namespace ab
