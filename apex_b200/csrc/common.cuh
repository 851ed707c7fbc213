// apex_b200 — shared device utilities for the sm_100a kernels (torch-free; C ABI launchers live in the .cu files).
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>

#define AB_API extern "C" __attribute__((visibility("default")))

namespace ab {

using bf16 = __nv_bfloat16;
using f16 = __half;

// dtype codes shared with python (apex_b200/_lib.py)
enum DType : int { kF32 = 0, kF16 = 1, kBF16 = 2, kF64 = 3, kU8 = 4, kI32 = 5, kI64 = 6, kI16 = 7, kE4M3 = 8, kE5M2 = 9 };

constexpr int kNumSMs = 148;  // B200

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<f16>(f16 v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<double>(double v) { return (float)v; }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ f16 from_f<f16>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ double from_f<double>(float v) { return (double)v; }

// ---- V-element vector access (V*sizeof(T) is 16 or 32 bytes; issued as 16-byte LDG/STG) -------------------------
template <typename T, int V> struct alignas(sizeof(T) * V > 16 ? 16 : sizeof(T) * V) Pack { T v[V]; };

// V elements of type T packed in 32-bit words -> fp32. bf16 by hand: (w << 16) and (w & 0xffff0000) are ONE instruction per element; the
// library conversion of the high half of a word compiles to a PRMT plus a shift, and the 16-bit row kernels are issue-bound before
// they are byte-bound (LayerNorm forward: 19 issued instructions per element, profiles/ln_fwd_now.md).
template <typename T, int V> __device__ __forceinline__ void words_to_float(const uint32_t* w, float (&r)[V]) {
  if constexpr (std::is_same<T, bf16>::value && V % 2 == 0) {
#pragma unroll
    for (int q = 0; q < V / 2; q++) { r[2 * q] = __uint_as_float(w[q] << 16); r[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u); }
  } else if constexpr (std::is_same<T, f16>::value && V % 2 == 0) {
#pragma unroll
    for (int q = 0; q < V / 2; q++) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
      r[2 * q] = t.x; r[2 * q + 1] = t.y;
    }
  } else {
    const T* e = reinterpret_cast<const T*>(w);
#pragma unroll
    for (int i = 0; i < V; i++) r[i] = to_f<T>(e[i]);
  }
}

template <typename T, int V>
__device__ __forceinline__ void load_vec(float (&r)[V], const T* __restrict__ p) {
  constexpr int BYTES = sizeof(T) * V;
  static_assert(BYTES % 16 == 0 || BYTES == 8 || BYTES == 4, "vector width");
  if constexpr (BYTES == 4) {
    uint32_t raw = *reinterpret_cast<const uint32_t*>(p);
    words_to_float<T, V>(&raw, r);
  } else if constexpr (BYTES == 8) {
    uint2 raw = *reinterpret_cast<const uint2*>(p);
    words_to_float<T, V>(reinterpret_cast<const uint32_t*>(&raw), r);
  } else {
    constexpr int N16 = BYTES / 16;
    uint4 raw[N16];
#pragma unroll
    for (int i = 0; i < N16; i++) raw[i] = reinterpret_cast<const uint4*>(p)[i];
    words_to_float<T, V>(reinterpret_cast<const uint32_t*>(raw), r);
  }
}

template <typename T, int V>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&r)[V]) {
  constexpr int BYTES = sizeof(T) * V;
  if constexpr (BYTES == 4) {
    uint32_t raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < V; i++) e[i] = from_f<T>(r[i]);
    *reinterpret_cast<uint32_t*>(p) = raw;
  } else if constexpr (BYTES == 8) {
    uint2 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < V; i++) e[i] = from_f<T>(r[i]);
    *reinterpret_cast<uint2*>(p) = raw;
  } else {
    constexpr int N16 = BYTES / 16;
    uint4 raw[N16];
    T* e = reinterpret_cast<T*>(raw);
#pragma unroll
    for (int i = 0; i < V; i++) e[i] = from_f<T>(r[i]);
#pragma unroll
    for (int i = 0; i < N16; i++) reinterpret_cast<uint4*>(p)[i] = raw[i];
  }
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- warp / block reductions -------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum, result valid in every thread. `red` is >= 33 floats of shared memory. Safe to call repeatedly.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from the previous call's readers
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

__device__ __forceinline__ bool finite_f(float v) { return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u; }

}  // namespace ab

// ---- dtype dispatch helpers (host) ----------------------------------------------------------------------------------
#define AB_DISPATCH_FLOAT3(code, NAME, ...)                         \
  switch (code) {                                                   \
    case ab::kF32: { using NAME = float; __VA_ARGS__; break; }      \
    case ab::kF16: { using NAME = ab::f16; __VA_ARGS__; break; }    \
    case ab::kBF16: { using NAME = ab::bf16; __VA_ARGS__; break; }  \
    default: return -1;                                             \
  }

#define AB_CHECK_LAUNCH() \
  do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)
