// Per-tensor dynamic fp8 quantisation for the fp8 GEMM path: amax in one pass, scale + cast in a second, everything stays on the
// device (the dequantisation scale is a device float the GEMM epilogue reads), no host synchronisation.
//   q = cast_fp8(x * (FMAX / amax)),   inv_scale = amax / FMAX,   x ~= q * inv_scale
#include "common.cuh"

namespace ab {

template <typename T>
__global__ void __launch_bounds__(512) amax_kernel(const T* __restrict__ x, long long n, unsigned int* __restrict__ amax_bits) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float red[40];
  float m = 0.f;
  const long long nv = n / V;
  for (long long i = (long long)blockIdx.x * 512 + threadIdx.x; i < nv; i += (long long)gridDim.x * 512) {
    float f[V]; load_vec<T, V>(f, x + i * V);
#pragma unroll
    for (int j = 0; j < V; j++) m = fmaxf(m, fabsf(f[j]));
  }
  for (long long i = nv * V + (long long)blockIdx.x * 512 + threadIdx.x; i < n; i += (long long)gridDim.x * 512) m = fmaxf(m, fabsf(to_f<T>(x[i])));
  m = block_max(m, red);
  if (threadIdx.x == 0) atomicMax(amax_bits, __float_as_uint(m));  // non-negative floats order like their bit patterns
}

template <typename T, typename Q>
__global__ void __launch_bounds__(512) quant_kernel(const T* __restrict__ x, Q* __restrict__ q, long long n, const unsigned int* __restrict__ amax_bits,
                                                   float fmax, float* __restrict__ inv_scale) {
  constexpr int V = 16 / sizeof(T);
  const float amax = fmaxf(__uint_as_float(*amax_bits), 1e-12f);
  const float scale = fmax / amax;
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = amax / fmax;
  const long long nv = n / V;
  for (long long i = (long long)blockIdx.x * 512 + threadIdx.x; i < nv; i += (long long)gridDim.x * 512) {
    float f[V]; load_vec<T, V>(f, x + i * V);
    Q o[V];
#pragma unroll
    for (int j = 0; j < V; j++) o[j] = Q(f[j] * scale);
    if constexpr (V == 8) *reinterpret_cast<uint2*>(q + i * V) = *reinterpret_cast<const uint2*>(o);
    else *reinterpret_cast<uint32_t*>(q + i * V) = *reinterpret_cast<const uint32_t*>(o);
  }
  for (long long i = nv * V + (long long)blockIdx.x * 512 + threadIdx.x; i < n; i += (long long)gridDim.x * 512) q[i] = Q(to_f<T>(x[i]) * scale);
}

// Quantise a row-major [R, C] 16-bit / fp32 matrix with ONE read into its fp8 copy q [R, C] and / or its TRANSPOSED fp8 copy qt [C, R]
// (the K-major operands the backward GEMMs need: dgrad reduces over N, wgrad over M, and kind::f8f6f4 takes 8-bit operands K-major only).
// 64 x 64 tiles through shared memory: 16-byte loads along C, 16-byte stores along R of the transposed copy.
template <typename T, typename Q>
__global__ void __launch_bounds__(256) quant_dual_kernel(const T* __restrict__ x, Q* __restrict__ q, Q* __restrict__ qt, int R, int C,
                                                        const unsigned int* __restrict__ amax_bits, float fmax, float* __restrict__ inv_scale) {
  __shared__ __align__(16) unsigned char tile[64][64 + 16];
  const float amax = fmaxf(__uint_as_float(*amax_bits), 1e-12f);
  const float scale = fmax / amax;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *inv_scale = amax / fmax;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  // load + convert: thread t handles row t / 4 (and + 32... two rows per pass), 16 columns (two 8-element vectors for 16-bit types)
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {        // 64 rows x 8 groups of 8 columns
    const int r = i >> 3, cg = (i & 7) * 8;
    const int gr = r0 + r, gc = c0 + cg;
    Q o[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float v = (gr < R && gc + j < C) ? to_f<T>(x[(size_t)gr * C + gc + j]) : 0.f;
      o[j] = Q(v * scale);
    }
    *reinterpret_cast<uint2*>(&tile[r][cg]) = *reinterpret_cast<const uint2*>(o);
    if (q && gr < R) {
      if (gc + 8 <= C && (C % 8) == 0) *reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(q) + (size_t)gr * C + gc) = *reinterpret_cast<const uint2*>(o);
      else for (int j = 0; j < 8; j++) if (gc + j < C) q[(size_t)gr * C + gc + j] = o[j];
    }
  }
  __syncthreads();
  if (qt) {
    for (int i = threadIdx.x; i < 64 * 4; i += 256) {       // 64 output rows (= input columns) x 4 groups of 16 bytes (= 16 input rows)
      const int c = i >> 2, rg = (i & 3) * 16;
      const int gc = c0 + c, gr = r0 + rg;
      if (gc >= C) continue;
      unsigned char o[16];
#pragma unroll
      for (int j = 0; j < 16; j++) o[j] = tile[rg + j][c];
      unsigned char* dst = reinterpret_cast<unsigned char*>(qt) + (size_t)gc * R + gr;
      if (gr + 16 <= R && (R % 16) == 0) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
      else for (int j = 0; j < 16; j++) if (gr + j < R) dst[j] = o[j];
    }
  }
}

}  // namespace ab

using namespace ab;

// x [R, C] row-major (f32 / f16 / bf16) -> q [R, C] and / or qt [C, R] (kE4M3 / kE5M2; either may be null), one shared per-tensor scale.
AB_API int ab_fp8_quantize_dual(const void* x, void* q, void* qt, int R, int C, void* amax_scratch, float* inv_scale, int dt_in, int dt_q,
                                cudaStream_t st) {
  if (R <= 0 || C <= 0) return 0;
  if (dt_q != kE4M3 && dt_q != kE5M2) return -2;
  cudaError_t e = cudaMemsetAsync(amax_scratch, 0, 4, st);
  if (e != cudaSuccess) return (int)e;
  const long long n = (long long)R * C;
  const long long want = (n / 8 + 511) / 512;
  const int grid = (int)(want < kNumSMs * 4 ? (want < 1 ? 1 : want) : kNumSMs * 4);
  unsigned int* ab = reinterpret_cast<unsigned int*>(amax_scratch);
  const dim3 tg((C + 63) / 64, (R + 63) / 64);
#define FQD_GO(T)                                                                                                                    \
  do {                                                                                                                               \
    if (!aligned16(x)) return -3;                                                                                                    \
    amax_kernel<T><<<grid, 512, 0, st>>>((const T*)x, n, ab);                                                                        \
    if (dt_q == kE4M3) quant_dual_kernel<T, __nv_fp8_e4m3><<<tg, 256, 0, st>>>((const T*)x, (__nv_fp8_e4m3*)q, (__nv_fp8_e4m3*)qt, R, C, ab, 448.f, inv_scale); \
    else quant_dual_kernel<T, __nv_fp8_e5m2><<<tg, 256, 0, st>>>((const T*)x, (__nv_fp8_e5m2*)q, (__nv_fp8_e5m2*)qt, R, C, ab, 57344.f, inv_scale);            \
  } while (0)
  if (dt_in == kF32) FQD_GO(float);
  else if (dt_in == kF16) FQD_GO(f16);
  else if (dt_in == kBF16) FQD_GO(bf16);
  else return -2;
  return (int)cudaGetLastError();
}

// x [n] (f32 / f16 / bf16, 16-byte aligned) -> q [n] (kE4M3 or kE5M2); amax_scratch: one uint (zeroed here); inv_scale: one float.
AB_API int ab_fp8_quantize(const void* x, void* q, long long n, void* amax_scratch, float* inv_scale, int dt_in, int dt_q, cudaStream_t st) {
  if (n <= 0) return 0;
  if (!aligned16(x) || ((uintptr_t)q % 8) != 0) return -3;
  cudaError_t e = cudaMemsetAsync(amax_scratch, 0, 4, st);
  if (e != cudaSuccess) return (int)e;
  const long long want = (n / 8 + 511) / 512;
  const int grid = (int)(want < kNumSMs * 4 ? (want < 1 ? 1 : want) : kNumSMs * 4);
  unsigned int* ab = reinterpret_cast<unsigned int*>(amax_scratch);
#define FQ_GO(T)                                                                                                      \
  do {                                                                                                                \
    amax_kernel<T><<<grid, 512, 0, st>>>((const T*)x, n, ab);                                                         \
    if (dt_q == kE4M3) quant_kernel<T, __nv_fp8_e4m3><<<grid, 512, 0, st>>>((const T*)x, (__nv_fp8_e4m3*)q, n, ab, 448.f, inv_scale); \
    else quant_kernel<T, __nv_fp8_e5m2><<<grid, 512, 0, st>>>((const T*)x, (__nv_fp8_e5m2*)q, n, ab, 57344.f, inv_scale);             \
  } while (0)
  if (dt_q != kE4M3 && dt_q != kE5M2) return -2;
  if (dt_in == kF32) FQ_GO(float);
  else if (dt_in == kF16) FQ_GO(f16);
  else if (dt_in == kBF16) FQ_GO(bf16);
  else return -2;
  return (int)cudaGetLastError();
}
